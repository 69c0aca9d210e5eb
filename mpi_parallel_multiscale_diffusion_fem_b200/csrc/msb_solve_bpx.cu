// msb_solve_bpx.cu -- shared-memory-resident PCG with a multilevel (BPX / multilevel
// diagonal scaling) preconditioner.  Default kernel of the shared-memory tier.
//
// Same system as msb_solve_smem.cu: the condensed interior block K_II phi_I = -K_IB g_B of
// one coarse cell (diffusion_problem_basis.tpp:450-465), scaled to unit diagonal.  The
// reference preconditions with SSOR(1.6), whose sweeps are sequential in DoF order; its
// parallel (4-colour) form loses most of the benefit (measured offline: 137 instead of 47
// iterations at n=64).  What maps to one CTA with everything on chip is the additive
// multilevel preconditioner
//       M^-1 = D^-1 + sum_{l>=1} P_l D_l^-1 P_l^T ,   D_l = diag(P_l^T K_II P_l)
// (bilinear prolongations P_l onto the 2^l-times coarsened interior grids, exact Galerkin
// diagonals): it needs NO coarse operators during the iteration, only one restriction and
// one prolongation sweep through tiny arrays, and brings k from ~200 (Jacobi) to ~28 at
// n=64, mesh-independently (39 at n=128).  Any SPD preconditioner gives the same converged
// phi; the stopping rule stays the reference's ||r||_2 <= tol on the unscaled, unpreconditioned
// residual, evaluated exactly every iteration (basis.tpp:297).
//
// Per iteration and fine DoF the CTA moves ~19 doubles through shared memory (stencil 10,
// staging/restriction/prolongation 7, direction update 2) and executes 7 block barriers.
#include <math.h>

#include <type_traits>

#include "msb_internal.cuh"

// Optional per-stage cycle timers (profiling build only: make EXTRA=-DMSB_STAGE_TIMERS).
#ifdef MSB_STAGE_TIMERS
__device__ unsigned long long g_msb_stage_cycles[16];
#  define ST_DECL long long st_t0 = clock64(), st_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#  define ST_MARK(i)                       \
    if (threadIdx.x == 0)                  \
      {                                    \
        const long long st_t1 = clock64(); \
        st_acc[i] += st_t1 - st_t0;        \
        st_t0 = st_t1;                     \
      }
#  define ST_FLUSH                                                              \
    if (threadIdx.x == 0)                                                       \
      for (int st_i = 0; st_i < 12; ++st_i)                                     \
        atomicAdd(&g_msb_stage_cycles[st_i], (unsigned long long)st_acc[st_i]);
extern "C" int
msb_debug_stage_cycles(unsigned long long *out, int reset)
{
  cudaError_t e = cudaMemcpyFromSymbol(out, g_msb_stage_cycles, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset)
    {
      unsigned long long z[16] = {0};
      e = cudaMemcpyToSymbol(g_msb_stage_cycles, z, sizeof z);
    }
  return (int)e;
}
#else
#  define ST_DECL
#  define ST_MARK(i)
#  define ST_FLUSH
#endif

namespace msb
{
  struct BpxParams
  {
    const double *corners; // [C][8]
    const double *q1coef;  // [C][16]
    const double *sten;    // [C][6][N]
    double       *phi;     // [C][4][N]
    int32_t      *iters;   // [C][4]
    double       *res;     // [C][4]
    int32_t      *fail;
    double        tol2;
    int           max_iter;
    int           n_cells;
  };

  namespace bpx
  {
    template <int NRHS>
    __device__ __forceinline__ void
    ldv(const double *p, int idx, double (&o)[NRHS])
    {
      if constexpr (NRHS == 1)
        o[0] = p[idx];
      else
        {
#pragma unroll
          for (int k = 0; k < NRHS; k += 2)
            {
              const double2 t = *reinterpret_cast<const double2 *>(p + (size_t)idx * NRHS + k);
              o[k]     = t.x;
              o[k + 1] = t.y;
            }
        }
    }

    template <int NRHS>
    __device__ __forceinline__ void
    stv(double *p, int idx, const double (&o)[NRHS])
    {
      if constexpr (NRHS == 1)
        p[idx] = o[0];
      else
        {
#pragma unroll
          for (int k = 0; k < NRHS; k += 2)
            *reinterpret_cast<double2 *>(p + (size_t)idx * NRHS + k) = make_double2(o[k], o[k + 1]);
        }
    }

    // deterministic block-wide sums; stage 2 is one load per lane plus a shuffle butterfly
    // (every thread ends with bitwise identical totals)
    template <int NV, int NWARP>
    __device__ __forceinline__ void
    block_sum(double (&v)[NV], double *buf, int warp, int lane)
    {
      static_assert(NWARP <= 32, "one lane per warp partial");
#pragma unroll
      for (int k = 0; k < NV; ++k)
        {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1)
            v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        }
      if (lane == 0)
        {
#pragma unroll
          for (int k = 0; k < NV; ++k)
            buf[k * NWARP + warp] = v[k];
        }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NV; ++k)
        {
          double s = lane < NWARP ? buf[k * NWARP + lane] : 0.0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1)
            if (off < NWARP || NWARP == 32)
              s += __shfl_xor_sync(0xffffffffu, s, off);
          // lanes >= NWARP hold partial garbage sums of zeros and real values: broadcast lane 0
          v[k] = __shfl_sync(0xffffffffu, s, 0);
        }
    }

    // a / b for finite b > 0 without the FP64 division sequence (every thread needs alpha and
    // beta each iteration): 20-bit hardware reciprocal seed, three Newton steps (>= 53 bits),
    // one correction step on the quotient
    __device__ __forceinline__ double
    fast_div(double a, double b)
    {
      double y;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
      double e = fma(-b, y, 1.0);
      y        = fma(y, e, y);
      e        = fma(-b, y, 1.0);
      y        = fma(y, e, y);
      e        = fma(-b, y, 1.0);
      y        = fma(y, e, y);
      const double q = a * y;
      return fma(fma(-b, q, a), y, q);
    }

    // block-wide sums of TWO values with half the shuffles of two separate reductions: after
    // the first exchange the lower half-warp carries value 0, the upper half-warp value 1
    template <int NWARP>
    __device__ __forceinline__ void
    block_sum2(double &v0, double &v1, double *buf, int warp, int lane)
    {
      static_assert(NWARP <= 16, "one half-warp lane per warp partial");
      const bool   up   = lane & 16;
      const double recv = __shfl_xor_sync(0xffffffffu, up ? v0 : v1, 16);
      double       s    = (up ? v1 : v0) + recv;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, off);
      if ((lane & 15) == 0)
        buf[(lane >> 4) * NWARP + warp] = s;
      __syncthreads();
      double t = (lane & 15) < NWARP ? buf[(lane >> 4) * NWARP + (lane & 15)] : 0.0;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, off);
      v0 = __shfl_sync(0xffffffffu, t, 0);
      v1 = __shfl_sync(0xffffffffu, t, 16);
    }

    // symmetric 9-point stencil storage (the layout of Shard::d_sten, any level):
    // entry of row node (x,y) towards (x+ex, y+ey); np = nodes per direction, N = np*np
    __device__ __forceinline__ double
    sten_get(const double *S, int np, int N, int x, int y, int ex, int ey)
    {
      const int i = y * np + x;
      if (ey == 0)
        return ex == 0 ? S[ST_KC * N + i] : S[ST_KE * N + (ex > 0 ? i : i - 1)];
      if (ex == 0)
        return S[ST_KN * N + (ey > 0 ? i : i - np)];
      if (ex == ey)
        return S[ST_KD1 * N + (ex > 0 ? i : i - np - 1)];
      return S[ST_KD2 * N + (ey > 0 ? i - 1 : i - np)];
    }

    // One row of the Galerkin coarse operator P^T A P for bilinear P: the five entries of
    // coarse node I = (X,Y) towards d = (0,0) (1,0) (0,1) (1,1) (-1,1).  Every fine stencil
    // entry A(i, i+e), i = 2I + a, is loaded once and scattered (at compile time) into the
    // entries it contributes to: (P^T A P)(I, I+d) = sum_a sum_e w(a) w(b) A(2I+a, 2I+a+e)
    // with b = a + e - 2d, |b| <= 1.
    __device__ __forceinline__ void
    galerkin_row(const double *Sf, int npf, int Nf, int X, int Y, double (&acc)[5])
    {
      constexpr int ddx[5] = {0, 1, 0, 1, -1}, ddy[5] = {0, 0, 1, 1, 1};
#pragma unroll
      for (int d = 0; d < 5; ++d)
        acc[d] = 0.0;
#pragma unroll
      for (int ay = -1; ay <= 1; ++ay)
#pragma unroll
        for (int ax = -1; ax <= 1; ++ax)
          {
            const double wa = (ax == 0 ? 1.0 : 0.5) * (ay == 0 ? 1.0 : 0.5);
            const int    ix = 2 * X + ax, iy = 2 * Y + ay;
#pragma unroll
            for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
              for (int ex = -1; ex <= 1; ++ex)
                {
                  const double v = wa * sten_get(Sf, npf, Nf, ix, iy, ex, ey);
#pragma unroll
                  for (int d = 0; d < 5; ++d)
                    {
                      const int bx = ax + ex - 2 * ddx[d], by = ay + ey - 2 * ddy[d];
                      if (bx >= -1 && bx <= 1 && by >= -1 && by <= 1)
                        acc[d] = fma((bx == 0 ? 1.0 : 0.5) * (by == 0 ? 1.0 : 0.5), v, acc[d]);
                    }
                }
          }
    }

    // compile-time loops over levels (ascending / descending, inclusive bounds)
    template <int L0, int L1, class F>
    __device__ __forceinline__ void
    for_levels(F &&f)
    {
      if constexpr (L0 <= L1)
        {
          f(std::integral_constant<int, L0>{});
          for_levels<L0 + 1, L1>(f);
        }
    }
    template <int L1, int L0, class F>
    __device__ __forceinline__ void
    for_levels_down(F &&f)
    {
      if constexpr (L1 >= L0)
        {
          f(std::integral_constant<int, L1>{});
          for_levels_down<L1 - 1, L0>(f);
        }
    }

    template <int NL, int NRHS, int THREADS>
    struct Cfg
    {
      static constexpr int n     = 1 << NL;
      static constexpr int np    = n + 1;
      static constexpr int N     = np * np;
      static constexpr int NWARP = THREADS / 32;
      static constexpr int WX    = (n - 1 + 31) / 32;
      static constexpr int WY    = NWARP / WX;
      static constexpr int RPT   = (n - 1 + WY - 1) / WY;
      static constexpr int LEVELS = NL - 1; // coarse levels 1..NL-1 (the last has one unknown)
      // levels 1..LW have >= 15x15 unknowns and are swept by the whole CTA
      static constexpr int LW = NL >= 4 ? NL - 4 : 0;
      // coarse level arrays, all levels packed: level l has (n>>l)+1 nodes per direction
      __host__ __device__ static constexpr int
      lvl_np(int l)
      {
        return (n >> l) + 1;
      }
      __host__ __device__ static constexpr int
      lvl_off(int l) // offset (in nodes) of level l >= 1 inside the packed arrays
      {
        int o = 0;
        for (int k = 1; k < l; ++k)
          o += lvl_np(k) * lvl_np(k);
        return o;
      }
      static constexpr int    CN  = lvl_off(NL); // total coarse nodes
      static constexpr int    RED = 3 * NRHS * NWARP;
      static constexpr size_t smem_doubles =
        4 * (size_t)n * n + 2 * (size_t)NRHS * N + (size_t)(NRHS + 1) * CN + 2 * RED + 8;
      static_assert(NWARP % WX == 0, "warp grid");
      static_assert(5 * CN <= 2 * NRHS * N, "Galerkin scratch must fit the p/u buffers");
    };

    template <int NL, int NRHS, int THREADS>
    __global__ void __launch_bounds__(THREADS, (NL <= 5 && NRHS == 1 && THREADS <= 128) ? 3 : 1)
    solve_bpx_kernel(BpxParams P)
    {
      using C             = Cfg<NL, NRHS, THREADS>;
      constexpr int n     = C::n, np = C::np, N = C::N;
      constexpr int NWARP = C::NWARP, WX = C::WX, RPT = C::RPT;
      constexpr int GROUPS = 4 / NRHS;
      constexpr int CN     = C::CN;

      extern __shared__ __align__(16) double smem[];
      double *sE   = smem;
      double *sN   = sE + n * n;
      double *sD1  = sN + n * n;
      double *sD2  = sD1 + n * n;
      double *sP   = sD2 + n * n;             // [N][NRHS] search direction, zero halo
      double *sU   = sP + (size_t)NRHS * N;   // [N][NRHS] unscaled residual staging
      double *sV   = sU + (size_t)NRHS * N;   // [CN][NRHS] coarse residuals / corrections
      double *sDi  = sV + (size_t)NRHS * CN;  // [CN] 1 / Galerkin diagonal
      double *sRed = sDi + CN;                // 2 reduction buffers

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int cell = blockIdx.x; // one CTA per coarse cell; its 4/NRHS groups of bases run in turn

      const double *S   = P.sten + (size_t)cell * ST_NARR * N;
      const double *KC  = S + ST_KC * N;
      const double *crn = P.corners + 8 * (size_t)cell;
      const double *q1  = P.q1coef + 16 * (size_t)cell;

      ST_DECL
      // ---------------------------------------------------------------- prologue
      // (a) Galerkin hierarchy of the UNSCALED interior operator; scratch = p/u buffers.
      //     Level 1 reads the raw stencil from global memory, level l+1 reads level l.
      {
        double       *G  = sP; // packed like d_sten per level: [5][npl*npl]
        const double *Sf = S;
        int           npf = np, Nf = N, goff = 0;
#pragma unroll 1
        for (int l = 1; l <= C::LEVELS; ++l)
          {
            const int npl = (n >> l) + 1, Nl = npl * npl, nin = npl - 2;
            double   *Gl = G + goff;
            // zero the level, then every interior coarse node writes its own row; the (-1,1)
            // entry belongs to the cell whose lower-left node is (X-1,Y)
            for (int t = tid; t < 5 * Nl; t += THREADS)
              Gl[t] = 0.0;
            for (int t = tid; t < Nl; t += THREADS)
              sDi[goff / 5 + t] = 0.0;
            __syncthreads();
            for (int t = tid; t < nin * nin; t += THREADS)
              {
                const int X = 1 + t % nin, Y = 1 + t / nin, i = Y * npl + X;
                double    a[5];
                galerkin_row(Sf, npf, Nf, X, Y, a);
                Gl[ST_KC * Nl + i] = a[0];
                if (X < nin)
                  Gl[ST_KE * Nl + i] = a[1];
                if (Y < nin)
                  Gl[ST_KN * Nl + i] = a[2];
                if (X < nin && Y < nin)
                  Gl[ST_KD1 * Nl + i] = a[3];
                if (X > 1 && Y < nin)
                  Gl[ST_KD2 * Nl + i - 1] = a[4];
                sDi[goff / 5 + i] = 1.0 / a[0];
              }
            __syncthreads();
            Sf   = Gl;
            npf  = npl;
            Nf   = Nl;
            goff += 5 * Nl;
          }
      }
      // (b) s = d^-1/2 on every node into the u buffer (p buffer still holds the hierarchy,
      //     which is dead from here on)
      double *sS = sU;
      for (int i = tid; i < N; i += THREADS)
        sS[i] = rsqrt(KC[i]);
      __syncthreads();
      // (c) scaled edge coefficients
      for (int i = tid; i < n * n; i += THREADS)
        {
          const int    x = i % n, y = i / n, g = y * np + x;
          const double s00 = sS[g], s10 = sS[g + 1], s01 = sS[g + np], s11 = sS[g + np + 1];
          sE[i]  = S[ST_KE * N + g] * s00 * s10;
          sN[i]  = S[ST_KN * N + g] * s00 * s01;
          sD1[i] = S[ST_KD1 * N + g] * s00 * s11;
          sD2[i] = S[ST_KD2 * N + g] * s10 * s01;
        }
      __syncthreads();
      ST_MARK(0)
      // The prologue above (scaled operator + Galerkin diagonals) is shared by all 2^dim bases
      // of the cell: they are solved one group of NRHS after the other by this CTA.
#pragma unroll 1
      for (int grp = 0; grp < GROUPS; ++grp)
      {
      const int rhs0 = grp * NRHS;
      // (d) clear p, u and the coarse vectors (their halos stay zero for the whole solve)
      for (int i = tid; i < 2 * NRHS * N + NRHS * CN; i += THREADS)
        sP[i] = 0.0;
      __syncthreads();

      // ---------------------------------------------------------------- ownership
      const int  wx = warp % WX, wy = warp / WX;
      const int  X  = 1 + 32 * wx + lane;
      const int  Y0 = 1 + RPT * wy;
      const bool colok = X <= n - 1;

      double x[RPT][NRHS], r[RPT][NRHS], q[RPT][NRHS], sq[RPT];

      // (e) rhat_0 = -D^-1/2 K_IB g_B, x = 0
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        {
          const int y = Y0 + j;
          sq[j]       = 0.0;
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            x[j][k] = 0.0, r[j][k] = 0.0, q[j][k] = 0.0;
          if (colok && y <= n - 1)
            {
              const int    i = y * np + X;
              const double d = KC[i];
              sq[j]          = sqrt(d);
              if (X == 1 || X == n - 1 || y == 1 || y == n - 1)
                {
                  double acc[NRHS];
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    acc[k] = 0.0;
#pragma unroll
                  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx)
                      {
                        const int bx = X + dx, by = y + dy;
                        if ((dx == 0 && dy == 0) || !(bx == 0 || bx == n || by == 0 || by == n))
                          continue;
                        const double kij = sten_get(S, np, N, X, y, dx, dy);
                        double       px, py;
                        fine_vertex(crn, n, bx, by, px, py);
#pragma unroll
                        for (int k = 0; k < NRHS; ++k)
                          acc[k] += kij * basis_q1_value(q1, rhs0 + k, px, py);
                      }
                  const double s = rsqrt(d);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    r[j][k] = -s * acc[k];
                }
            }
        }

      // ------------------------------------------------ the preconditioner zhat = Mhat^-1 rhat
      // returns z in `q` (register array reuse), accumulates rz = rhat.zhat and rr = ||r||^2
      auto precondition = [&](double (&z)[RPT][NRHS], double (&rz)[NRHS], double (&rr)[NRHS]) {
        // u = D^1/2 rhat = unscaled residual, staged for the restriction
#pragma unroll
        for (int j = 0; j < RPT; ++j)
          {
            const int y = Y0 + j;
            if (colok && y <= n - 1)
              {
                double u[NRHS];
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  {
                    u[k]  = sq[j] * r[j][k];
                    rr[k] = fma(u[k], u[k], rr[k]);
                  }
                stv<NRHS>(sU, y * np + X, u);
              }
          }
        __syncthreads();
        ST_MARK(4)
        // Level sweeps.  "Wide" levels (>= 15x15 unknowns) are done by the whole CTA with a block
        // barrier each; the remaining tiny levels form a short serial chain on warp 0.
        // restrict(l): r_l = P^T r_{l-1} (full weighting); source of level 1 is the staged u.
        auto restrict_level = [&](auto lc, int first, int nthr) {
          constexpr int l   = decltype(lc)::value;
          constexpr int W   = n >> l, LG = NL - l, npl = W + 1;
          constexpr int npf = (n >> (l - 1)) + 1;
          const double *Vf  = l == 1 ? sU : sV + (size_t)NRHS * C::lvl_off(l - 1);
          double       *Vl  = sV + (size_t)NRHS * C::lvl_off(l);
          for (int t = first; t < W * W; t += nthr)
            {
              const int cx = 1 + (t & (W - 1)), cy = 1 + (t >> LG);
              if (cx > W - 1 || cy > W - 1)
                continue;
              // three independent row sums, then combined (short dependency chains)
              double row[3][NRHS], acc[NRHS];
#pragma unroll
              for (int ay = -1; ay <= 1; ++ay)
                {
                  double a[NRHS], b[NRHS], c[NRHS];
                  ldv<NRHS>(Vf, (2 * cy + ay) * npf + 2 * cx - 1, a);
                  ldv<NRHS>(Vf, (2 * cy + ay) * npf + 2 * cx, b);
                  ldv<NRHS>(Vf, (2 * cy + ay) * npf + 2 * cx + 1, c);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    row[ay + 1][k] = fma(0.5, a[k] + c[k], b[k]);
                }
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                acc[k] = fma(0.5, row[0][k] + row[2][k], row[1][k]);
              stv<NRHS>(Vl, cy * npl + cx, acc);
            }
        };
        // prolong(l): z_l = r_l / D_l + P z_{l+1}  (coarsest level: z = r / D)
        auto prolong_level = [&](auto lc, int first, int nthr) {
          constexpr int l   = decltype(lc)::value;
          constexpr int W   = n >> l, LG = NL - l, npl = W + 1;
          double       *Vl  = sV + (size_t)NRHS * C::lvl_off(l);
          const double *Dl  = sDi + C::lvl_off(l);
          for (int t = first; t < W * W; t += nthr)
            {
              const int fx = 1 + (t & (W - 1)), fy = 1 + (t >> LG);
              if (fx > W - 1 || fy > W - 1)
                continue;
              const int i = fy * npl + fx;
              double    v[NRHS];
              ldv<NRHS>(Vl, i, v);
              const double di = Dl[i];
              if constexpr (l < C::LEVELS)
                {
                  constexpr int npc = (n >> (l + 1)) + 1;
                  const double *Vc  = sV + (size_t)NRHS * C::lvl_off(l + 1);
                  const int xl = fx >> 1, xh = (fx + 1) >> 1, yl = fy >> 1, yh = (fy + 1) >> 1;
                  double    a[NRHS], b[NRHS], c[NRHS], d[NRHS];
                  ldv<NRHS>(Vc, yl * npc + xl, a);
                  ldv<NRHS>(Vc, yl * npc + xh, b);
                  ldv<NRHS>(Vc, yh * npc + xl, c);
                  ldv<NRHS>(Vc, yh * npc + xh, d);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    v[k] = fma(v[k], di, 0.25 * ((a[k] + b[k]) + (c[k] + d[k])));
                }
              else
                {
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    v[k] *= di;
                }
              stv<NRHS>(Vl, i, v);
            }
        };
        // down: wide levels
        for_levels<1, C::LW>([&](auto lc) {
          restrict_level(lc, tid, THREADS);
          __syncthreads();
        });
        ST_MARK(5)
        // the three tiny levels below the 15x15 level B = LW (7x7, 3x3, 1x1 unknowns)
        if constexpr (C::LW >= 1)
          {
            // No serial chain: their residuals are restricted DIRECTLY from level B (a product
            // of full-weighting restrictions is the restriction with the nested hat function) by
            // different warps in parallel, and their corrections are interpolated DIRECTLY back
            // to level B (a product of bilinear interpolations is bilinear on the coarser grid).
            static_assert(C::LEVELS == C::LW + 3, "levels below the 15x15 level");
            constexpr int B   = C::LW, npB = 17;
            double       *VB  = sV + (size_t)NRHS * C::lvl_off(B);
            double       *V1  = sV + (size_t)NRHS * C::lvl_off(B + 1); // 9x9 nodes
            double       *V2  = sV + (size_t)NRHS * C::lvl_off(B + 2); // 5x5 nodes
            double       *V3  = sV + (size_t)NRHS * C::lvl_off(B + 3); // 3x3 nodes
            const double *DB  = sDi + C::lvl_off(B), *D1 = sDi + C::lvl_off(B + 1);
            const double *D2  = sDi + C::lvl_off(B + 2), *D3 = sDi + C::lvl_off(B + 3);
            for (int task = warp; task < 12; task += NWARP)
              {
                if (task >= 10)
                  {
                    // level B+1: one thread per node, 3x3 window; stores t = r / D
                    const int t = (task - 10) * 32 + lane;
                    if (t < 49)
                      {
                        const int cx = 1 + t % 7, cy = 1 + t / 7, i = cy * 9 + cx;
                        double    row[3][NRHS];
#pragma unroll
                        for (int ay = -1; ay <= 1; ++ay)
                          {
                            double a[NRHS], b[NRHS], c[NRHS];
                            ldv<NRHS>(VB, (2 * cy + ay) * npB + 2 * cx - 1, a);
                            ldv<NRHS>(VB, (2 * cy + ay) * npB + 2 * cx, b);
                            ldv<NRHS>(VB, (2 * cy + ay) * npB + 2 * cx + 1, c);
#pragma unroll
                            for (int k = 0; k < NRHS; ++k)
                              row[ay + 1][k] = fma(0.5, a[k] + c[k], b[k]);
                          }
                        const double di = D1[i];
                        double       o[NRHS];
#pragma unroll
                        for (int k = 0; k < NRHS; ++k)
                          o[k] = fma(0.5, row[0][k] + row[2][k], row[1][k]) * di;
                        stv<NRHS>(V1, i, o);
                      }
                  }
                else
                  {
                    // level B+2 (task 0..8: node of the 3x3 grid, 7x7 window, hat of width 4) or
                    // level B+3 (task 9: the single node, 15x15 window, hat of width 8): one warp
                    // per node, lane <-> (window column, row group), constant trip counts, then a
                    // shuffle reduction.  The hat weights are separable: w = hx(ax) * hy(ay).
                    const bool top = task == 9;
                    double     acc[NRHS];
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      acc[k] = 0.0;
                    if (top)
                      {
                        // 30 lanes: column ax = lane % 15 - 7, rows [-7,0] (lanes < 15) or [1,7]
                        const int  col = lane % 15, grp = lane / 15;
                        if (grp < 2)
                          {
                            const double hx = 1.0 - abs(col - 7) * 0.125;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                              {
                                const int ay = grp ? 1 + j : j - 7;
                                if (grp && j == 7)
                                  break;
                                const double hy = grp ? 1.0 - (1 + j) * 0.125 : 1.0 - (7 - j) * 0.125;
                                double       u[NRHS];
                                ldv<NRHS>(VB, (8 + ay) * npB + 1 + col, u);
#pragma unroll
                                for (int k = 0; k < NRHS; ++k)
                                  acc[k] = fma(hy * hx, u[k], acc[k]);
                              }
                          }
                      }
                    else
                      {
                        // 28 lanes: column ax = lane % 7 - 3, rows ay = -3 + grp + 4 j, j = 0,1
                        const int col = lane % 7, grp = lane / 7;
                        const int cx = 4 * (1 + task % 3), cy = 4 * (1 + task / 3);
                        if (grp < 4)
                          {
                            const double hx = 1.0 - abs(col - 3) * 0.25;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                              {
                                const int ay = -3 + grp + 4 * j;
                                if (ay > 3)
                                  break;
                                const double hy = 1.0 - abs(ay) * 0.25;
                                double       u[NRHS];
                                ldv<NRHS>(VB, (cy + ay) * npB + cx - 3 + col, u);
#pragma unroll
                                for (int k = 0; k < NRHS; ++k)
                                  acc[k] = fma(hy * hx, u[k], acc[k]);
                              }
                          }
                      }
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
#pragma unroll
                      for (int off = 16; off > 0; off >>= 1)
                        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
                    if (lane == 0)
                      {
                        const int    i  = top ? 4 : (1 + task / 3) * 5 + 1 + task % 3;
                        const double di = top ? D3[4] : D2[i];
#pragma unroll
                        for (int k = 0; k < NRHS; ++k)
                          acc[k] *= di;
                        stv<NRHS>(top ? V3 : V2, i, acc);
                      }
                  }
              }
            __syncthreads();
            // z_B = r_B / D_B + interpolants of t_{B+1}, t_{B+2}, t_{B+3} at the level-B nodes
            for (int t = tid; t < 256; t += THREADS)
              {
                const int fx = 1 + (t & 15), fy = 1 + (t >> 4);
                if (fx > 15 || fy > 15)
                  continue;
                const int i = fy * npB + fx;
                double    v[NRHS];
                ldv<NRHS>(VB, i, v);
                const double di = DB[i];
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  v[k] *= di;
#pragma unroll
                for (int m = 1; m <= 3; ++m)
                  {
                    const int     R = 1 << m, npm = (16 >> m) + 1;
                    const double *Vm = m == 1 ? V1 : (m == 2 ? V2 : V3);
                    const int     cx = fx >> m, cy = fy >> m;
                    const double  gx = (fx & (R - 1)) * (1.0 / R), gy = (fy & (R - 1)) * (1.0 / R);
                    double        a[NRHS], b[NRHS], c[NRHS], d[NRHS];
                    ldv<NRHS>(Vm, cy * npm + cx, a);
                    ldv<NRHS>(Vm, cy * npm + cx + 1, b);
                    ldv<NRHS>(Vm, (cy + 1) * npm + cx, c);
                    ldv<NRHS>(Vm, (cy + 1) * npm + cx + 1, d);
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      {
                        const double lo = fma(gx, b[k] - a[k], a[k]), hi = fma(gx, d[k] - c[k], c[k]);
                        v[k] += fma(gy, hi - lo, lo);
                      }
                  }
                stv<NRHS>(VB, i, v);
              }
            __syncthreads();
            ST_MARK(6)
            // up: the wide levels above B
            for_levels_down<C::LW - 1, 1>([&](auto lc) {
              prolong_level(lc, tid, THREADS);
              __syncthreads();
            });
          }
        else
          {
            // small local meshes (n <= 16): every coarse level on warp 0, serially
            if (warp == 0)
              {
                for_levels<1, C::LEVELS>([&](auto lc) {
                  restrict_level(lc, lane, 32);
                  __syncwarp();
                });
                for_levels_down<C::LEVELS, 1>([&](auto lc) {
                  prolong_level(lc, lane, 32);
                  __syncwarp();
                });
              }
            __syncthreads();
            ST_MARK(6)
          }
        ST_MARK(7)
        // level 0: zhat = rhat + D^1/2 (P z_1).  The strip of RPT fine rows (first row odd, RPT
        // even) lies under RPT/2+1 coarse rows: their horizontal averages are loaded once and kept
        // in registers (2 loads per coarse row instead of 4 per fine row).
        {
          static_assert(RPT % 2 == 0, "strip must start on an odd row");
          constexpr int np1 = C::lvl_np(1);
          if (colok)
            {
              const int xl = X >> 1, xh = (X + 1) >> 1, cr0 = (Y0 - 1) >> 1;
              double    h[RPT / 2 + 1][NRHS]; // 0.5 * (V1[cr][xl] + V1[cr][xh]); rows beyond n/2 are halo zeros
#pragma unroll
              for (int c = 0; c <= RPT / 2; ++c)
                {
                  const int cr = cr0 + c <= n / 2 ? cr0 + c : n / 2;
                  double    a[NRHS], b[NRHS];
                  ldv<NRHS>(sV, cr * np1 + xl, a);
                  ldv<NRHS>(sV, cr * np1 + xh, b);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    h[c][k] = 0.5 * (a[k] + b[k]);
                }
#pragma unroll
              for (int j = 0; j < RPT; ++j)
                {
                  const int y = Y0 + j;
                  if (y <= n - 1)
                    {
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        {
                          // j even: y odd, between coarse rows j/2 and j/2+1; j odd: y even, on row (j+1)/2
                          const double cc = (j & 1) ? h[(j + 1) / 2][k] : 0.5 * (h[j / 2][k] + h[j / 2 + 1][k]);
                          z[j][k]         = fma(sq[j], cc, r[j][k]);
                          rz[k]           = fma(r[j][k], z[j][k], rz[k]);
                        }
                    }
                }
            }
        }
      };

      // (f) z_0, p_0 = z_0, rho = r.z, initial residual norm
      double rho[NRHS], exact[NRHS];
      {
        double rz[NRHS], rr[NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          rz[k] = 0.0, rr[k] = 0.0;
        precondition(q, rz, rr);
        double both[2 * NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          both[k] = rz[k], both[NRHS + k] = rr[k];
        if constexpr (NRHS == 1 && NWARP <= 16)
          block_sum2<NWARP>(both[0], both[1], sRed, warp, lane);
        else
          block_sum<2 * NRHS, NWARP>(both, sRed, warp, lane);
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          rho[k] = both[k], exact[k] = both[NRHS + k];
#pragma unroll
        for (int j = 0; j < RPT; ++j)
          {
            const int y = Y0 + j;
            if (colok && y <= n - 1)
              stv<NRHS>(sP, y * np + X, q[j]);
          }
      }
      __syncthreads();

      ST_MARK(1)
      bool done[NRHS];
      int  kit[NRHS];
      bool all_done = true;
#pragma unroll
      for (int k = 0; k < NRHS; ++k)
        {
          done[k]  = exact[k] <= P.tol2;
          kit[k]   = 0;
          all_done = all_done && done[k];
        }

      // ---------------------------------------------------------------- PCG iterations
      int it = 0;
      while (!all_done && it < P.max_iter)
        {
          ++it;
          // ---- q = Ahat p on the owned strip, marching up the rows
          double pq[NRHS];
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            pq[k] = 0.0;
          if (colok)
            {
              double a0[NRHS], a1[NRHS], a2[NRHS];
              double b0[NRHS], b1[NRHS], b2[NRHS];
              ldv<NRHS>(sP, (Y0 - 1) * np + X - 1, a0);
              ldv<NRHS>(sP, (Y0 - 1) * np + X, a1);
              ldv<NRHS>(sP, (Y0 - 1) * np + X + 1, a2);
              ldv<NRHS>(sP, Y0 * np + X - 1, b0);
              ldv<NRHS>(sP, Y0 * np + X, b1);
              ldv<NRHS>(sP, Y0 * np + X + 1, b2);
              double cS = sN[(Y0 - 1) * n + X];
#pragma unroll
              for (int j = 0; j < RPT; ++j)
                {
                  const int y = Y0 + j;
                  if (y <= n - 1)
                    {
                      double c0[NRHS], c1[NRHS], c2[NRHS];
                      ldv<NRHS>(sP, (y + 1) * np + X - 1, c0);
                      ldv<NRHS>(sP, (y + 1) * np + X, c1);
                      ldv<NRHS>(sP, (y + 1) * np + X + 1, c2);
                      const double cE = sE[y * n + X], cW = sE[y * n + X - 1];
                      const double cN = sN[y * n + X];
                      const double cNE = sD1[y * n + X], cSW = sD1[(y - 1) * n + X - 1];
                      const double cNW = sD2[y * n + X - 1], cSE = sD2[(y - 1) * n + X];
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        {
                          double t = b1[k];
                          t        = fma(cE, b2[k], t);
                          t        = fma(cW, b0[k], t);
                          t        = fma(cN, c1[k], t);
                          t        = fma(cS, a1[k], t);
                          t        = fma(cNE, c2[k], t);
                          t        = fma(cSW, a0[k], t);
                          t        = fma(cNW, c0[k], t);
                          t        = fma(cSE, a2[k], t);
                          q[j][k]  = t;
                          pq[k]    = fma(b1[k], t, pq[k]);
                          // x += alpha p needs p_old later; keep it in the x update below
                          a0[k] = b0[k], a1[k] = b1[k], a2[k] = b2[k];
                          b0[k] = c0[k], b1[k] = c1[k], b2[k] = c2[k];
                        }
                      cS = cN;
                    }
                }
            }
          ST_MARK(2)
          block_sum<NRHS, NWARP>(pq, sRed, warp, lane);
          ST_MARK(3)

          double alpha[NRHS];
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            alpha[k] = done[k] ? 0.0 : fast_div(rho[k], pq[k]);

          // ---- r -= alpha q ; z = M^-1 r ; rho' = r.z ; ||r||^2
#pragma unroll
          for (int j = 0; j < RPT; ++j)
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              r[j][k] = fma(-alpha[k], q[j][k], r[j][k]);
          double rz[NRHS], rr[NRHS];
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            rz[k] = 0.0, rr[k] = 0.0;
          precondition(q, rz, rr); // q now holds z
          {
            double both[2 * NRHS];
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              both[k] = rz[k], both[NRHS + k] = rr[k];
            if constexpr (NRHS == 1 && NWARP <= 16)
              block_sum2<NWARP>(both[0], both[1], sRed + C::RED, warp, lane);
            else
              block_sum<2 * NRHS, NWARP>(both, sRed + C::RED, warp, lane);
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              rz[k] = both[k], rr[k] = both[NRHS + k];
          }

          ST_MARK(8)
          double beta[NRHS];
          all_done = true;
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            {
              beta[k] = done[k] ? 0.0 : fast_div(rz[k], rho[k]);
              if (!done[k])
                {
                  rho[k]   = rz[k];
                  exact[k] = rr[k];
                  if (rr[k] <= P.tol2)
                    {
                      done[k] = true;
                      kit[k]  = it;
                    }
                  else if (it >= P.max_iter)
                    kit[k] = it;
                }
              all_done = all_done && done[k];
            }

          // ---- x += alpha p_old ; p = z + beta p_old
#pragma unroll
          for (int j = 0; j < RPT; ++j)
            {
              const int y = Y0 + j;
              if (colok && y <= n - 1)
                {
                  double po[NRHS];
                  ldv<NRHS>(sP, y * np + X, po);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    {
                      x[j][k] = fma(alpha[k], po[k], x[j][k]);
                      po[k]   = done[k] && beta[k] == 0.0 ? po[k] : fma(beta[k], po[k], q[j][k]);
                    }
                  stv<NRHS>(sP, y * np + X, po);
                }
            }
          __syncthreads();
          ST_MARK(9)
        }

      // ---------------------------------------------------------------- epilogue
      double *out = P.phi + ((size_t)cell * 4 + rhs0) * N;
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        {
          const int y = Y0 + j;
          if (colok && y <= n - 1)
            {
              const int    i = y * np + X;
              const double s = 1.0 / sq[j];
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                out[(size_t)k * N + i] = s * x[j][k];
            }
        }
      for (int t = tid; t < 4 * n; t += THREADS)
        {
          int jx, jy;
          if (t < n)
            jx = t, jy = 0;
          else if (t < 2 * n)
            jx = n, jy = t - n;
          else if (t < 3 * n)
            jx = n - (t - 2 * n), jy = n;
          else
            jx = 0, jy = n - (t - 3 * n);
          double px, py;
          fine_vertex(crn, n, jx, jy, px, py);
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            out[(size_t)k * N + jy * np + jx] = basis_q1_value(q1, rhs0 + k, px, py);
        }
      if (tid == 0)
        {
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            {
              const int sidx = cell * 4 + rhs0 + k;
              P.iters[sidx]  = kit[k];
              P.res[sidx]    = sqrt(exact[k]);
              if (!done[k])
                atomicMin(P.fail, sidx);
            }
        }
      __syncthreads(); // shared buffers are reused by the next group of bases
      ST_MARK(10)
      } // grp
      ST_FLUSH
    }

    template <int NL, int NRHS, int THREADS>
    static cudaError_t
    launch_one(const BpxParams &P, cudaStream_t st)
    {
      using C            = Cfg<NL, NRHS, THREADS>;
      const size_t bytes = C::smem_doubles * sizeof(double);
      auto         kern  = solve_bpx_kernel<NL, NRHS, THREADS>;
      cudaError_t  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess)
        return e;
      kern<<<P.n_cells, THREADS, bytes, st>>>(P);
      return cudaGetLastError();
    }
  } // namespace bpx

  cudaError_t
  launch_solve_bpx(const Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches)
  {
    BpxParams P;
    P.corners  = s.d_corners;
    P.q1coef   = s.d_q1coef;
    P.sten     = s.d_sten;
    P.phi      = s.d_phi;
    P.iters    = s.d_iters;
    P.res      = s.d_res;
    P.fail     = s.d_fail;
    P.tol2     = tol * tol;
    P.max_iter = max_iter;
    P.n_cells  = s.n_cells;
    ++*n_launches;
    switch (s.l)
      {
        case 3:
          return bpx::launch_one<3, 4, 64>(P, st);
        case 4:
          return bpx::launch_one<4, 4, 128>(P, st);
        case 5:
          // small CTAs, two resident per SM, hide the barrier latency (measured: 1.52M vs 1.03M
          // solves/s on cfg4 against one 256-thread CTA with all four bases)
          if (s.variant == 1)
            return bpx::launch_one<5, 2, 256>(P, st);
          if (s.variant == 2)
            return bpx::launch_one<5, 1, 128>(P, st); // 3 CTAs per SM
          if (s.variant == 3)
            return bpx::launch_one<5, 4, 256>(P, st);
          return bpx::launch_one<5, 2, 128>(P, st);
        case 6:
          if (s.variant == 1)
            return bpx::launch_one<6, 1, 256>(P, st);
          return bpx::launch_one<6, 1, 512>(P, st);
        default:
          return cudaErrorInvalidValue;
      }
  }
} // namespace msb
