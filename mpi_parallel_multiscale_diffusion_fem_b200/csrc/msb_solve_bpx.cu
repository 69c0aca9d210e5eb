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
#define MSB_STAGE_ARRAY g_msb_stage_cycles
#include "msb_bpx_common.cuh"
#include "msb_fused.cuh"

#ifdef MSB_STAGE_TIMERS
__device__ unsigned long long g_msb_stage_cycles[16];
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
#endif

namespace msb
{
  namespace bpx
  {
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
      // n = 32: the 7x7-unknown level is solved exactly with a dense inverse in shared memory
      // (bpx::exact7_build); at n = 64 shared memory is full and n <= 16 has no such level
      static constexpr bool   EXACT7 = NL == 5;
      static constexpr size_t smem_doubles =
        4 * (size_t)n * n + 2 * (size_t)NRHS * N + (size_t)(NRHS + 1) * CN + 2 * RED + 8 + (EXACT7 ? 49 * 49 : 0);
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
      constexpr int PRESUM_RPT = C::LW >= 1 ? RPT : 0; // pre-summed residual staging (n >= 32)
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
      double *sGi  = sRed + 2 * C::RED + 8;   // [49][49] inverse of the 7x7-level operator (EXACT7)

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
      if constexpr (C::EXACT7)
        {
          // scratch: the first coefficient array (filled in (c) below; the p/u buffers hold the hierarchy)
          static_assert(EXACT7_SCRATCH <= n * n, "band scratch must fit a coefficient array");
          exact7_build<THREADS>(sP + 5 * C::lvl_off(C::LW + 1), sGi, sE, tid);
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
        // u = D^1/2 rhat = unscaled residual, staged for the restriction: as pre-summed strips
        // (Presum) when level 1 is a wide level, else as plain values
        {
          using PS = Presum<NL, NRHS, RPT>;
          double    acc[NRHS];
          const int pc = PS::col(X);
#pragma unroll
          for (int j = 0; j < RPT; ++j)
            {
              double u[NRHS];
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                {
                  u[k]  = sq[j] * r[j][k]; // zero on rows / columns beyond the mesh
                  rr[k] = fma(u[k], u[k], rr[k]);
                }
              if constexpr (PRESUM_RPT > 0)
                PS::push(sU, j, u, acc, pc, wy, colok);
              else if (colok && Y0 + j <= n - 1)
                stv<NRHS>(sU, (Y0 + j) * np + X, u);
            }
          (void)acc, (void)pc;
        }
        __syncthreads();
        ST_MARK(4)
        ST_MARK(5)
        coarse_correction<NL, NRHS, THREADS, PRESUM_RPT, C::EXACT7>(
          sU, sV, sDi, tid, warp, lane,
          [&](int st_k) {
            (void)st_k;
            ST_MARK(st_k)
          },
          [&](int c, double(&g)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              g[i] = Exact7<THREADS>::fetch(sGi, tid, c, i);
          });
        ST_MARK(11)
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
          // default: two bases in flight per CTA, tensor memory as spill space (1.07M vs 0.89M
          // solves/s on the target configuration against one basis per pass) + exact solve of the
          // 7x7 coarse level (26.0 instead of 27.8 iterations; pays since its inverse is built by the
          // banded factorisation: 16.94 vs 17.44 ms on 5920 cells; it lost with the Gauss-Jordan sweep)
          // (variant 0 on axis-aligned cells with an analytic coefficient never gets here: msb_run takes the
          //  fused one-kernel stage, msb_solve_fused.cu; variant 9 = this three-kernel path for A/B)
          if (s.variant == 5)
            return launch_solve_bpx_tm(P, 256, false, st);
          if (s.variant == 7)
            return launch_solve_bpx_tm(P, 512, false, st); // without the exact 7x7 coarse solve (A/B)
          if (s.variant == 6)
            return bpx::launch_one<6, 1, 512>(P, st);
          if (s.variant == 1)
            return bpx::launch_one<6, 1, 256>(P, st);
          return launch_solve_bpx_tm(P, 512, true, st);
        default:
          return cudaErrorInvalidValue;
      }
  }
  // the fused one-kernel stage (msb_solve_fused.cu) for the cells [c0, c0 + nc) of the shard
  cudaError_t
  launch_stage_fused_range(const Shard &s, int c0, int nc, double tol, int max_iter, cudaStream_t st, int *n_launches,
                           bool split)
  {
    FusedParams P;
    P.split     = split ? 1 : 0;
    P.corners   = s.d_corners + 8 * (size_t)c0;
    P.q1coef    = s.d_q1coef + 16 * (size_t)c0;
    P.phi       = s.d_phi + (size_t)c0 * 4 * s.N;
    P.M         = s.d_M + 16 * (size_t)c0;
    P.b         = s.d_b + 4 * (size_t)c0;
    P.iters     = s.d_iters + 4 * (size_t)c0;
    P.res       = s.d_res + 4 * (size_t)c0;
    P.fail      = s.d_fail;
    P.fail_base = 4 * c0;
    P.tol2      = tol * tol;
    P.max_iter  = max_iter;
    P.n_cells   = nc;
    P.rhs_value = s.rhs_value;
    P.flavor    = s.variant >= 10 && s.variant <= 12 ? s.variant - 10 : 0; // (13: flavour 0 without tail balancing)
    P.coef      = make_coeff_eval(s.coeff);
    ++*n_launches;
    return launch_solve_fused(P, s.l, st);
  }

  cudaError_t
  launch_stage_fused(const Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches)
  {
    // Tail balancing (as in the cluster tier): `slots` CTAs are co-resident (one per SM at n = 64, three at n = 32); when
    // the last wave of a shard is at most half full, its cells go to two CTAs each, one per pair of bases, and the wave
    // takes ~0.6 of a full one.  Matters for shards of a few waves (cfg2: 1024 cells on 444 slots); the 443 waves of the
    // target workload end in a wave that is more than half full and stay ONE launch.  Variant 13: off (A/B).
    int tail = 0;
    if (s.variant != 13)
      {
        static int sms = 0;
        if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device) != cudaSuccess)
          sms = 0, (void)cudaGetLastError();
        const int slots = sms * (s.l == 5 ? 3 : 1);
        if (slots > 0)
          {
            const int t = s.n_cells % slots;
            if (t > 0 && 2 * t <= slots)
              tail = t;
          }
      }
    cudaError_t e = cudaSuccess;
    if (s.n_cells - tail > 0)
      e = launch_stage_fused_range(s, 0, s.n_cells - tail, tol, max_iter, st, n_launches, false);
    if (e == cudaSuccess && tail > 0)
      e = launch_stage_fused_range(s, s.n_cells - tail, tail, tol, max_iter, st, n_launches, true);
    return e;
  }
} // namespace msb
