// msb_solve_smem.cu -- shared-memory-resident tier of the local multiscale-basis solves.
//
// Replaces, for one coarse cell and NRHS of its 2^dim vertex right-hand sides at a time,
// the reference's per-basis sequence (diffusion_problem_basis.tpp:450-465)
//     system_matrix = diffusion_matrix; condense(); SolverCG+PreconditionSSOR; distribute()
// by ONE CTA that keeps the whole solve on chip:
//   * the condensed system is the interior block K_II phi_I = -K_IB g_B (what condense()
//     produces), symmetrically scaled to unit diagonal:
//         Ahat = D^-1/2 K_II D^-1/2,  yhat = D^1/2 phi_I,  bhat = D^-1/2 b.
//     CG on Ahat is exactly Jacobi-preconditioned CG on K_II, needs no diagonal, no z vector
//     and only two dot products per iteration;
//   * the four scaled edge-coefficient arrays (E, N, D1, D2; n*n doubles each) and the search
//     direction p (with a zero halo on the constrained boundary) live in shared memory, shared
//     by the NRHS right-hand sides which are interleaved so one 128-bit LDS feeds two of them;
//   * x, r and A p never leave registers: every thread owns a fixed vertical strip of DoFs;
//   * dot products: warp shuffles + one shared-memory stage, no grid sync, no host round trip;
//   * stopping rule = the reference's (basis.tpp:297): ||r||_2 <= tol on the UNSCALED
//     residual, tested every iteration.  ||r||^2 = sum_i d_i rhat_i^2 is bracketed for free by
//     d_min*rho <= ||r||^2 <= d_max*rho and only evaluated exactly inside the bracket.
#include <math.h>

#include "msb_internal.cuh"

namespace msb
{
  struct SolveParams
  {
    const double *corners; // [C][8]
    const double *q1coef;  // [C][16]
    const double *sten;    // [C][6][N]
    double       *phi;     // [C][4][N]
    int32_t      *iters;   // [C][4]
    double       *res;     // [C][4]
    int32_t      *fail;    // [0] = min failing solve index
    double        tol2;    // tol^2
    int           max_iter;
    int           n_cells;
  };

  template <int NV>
  struct Vec
  {
    double v[NV];
  };

  // NRHS interleaved doubles at p[idx*NRHS ..]; 128-bit accesses where possible
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

  // deterministic block-wide sum of NV values; every thread returns bitwise the same totals
  template <int NV, int NWARP>
  __device__ __forceinline__ void
  block_sum(double (&v)[NV], double *buf, int warp, int lane)
  {
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
          buf[warp * NV + k] = v[k];
      }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k)
      {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w)
          s += buf[w * NV + k];
        v[k] = s;
      }
  }

  template <int NL, int NRHS, int THREADS>
  struct SmemCfg
  {
    static constexpr int n     = 1 << NL;
    static constexpr int np    = n + 1;
    static constexpr int N     = np * np;
    static constexpr int NWARP = THREADS / 32;
    static constexpr int WX    = (n - 1 + 31) / 32;     // warps across a row of interior DoFs
    static constexpr int WY    = NWARP / WX;            // warp rows
    static constexpr int RPT   = (n - 1 + WY - 1) / WY; // rows per thread
    static constexpr int RED   = NWARP * NRHS;          // one reduction buffer
    static constexpr size_t smem_doubles = 4 * (size_t)n * n + (size_t)NRHS * N + 3 * RED + 8;
    static_assert(NWARP % WX == 0, "warp grid");
  };

  template <int NL, int NRHS, int THREADS>
  __global__ void __launch_bounds__(THREADS, 1)
  solve_smem_kernel(SolveParams P)
  {
    using C             = SmemCfg<NL, NRHS, THREADS>;
    constexpr int n     = C::n, np = C::np, N = C::N;
    constexpr int NWARP = C::NWARP, WX = C::WX, RPT = C::RPT;
    constexpr int GROUPS = 4 / NRHS;

    extern __shared__ __align__(16) double smem[];
    double *sE   = smem;
    double *sN   = sE + n * n;
    double *sD1  = sN + n * n;
    double *sD2  = sD1 + n * n;
    double *sP   = sD2 + n * n;              // [N][NRHS]
    double *sRed = sP + (size_t)NRHS * N;    // 3 buffers of RED doubles
    double *sMM  = sRed + 3 * C::RED;        // d_min, d_max staging

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cell = blockIdx.x / GROUPS, rhs0 = (blockIdx.x % GROUPS) * NRHS;

    const double *S    = P.sten + (size_t)cell * ST_NARR * N;
    const double *KC   = S + ST_KC * N;
    const double *crn  = P.corners + 8 * (size_t)cell;
    const double *q1   = P.q1coef + 16 * (size_t)cell;

    // ------------------------------------------------------------------ prologue
    // (a) s = d^-1/2 for every node into the (still unused) p region; d_min / d_max over
    //     the interior (the unknowns of the condensed system)
    double *sS   = sP;
    double  dmin = 1e300, dmax = 0.0;
    for (int i = tid; i < N; i += THREADS)
      {
        const double d  = KC[i];
        const int    jx = i % np, jy = i / np;
        sS[i]           = rsqrt(d);
        if (jx > 0 && jx < n && jy > 0 && jy < n)
          {
            dmin = fmin(dmin, d);
            dmax = fmax(dmax, d);
          }
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      {
        dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, off));
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, off));
      }
    if (lane == 0)
      {
        sRed[warp]         = dmin;
        sRed[NWARP + warp] = dmax;
      }
    __syncthreads();
    for (int w = 0; w < NWARP; ++w)
      {
        dmin = fmin(dmin, sRed[w]);
        dmax = fmax(dmax, sRed[NWARP + w]);
      }
    // (b) scaled edge coefficients, indexed by the lower-left node / fine cell (x,y) in [0,n)^2
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
    // (c) clear p (its boundary halo stays zero for the whole solve: constrained DoFs are
    //     decoupled by condense())
    for (int i = tid; i < NRHS * N; i += THREADS)
      sP[i] = 0.0;
    __syncthreads();

    // ------------------------------------------------------------------ ownership
    const int  wx = warp % WX, wy = warp / WX;
    const int  X  = 1 + 32 * wx + lane; // node column
    const int  Y0 = 1 + RPT * wy;       // first node row
    const bool colok = X <= n - 1;

    double x[RPT][NRHS], r[RPT][NRHS], q[RPT][NRHS];

    // (d) rhat_0 = bhat = -D^-1/2 K_IB g_B (condense(), basis.tpp:461), x = 0, p = r
    double part[NRHS], exact[NRHS];
#pragma unroll
    for (int k = 0; k < NRHS; ++k)
      part[k] = 0.0, exact[k] = 0.0;
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      {
        const int y = Y0 + j;
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          x[j][k] = 0.0, r[j][k] = 0.0, q[j][k] = 0.0;
        if (colok && y <= n - 1)
          {
            const int i = y * np + X;
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
                      // raw coupling K(i, neighbour)
                      double kij;
                      if (dy == 0)
                        kij = S[ST_KE * N + (dx > 0 ? i : i - 1)];
                      else if (dx == 0)
                        kij = S[ST_KN * N + (dy > 0 ? i : i - np)];
                      else if (dx == dy)
                        kij = S[ST_KD1 * N + (dx > 0 ? i : i - np - 1)];
                      else
                        kij = S[ST_KD2 * N + (dy > 0 ? i - 1 : i - np)];
                      double px, py;
                      fine_vertex(crn, n, bx, by, px, py);
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        acc[k] += kij * basis_q1_value(q1, rhs0 + k, px, py);
                    }
                const double d = KC[i], s = rsqrt(d);
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  {
                    r[j][k] = -s * acc[k];
                    part[k] += r[j][k] * r[j][k];
                    exact[k] += r[j][k] * r[j][k] * d;
                  }
              }
            stv<NRHS>(sP, i, r[j]);
          }
      }
    double rho[NRHS];
    {
      double both[2 * NRHS];
#pragma unroll
      for (int k = 0; k < NRHS; ++k)
        both[k] = part[k], both[NRHS + k] = exact[k];
      // the three reduction buffers hold 3*RED >= 2*NRHS*NWARP doubles
      block_sum<2 * NRHS, NWARP>(both, sRed, warp, lane);
#pragma unroll
      for (int k = 0; k < NRHS; ++k)
        rho[k] = both[k], exact[k] = both[NRHS + k];
    }
    __syncthreads(); // p visible; reduction buffers free again

    bool done[NRHS];
    int  kit[NRHS];
    bool all_done = true;
#pragma unroll
    for (int k = 0; k < NRHS; ++k)
      {
        done[k] = exact[k] <= P.tol2; // SolverCG: initial residual already below tol
        kit[k]  = 0;
        all_done = all_done && done[k];
      }

    // ------------------------------------------------------------------ CG iterations
    int it = 0;
    while (!all_done && it < P.max_iter)
      {
        ++it;
        // ---- phase 1: q = Ahat p on the owned strip, marching up the rows
        double pq[NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          pq[k] = 0.0;
        if (colok)
          {
            double a0[NRHS], a1[NRHS], a2[NRHS]; // row y-1: x-1, x, x+1
            double b0[NRHS], b1[NRHS], b2[NRHS]; // row y
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
                    double c0[NRHS], c1[NRHS], c2[NRHS]; // row y+1
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
                        a0[k] = b0[k], a1[k] = b1[k], a2[k] = b2[k];
                        b0[k] = c0[k], b1[k] = c1[k], b2[k] = c2[k];
                      }
                    cS = cN;
                  }
              }
          }
        block_sum<NRHS, NWARP>(pq, sRed, warp, lane);

        double alpha[NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          alpha[k] = done[k] ? 0.0 : rho[k] / pq[k];

        // ---- phase 2a: r -= alpha q, rho' = r.r
        double rr[NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          rr[k] = 0.0;
#pragma unroll
        for (int j = 0; j < RPT; ++j)
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            {
              r[j][k] = fma(-alpha[k], q[j][k], r[j][k]);
              rr[k]   = fma(r[j][k], r[j][k], rr[k]);
            }
        block_sum<NRHS, NWARP>(rr, sRed + C::RED, warp, lane);

        // ---- stopping rule on the unscaled residual: d_min rho' <= ||r||^2 <= d_max rho'
        bool need_exact = false;
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          need_exact = need_exact || (!done[k] && rr[k] * dmin <= P.tol2 && rr[k] * dmax > P.tol2);
        double ex[NRHS];
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          ex[k] = rr[k] * dmax; // upper bound unless evaluated exactly
        if (need_exact)
          {
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              ex[k] = 0.0;
#pragma unroll
            for (int j = 0; j < RPT; ++j)
              {
                const int y = Y0 + j;
                if (colok && y <= n - 1)
                  {
                    const double d = KC[y * np + X];
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      ex[k] = fma(r[j][k] * r[j][k], d, ex[k]);
                  }
              }
            block_sum<NRHS, NWARP>(ex, sRed + 2 * C::RED, warp, lane);
          }

        double beta[NRHS];
        all_done = true;
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          {
            beta[k] = done[k] ? 0.0 : rr[k] / rho[k];
            if (!done[k])
              {
                rho[k] = rr[k];
                if (ex[k] <= P.tol2)
                  {
                    done[k]  = true;
                    kit[k]   = it;
                    exact[k] = ex[k];
                  }
                else if (it >= P.max_iter)
                  {
                    kit[k]   = it;
                    exact[k] = ex[k];
                  }
              }
            all_done = all_done && done[k];
          }

        // ---- phase 2b: x += alpha p_old; p = r + beta p_old
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
                    po[k]   = fma(beta[k], po[k], r[j][k]);
                  }
                stv<NRHS>(sP, y * np + X, po);
              }
          }
        __syncthreads();
      }

    // ------------------------------------------------------------------ epilogue
    // distribute() (basis.tpp:308): interior phi = D^-1/2 yhat, boundary phi = g
    double *out = P.phi + ((size_t)cell * 4 + rhs0) * N;
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      {
        const int y = Y0 + j;
        if (colok && y <= n - 1)
          {
            const int    i = y * np + X;
            const double s = rsqrt(KC[i]);
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
    (void)sMM;
  }

  // --------------------------------------------------------------------------- dispatch
  template <int NL, int NRHS, int THREADS>
  static cudaError_t
  launch_one(const SolveParams &P, cudaStream_t st)
  {
    using C            = SmemCfg<NL, NRHS, THREADS>;
    const size_t bytes = C::smem_doubles * sizeof(double);
    auto         kern  = solve_smem_kernel<NL, NRHS, THREADS>;
    cudaError_t  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess)
      return e;
    kern<<<P.n_cells * (4 / NRHS), THREADS, bytes, st>>>(P);
    return cudaGetLastError();
  }

  bool
  smem_tier_supported(int l)
  {
    return l >= 3 && l <= 6;
  }

  // variant 0 = default per local mesh size; other values select experimental shapes
  cudaError_t
  launch_solve_smem(const Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches)
  {
    SolveParams P;
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
    const int variant = s.variant >= 100 ? s.variant - 100 : s.variant;
    switch (s.l)
      {
        case 3:
          return launch_one<3, 4, 64>(P, st);
        case 4:
          return launch_one<4, 4, 128>(P, st);
        case 5:
          if (variant == 1)
            return launch_one<5, 2, 256>(P, st);
          if (variant == 2)
            return launch_one<5, 4, 128>(P, st);
          return launch_one<5, 4, 256>(P, st);
        case 6:
          if (variant == 1)
            return launch_one<6, 2, 512>(P, st);
          if (variant == 2)
            return launch_one<6, 2, 256>(P, st);
          if (variant == 3)
            return launch_one<6, 1, 256>(P, st);
          return launch_one<6, 1, 512>(P, st);
        default:
          return cudaErrorInvalidValue;
      }
  }
} // namespace msb
