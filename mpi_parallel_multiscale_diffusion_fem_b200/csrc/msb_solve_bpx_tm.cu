// msb_solve_bpx_tm.cu -- the 64x64 local mesh with TWO right-hand sides in flight per CTA.
//
// Same algorithm as msb_solve_bpx.cu (multilevel-preconditioned CG on the unit-diagonal
// condensed system of one coarse cell, diffusion_problem_basis.tpp:450-465), but a CTA
// advances two of the cell's four bases together, so the seven coefficient loads of every
// stencil row, all index arithmetic, every block barrier and every reduction are shared by
// two solves.  At n = 64 that does not fit the classic way: the scaled operator alone is
// 128 KB of the 227 KB of shared memory and the register file is full with r, z and the
// stencil window of two right-hand sides.  Blackwell's third on-chip memory makes it fit:
//
//   shared memory  (224 KB)  E,N,D1,D2 coefficient arrays; ONE vector buffer that carries the
//                            search direction p during the stencil and the unscaled residual
//                            during the restriction; coarse-level vectors; reciprocal
//                            Galerkin diagonals (float: the preconditioner tolerates it)
//   registers      (128/thr) r, z / Ap, the 3x3x2 stencil window
//   tensor memory  (160 KB)  per-thread private spill space accessed with tcgen05.ld/st
//                            (32x32b shape: one warp owns a 32-lane quarter, a thread owns
//                            its lane's columns): the iterate x, the saved search direction
//                            p_old and sqrt(d).  Each is touched once per iteration.
//
// Measured on B200 (scripts/probes/tmem_probe.cu): tcgen05.st + tcgen05.ld round trips are
// bit exact and sustain ~556 B/clk/SM, four times the shared-memory crossbar.
// No tensor-core instruction is issued; TMEM is used as memory only.
#define MSB_STAGE_ARRAY g_msb_stage_cycles_tm
#include "msb_bpx_common.cuh"

#ifdef MSB_STAGE_TIMERS
__device__ unsigned long long g_msb_stage_cycles_tm[16];
extern "C" int
msb_debug_stage_cycles_tm(unsigned long long *out, int reset)
{
  cudaError_t e = cudaMemcpyFromSymbol(out, g_msb_stage_cycles_tm, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset)
    {
      unsigned long long z[16] = {0};
      e = cudaMemcpyToSymbol(g_msb_stage_cycles_tm, z, sizeof z);
    }
  return (int)e;
}
#endif

namespace msb
{
  namespace bpx
  {
    // ------------------------------------------------------------------ tensor memory access
    // 8 doubles = 16 consecutive 32-bit columns of this thread's TMEM lane
    __device__ __forceinline__ void
    tmem_ld8(uint32_t taddr, double (&d)[8])
    {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                     "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                     "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d[i] = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
    }

    // two independent loads in flight, one wait
    __device__ __forceinline__ void
    tmem_ld8x2(uint32_t ta, uint32_t tb, double (&da)[8], double (&db)[8])
    {
      uint32_t v[16], w[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                     "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                     "=r"(v[14]), "=r"(v[15])
                   : "r"(ta)
                   : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                     "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]),
                     "=r"(w[14]), "=r"(w[15])
                   : "r"(tb)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          da[i] = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
          db[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
        }
    }

    __device__ __forceinline__ void
    tmem_st8(uint32_t taddr, const double (&d)[8])
    {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          v[2 * i]     = (uint32_t)__double2loint(d[i]);
          v[2 * i + 1] = (uint32_t)__double2hiint(d[i]);
        }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                   "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                   :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
                   "r"(v[15])
                   : "memory");
    }

    __device__ __forceinline__ void
    tmem_wait_st()
    {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    template <int THREADS>
    struct TmCfg
    {
      static constexpr int NL = 6, NRHS = 2;
      using L                    = Levels<NL>;
      static constexpr int n     = 64, np = 65, N = np * np;
      static constexpr int NWARP = THREADS / 32;
      static constexpr int WX    = 2;
      static constexpr int WY    = NWARP / WX;
      static constexpr int RPT   = (n - 1 + WY - 1) / WY; // 8 (512 threads) or 16 (256 threads)
      static constexpr int NCH   = RPT / 4;               // TMEM chunks of 4 rows x 2 bases = 8 doubles
      static constexpr int CN    = L::CN;
      // per-thread TMEM columns: x | p_old | sqrt(d)
      static constexpr int XOFF = 0, POFF = 2 * RPT * NRHS, SOFF = 4 * RPT * NRHS;
      static constexpr int TCOLS = SOFF + 2 * RPT;
      static constexpr int TMEM_COLS = 512;
      // behind the per-warp column blocks: the dense inverse of the 7x7-level Galerkin operator,
      // the Exact7<THREADS> pieces of its rows in the tensor-memory lanes of warps 0..WARPS-1
      using X7 = Exact7<THREADS>;
      static constexpr int MOFF = (NWARP / 4) * TCOLS, MCOLS = ((X7::WARPS + 3) / 4) * 16 * X7::NCHK;
      static constexpr int RED   = 4 * NWARP;
      static constexpr size_t smem_bytes =
        sizeof(double) * (4 * (size_t)n * n + (size_t)NRHS * N + (size_t)NRHS * CN + 2 * RED + 8) +
        sizeof(float) * (size_t)CN;
      static_assert(RPT % 4 == 0 && NWARP % WX == 0, "strip shape");
      static_assert(MOFF + MCOLS <= TMEM_COLS, "tensor memory columns");
      static_assert(5 * CN <= NRHS * N, "Galerkin scratch must fit the vector buffer");
      static_assert(smem_bytes <= 232448, "shared memory");
    };

    // EXACT: the 7x7 coarse level is solved exactly (bpx::exact7_build, inverse in tensor memory)
    template <int THREADS, bool EXACT>
    __global__ void __launch_bounds__(THREADS, 1)
    solve_bpx_tm_kernel(BpxParams P)
    {
      using C             = TmCfg<THREADS>;
      using L             = typename C::L;
      constexpr int NL = C::NL, NRHS = C::NRHS, n = C::n, np = C::np, N = C::N;
      constexpr int NWARP = C::NWARP, WX = C::WX, RPT = C::RPT, NCH = C::NCH, CN = C::CN;

      extern __shared__ __align__(16) double smem[];
      double *sE   = smem;
      double *sN   = sE + n * n;
      double *sD1  = sN + n * n;
      double *sD2  = sD1 + n * n;
      double *sP   = sD2 + n * n;             // [N][2]: p (stencil) or u (restriction), zero halo
      double *sV   = sP + (size_t)NRHS * N;   // [CN][2] coarse residuals / corrections
      double *sRed = sV + (size_t)NRHS * CN;  // 2 reduction buffers
      float  *sDi  = reinterpret_cast<float *>(sRed + 2 * C::RED + 8); // [CN] 1 / Galerkin diagonal
      __shared__ uint32_t s_tmem_base;

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int cell = blockIdx.x;

      const double *S   = P.sten + (size_t)cell * ST_NARR * N;
      const double *KC  = S + ST_KC * N;
      const double *crn = P.corners + 8 * (size_t)cell;
      const double *q1  = P.q1coef + 16 * (size_t)cell;

      ST_DECL
      // ---------------------------------------------------------------- tensor memory
      if (warp == 0)
        {
          const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&s_tmem_base);
          asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr),
                       "r"(C::TMEM_COLS));
          asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t tmem_base = s_tmem_base;
      // this warp's 32-lane quarter (bits 31:16) and this warp's column block (bits 15:0)
      const uint32_t tm = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * C::TCOLS);

      // ---------------------------------------------------------------- prologue
      // (a) Galerkin hierarchy of the UNSCALED interior operator; scratch = the vector buffer
      {
        double       *G  = sP;
        const double *Sf = S;
        int           npf = np, Nf = N, goff = 0;
#pragma unroll 1
        for (int l = 1; l <= L::LEVELS; ++l)
          {
            const int npl = (n >> l) + 1, Nl = npl * npl, nin = npl - 2;
            double   *Gl = G + goff;
            for (int t = tid; t < 5 * Nl; t += THREADS)
              Gl[t] = 0.0;
            for (int t = tid; t < Nl; t += THREADS)
              sDi[goff / 5 + t] = 0.0f;
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
                sDi[goff / 5 + i] = (float)(1.0 / a[0]);
              }
            __syncthreads();
            Sf   = Gl;
            npf  = npl;
            Nf   = Nl;
            goff += 5 * Nl;
          }
      }
      // (a') exact coarse solve: invert the 49 x 49 operator of the 7x7 level in the (still unused)
      //      coefficient area and park it in tensor memory
      using X7            = typename C::X7;
      const uint32_t tmat = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) +
                            (uint32_t)(C::MOFF + (warp >> 2) * 16 * X7::NCHK);
      if constexpr (EXACT)
        {
          exact7_build<THREADS>(sP + 5 * L::lvl_off(L::LW + 1), sE, sN, tid); // scratch: the next coefficient array
          if (warp < X7::WARPS)
            {
#pragma unroll
              for (int c = 0; c < X7::NCHK; ++c)
                {
                  double g[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    g[i] = X7::fetch(sE, tid, c, i);
                  tmem_st8(tmat + 16 * c, g);
                }
              tmem_wait_st();
            }
          __syncthreads();
        }
      // (b) s = d^-1/2 on every node (the hierarchy scratch is dead from here on)
      double *sS = sP;
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

      // ---------------------------------------------------------------- ownership
      const int  wx = warp % WX, wy = warp / WX;
      const int  X  = 1 + 32 * wx + lane;
      const int  Y0 = 1 + RPT * wy;
      const bool colok = X <= n - 1;

      // sqrt(d) of the owned DoFs -> tensor memory (kept for all four bases of the cell)
#pragma unroll
      for (int c = 0; c < RPT / 8; ++c)
        {
          double sq8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            {
              const int y = Y0 + 8 * c + j;
              sq8[j]      = (colok && y <= n - 1) ? sqrt(KC[y * np + X]) : 0.0;
            }
          tmem_st8(tm + C::SOFF + 16 * c, sq8);
        }
      tmem_wait_st();
      ST_MARK(0)

#pragma unroll 1
      for (int grp = 0; grp < 4 / NRHS; ++grp)
        {
          const int rhs0 = grp * NRHS;
          // (d) clear the vector buffer and the coarse vectors (halos stay zero for the solve)
          for (int i = tid; i < NRHS * N + NRHS * CN; i += THREADS)
            sP[i] = 0.0;
          __syncthreads();

          double r[RPT][NRHS], q[RPT][NRHS];
          // the pre-summed residual strips share the vector buffer with p and overwrite parts of
          // p's zero halo: restore the halo whenever p is rewritten
          auto zero_halo = [&]() {
            const double zero2[NRHS] = {0.0, 0.0};
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
                stv<NRHS>(sP, jy * np + jx, zero2);
              }
          };

          // (e) rhat_0 = -D^-1/2 K_IB g_B ; x = 0 in tensor memory
#pragma unroll
          for (int j = 0; j < RPT; ++j)
            {
              const int y = Y0 + j;
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                r[j][k] = 0.0, q[j][k] = 0.0;
              if (colok && y <= n - 1 && (X == 1 || X == n - 1 || y == 1 || y == n - 1))
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
                  const double s = rsqrt(KC[y * np + X]);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    r[j][k] = -s * acc[k];
                }
            }
          {
            const double zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              tmem_st8(tm + C::XOFF + 16 * c, zero8);
          }

          // ------------------------------------------ zhat = Mhat^-1 rhat (into z), r.z and ||r||^2
          auto precondition = [&](double (&z)[RPT][NRHS], double (&rz)[NRHS], double (&rr)[NRHS]) {
            // u = D^1/2 rhat, pre-summed per strip (Presum) into the vector buffer (p is dead there:
            // it is saved in tensor memory)
            {
              using PS = Presum<NL, NRHS, RPT>;
              double    acc[NRHS];
              const int pc = PS::col(X);
#pragma unroll
              for (int c = 0; c < RPT / 8; ++c)
                {
                  double sq8[8];
                  tmem_ld8(tm + C::SOFF + 16 * c, sq8);
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj)
                    {
                      double u[NRHS];
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        {
                          u[k]  = sq8[jj] * r[8 * c + jj][k]; // zero beyond the mesh
                          rr[k] = fma(u[k], u[k], rr[k]);
                        }
                      PS::push(sP, 8 * c + jj, u, acc, pc, wy, colok);
                    }
                }
            }
            __syncthreads();
            ST_MARK(4)
            if constexpr (EXACT)
              coarse_correction<NL, NRHS, THREADS, RPT, true>(
                sP, sV, sDi, tid, warp, lane,
                [&](int st_k) {
                  (void)st_k;
                  ST_MARK(st_k)
                },
                [&](int c, double(&g)[8]) { tmem_ld8(tmat + 16 * c, g); });
            else
              coarse_correction<NL, NRHS, THREADS, RPT>(sP, sV, sDi, tid, warp, lane, [&](int st_k) {
                (void)st_k;
                ST_MARK(st_k)
              });
            ST_MARK(11)
            // level 0: zhat = rhat + D^1/2 (P z_1), coarse-row averages cached in registers
            {
              constexpr int np1 = L::lvl_np(1);
              const int     xl = X >> 1, xh = (X + 1) >> 1, cr0 = (Y0 - 1) >> 1;
#pragma unroll
              for (int c = 0; c < RPT / 8; ++c)
                {
                  double sq8[8];
                  tmem_ld8(tm + C::SOFF + 16 * c, sq8);
                  if (colok)
                    {
                      double h[5][NRHS];
#pragma unroll
                      for (int m = 0; m <= 4; ++m)
                        {
                          const int cr = cr0 + 4 * c + m <= n / 2 ? cr0 + 4 * c + m : n / 2;
                          double    a[NRHS], b[NRHS];
                          ldv<NRHS>(sV, cr * np1 + xl, a);
                          ldv<NRHS>(sV, cr * np1 + xh, b);
#pragma unroll
                          for (int k = 0; k < NRHS; ++k)
                            h[m][k] = 0.5 * (a[k] + b[k]);
                        }
#pragma unroll
                      for (int jj = 0; jj < 8; ++jj)
                        {
                          const int j = 8 * c + jj, y = Y0 + j;
                          if (y <= n - 1)
                            {
#pragma unroll
                              for (int k = 0; k < NRHS; ++k)
                                {
                                  const double cc =
                                    (jj & 1) ? h[(jj + 1) / 2][k] : 0.5 * (h[jj / 2][k] + h[jj / 2 + 1][k]);
                                  z[j][k] = fma(sq8[jj], cc, r[j][k]);
                                  rz[k]   = fma(r[j][k], z[j][k], rz[k]);
                                }
                            }
                        }
                    }
                }
            }
          };

          // (f) z_0, p_0 = z_0 (vector buffer + tensor memory), rho = r.z, initial ||r||
          double rho[NRHS], exact[NRHS];
          {
            double rz[NRHS] = {0.0, 0.0}, rr[NRHS] = {0.0, 0.0};
            precondition(q, rz, rr);
            double four[4] = {rz[0], rz[1], rr[0], rr[1]};
            block_sum4<NWARP>(four, sRed, warp, lane);
            rho[0] = four[0], rho[1] = four[1], exact[0] = four[2], exact[1] = four[3];
            // the restriction has finished reading u (barriers inside coarse_correction)
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              {
                double p8[8];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  {
                    const int j = 4 * c + jj, y = Y0 + j;
                    p8[2 * jj] = q[j][0], p8[2 * jj + 1] = q[j][1];
                    if (colok && y <= n - 1)
                      stv<NRHS>(sP, y * np + X, q[j]);
                  }
                tmem_st8(tm + C::POFF + 16 * c, p8);
              }
            zero_halo();
            tmem_wait_st();
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

          // -------------------------------------------------------------- PCG iterations
          int it = 0;
          while (!all_done && it < P.max_iter)
            {
              ++it;
              // ---- q = Ahat p for both bases: one set of coefficient loads per stencil row
              double pq[NRHS] = {0.0, 0.0};
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
                              a0[k] = b0[k], a1[k] = b1[k], a2[k] = b2[k];
                              b0[k] = c0[k], b1[k] = c1[k], b2[k] = c2[k];
                            }
                          cS = cN;
                        }
                    }
                }
              ST_MARK(2)
              block_sum2<NWARP>(pq[0], pq[1], sRed, warp, lane);
              ST_MARK(3)

              double alpha[NRHS];
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                alpha[k] = done[k] ? 0.0 : fast_div(rho[k], pq[k]);

              // ---- r -= alpha q ; z = M^-1 r (into q) ; rho' = r.z ; ||r||^2
#pragma unroll
              for (int j = 0; j < RPT; ++j)
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  r[j][k] = fma(-alpha[k], q[j][k], r[j][k]);
              double rz[NRHS] = {0.0, 0.0}, rr[NRHS] = {0.0, 0.0};
              precondition(q, rz, rr);
              {
                double four[4] = {rz[0], rz[1], rr[0], rr[1]};
                block_sum4<NWARP>(four, sRed + C::RED, warp, lane);
                rz[0] = four[0], rz[1] = four[1], rr[0] = four[2], rr[1] = four[3];
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

              // ---- x += alpha p_old ; p = z + beta p_old: x and p_old live in tensor memory
#pragma unroll
              for (int c = 0; c < NCH; ++c)
                {
                  double x8[8], p8[8];
                  // (two loads in flight would need 32 more live registers and spill: measured slower)
                  tmem_ld8(tm + C::XOFF + 16 * c, x8);
                  tmem_ld8(tm + C::POFF + 16 * c, p8);
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj)
                    {
                      const int j = 4 * c + jj, y = Y0 + j;
                      double    pn[NRHS];
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        {
                          const double po = p8[2 * jj + k];
                          x8[2 * jj + k]  = fma(alpha[k], po, x8[2 * jj + k]);
                          pn[k]           = (done[k] && beta[k] == 0.0) ? po : fma(beta[k], po, q[j][k]);
                          p8[2 * jj + k]  = pn[k];
                        }
                      if (colok && y <= n - 1)
                        stv<NRHS>(sP, y * np + X, pn);
                    }
                  tmem_st8(tm + C::XOFF + 16 * c, x8);
                  tmem_st8(tm + C::POFF + 16 * c, p8);
                }
              zero_halo();
              tmem_wait_st();
              __syncthreads();
              ST_MARK(9)
            }

          // -------------------------------------------------------------- epilogue
          // distribute() (basis.tpp:308): interior phi = D^-1/2 yhat, boundary phi = g
          double *out = P.phi + ((size_t)cell * 4 + rhs0) * N;
#pragma unroll
          for (int c = 0; c < RPT / 8; ++c)
            {
              double sq8[8], xa[8], xb[8];
              tmem_ld8(tm + C::SOFF + 16 * c, sq8);
              tmem_ld8x2(tm + C::XOFF + 32 * c, tm + C::XOFF + 32 * c + 16, xa, xb); // rows 8c.., 8c+4..
#pragma unroll
              for (int jj = 0; jj < 8; ++jj)
                {
                  const int y = Y0 + 8 * c + jj;
                  if (colok && y <= n - 1)
                    {
                      const int    i = y * np + X;
                      const double s = 1.0 / sq8[jj];
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        out[(size_t)k * N + i] = s * (jj < 4 ? xa[2 * jj + k] : xb[2 * (jj - 4) + k]);
                    }
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
          __syncthreads(); // shared buffers are reused by the next pair of bases
          ST_MARK(10)
        }
      ST_FLUSH
      // ---------------------------------------------------------------- release tensor memory
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
    }

    template <int THREADS, bool EXACT>
    static cudaError_t
    launch_tm(const BpxParams &P, cudaStream_t st)
    {
      using C           = TmCfg<THREADS>;
      auto        kern  = solve_bpx_tm_kernel<THREADS, EXACT>;
      cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
      if (e != cudaSuccess)
        return e;
      kern<<<P.n_cells, THREADS, C::smem_bytes, st>>>(P);
      return cudaGetLastError();
    }
  } // namespace bpx

  // variant 4: 512 threads (8 DoFs x 2 bases per thread); variant 5: 256 threads (16 x 2)
  cudaError_t
  launch_solve_bpx_tm(const BpxParams &P, int threads, bool exact7, cudaStream_t st)
  {
    if (threads == 256)
      return bpx::launch_tm<256, false>(P, st);
    return exact7 ? bpx::launch_tm<512, true>(P, st) : bpx::launch_tm<512, false>(P, st);
  }
} // namespace msb
