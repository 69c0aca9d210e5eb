// msb_bpx_common.cuh -- device helpers shared by the shared-memory-resident multilevel PCG
// kernels (msb_solve_bpx.cu: one right-hand side per pass; msb_solve_bpx_tm.cu: two, with
// tensor memory as per-thread spill space).
#pragma once

#include <math.h>

#include <type_traits>

#include "msb_internal.cuh"
#include "msb_tmem.cuh"

// Optional per-stage cycle timers (profiling build only: make EXTRA=-DMSB_STAGE_TIMERS).
#ifdef MSB_STAGE_TIMERS
// the including .cu defines MSB_STAGE_ARRAY (its own __device__ unsigned long long [16])
#  define ST_DECL long long st_t0 = clock64(), st_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#  define ST_MARK(i)                       \
    if (threadIdx.x == 0)                  \
      {                                    \
        const long long st_t1 = clock64(); \
        st_acc[i] += st_t1 - st_t0;        \
        st_t0 = st_t1;                     \
      }
#  define ST_FLUSH                                                              \
    if (threadIdx.x == 0)                                                       \
      for (int st_i = 0; st_i < 16; ++st_i)                                     \
        atomicAdd(&MSB_STAGE_ARRAY[st_i], (unsigned long long)st_acc[st_i]);
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

  // two right-hand sides per pass with tensor memory as spill space (msb_solve_bpx_tm.cu)
  cudaError_t launch_solve_bpx_tm(const BpxParams &P, int threads, bool exact7, cudaStream_t st);

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
#ifndef MSB_EMU
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
#else
      y = (double)(1.0f / (float)b); // host emulation of the cluster kernel (scripts/emu)
#endif
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

    // block-wide sums of FOUR values: two transposing exchanges leave one value per quarter
    // warp, so the butterfly costs 6 double shuffles instead of 20
    template <int NWARP>
    __device__ __forceinline__ void
    block_sum4(double (&v)[4], double *buf, int warp, int lane)
    {
      static_assert(NWARP <= 16, "two partials per lane in the second stage");
      const bool up16 = lane & 16, up8 = lane & 8;
      double     a0 = up16 ? v[2] : v[0], a1 = up16 ? v[3] : v[1];
      a0 += __shfl_xor_sync(0xffffffffu, up16 ? v[0] : v[2], 16);
      a1 += __shfl_xor_sync(0xffffffffu, up16 ? v[1] : v[3], 16);
      double t = (up8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a1, 8);
#pragma unroll
      for (int off = 4; off > 0; off >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, off);
      // layout [warp / 8][value][warp % 8]: the four stores of a warp fall into different banks and the second
      // stage reads 32 consecutive doubles (buf holds max(32, 4 NWARP) entries); [value][warp] was a 4-way
      // conflict on both sides
      if ((lane & 7) == 0)
        buf[(warp >> 3) * 32 + lane + (warp & 7)] = t;
      __syncthreads();
      const int w = lane & 7;
      double    u = w < NWARP ? buf[lane] : 0.0;
      if (w + 8 < NWARP)
        u += buf[32 + lane];
#pragma unroll
      for (int off = 4; off > 0; off >>= 1)
        u += __shfl_xor_sync(0xffffffffu, u, off);
      v[0] = __shfl_sync(0xffffffffu, u, 0);
      v[1] = __shfl_sync(0xffffffffu, u, 8);
      v[2] = __shfl_sync(0xffffffffu, u, 16);
      v[3] = __shfl_sync(0xffffffffu, u, 24);
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
    // get(ix, iy, ex, ey): entry of fine row node (ix,iy) towards (ix+ex, iy+ey)
    template <class Get>
    __device__ __forceinline__ void
    galerkin_row_of(Get &&get, int X, int Y, double (&acc)[5])
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
                  const double v = wa * get(ix, iy, ex, ey);
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

    __device__ __forceinline__ void
    galerkin_row(const double *Sf, int npf, int Nf, int X, int Y, double (&acc)[5])
    {
      galerkin_row_of([&](int ix, int iy, int ex, int ey) { return sten_get(Sf, npf, Nf, ix, iy, ex, ey); }, X, Y,
                      acc);
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


    // level bookkeeping of the coarsened interior grids of an n = 2^NL local mesh
    template <int NL>
    struct Levels
    {
      static constexpr int n      = 1 << NL;
      static constexpr int LEVELS = NL - 1;               // coarse levels 1..NL-1 (last: one unknown)
      static constexpr int LW     = NL >= 4 ? NL - 4 : 0; // levels 1..LW have >= 15x15 unknowns
      __host__ __device__ static constexpr int
      lvl_np(int l)
      {
        return (n >> l) + 1;
      }
      __host__ __device__ static constexpr int
      lvl_off(int l) // offset (in nodes) of level l >= 1 inside the packed level arrays
      {
        int o = 0;
        for (int k = 1; k < l; ++k)
          o += lvl_np(k) * lvl_np(k);
        return o;
      }
      static constexpr int CN = lvl_off(NL); // total coarse nodes
    };

    // The coarse part of the multilevel preconditioner, entirely in shared memory:
    // on entry sU holds the staged unscaled fine residual (stored and block-synchronised);
    // on exit level 1 of sV holds z_1 = sum_{l>=1} P_{l->1} D_l^-1 P_l^T u (block-synchronised).
    // sDi: reciprocal Galerkin diagonals of all levels (double or float storage).
    // Staging of the fine residual for the restriction to level 1 (local meshes with n >= 32):
    // the thread that owns column X and fine rows Y0 .. Y0+RPT-1 (Y0 odd, RPT even) forms the
    // VERTICAL full-weighting sums of its own strip in registers and stores
    //   - the RPT/2-1 complete coarse rows and the partial sum of the coarse row on the upper
    //     strip edge into T[cy][col(X)],
    //   - half its first row (the missing part of the coarse row on the lower strip edge)
    //     into TB[strip][col(X)],
    // with the columns de-interleaved (odd X in the upper half of a row) so that the horizontal
    // 3-point combination of the level-1 stage reads consecutive addresses.
    // PAD > 0 shifts the odd-column half of a row by PAD entries: a quarter warp stores four even and four
    // odd columns at once, and with PAD = 5 (mod 8) entries of 16 bytes the two groups fall into disjoint
    // banks (PAD = 0: 2-way conflicts on every store of two right-hand sides, 20 % of the store wavefronts
    // of the round-1 kernel, profiles/r01u_ncu_full_*).
    template <int NL, int NRHS, int RPT, int PAD = 0>
    struct Presum
    {
      static constexpr int n = 1 << NL, ROW = n + PAD, HALF = RPT / 2;
      static constexpr int HOFF = n / 2 + PAD; // entry offset of the odd columns inside a row
      static constexpr int TB = (n / 2) * ROW; // entry offset of the strip-edge contributions
      static constexpr int ENTRIES = (n / 2 + (n - 1 + RPT - 1) / RPT) * ROW; // staging footprint
      __device__ static __forceinline__ int
      col(int X)
      {
        return (X & 1) * HOFF + (X >> 1);
      }
      // Streaming form: call for j = 0 .. RPT-1 in order (j is a compile-time constant after
      // unrolling) with u_j = the unscaled residual of row Y0+j (zero beyond the mesh); `acc`
      // carries the running vertical sum, so only one row is live at a time.
      __device__ static __forceinline__ void
      push(double *sT, int j, const double (&uj)[NRHS], double (&acc)[NRHS], int c, int wy, bool colok)
      {
        const int cr0 = HALF * wy;
        if ((j & 1) == 0)
          {
            if (j > 0)
              {
                // row j closes coarse row cr0 + j/2:  0.5 u[j-2] + u[j-1] + 0.5 u[j]
                double v[NRHS];
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  v[k] = fma(0.5, uj[k], acc[k]);
                if (colok && cr0 + j / 2 <= n / 2 - 1)
                  stv<NRHS>(sT, (cr0 + j / 2) * ROW + c, v);
              }
            else if (colok && wy >= 1)
              {
                // half the first row: the missing part of the coarse row on the lower strip edge
                double v[NRHS];
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  v[k] = 0.5 * uj[k];
                stv<NRHS>(sT, TB + wy * ROW + c, v);
              }
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              acc[k] = 0.5 * uj[k];
          }
        else
          {
#pragma unroll
            for (int k = 0; k < NRHS; ++k)
              acc[k] += uj[k];
            if (j == RPT - 1 && colok && cr0 + HALF <= n / 2 - 1)
              stv<NRHS>(sT, (cr0 + HALF) * ROW + c, acc); // partial sum of the upper strip edge
          }
      }
    };

    // Full weighting of coarse node (cx, cy) from a finer level vector Vf ([node][NRHS], npf = 2^m + 1 >= 9 nodes per row).
    // The lanes of a quarter warp read every second node of a row: with 16-byte nodes (NRHS = 2) that is a 32-byte
    // stride, lanes cx and cx + 4 in the same banks -- a 2-way conflict on each of the nine loads (three quarters of the
    // conflict wavefronts that were left in the fused n = 64 kernel, profiles/r02f_*).  A row is 16 bytes longer than a
    // multiple of 128, and so is one step in x: the lanes with rot = 1 (cx in 5..8, 13..16) visit the nine points of
    // the window ONE POSITION AHEAD in the lexicographic order, which puts them an odd multiple of 16 bytes away from
    // the other lanes in eight of the nine load instructions (the ninth pairs (1,1) with (-1,-1)).  Weighted
    // accumulation in three partial sums instead of three row sums: the result differs from the unrotated order in the
    // last bits only (the preconditioner stays a fixed linear operator).
    template <int NRHS>
    __device__ __forceinline__ void
    full_weighting_rot(const double *Vf, int npf, int cx, int cy, int rot, double (&acc)[NRHS])
    {
      double part[3][NRHS];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int k = 0; k < NRHS; ++k)
          part[q][k] = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i)
        {
          constexpr int AY[10] = {-1, -1, -1, 0, 0, 0, 1, 1, 1, -1}, DX[10] = {-1, 0, 1, -1, 0, 1, -1, 0, 1, -1};
          const int     ay = rot ? AY[i + 1] : AY[i], dx = rot ? DX[i + 1] : DX[i];
          const double  w0 = (AY[i] == 0 ? 1.0 : 0.5) * (DX[i] == 0 ? 1.0 : 0.5);
          const double  w1 = (AY[i + 1] == 0 ? 1.0 : 0.5) * (DX[i + 1] == 0 ? 1.0 : 0.5);
          const double  w  = rot ? w1 : w0;
          double        v[NRHS];
          ldv<NRHS>(Vf, (2 * cy + ay) * npf + 2 * cx + dx, v);
#pragma unroll
          for (int k = 0; k < NRHS; ++k)
            part[i % 3][k] = fma(w, v[k], part[i % 3][k]);
        }
#pragma unroll
      for (int k = 0; k < NRHS; ++k)
        acc[k] = (part[0][k] + part[1][k]) + part[2][k];
    }

    // mark(i): optional stage-timer hook (a no-op lambda in production builds).
    // How the 49 x 49 inverse is spread over threads for the matvec: row r = tid / PARTS, the
    // columns of a row in PARTS contiguous pieces of CH entries, fetched in NCHK chunks of 8.
    template <int THREADS>
    struct Exact7
    {
      // 128-thread kernels: one thread per row (measured: better than 2 lanes per row); 512 threads:
      // 8 lanes per row, one chunk of 7 entries each, so the whole CTA shares the chain
      static constexpr int PARTS = THREADS >= 512 ? 8 : (THREADS >= 256 ? 4 : 1);
      static constexpr int CH    = (49 + PARTS - 1) / PARTS; // entries per piece
      static constexpr int NCHK  = (CH + 7) / 8;             // chunks of 8
      static constexpr int WARPS = (49 * PARTS + 31) / 32;   // participating warps
      // entry (chunk c, slot i) of thread tid: element [row][col] of the (symmetric) inverse
      __device__ static __forceinline__ double
      fetch(const double *Gi, int tid, int c, int i)
      {
        const int r = tid / PARTS, jl = 8 * c + i, j = (tid % PARTS) * CH + jl;
        return (r < 49 && jl < CH && j < 49) ? Gi[j * 49 + r] : 0.0;
      }
      // the same entry from the packed lower triangle T[i (i + 1) / 2 + j], j <= i (exact7_build<THREADS, true>):
      // half the footprint (1225 doubles).  For a fixed column the lanes below it read consecutive words and the
      // lanes from it on read T[r (r + 1) / 2 + j]: the triangular numbers of 16 consecutive r are distinct mod 16,
      // so a half warp touches 16 different 8-byte banks
      __device__ static __forceinline__ double
      fetch_tri(const double *T, int tid, int c, int i)
      {
        const int r = tid / PARTS, jl = 8 * c + i, j = (tid % PARTS) * CH + jl;
        const int hi = r > j ? r : j, lo = r > j ? j : r;
        return (r < 49 && jl < CH && j < 49) ? T[hi * (hi + 1) / 2 + lo] : 0.0;
      }
    };
    constexpr int EXACT7_TRI = 49 * 50 / 2; // doubles of the packed inverse

    // Dense inverse of the Galerkin operator of the 7x7-unknown level (9-point stencil Gl on 9x9
    // nodes, layout of Shard::d_sten) into sGi[49][49].  With it the levels below 15x15 are solved
    // EXACTLY instead of being diagonally scaled: the preconditioner becomes
    // D^-1 + ... + P_7 A_7^-1 P_7^T, which removes the contrast dependence of the coarse part
    // (cfg4: 67 -> 38 iterations).
    // The operator is SPD and banded (lexicographic 7x7 grid, 9-point stencil: half bandwidth 8), so
    // the inverse is built from a banded L D L^T factorisation (warp 0: 49 steps of <= 36 entry updates,
    // __syncwarp only) followed by 49 independent pairs of band substitutions, one column of the
    // inverse per thread: ~15 k cycles instead of the ~55 k of a dense Gauss-Jordan sweep with two CTA
    // barriers per pivot, which was 20 % of the n = 32 kernel (profiles/r01s_stage_timers_cfg4_2368.txt).
    // sBand: scratch of EXACT7_SCRATCH doubles (band storage [49 + 8][9], zero padded).
    constexpr int EXACT7_SCRATCH = (49 + 8) * 9;
    // TRI: sGi receives only the lower triangle, packed (EXACT7_TRI doubles, Exact7::fetch_tri): column c of the
    // inverse vanishes above row c after the forward substitution, and the entries above the diagonal that the
    // backward substitution produces are the transposes of entries other threads store.
    template <int THREADS, bool TRI = false>
    __device__ __forceinline__ void
    exact7_build(const double *Gl, double *sGi, double *sBand, int tid)
    {
      auto at = [](int i, int c) { return TRI ? i * (i + 1) / 2 + c : i * 49 + c; };
      constexpr int BW = 8, LD = BW + 1;
      // lower band: sBand[i * LD + b] = A(i, i - b), b = 0 .. 8, i = (Y-1) * 7 + (X-1)
      for (int t = tid; t < EXACT7_SCRATCH; t += THREADS)
        sBand[t] = 0.0;
      __syncthreads();
      if (tid < 49)
        {
          const int X = 1 + tid % 7, Y = 1 + tid / 7;
          sBand[tid * LD + 0] = sten_get(Gl, 9, 81, X, Y, 0, 0);
          if (X > 1)
            sBand[tid * LD + 1] = sten_get(Gl, 9, 81, X, Y, -1, 0);
          if (Y > 1)
            {
              if (X < 7)
                sBand[tid * LD + 6] = sten_get(Gl, 9, 81, X, Y, 1, -1);
              sBand[tid * LD + 7] = sten_get(Gl, 9, 81, X, Y, 0, -1);
              if (X > 1)
                sBand[tid * LD + 8] = sten_get(Gl, 9, 81, X, Y, -1, -1);
            }
        }
      __syncthreads();
      // ---- A = L D L^T in place (right looking): after step k column k holds L(.,k), sBand[k*LD] = 1/D_k
      if (tid < 32)
        {
          // a step updates the entries (k+a, k+c), 1 <= c <= a <= 8: 36 pairs, lanes 0..3 take two
          int pa[2], pc[2];
#pragma unroll
          for (int u = 0; u < 2; ++u)
            {
              int q = tid + 32 * u, a = 1;
              while (q >= a)
                q -= a, ++a;
              pa[u] = a, pc[u] = q + 1; // (u = 1 is only meaningful for lanes 0..3: a = 8)
            }
#pragma unroll 1
          for (int k = 0; k < 49; ++k)
            {
              const double rd = fast_div(1.0, sBand[k * LD]);
#pragma unroll
              for (int u = 0; u < 2; ++u)
                if ((u == 0 || tid < 4) && k + pa[u] < 49)
                  {
                    const int a = pa[u], c = pc[u];
                    sBand[(k + a) * LD + (a - c)] =
                      fma(-(sBand[(k + a) * LD + a] * rd), sBand[(k + c) * LD + c], sBand[(k + a) * LD + (a - c)]);
                  }
              __syncwarp();
              if (tid >= 1 && tid <= BW && k + tid < 49)
                sBand[(k + tid) * LD + tid] *= rd;
              if (tid == 0)
                sBand[k * LD] = rd;
              __syncwarp();
            }
        }
      __syncthreads();
      // ---- column c of the inverse: L y = e_c, z = D^-1 y, L^T x = z (window of the last 8 values in registers)
      if (tid < 49)
        {
          const int c = tid;
          double    w[BW];
#pragma unroll
          for (int b = 0; b < BW; ++b)
            w[b] = 0.0;
#pragma unroll 7
          for (int i = 0; i < 49; ++i)
            {
              // two partial sums (even / odd offsets), the newest value last: half the dependency chain
              double s = i == c ? 1.0 : 0.0, s2 = 0.0;
#pragma unroll
              for (int b = BW; b >= 2; b -= 2)
                {
                  s2 = fma(-sBand[i * LD + b], w[b - 1], s2);
                  s  = fma(-sBand[i * LD + b - 1], w[b - 2], s);
                }
              s += s2;
#pragma unroll
              for (int b = BW - 1; b >= 1; --b)
                w[b] = w[b - 1];
              w[0]            = s;
              if (!TRI || i >= c)
                sGi[at(i, c)] = s * sBand[i * LD];
            }
#pragma unroll
          for (int b = 0; b < BW; ++b)
            w[b] = 0.0;
#pragma unroll 7
          for (int i = 48; i >= 0; --i)
            {
              if (TRI && i < c)
                break; // (rows above the diagonal: stored by the threads of those columns)
              double s = sGi[at(i, c)], s2 = 0.0;
#pragma unroll
              for (int b = BW; b >= 2; b -= 2) // rows beyond 48 are the zero padding of sBand
                {
                  s2 = fma(-sBand[(i + b) * LD + b], w[b - 1], s2);
                  s  = fma(-sBand[(i + b - 1) * LD + b - 1], w[b - 2], s);
                }
              s += s2;
#pragma unroll
              for (int b = BW - 1; b >= 1; --b)
                w[b] = w[b - 1];
              w[0]          = s;
              sGi[at(i, c)] = s;
            }
        }
      __syncthreads();
    }

    // RPT > 0: sU holds the pre-summed strips of Presum<NL,NRHS,RPT>; RPT == 0: sU holds u itself.
    // EXACT7: gi(c, g) delivers rows 8c..8c+7 of column tid of the inverse built by exact7_build
    // (from shared memory or from tensor memory; called by all lanes of warps 0-1); the 3x3 and
    // 1x1 levels are not used.
    // PAD: the Presum padding the staging was written with.  dot1 != nullptr: the caller wants
    // sum_{level-1 nodes} u_1 z_1 (restricted residual times final level-1 correction = r.z minus the fine-level
    // part, see msb_solve_fused.cu) accumulated into dot1[NRHS] per thread, and does the block barrier that
    // publishes z_1 itself (inside its block reduction): the final __syncthreads is skipped.
    template <int NL, int NRHS, int THREADS, int RPT, bool EXACT7 = false, int PAD = 0, class DiT, class Mark,
              class GiChunk = int>
    __device__ __forceinline__ void
    coarse_correction(const double *sU, double *sV, const DiT *sDi, int tid, int warp, int lane, Mark &&mark,
                      GiChunk &&gi = 0, double *dot1 = nullptr)
    {
      using L             = Levels<NL>;
      constexpr int NWARP = THREADS / 32;
      constexpr int n     = L::n;
      constexpr int np    = n + 1;
        // Level sweeps.  "Wide" levels (>= 15x15 unknowns) are done by the whole CTA with a block
      // barrier each; the remaining tiny levels form a short serial chain on warp 0.
      // restrict(l): r_l = P^T r_{l-1} (full weighting); source of level 1 is the staged u.
      auto restrict_level = [&](auto lc, int first, int nthr) {
        constexpr int l   = decltype(lc)::value;
        constexpr int W   = n >> l, LG = NL - l, npl = W + 1;
        constexpr int npf = (n >> (l - 1)) + 1;
        const double *Vf  = l == 1 ? sU : sV + (size_t)NRHS * L::lvl_off(l - 1);
        double       *Vl  = sV + (size_t)NRHS * L::lvl_off(l);
        for (int t = first; t < W * W; t += nthr)
          {
            const int cx = 1 + (t & (W - 1)), cy = 1 + (t >> LG);
            if (cx > W - 1 || cy > W - 1)
              continue;
            double acc[NRHS];
            if constexpr (NRHS == 2 && npf >= 9)
              full_weighting_rot<NRHS>(Vf, npf, cx, cy, ((cx - 1) >> 2) & 1, acc);
            else
              {
                // three independent row sums, then combined (short dependency chains)
                double row[3][NRHS];
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
              }
            stv<NRHS>(Vl, cy * npl + cx, acc);
          }
      };
      // prolong(l): z_l = r_l / D_l + P z_{l+1}  (coarsest level: z = r / D)
      auto prolong_level = [&](auto lc, int first, int nthr) {
        constexpr int l   = decltype(lc)::value;
        constexpr int W   = n >> l, LG = NL - l, npl = W + 1;
        double       *Vl  = sV + (size_t)NRHS * L::lvl_off(l);
        const DiT *Dl  = sDi + L::lvl_off(l);
        for (int t = first; t < W * W; t += nthr)
          {
            const int fx = 1 + (t & (W - 1)), fy = 1 + (t >> LG);
            if (fx > W - 1 || fy > W - 1)
              continue;
            const int i = fy * npl + fx;
            double    v[NRHS];
            ldv<NRHS>(Vl, i, v);
            [[maybe_unused]] double v_in[NRHS];
            if constexpr (l == 1)
              {
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  v_in[k] = v[k];
              }
            const double di = Dl[i];
            if constexpr (l < L::LEVELS)
              {
                constexpr int npc = (n >> (l + 1)) + 1;
                const double *Vc  = sV + (size_t)NRHS * L::lvl_off(l + 1);
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
            if constexpr (l == 1)
              {
                if (dot1)
                  {
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      dot1[k] = fma(v_in[k], v[k], dot1[k]);
                  }
              }
          }
      };
      // down: wide levels
      if constexpr (RPT > 0)
        {
          // level 1 from the pre-summed strips: horizontal 3-point combination, conflict-free
          static_assert(L::LW >= 1, "pre-summed staging needs a wide level 1");
          using PS          = Presum<NL, NRHS, RPT, PAD>;
          constexpr int W   = n >> 1, LG = NL - 1, np1 = W + 1;
          for (int t = tid; t < W * W; t += THREADS)
            {
              const int cx = 1 + (t & (W - 1)), cy = 1 + (t >> LG);
              if (cx > W - 1 || cy > W - 1)
                continue;
              const double *row = sU + (size_t)NRHS * (cy * PS::ROW);
              double        a[NRHS], b[NRHS], c[NRHS];
              ldv<NRHS>(row, PS::HOFF + cx - 1, a); // X = 2cx-1
              ldv<NRHS>(row, cx, b);                // X = 2cx
              ldv<NRHS>(row, PS::HOFF + cx, c);     // X = 2cx+1
              if (cy % PS::HALF == 0)
                {
                  const double *rb = sU + (size_t)NRHS * (PS::TB + (cy / PS::HALF) * PS::ROW);
                  double        a2[NRHS], b2[NRHS], c2[NRHS];
                  ldv<NRHS>(rb, PS::HOFF + cx - 1, a2);
                  ldv<NRHS>(rb, cx, b2);
                  ldv<NRHS>(rb, PS::HOFF + cx, c2);
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    a[k] += a2[k], b[k] += b2[k], c[k] += c2[k];
                }
              double o[NRHS];
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                o[k] = fma(0.5, a[k] + c[k], b[k]);
              stv<NRHS>(sV, cy * np1 + cx, o);
            }
          __syncthreads();
          for_levels<2, L::LW>([&](auto lc) {
            restrict_level(lc, tid, THREADS);
            __syncthreads();
          });
        }
      else
        {
          for_levels<1, L::LW>([&](auto lc) {
            restrict_level(lc, tid, THREADS);
            __syncthreads();
          });
        }
      mark(5);
      // the three tiny levels below the 15x15 level B = LW (7x7, 3x3, 1x1 unknowns)
      if constexpr (L::LW >= 1)
        {
          // No serial chain: their residuals are restricted DIRECTLY from level B (a product
          // of full-weighting restrictions is the restriction with the nested hat function) by
          // different warps in parallel, and their corrections are interpolated DIRECTLY back
          // to level B (a product of bilinear interpolations is bilinear on the coarser grid).
          static_assert(L::LEVELS == L::LW + 3, "levels below the 15x15 level");
          constexpr int B   = L::LW, npB = 17;
          double       *VB  = sV + (size_t)NRHS * L::lvl_off(B);
          double       *V1  = sV + (size_t)NRHS * L::lvl_off(B + 1); // 9x9 nodes
          double       *V2  = sV + (size_t)NRHS * L::lvl_off(B + 2); // 5x5 nodes
          double       *V3  = sV + (size_t)NRHS * L::lvl_off(B + 3); // 3x3 nodes
          const DiT *DB  = sDi + L::lvl_off(B), *D1 = sDi + L::lvl_off(B + 1);
          const DiT *D2  = sDi + L::lvl_off(B + 2), *D3 = sDi + L::lvl_off(B + 3);
          if constexpr (EXACT7)
            {
              // r_7 = P^T r_15 by 49 threads, then t_7 = A_7^-1 r_7 with every row of the dense
              // inverse split over X7::PARTS neighbouring lanes (row = tid / PARTS); the phases are
              // fenced by a named barrier over the X7::WARPS participating warps only
              using X = Exact7<THREADS>;
              static_assert(32 * X::WARPS <= THREADS, "not enough threads for the exact coarse solve");
              if (warp < X::WARPS)
                {
                  if (tid < 49)
                    {
                      const int cx = 1 + tid % 7, cy = 1 + tid / 7;
                      double    o[NRHS];
                      if constexpr (NRHS == 2)
                        full_weighting_rot<NRHS>(VB, npB, cx, cy, ((cx - 1) >> 2) & 1, o);
                      else
                        {
                          double row[3][NRHS];
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
#pragma unroll
                          for (int k = 0; k < NRHS; ++k)
                            o[k] = fma(0.5, row[0][k] + row[2][k], row[1][k]);
                        }
                      stv<NRHS>(V1, cy * 9 + cx, o);
                    }
                  named_barrier<1, 32 * X::WARPS>();
                  const int r7 = tid / X::PARTS, q7 = tid % X::PARTS;
                  double    acc0[NRHS], acc1[NRHS];
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    acc0[k] = acc1[k] = 0.0;
#pragma unroll
                  for (int c = 0; c < X::NCHK; ++c)
                    {
                      double g[8];
                      gi(c, g); // entries q7 * CH + 8c .. +7 of row r7 (zero where there is none)
#pragma unroll
                      for (int jj = 0; jj < 8; ++jj)
                        {
                          if (8 * c + jj >= X::CH)
                            continue;
                          // (entries beyond the 49th belong to the padding lanes: their matrix entries are zero
                          //  and they read the zero halo row 8 of the 9 x 9 level -- clamping them to entry 48 put
                          //  a second address into the banks of five of the seven loads, profiles/r02e_*)
                          const int j = q7 * X::CH + 8 * c + jj;
                          double    u[NRHS];
                          ldv<NRHS>(V1, (1 + j / 7) * 9 + 1 + j % 7, u);
#pragma unroll
                          for (int k = 0; k < NRHS; ++k)
                            {
                              if (jj & 1)
                                acc1[k] = fma(g[jj], u[k], acc1[k]);
                              else
                                acc0[k] = fma(g[jj], u[k], acc0[k]);
                            }
                        }
                    }
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    {
                      acc0[k] += acc1[k];
#pragma unroll
                      for (int off = 1; off < X::PARTS; off <<= 1)
                        acc0[k] += __shfl_xor_sync(0xffffffffu, acc0[k], off);
                    }
                  named_barrier<1, 32 * X::WARPS>();
                  if (r7 < 49 && q7 == 0)
                    stv<NRHS>(V1, (1 + r7 / 7) * 9 + 1 + r7 % 7, acc0);
                }
            }
          else
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
          mark(6);
          // z_B = r_B / D_B + interpolants of t_{B+1}, t_{B+2}, t_{B+3} at the level-B nodes
          for (int t = tid; t < 256; t += THREADS)
            {
              const int fx = 1 + (t & 15), fy = 1 + (t >> 4);
              if (fx > 15 || fy > 15)
                continue;
              const int i = fy * npB + fx;
              double    v[NRHS];
              ldv<NRHS>(VB, i, v);
              [[maybe_unused]] double v_in[NRHS]; // n = 32: level B IS level 1, the u_1 . z_1 sum is formed here
              if constexpr (B == 1)
                {
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    v_in[k] = v[k];
                }
              const double di = DB[i];
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                v[k] *= di;
#pragma unroll
              for (int m = 1; m <= (EXACT7 ? 1 : 3); ++m)
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
              if constexpr (B == 1)
                {
                  if (dot1)
                    {
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        dot1[k] = fma(v_in[k], v[k], dot1[k]);
                    }
                }
            }
          __syncthreads();
          mark(7);
          // up: the wide levels above B
          for_levels_down<L::LW - 1, 1>([&](auto lc) {
            prolong_level(lc, tid, THREADS);
            if (!(decltype(lc)::value == 1 && dot1))
              __syncthreads();
          });
        }
      else
        {
          // small local meshes (n <= 16): every coarse level on warp 0, serially
          if (warp == 0)
            {
              for_levels<1, L::LEVELS>([&](auto lc) {
                restrict_level(lc, lane, 32);
                __syncwarp();
              });
              for_levels_down<L::LEVELS, 1>([&](auto lc) {
                prolong_level(lc, lane, 32);
                __syncwarp();
              });
            }
          __syncthreads();
        }
    }
  } // namespace bpx
} // namespace msb
