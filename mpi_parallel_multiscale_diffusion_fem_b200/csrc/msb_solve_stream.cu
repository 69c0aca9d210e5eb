// msb_solve_stream.cu -- HBM/L2-streamed tier: batched Jacobi-PCG for local meshes that do
// not fit the shared-memory tier (n >= 128; BASELINE cfg1 and cfg5).
//
// Same mathematics as the reference's per-basis sequence (diffusion_problem_basis.tpp:
// 450-465 + solve_iterative :293-317) on the condensed system: unknowns are the interior
// DoFs, the boundary DoFs carry the BasisQ1 data g (distribute(), :308) and never move.
// All (cell, basis) solves of the shard advance together, three kernels per iteration:
//     K1  p = D^-1 r + beta p                (beta from the previous iteration's dots)
//     K2  q = K p  (matrix-free 9-point stencil, coefficients shared by the 4 bases),  p.q
//     K3  x += alpha p ; r -= alpha q ; r.D^-1 r and r.r
// Dot products are two-level and deterministic: every CTA writes one partial per basis,
// every consumer CTA re-adds the partials of its cell in a fixed order, so all CTAs of a
// cell take bitwise identical decisions and no atomics are needed.  The stopping rule is
// the reference's: ||r||_2 <= tol every iteration (basis.tpp:297); a converged solve is
// frozen (its CTAs exit at once) and its iteration count recorded on the device.
#include <limits.h>
#include <math.h>

#include "msb_internal.cuh"

namespace msb
{
  constexpr int STREAM_THREADS = 256;
  constexpr int STREAM_MAXBLK  = 32; // max CTAs per coarse cell

  struct StreamParams
  {
    int           n, nblk, rows; // rows of nodes per CTA
    const double *corners, *q1coef, *sten;
    double       *x;             // phi buffer [C][4][N]
    double       *r, *p, *q;     // [C][4][N]
    double       *part;          // [C][4][2 parity][3][STREAM_MAXBLK]: 0 rz, 1 pq, 2 rr;
                                 // iteration `it` reads parity (it-1)&1 and writes it&1, so the
                                 // CTAs of one cell never read partials a sibling is rewriting
    double       *rzprev;        // [C][4]
    int32_t      *iters;         // [C][4]  (-1 while running)
    double       *res;           // [C][4]
    int32_t      *remaining;     // cells with an unfinished solve
    double        tol2;
    int           it;            // iteration about to be executed (1-based)
  };

  constexpr int PART_STRIDE = 2 * 3 * STREAM_MAXBLK; // doubles per solve

  __device__ __forceinline__ double *
  part_ptr(double *part, int sidx, int parity, int which)
  {
    return part + (size_t)sidx * PART_STRIDE + (parity * 3 + which) * STREAM_MAXBLK;
  }

  __device__ __forceinline__ double
  sum_part(const double *part, int nblk)
  {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b)
      s += part[b];
    return s;
  }

  // block-wide deterministic sums of NV values -> written by thread 0
  template <int NV>
  __device__ __forceinline__ void
  block_sum_to(double (&v)[NV], double *sbuf /*[8][NV]*/)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0)
          sbuf[warp * NV + k] = v[k];
      }
    __syncthreads();
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int k = 0; k < NV; ++k)
          {
            double s = 0.0;
            for (int w = 0; w < STREAM_THREADS / 32; ++w)
              s += sbuf[w * NV + k];
            v[k] = s;
          }
      }
  }

  // done(solve) as seen by every CTA of a cell during iteration `it`: either recorded in an
  // earlier launch (iters >= 0) or implied by the r.r partials of iteration it-1.  A CTA
  // that races with the recording CTA re-derives the same answer from the partials.
  __device__ __forceinline__ int
  solve_done(const StreamParams &P, int sidx, int parity, double *rr_out)
  {
    const double rr = sum_part(part_ptr(P.part, sidx, parity, 2), P.nblk);
    if (rr_out)
      *rr_out = rr;
    return (P.iters[sidx] >= 0) || (rr <= P.tol2);
  }

  // r = b = -K_IB g_B on interior rows, 0 on constrained rows (condense);
  // x = g on the boundary, 0 inside; p = 0; partial r.z and r.r into parity 0
  __global__ void __launch_bounds__(STREAM_THREADS)
  stream_init_kernel(StreamParams P)
  {
    const int     n = P.n, np = n + 1, N = np * np, cell = blockIdx.y, blk = blockIdx.x;
    const double *S = P.sten + (size_t)cell * ST_NARR * N;
    const double *c = P.corners + 8 * (size_t)cell, *q1 = P.q1coef + 16 * (size_t)cell;
    const int     y0 = blk * P.rows, y1 = min(np, y0 + P.rows);
    double        acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] = 0.0;
    for (int t = y0 * np + threadIdx.x; t < y1 * np; t += STREAM_THREADS)
      {
        const int  jx = t % np, jy = t / np;
        const bool bd = jx == 0 || jy == 0 || jx == n || jy == n;
        double     rv[4] = {0, 0, 0, 0}, xv[4] = {0, 0, 0, 0};
        if (bd)
          {
            double px, py;
            fine_vertex(c, n, jx, jy, px, py);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              xv[k] = basis_q1_value(q1, k, px, py);
          }
        else if (jx == 1 || jy == 1 || jx == n - 1 || jy == n - 1)
          {
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx)
                {
                  const int bx = jx + dx, by = jy + dy;
                  if ((dx == 0 && dy == 0) || !(bx == 0 || bx == n || by == 0 || by == n))
                    continue;
                  double kij;
                  if (dy == 0)
                    kij = S[ST_KE * N + (dx > 0 ? t : t - 1)];
                  else if (dx == 0)
                    kij = S[ST_KN * N + (dy > 0 ? t : t - np)];
                  else if (dx == dy)
                    kij = S[ST_KD1 * N + (dx > 0 ? t : t - np - 1)];
                  else
                    kij = S[ST_KD2 * N + (dy > 0 ? t - 1 : t - np)];
                  double px, py;
                  fine_vertex(c, n, bx, by, px, py);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    rv[k] -= kij * basis_q1_value(q1, k, px, py);
                }
            const double dinv = 1.0 / S[ST_KC * N + t];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              {
                acc[k] += rv[k] * rv[k] * dinv;
                acc[4 + k] += rv[k] * rv[k];
              }
          }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            const size_t o = ((size_t)cell * 4 + k) * N + t;
            P.x[o]         = xv[k];
            P.r[o]         = rv[k];
            P.p[o]         = 0.0;
          }
      }
    __shared__ double sbuf[8 * 8];
    block_sum_to<8>(acc, sbuf);
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            part_ptr(P.part, cell * 4 + k, 0, 0)[blk] = acc[k];
            part_ptr(P.part, cell * 4 + k, 0, 2)[blk] = acc[4 + k];
          }
      }
  }

  // K1: convergence bookkeeping of the previous iteration, then p = z + beta p
  __global__ void __launch_bounds__(STREAM_THREADS)
  stream_k1_kernel(StreamParams P)
  {
    const int n = P.n, np = n + 1, N = np * np, cell = blockIdx.y, blk = blockIdx.x;
    const int par = (P.it - 1) & 1;
    __shared__ double sbeta[4];
    __shared__ int    sdone[4];
    if (threadIdx.x < 4)
      {
        const int    k = threadIdx.x, sidx = cell * 4 + k;
        double       rr;
        const int    dn = solve_done(P, sidx, par, &rr);
        const double rz = sum_part(part_ptr(P.part, sidx, par, 0), P.nblk);
        sdone[k]        = dn;
        sbeta[k]        = P.it == 1 ? 0.0 : rz / P.rzprev[sidx];
        if (blk == 0 && dn && P.iters[sidx] < 0)
          {
            P.iters[sidx] = P.it - 1;
            P.res[sidx]   = sqrt(rr);
          }
      }
    __syncthreads();
    if (sdone[0] && sdone[1] && sdone[2] && sdone[3])
      return;
    const double *KC = P.sten + (size_t)cell * ST_NARR * N + ST_KC * N;
    const int     y0 = blk * P.rows, y1 = min(np, y0 + P.rows);
    for (int t = y0 * np + threadIdx.x; t < y1 * np; t += STREAM_THREADS)
      {
        const int jx = t % np, jy = t / np;
        if (jx == 0 || jy == 0 || jx == n || jy == n)
          continue;
        const double dinv = 1.0 / KC[t];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            if (sdone[k])
              continue;
            const size_t o = ((size_t)cell * 4 + k) * N + t;
            P.p[o]         = fma(sbeta[k], P.p[o], P.r[o] * dinv);
          }
      }
  }

  // K2: q = K p on interior rows (p is zero on the boundary), partial p.q
  __global__ void __launch_bounds__(STREAM_THREADS)
  stream_k2_kernel(StreamParams P)
  {
    const int n = P.n, np = n + 1, N = np * np, cell = blockIdx.y, blk = blockIdx.x;
    const int par = (P.it - 1) & 1;
    __shared__ int    sdone[4];
    __shared__ double sbuf[8 * 4];
    if (threadIdx.x < 4)
      sdone[threadIdx.x] = solve_done(P, cell * 4 + threadIdx.x, par, nullptr);
    __syncthreads();
    if (sdone[0] && sdone[1] && sdone[2] && sdone[3])
      return;
    const double *S  = P.sten + (size_t)cell * ST_NARR * N;
    const int     y0 = blk * P.rows, y1 = min(np, y0 + P.rows);
    double        acc[4] = {0, 0, 0, 0};
    for (int t = y0 * np + threadIdx.x; t < y1 * np; t += STREAM_THREADS)
      {
        const int jx = t % np, jy = t / np;
        if (jx == 0 || jy == 0 || jx == n || jy == n)
          continue;
        const double kc = S[ST_KC * N + t], kE = S[ST_KE * N + t], kW = S[ST_KE * N + t - 1];
        const double kN = S[ST_KN * N + t], kS = S[ST_KN * N + t - np];
        const double kNE = S[ST_KD1 * N + t], kSW = S[ST_KD1 * N + t - np - 1];
        const double kNW = S[ST_KD2 * N + t - 1], kSE = S[ST_KD2 * N + t - np];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            if (sdone[k])
              continue;
            const double *p  = P.p + ((size_t)cell * 4 + k) * N + t;
            const double  pc = p[0];
            double        y  = kc * pc;
            y = fma(kE, p[1], y);
            y = fma(kW, p[-1], y);
            y = fma(kN, p[np], y);
            y = fma(kS, p[-np], y);
            y = fma(kNE, p[np + 1], y);
            y = fma(kSW, p[-np - 1], y);
            y = fma(kNW, p[np - 1], y);
            y = fma(kSE, p[-np + 1], y);
            P.q[((size_t)cell * 4 + k) * N + t] = y;
            acc[k] = fma(pc, y, acc[k]);
          }
      }
    block_sum_to<4>(acc, sbuf);
    if (threadIdx.x == 0)
      {
        // p.q is produced and consumed inside one iteration: parity slot of this iteration
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!sdone[k])
            part_ptr(P.part, cell * 4 + k, P.it & 1, 1)[blk] = acc[k];
      }
  }

  // K3: alpha = rz/pq ; x += alpha p ; r -= alpha q ; partial r.z and r.r
  __global__ void __launch_bounds__(STREAM_THREADS)
  stream_k3_kernel(StreamParams P)
  {
    const int n = P.n, np = n + 1, N = np * np, cell = blockIdx.y, blk = blockIdx.x;
    const int par = (P.it - 1) & 1;
    __shared__ int    sdone[4];
    __shared__ double salpha[4];
    __shared__ double sbuf[8 * 8];
    if (threadIdx.x < 4)
      {
        const int sidx      = cell * 4 + threadIdx.x;
        const int dn        = solve_done(P, sidx, par, nullptr);
        sdone[threadIdx.x]  = dn;
        const double rz     = sum_part(part_ptr(P.part, sidx, par, 0), P.nblk);
        const double pq     = sum_part(part_ptr(P.part, sidx, P.it & 1, 1), P.nblk);
        salpha[threadIdx.x] = dn ? 0.0 : rz / pq;
      }
    __syncthreads();
    if (sdone[0] && sdone[1] && sdone[2] && sdone[3])
      return;
    const double *KC = P.sten + (size_t)cell * ST_NARR * N + ST_KC * N;
    const int     y0 = blk * P.rows, y1 = min(np, y0 + P.rows);
    double        acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] = 0.0;
    for (int t = y0 * np + threadIdx.x; t < y1 * np; t += STREAM_THREADS)
      {
        const int jx = t % np, jy = t / np;
        if (jx == 0 || jy == 0 || jx == n || jy == n)
          continue;
        const double dinv = 1.0 / KC[t];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            if (sdone[k])
              continue;
            const size_t o  = ((size_t)cell * 4 + k) * N + t;
            const double a  = salpha[k];
            P.x[o]          = fma(a, P.p[o], P.x[o]);
            const double rn = fma(-a, P.q[o], P.r[o]);
            P.r[o]          = rn;
            acc[k]          = fma(rn * dinv, rn, acc[k]);
            acc[4 + k]      = fma(rn, rn, acc[4 + k]);
          }
      }
    block_sum_to<8>(acc, sbuf);
    if (threadIdx.x == 0)
      {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!sdone[k])
            {
              const int sidx = cell * 4 + k;
              part_ptr(P.part, sidx, P.it & 1, 0)[blk] = acc[k];
              part_ptr(P.part, sidx, P.it & 1, 2)[blk] = acc[4 + k];
              // r.z of the iteration just consumed becomes "previous" for the next K1
              // (every CTA of the cell writes the same value; nobody reads it in this launch)
              if (blk == 0)
                P.rzprev[sidx] = sum_part(part_ptr(P.part, sidx, par, 0), P.nblk);
            }
      }
  }

  // after the loop (P.it = last executed iteration): record every solve not yet recorded
  __global__ void
  stream_finalize_kernel(StreamParams P, int n_solves, int32_t *fail)
  {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= n_solves || P.iters[sidx] >= 0)
      return;
    const double rr = sum_part(part_ptr(P.part, sidx, P.it & 1, 2), P.nblk);
    P.iters[sidx]   = P.it;
    P.res[sidx]     = sqrt(rr);
    if (!(rr <= P.tol2))
      atomicMin(fail, sidx);
  }

  // number of solves still running after iteration P.it
  __global__ void
  stream_count_kernel(StreamParams P, int n_solves, int32_t *remaining)
  {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= n_solves || P.iters[sidx] >= 0)
      return;
    if (!(sum_part(part_ptr(P.part, sidx, P.it & 1, 2), P.nblk) <= P.tol2))
      atomicAdd(remaining, 1);
  }

  size_t
  streamed_workspace_doubles(const Shard &s)
  {
    return 3 * (size_t)s.n_cells * 4 * s.N;
  }

  static StreamParams
  shifted(const StreamParams &P, const Shard &s, int c0)
  {
    // all per-cell pointers are indexed by blockIdx.y inside the kernels
    StreamParams Q = P;
    Q.corners += 8 * (size_t)c0, Q.q1coef += 16 * (size_t)c0, Q.sten += (size_t)c0 * ST_NARR * s.N;
    Q.x += (size_t)c0 * 4 * s.N, Q.r += (size_t)c0 * 4 * s.N, Q.p += (size_t)c0 * 4 * s.N;
    Q.q += (size_t)c0 * 4 * s.N, Q.part += (size_t)c0 * 4 * PART_STRIDE;
    Q.rzprev += 4 * (size_t)c0, Q.iters += 4 * (size_t)c0, Q.res += 4 * (size_t)c0;
    return Q;
  }

  cudaError_t
  launch_solve_streamed(Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches)
  {
    StreamParams P;
    P.n = s.n;
    // 16 node rows per CTA (9 CTAs per cell at n=128); never more than STREAM_MAXBLK CTAs
    int rows = 16;
    while ((s.np + rows - 1) / rows > STREAM_MAXBLK)
      rows *= 2;
    P.rows      = rows;
    P.nblk      = (s.np + rows - 1) / rows;
    P.corners   = s.d_corners;
    P.q1coef    = s.d_q1coef;
    P.sten      = s.d_sten;
    P.x         = s.d_phi;
    P.r         = s.d_wr;
    P.p         = s.d_wp;
    P.q         = s.d_wq;
    P.part      = s.d_part;
    P.rzprev    = s.d_scal;
    P.iters     = s.d_iters;
    P.res       = s.d_res;
    P.remaining = s.d_flags;
    P.tol2      = tol * tol;
    P.it        = 0;
    const int   n_solves = 4 * s.n_cells;
    cudaError_t e;
    if ((e = cudaMemsetAsync(s.d_iters, 0xff, sizeof(int32_t) * n_solves, st)) != cudaSuccess)
      return e;
    if ((e = cudaMemsetAsync(s.d_part, 0, sizeof(double) * (size_t)n_solves * PART_STRIDE, st)) !=
        cudaSuccess)
      return e;
    for (int c0 = 0; c0 < s.n_cells; c0 += 65535)
      {
        const int nc = s.n_cells - c0 < 65535 ? s.n_cells - c0 : 65535;
        stream_init_kernel<<<dim3(P.nblk, nc), STREAM_THREADS, 0, st>>>(shifted(P, s, c0));
        ++*n_launches;
      }
    int32_t   h_remaining = 1;
    int       it          = 0;
    const int check_every = 8;
    while (it < max_iter)
      {
        // host poll: how many solves are still running after iteration `it`?
        if (it % check_every == 0)
          {
            P.it = it;
            if ((e = cudaMemsetAsync(s.d_flags, 0, sizeof(int32_t), st)) != cudaSuccess)
              return e;
            stream_count_kernel<<<(n_solves + 255) / 256, 256, 0, st>>>(P, n_solves, s.d_flags);
            ++*n_launches;
            if ((e = cudaMemcpyAsync(&h_remaining, s.d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                     st)) != cudaSuccess)
              return e;
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess)
              return e;
            if (h_remaining == 0)
              break;
          }
        ++it;
        P.it = it;
        for (int c0 = 0; c0 < s.n_cells; c0 += 65535)
          {
            const int          nc = s.n_cells - c0 < 65535 ? s.n_cells - c0 : 65535;
            const StreamParams Q  = shifted(P, s, c0);
            const dim3         g(P.nblk, nc);
            stream_k1_kernel<<<g, STREAM_THREADS, 0, st>>>(Q);
            stream_k2_kernel<<<g, STREAM_THREADS, 0, st>>>(Q);
            stream_k3_kernel<<<g, STREAM_THREADS, 0, st>>>(Q);
            *n_launches += 3;
          }
      }
    P.it = it;
    stream_finalize_kernel<<<(n_solves + 255) / 256, 256, 0, st>>>(P, n_solves, s.d_fail);
    ++*n_launches;
    return cudaGetLastError();
  }
} // namespace msb
