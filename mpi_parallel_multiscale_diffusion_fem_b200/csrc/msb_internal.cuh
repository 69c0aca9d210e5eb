// msb_internal.cuh -- shared declarations of the sm_100a multiscale-basis library.
//
// Data layout in HBM (one shard = the coarse cells owned by one GPU), all FP64,
// node index "lex" = jy*(n+1)+jx on the implicit structured fine grid:
//   corners  [C][4][2]      coarse-cell vertices, deal.II vertex order
//   q1coef   [C][16]        BasisQ1 coefficient matrix (basis_q1.tpp:26-47)
//   sten     [C][6][np*np]  symmetric 9-point stencil of the unconstrained fine
//                           stiffness matrix + load vector:
//                           0 KC diag, 1 KE (jx,jy)-(jx+1,jy), 2 KN (jx,jy)-(jx,jy+1),
//                           3 KD1 (jx,jy)-(jx+1,jy+1), 4 KD2 (jx+1,jy)-(jx,jy+1), 5 F
//   phi      [C][4][np*np]  the multiscale bases, lexicographic node order
//   M [C][16], b [C][4], iters [C][4], res [C][4]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "msfem_basis.h"

namespace msb
{
  enum
  {
    ST_KC  = 0,
    ST_KE  = 1,
    ST_KN  = 2,
    ST_KD1 = 3,
    ST_KD2 = 4,
    ST_F   = 5,
    ST_NARR = 6
  };
  // dim = 3 (msb_dim3.cu): diagonal, 13 forward couplings, load vector
  enum
  {
    ST3_F    = 14,
    ST3_NARR = 15
  };

  struct Shard
  {
    int     dim = 2, nb = 4, nst = ST_NARR; // bases per cell 2^dim, stencil arrays per cell
    int     l, n, np, N; // n = 2^l, np = n+1, N = np^dim
    int     n_cells;
    int     device;
    int     tier, variant;
    msb_coeff_desc coeff;
    double  rhs_value;

    double       *d_corners = nullptr;
    double       *d_q1coef  = nullptr;
    double       *d_table   = nullptr;
    double       *d_sten    = nullptr;
    double       *d_phi     = nullptr;
    double       *d_M       = nullptr;
    double       *d_b       = nullptr;
    int32_t      *d_iters   = nullptr;
    double       *d_res     = nullptr;
    int32_t      *d_fail    = nullptr; // [2]: first failing solve index (cell*4+ib) or INT_MAX, flag
    uint32_t     *d_dofmap  = nullptr; // lex -> deal.II dof
    uint32_t     *d_invmap  = nullptr; // deal.II dof -> lex
    double       *d_gsol    = nullptr; // [C][N] global solution (after set_global_weights)
    double       *d_tmp     = nullptr; // 2*N scratch for single-vector calls
    double       *d_w       = nullptr; // [C][nb] coarse weights (set_global_weights)
    // bulk accessors (msb_get_bases / msb_get_global_solutions): two staging buffers for the
    // reordered vectors, each on its own stream so that the device->host copy of one chunk
    // overlaps the reordering launch of the next
    double       *d_stage[2]   = {nullptr, nullptr};
    size_t        stage_vecs   = 0;    // vectors of N doubles per staging buffer
    cudaStream_t  stage_stream[2] = {nullptr, nullptr};
    // streamed tier work vectors [C][4][N] each
    double       *d_wr = nullptr, *d_wp = nullptr, *d_wq = nullptr, *d_wz = nullptr;
    double       *d_wr2 = nullptr;    // 2D streamed tier: second residual buffer (fused iteration kernels)
    double       *d_wv = nullptr;     // streamed tier coarse-level vectors [C][4][cn]
    double       *d_dinv = nullptr;   // streamed tier reciprocal Galerkin diagonals [C][cn]
    double       *d_gal = nullptr;    // streamed tier Galerkin scratch (two coarse stencil buffers)
    double       *d_scal = nullptr;   // streamed tier per-solve scalars
    double       *d_part = nullptr;   // streamed tier partial sums
    int32_t      *d_flags = nullptr;  // streamed tier per-solve state

    cudaStream_t stream = nullptr;    // library-owned stream
    cudaStream_t run_stream = nullptr;
    cudaEvent_t  ev[4]  = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t  ev_chunk[2] = {nullptr, nullptr}; // msb_run_with_bases: "chunk solved" per staging buffer
    bool         assembled = false, ran = false, weights_set = false, run_pending = false;
    bool         valid = true;        // false after a failed msb_set_cells: only set_cells / destroy work
    bool         bricks = false;      // dim 3: every coarse cell is an axis-aligned brick
    bool         aligned = false;     // dim 2: every coarse cell is an axis-aligned rectangle
    bool         fused_last = false;  // the last run used the fused one-kernel stage
    int          n_launches = 0;
    int          tier_used  = 0;
    int          last_status = 0;
  };

  // ---- launchers implemented in the .cu files -------------------------------------------
  cudaError_t launch_dofmap(const Shard &s, cudaStream_t st);
  cudaError_t launch_basis_q1(const Shard &s, cudaStream_t st, int32_t *h_bad);
  cudaError_t launch_assemble(const Shard &s, cudaStream_t st, int *n_launches);
  cudaError_t launch_solve_smem(const Shard &s, double tol, int max_iter, cudaStream_t st,
                                int *n_launches);
  cudaError_t launch_solve_bpx(const Shard &s, double tol, int max_iter, cudaStream_t st,
                               int *n_launches);
  cudaError_t launch_solve_streamed(Shard &s, double tol, int max_iter, cudaStream_t st,
                                    int *n_launches);
  // cluster / DSMEM tier (msb_solve_cluster.cu); called by launch_solve_streamed after its setup
  cudaError_t launch_solve_cluster(const Shard &s, double tol, int max_iter, bool tmem, cudaStream_t st,
                                   int *n_launches);
  bool        cluster_tier_supported(int l);
  bool        streamed_tier_uses_cluster(int l, int variant);
  cudaError_t launch_element_matrices(const Shard &s, cudaStream_t st, int *n_launches);
  cudaError_t launch_apply_operator(const Shard &s, int cell, const double *d_x_lex, double *d_y_lex,
                                    cudaStream_t st);
  cudaError_t launch_permute(const Shard &s, const double *d_src, double *d_dst, bool lex_to_dof,
                             cudaStream_t st);
  // n_vec consecutive vectors of N doubles, lexicographic -> deal.II DoF order, one launch
  cudaError_t launch_permute_batch(const Shard &s, const double *d_src, double *d_dst, size_t n_vec,
                                   cudaStream_t st);
  cudaError_t launch_global_solution(const Shard &s, const double *d_w, cudaStream_t st);
  cudaError_t launch_constraints(const Shard &s, int cell, int ib, uint32_t *d_dofs, double *d_vals,
                                 cudaStream_t st);
  // dim = 3 (msb_dim3.cu)
  cudaError_t launch_dofmap3(const Shard &s, cudaStream_t st);
  cudaError_t launch_assemble3(const Shard &s, cudaStream_t st, int *n_launches);
  cudaError_t launch_solve3(Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches);
  cudaError_t launch_element_matrices3(const Shard &s, cudaStream_t st, int *n_launches);
  cudaError_t launch_apply_operator3(const Shard &s, int cell, const double *d_x_lex, double *d_y_lex,
                                     cudaStream_t st);
  cudaError_t launch_constraints3(const Shard &s, int cell, int ib, uint32_t *d_dofs, double *d_vals,
                                  cudaStream_t st);
  size_t      dim3_coarse_nodes(int l);
  int         dim3_part_stride();
  bool        smem_tier_supported(int l);
  // the fused one-kernel stage (msb_solve_fused.cu): 64 x 64 local meshes, axis-aligned cells,
  // analytic coefficient, default variant
  cudaError_t launch_stage_fused(const Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches);
  // the same for the cells [c0, c0 + nc) of the shard only (pipelined msb_run_with_bases)
  cudaError_t launch_stage_fused_range(const Shard &s, int c0, int nc, double tol, int max_iter, cudaStream_t st,
                                       int *n_launches, bool split = false);
  size_t      streamed_coarse_nodes(int l);
  size_t      streamed_galerkin_scratch_doubles(int l, int n_cells);

  // ---- small device helpers -----------------------------------------------------------------
  __host__ __device__ inline uint32_t
  morton_compact(uint32_t m)
  {
    uint32_t x = m & 0x55555555u;
    x          = (x | (x >> 1)) & 0x33333333u;
    x          = (x | (x >> 2)) & 0x0f0f0f0fu;
    x          = (x | (x >> 4)) & 0x00ff00ffu;
    x          = (x | (x >> 8)) & 0x0000ffffu;
    return x;
  }

  __host__ __device__ inline uint32_t
  morton_spread(uint32_t x)
  {
    x &= 0x0000ffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
  }

  // Morton index of fine cell (ix,iy), x the low bit (deal.II child order)
  __host__ __device__ inline uint32_t
  morton_encode(uint32_t ix, uint32_t iy)
  {
    return morton_spread(ix) | (morton_spread(iy) << 1);
  }

  // fine vertex (jx,jy) of the refined general_cell (basis.tpp:94-98): bilinear image of
  // the uniform grid, exact for axis-aligned dyadic cells
  __device__ inline void
  fine_vertex(const double *__restrict__ c, int n, int jx, int jy, double &px, double &py)
  {
    // n is a power of two: multiplying by 1/n is exact and avoids two FP64 divisions
    const double rn = 1.0 / (double)n, s = (double)jx * rn, t = (double)jy * rn;
    const double st = s * t;
    px = c[0] + s * (c[2] - c[0]) + t * (c[4] - c[0]) + st * ((c[6] - c[4]) - (c[2] - c[0]));
    py = c[1] + s * (c[3] - c[1]) + t * (c[5] - c[1]) + st * ((c[7] - c[5]) - (c[3] - c[1]));
  }

  // BasisQ1<2>::value (basis_q1.tpp:86-96), coef[r*4+ib]
  __device__ inline double
  basis_q1_value(const double *__restrict__ coef, int ib, double x, double y)
  {
    return coef[0 * 4 + ib] + coef[1 * 4 + ib] * x + coef[2 * 4 + ib] * y + coef[3 * 4 + ib] * x * y;
  }
} // namespace msb
