/*
 * msfem_basis.h -- C ABI of the B200-native multiscale-basis stage.
 *
 * One handle owns a SHARD of coarse cells on one GPU and replaces, for those
 * cells, the reference's serial hot loop
 *
 *     for (it_basis ...) (it_basis->second).run();
 *                    /root/reference/include/base/diffusion_problem_ms.tpp:81-87
 *
 * i.e. DiffusionProblemBasis<dim>::run() for every locally owned coarse cell
 * (/root/reference/include/base/diffusion_problem_basis.tpp:438-474), and the
 * accessors the caller uses afterwards (diffusion_problem_basis.hpp:72-138).
 *
 * Conventions
 *  - plain C, no torch / CUDA types in any signature; a CUDA stream crosses the
 *    boundary as an opaque void* (cudaStream_t), NULL = the library's own stream;
 *  - the caller owns every host buffer, the library owns all device memory
 *    behind the opaque handle;
 *  - every call returns MSB_OK (0) or a negative msb_status; nothing throws
 *    across the boundary (the reference throws SolverControl::NoConvergence out
 *    of basis.tpp:303; here msb_run returns MSB_ERR_NO_CONVERGENCE and
 *    msb_get_failure names the first failing (cell, basis));
 *  - all calls are synchronous on return except msb_run_async;
 *  - handles are independent: one handle per device / host thread.
 *  - there is NO CPU fallback: without a CUDA device msb_create fails with
 *    MSB_ERR_NO_DEVICE.
 *
 * Index conventions follow deal.II: coarse-cell vertices
 * v0=(x0,y0) v1=(x1,y0) v2=(x0,y1) v3=(x1,y1); fine DoFs in the first-touch
 * numbering of the Morton-ordered refined cell; FullMatrix row-major.
 */
#ifndef MSFEM_BASIS_H
#define MSFEM_BASIS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB_ABI_VERSION 1

typedef enum
{
  MSB_OK                 = 0,
  MSB_ERR_INVALID_ARG    = -1,
  MSB_ERR_UNSUPPORTED    = -2, /* dim/tier/coefficient/n_refine_local not built   */
  MSB_ERR_NO_DEVICE      = -3, /* no CUDA device / wrong architecture            */
  MSB_ERR_CUDA           = -4, /* a CUDA call failed, see msb_last_error()       */
  MSB_ERR_NO_CONVERGENCE = -5, /* some local solve hit max_iter (basis.tpp:297)  */
  MSB_ERR_STATE          = -6  /* call order: run before get, weights before...  */
} msb_status;

/* Diffusion coefficient kinds.  MSB_COEFF_REFERENCE restates
 * Coefficients::MatrixCoeff<2> verbatim (include/coefficients/matrix_coeff.tpp:
 * 17-25,66-91; constants matrix_coeff.hpp:45-48; PI_D typo coefficients.h:21).
 * PERIODIC / INCLUSIONS / CONSTANT are the BASELINE.md section-4 synthetic
 * coefficients.  TABLE takes the tensor values a user-defined
 * TensorFunction<2,dim>::value_list produced at the fine quadrature points, so
 * any host coefficient class can be used unchanged. */
typedef enum
{
  MSB_COEFF_REFERENCE  = 0,
  MSB_COEFF_PERIODIC   = 1, /* par[0]=eps, par[1]=scale (0.9999)                 */
  MSB_COEFF_INCLUSIONS = 2, /* par[0]=block size, par[1]=probability,
                               par[2]=a_inclusion, par[3]=a_background; seed     */
  MSB_COEFF_CONSTANT   = 3, /* par[0]=a0                                         */
  MSB_COEFF_TABLE      = 4  /* coeff_table of msb_create                         */
} msb_coeff_kind;

typedef enum
{
  MSB_TIER_AUTO     = 0, /* shared-memory resident when the local mesh fits      */
  MSB_TIER_SMEM     = 1, /* one CTA per (cell, right-hand-side group)            */
  MSB_TIER_STREAMED = 2  /* any local mesh size: vectors spread over the shared
                            memory of a thread-block cluster (128x128 fine cells) or
                            streamed through HBM/L2 (larger, and on request)       */
} msb_tier;

typedef struct
{
  int32_t kind; /* msb_coeff_kind */
  int32_t seed;
  double  par[6];
} msb_coeff_desc;

typedef struct
{
  int32_t        abi_version;    /* MSB_ABI_VERSION                               */
  int32_t        dim;            /* 2 or 3 (diffusion_problem_basis.inst.cc:15-16) */
  int32_t        n_refine_local; /* l: n = 2^l fine cells per direction, 1..9     */
  int32_t        n_cells;        /* coarse cells of this shard                    */
  int32_t        device_id;      /* CUDA device ordinal                           */
  int32_t        tier;           /* msb_tier                                      */
  int32_t        variant;        /* kernel variant selector, 0 = default          */
  int32_t        reserved;
  double         rhs_value;      /* RightHandSide<dim> f (right_hand_side.tpp:27) */
  msb_coeff_desc coeff;
} msb_config;

typedef struct msb_handle_s *msb_handle;

/* Replaces the construction loop ms.tpp:50-73 (DiffusionProblemBasis ctor,
 * basis.tpp:18-50).  corners: [n_cells][2^dim][dim] doubles, deal.II vertex
 * order.  dim 3 (hexahedral cells, 8 bases, 1 <= n_refine_local <= 6) takes
 * MSB_COEFF_REFERENCE (MatrixCoeff<3>, matrix_coeff.tpp:28-41), MSB_COEFF_CONSTANT or
 * MSB_COEFF_TABLE and always runs in the streamed tier.  coeff_table: NULL unless
 * coeff.kind == MSB_COEFF_TABLE, then what TensorFunction<2,dim>::value_list returned at the fine
 * quadrature points (basis.tpp:202-203):
 *   dim 2: [n_cells][n*n fine cells, iy*n+ix][4 q-points, x fastest][a00,a01,a10,a11]
 *   dim 3: [n_cells][n^3 fine cells, (iz*n+iy)*n+ix][8 q-points, x fastest][a00,a01,a02,a10,...,a22]
 * A tensor that is not symmetric up to 1e-12 max|a| is rejected (MSB_ERR_INVALID_ARG): the stiffness
 * matrix must be symmetric for CG -- the reference's SolverCG needs that too -- and the 2D kernels
 * assemble from the symmetric part only. */
int msb_create(const msb_config *cfg, const double *corners, const double *coeff_table,
               msb_handle *out);

/* Re-targets an existing handle at another batch of n_cells coarse cells (same count,
 * same n_refine_local and coefficient): the handle is a reusable device workspace, so a
 * caller can stream many batches through one allocation.  Copies corners (and the table)
 * host -> device; invalidates the previous results. */
int msb_set_cells(msb_handle h, const double *corners, const double *coeff_table);

/* Replaces the hot loop ms.tpp:81-87 (run(): make_grid, setup_system,
 * assemble_system, 2^dim x {condense, PCG, distribute},
 * assemble_global_element_matrix).  tol_abs / max_iter are SolverControl's
 * arguments (basis.tpp:297: 1000, 1e-12; the stopping rule is the reference's:
 * absolute l2 norm of the unpreconditioned residual of the condensed system,
 * tested every iteration).  Initial guess: the reference starts SolverCG from the
 * zero vector (basis.tpp:152 reinit, :304); the fused, cluster, streamed and dim-3
 * kernels start from the coarse Q1 shape function g itself (the Dirichlet data
 * extended into the cell: the exact solution for a constant coefficient), which
 * changes the iteration counts msb_get_iteration_counts reports (0 is possible),
 * not the converged bases. */
int msb_run(msb_handle h, double tol_abs, int32_t max_iter);

/* Same, enqueued on `cuda_stream`; pair with msb_sync.  The shared-memory and cluster
 * tiers (n_refine_local <= 7 in 2D) are one asynchronous launch sequence: the call returns
 * without synchronising.  The HBM-streamed kernels (2D n_refine_local >= 8, dim 3) test
 * convergence from the host every few iterations and therefore block the calling thread
 * until the solves have finished (the stream is still the caller's). */
int msb_run_async(msb_handle h, double tol_abs, int32_t max_iter, void *cuda_stream);
int msb_sync(msb_handle h);

/* msb_run followed by msb_get_bases of every cell, PIPELINED: the shard is processed in chunks of coarse
 * cells, and while chunk k+1 is being solved the bases of chunk k are reordered to the deal.II DoF order and
 * copied to bases_out ([n_cells][2^dim][N]; page-locked memory makes the copy overlap the solves).  For a
 * caller that keeps the 2^dim solution_vectors host-side like the reference does (basis.hpp:216; all of them are
 * walked in output_global_fine, ms.tpp:386-393) this hides the PCIe time of the bases (8.9 GB for the
 * 256x256 x 64x64 problem) behind the stage instead of paying it afterwards.  Synchronous on return; same
 * return codes as msb_run.  Shards that do not run the fused one-kernel stage take msb_run + msb_get_bases. */
int msb_run_with_bases(msb_handle h, double tol_abs, int32_t max_iter, double *bases_out);

/* First non-converged solve of the last run (cell = -1 if none). */
int msb_get_failure(msb_handle h, int32_t *cell, int32_t *index_basis, double *residual);

/* get_global_element_matrix / get_global_element_rhs (basis.tpp:320-333) for
 * all cells: M [n_cells][2^dim][2^dim] row-major, b [n_cells][2^dim]. */
int msb_get_element_matrices(msb_handle h, double *M, double *b);

/* solver_control.last_step() of every solve (basis.tpp:310-316): [n_cells][2^dim];
 * residuals: final ||r||_2 of every solve, may be NULL. */
int msb_get_iteration_counts(msb_handle h, int32_t *iters, double *residuals);

/* solution_vector[index_basis] of one cell (basis.hpp:216), N doubles in the
 * deal.II DoF order. */
int msb_get_basis(msb_handle h, int32_t cell, int32_t index_basis, double *out);

/* Bulk form: the 2^dim solution_vectors (basis.hpp:216) of the cells [cell0, cell0 + n_cells)
 * in ONE call, out [n_cells][2^dim][N] in the deal.II DoF order.  The reference keeps these
 * vectors host-side inside every basis object; a caller that walks all cells
 * (output_global_fine, ms.tpp:386-393) gets them with one reordering launch per staging
 * chunk and chunked device->host copies that overlap it (fully asynchronous when `out` is
 * page-locked), instead of one launch + one synchronisation per (cell, basis). */
int msb_get_bases(msb_handle h, int32_t cell0, int32_t n_cells, double *out);

/* DoF map of the local mesh (basis.tpp:106): dof_of_vertex[jy*(n+1)+jx]
 * (dim 3: [(jz*(n+1)+jy)*(n+1)+jx]). */
int msb_get_dof_map(msb_handle h, uint32_t *dof_of_vertex);

/* Constraint set `index_basis` of one cell (basis.tpp:119-135): the
 * (n+1)^dim - (n-1)^dim constrained DoFs (4n for dim 2) in ascending order and
 * their inhomogeneities. */
int msb_get_constraints(msb_handle h, int32_t cell, int32_t index_basis, uint32_t *dofs,
                        double *values);

/* y = K x with the unconstrained fine stiffness matrix of one cell
 * (diffusion_matrix.vmult, basis.tpp:271), x and y in deal.II DoF order; and
 * the load vector global_rhs (basis.tpp:239).  For operator-level parity. */
int msb_apply_operator(msb_handle h, int32_t cell, const double *x, double *y);
int msb_get_load_vector(msb_handle h, int32_t cell, double *F);

/* set_global_weights for all cells (basis.tpp:352-377; caller
 * send_global_weights_to_cell ms.tpp:299-325): w [n_cells][2^dim]. */
int msb_set_global_weights(msb_handle h, const double *w);

/* global_solution of one cell (basis.hpp, basis.tpp:421-435), deal.II order. */
int msb_get_global_solution(msb_handle h, int32_t cell, double *out);

/* Bulk form for the walk over all cells in output_global_fine (ms.tpp:386-393):
 * global_solution of the cells [cell0, cell0 + n_cells), out [n_cells][N], deal.II order. */
int msb_get_global_solutions(msb_handle h, int32_t cell0, int32_t n_cells, double *out);

/* Device addresses (as integers of the handle's device) of the stage results of the last
 * run: M [n_cells][2^dim][2^dim], b [n_cells][2^dim], iteration counts [n_cells][2^dim]
 * (int32).  For the exchange that follows the stage -- the reference's
 * compress(VectorOperation::add), ms.tpp:253-254 -- so that an NCCL collective can read the
 * contributions where they are instead of through a host round trip.  The memory stays
 * owned by the handle and valid until msb_set_cells / msb_run / msb_destroy.  Any of the
 * out-pointers may be NULL. */
int msb_get_device_results(msb_handle h, uint64_t *d_M, uint64_t *d_b, uint64_t *d_iters);

/* Device-side timing of the last run (CUDA events on the run's stream), ms;
 * and the number of kernels the run launched. */
int msb_get_run_stats(msb_handle h, float *ms_total, float *ms_solve_kernel, int32_t *n_launches,
                      int32_t *tier_used);

/* Algorithmic bytes of the last run per SURVEY 8(d): sum over solves of
 * N*(96*k+16). */
int msb_get_algorithmic_bytes(msb_handle h, double *bytes, double *mean_iterations);

int msb_destroy(msb_handle h);

/* Thread-local description of the last error. */
const char *msb_last_error(void);

/* Library / device identification.  msb_build_id: first 16 hex digits of the SHA-256 of the
 * kernel sources this binary was compiled from (csrc/Makefile), so that a profile captured
 * on one binary is never quoted beside a run of another (bench.py roofline.traffic). */
int msb_device_count(void);
const char *msb_version(void);
const char *msb_build_id(void);

#ifdef __cplusplus
}
#endif
#endif /* MSFEM_BASIS_H */
