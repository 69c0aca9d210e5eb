// msb_api.cu -- the C ABI declared in include/msfem_basis.h.
//
// One handle = one shard of coarse cells on one GPU.  The construction loop
// diffusion_problem_ms.tpp:50-73 becomes msb_create, the hot loop :81-87 becomes
// msb_run / msb_run_async, the accessors of diffusion_problem_basis.hpp:72-138 become the
// msb_get_* / msb_set_* calls.  There is no CPU fallback anywhere in this file.
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "msb_internal.cuh"

using namespace msb;

struct msb_handle_s
{
  Shard s;
};

static thread_local char g_err[512] = "";

static int
fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(call)                                                                        \
  do                                                                                          \
    {                                                                                         \
      cudaError_t e_ = (call);                                                                \
      if (e_ != cudaSuccess)                                                                  \
        return fail(MSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                    __FILE__, __LINE__);                                                      \
    }                                                                                         \
  while (0)

extern "C" const char *
msb_last_error(void)
{
  return g_err;
}

extern "C" const char *
msb_version(void)
{
  return "msfem_basis 0.1 (sm_100a)";
}

#ifndef MSB_BUILD_ID
#  define MSB_BUILD_ID "unknown"
#endif
extern "C" const char *
msb_build_id(void)
{
  return MSB_BUILD_ID;
}

extern "C" int
msb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
  return n;
}

// dim 3: are all coarse cells axis-aligned bricks (vertex v = v0 + (bit0 hx, bit1 hy, bit2 hz))?
static bool
all_bricks(const double *corners, size_t n_cells)
{
  for (size_t k = 0; k < n_cells; ++k)
    {
      const double *c = corners + 24 * k;
      for (int v = 0; v < 8; ++v)
        if (c[3 * v] != ((v & 1) ? c[3] : c[0]) || c[3 * v + 1] != (((v >> 1) & 1) ? c[7] : c[1]) ||
            c[3 * v + 2] != ((v >> 2) ? c[14] : c[2]))
          return false;
      if (!(c[3] > c[0] && c[7] > c[1] && c[14] > c[2]))
        return false;
    }
  return true;
}

// dim 2: are all coarse cells axis-aligned rectangles with positive extent (the only kind the reference's
// refined hyper_cube produces, ms.tpp:97-99)?  Same test as assemble_kernel's.
static bool
all_rectangles(const double *corners, size_t n_cells)
{
  for (size_t k = 0; k < n_cells; ++k)
    {
      const double *c = corners + 8 * k;
      if (!(c[0] == c[4] && c[2] == c[6] && c[1] == c[3] && c[5] == c[7] && c[2] > c[0] && c[5] > c[1]))
        return false;
    }
  return true;
}

// does msb_run take the fused one-kernel stage?
static bool
fused_eligible(const Shard &s)
{
  return s.dim == 2 && (s.l == 5 || s.l == 6) && s.tier == MSB_TIER_SMEM && (s.variant == 0 || (s.variant >= 10 && s.variant <= 13)) &&
         s.aligned &&
         s.coeff.kind != MSB_COEFF_TABLE;
}

// the HBM copy of the stencil (203 KB per cell at n = 64) is only needed by the three-kernel path and by the
// operator-level accessors: allocated on first use
static cudaError_t
ensure_sten(Shard &s)
{
  if (s.d_sten)
    return cudaSuccess;
  return cudaMalloc((void **)&s.d_sten, sizeof(double) * (size_t)s.n_cells * s.nst * (size_t)s.N);
}

// MSB_COEFF_TABLE: the stiffness matrix is assembled from the symmetric part of the tensor
// (K must be symmetric for the reference's SolverCG as well, basis.tpp:299-306), so a table
// whose tensors are not symmetric up to rounding is rejected instead of being silently
// symmetrised.  The reference MatrixCoeff is asymmetric by ~6e-17 a.
// Returns the index of the first offending tensor or -1.
static long long
first_unsymmetric_tensor(const double *table, size_t n_tensors, int dim)
{
  const int e = dim * dim;
  for (size_t t = 0; t < n_tensors; ++t)
    {
      const double *a = table + e * t;
      double        big = 0.0;
      for (int i = 0; i < e; ++i)
        big = fabs(a[i]) > big ? fabs(a[i]) : big;
      for (int i = 0; i < dim; ++i)
        for (int j = i + 1; j < dim; ++j)
          if (!(fabs(a[dim * i + j] - a[dim * j + i]) <= 1e-12 * big))
            return (long long)t;
    }
  return -1;
}

// doubles of the coefficient table per coarse cell: fine cells x 2^dim q-points x dim*dim entries
static size_t
ncoef_table_doubles(const Shard &s)
{
  return ((size_t)1 << (s.dim * s.l)) * (size_t)s.nb * (size_t)(s.dim * s.dim);
}

static void
free_shard(Shard &s)
{
  cudaSetDevice(s.device);
  void *ptrs[] = {s.d_corners, s.d_q1coef, s.d_table, s.d_sten, s.d_phi,  s.d_M,    s.d_b,
                  s.d_iters,   s.d_res,    s.d_fail,  s.d_dofmap, s.d_invmap, s.d_gsol, s.d_tmp,
                  s.d_wr,      s.d_wp,     s.d_wq,    s.d_scal, s.d_part, s.d_flags,
                  s.d_wz,      s.d_wv,     s.d_dinv,  s.d_gal,  s.d_w,    s.d_stage[0], s.d_stage[1], s.d_wr2};
  for (void *p : ptrs)
    if (p)
      cudaFree(p);
  for (auto &q : s.stage_stream)
    if (q)
      cudaStreamDestroy(q);
  for (auto &e : s.ev)
    if (e)
      cudaEventDestroy(e);
  for (auto &e : s.ev_chunk)
    if (e)
      cudaEventDestroy(e);
  if (s.stream)
    cudaStreamDestroy(s.stream);
}

extern "C" int
msb_create(const msb_config *cfg, const double *corners, const double *coeff_table, msb_handle *out)
{
  if (!cfg || !corners || !out)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != MSB_ABI_VERSION)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: abi_version %d, library has %d", cfg->abi_version,
                MSB_ABI_VERSION);
  if (cfg->dim != 2 && cfg->dim != 3)
    return fail(MSB_ERR_UNSUPPORTED, "msb_create: dim=%d (the reference instantiates 2 and 3)", cfg->dim);
  const int max_l = cfg->dim == 2 ? 9 : 6;
  if (cfg->n_refine_local < 1 || cfg->n_refine_local > max_l)
    return fail(MSB_ERR_UNSUPPORTED, "msb_create: n_refine_local=%d outside 1..%d for dim=%d",
                cfg->n_refine_local, max_l, cfg->dim);
  if (cfg->dim == 3 && cfg->coeff.kind != MSB_COEFF_REFERENCE && cfg->coeff.kind != MSB_COEFF_CONSTANT &&
      cfg->coeff.kind != MSB_COEFF_TABLE)
    return fail(MSB_ERR_UNSUPPORTED,
                "msb_create: dim=3 supports MSB_COEFF_REFERENCE, MSB_COEFF_CONSTANT and MSB_COEFF_TABLE");
  if (cfg->dim == 3 && cfg->tier == MSB_TIER_SMEM)
    return fail(MSB_ERR_UNSUPPORTED, "msb_create: dim=3 runs in the streamed tier");
  if (cfg->n_cells < 1)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: n_cells=%d", cfg->n_cells);
  if (cfg->coeff.kind < MSB_COEFF_REFERENCE || cfg->coeff.kind > MSB_COEFF_TABLE)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: coefficient kind %d", cfg->coeff.kind);
  if (cfg->coeff.kind == MSB_COEFF_TABLE && !coeff_table)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: MSB_COEFF_TABLE needs coeff_table");
  if (cfg->coeff.kind == MSB_COEFF_TABLE)
    {
      const size_t    ncell_f = (size_t)1 << (cfg->dim * cfg->n_refine_local);
      const long long bad = first_unsymmetric_tensor(coeff_table, (size_t)cfg->n_cells * ncell_f * (1u << cfg->dim), cfg->dim);
      if (bad >= 0)
        return fail(MSB_ERR_INVALID_ARG,
                    "msb_create: coefficient tensor %lld of the table is not symmetric (|a_ij - a_ji| > 1e-12 max|a|): "
                    "the stage, like the reference's SolverCG, needs a symmetric stiffness matrix", bad);
    }
  if (cfg->coeff.kind == MSB_COEFF_PERIODIC && !(cfg->coeff.par[0] > 0.0))
    return fail(MSB_ERR_INVALID_ARG, "msb_create: periodic coefficient needs eps > 0");
  if (cfg->coeff.kind == MSB_COEFF_INCLUSIONS && !(cfg->coeff.par[0] > 0.0))
    return fail(MSB_ERR_INVALID_ARG, "msb_create: inclusion coefficient needs block size > 0");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
      cudaGetLastError();
      return fail(MSB_ERR_NO_DEVICE, "msb_create: no CUDA device (this library has no CPU path)");
    }
  if (cfg->device_id < 0 || cfg->device_id >= ndev)
    return fail(MSB_ERR_NO_DEVICE, "msb_create: device %d of %d", cfg->device_id, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device_id));
  if (prop.major != 10)
    return fail(MSB_ERR_NO_DEVICE, "msb_create: device %d is sm_%d%d, kernels are built for sm_100a",
                cfg->device_id, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(cfg->device_id));

  msb_handle h = new (std::nothrow) msb_handle_s();
  if (!h)
    return fail(MSB_ERR_INVALID_ARG, "msb_create: out of host memory");
  Shard &s    = h->s;
  s.dim       = cfg->dim;
  s.nb        = 1 << s.dim;
  s.nst       = s.dim == 2 ? (int)ST_NARR : (int)ST3_NARR;
  s.l         = cfg->n_refine_local;
  s.n         = 1 << s.l;
  s.np        = s.n + 1;
  s.N         = s.dim == 2 ? s.np * s.np : s.np * s.np * s.np;
  s.n_cells   = cfg->n_cells;
  s.device    = cfg->device_id;
  s.variant   = cfg->variant;
  s.coeff     = cfg->coeff;
  s.rhs_value = cfg->rhs_value;
  s.tier      = cfg->tier;
  if (s.dim == 3)
    s.tier = MSB_TIER_STREAMED;
  s.bricks  = s.dim == 3 && all_bricks(corners, (size_t)s.n_cells);
  s.aligned = s.dim == 2 && all_rectangles(corners, (size_t)s.n_cells);
  if (s.tier == MSB_TIER_AUTO)
    s.tier = smem_tier_supported(s.l) ? MSB_TIER_SMEM : MSB_TIER_STREAMED;
  if (s.tier == MSB_TIER_SMEM && !smem_tier_supported(s.l))
    {
      delete h;
      return fail(MSB_ERR_UNSUPPORTED, "msb_create: shared-memory tier needs 3 <= n_refine_local <= 6");
    }

  const size_t C = (size_t)s.n_cells, N = (size_t)s.N, NB = (size_t)s.nb, NCORN = NB * s.dim;

#define ALLOC(ptr, count)                                                                      \
  do                                                                                           \
    {                                                                                          \
      cudaError_t e_ = cudaMalloc((void **)&(ptr), sizeof(*(ptr)) * (size_t)(count));          \
      if (e_ != cudaSuccess)                                                                   \
        {                                                                                      \
          free_shard(s);                                                                       \
          delete h;                                                                            \
          return fail(MSB_ERR_CUDA, "cudaMalloc(%s, %zu bytes) failed: %s", #ptr,              \
                      sizeof(*(ptr)) * (size_t)(count), cudaGetErrorString(e_));               \
        }                                                                                      \
    }                                                                                          \
  while (0)

  ALLOC(s.d_corners, NCORN * C);
  ALLOC(s.d_q1coef, NB * NB * C);
  if (!fused_eligible(s)) // the fused stage keeps the stencil on chip
    ALLOC(s.d_sten, C * s.nst * N);
  ALLOC(s.d_phi, C * NB * N);
  ALLOC(s.d_M, NB * NB * C);
  ALLOC(s.d_b, NB * C);
  ALLOC(s.d_iters, NB * C);
  ALLOC(s.d_res, NB * C);
  ALLOC(s.d_fail, 2);
  ALLOC(s.d_dofmap, N);
  ALLOC(s.d_invmap, N);
  ALLOC(s.d_tmp, 3 * N);
  ALLOC(s.d_flags, 4);
  if (s.coeff.kind == MSB_COEFF_TABLE)
    ALLOC(s.d_table, C * ncoef_table_doubles(s));
  if (s.tier == MSB_TIER_STREAMED)
    {
      const size_t cn = s.dim == 2 ? streamed_coarse_nodes(s.l) : dim3_coarse_nodes(s.l);
      // the cluster kernel (n = 128) keeps every vector on chip: no HBM work vectors for it
      const bool on_chip = s.dim == 2 && streamed_tier_uses_cluster(s.l, s.variant);
      if (!on_chip)
        {
          ALLOC(s.d_wr, C * NB * N);
          ALLOC(s.d_wp, C * NB * N);
          ALLOC(s.d_wq, C * NB * N);
          ALLOC(s.d_wz, C * NB * N);
          ALLOC(s.d_wv, C * NB * cn);
          ALLOC(s.d_scal, NB * C);
          ALLOC(s.d_part, NB * C * (size_t)(s.dim == 2 ? 2 * 4 * 128 : dim3_part_stride()));
          if (s.dim == 2)
            ALLOC(s.d_wr2, C * NB * N);
        }
      ALLOC(s.d_dinv, C * cn);
      if (s.dim == 2)
        ALLOC(s.d_gal, streamed_galerkin_scratch_doubles(s.l, s.n_cells));
    }
#undef ALLOC

  cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i)
    e = cudaEventCreate(&s.ev[i]);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(s.d_corners, corners, sizeof(double) * NCORN * C, cudaMemcpyHostToDevice, s.stream);
  int32_t bad_cell = INT_MAX;
  if (e == cudaSuccess)
    e = launch_basis_q1(s, s.stream, &bad_cell); // BasisQ1 coefficient matrices, on the device
  if (e == cudaSuccess && s.d_table)
    e = cudaMemcpyAsync(s.d_table, coeff_table, sizeof(double) * C * ncoef_table_doubles(s),
                        cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(s.stream);
  if (e == cudaSuccess)
    e = s.dim == 2 ? launch_dofmap(s, s.stream) : launch_dofmap3(s, s.stream);
  if (e != cudaSuccess)
    {
      free_shard(s);
      delete h;
      return fail(MSB_ERR_CUDA, "msb_create: device setup failed: %s", cudaGetErrorString(e));
    }
  if (bad_cell != INT_MAX)
    {
      free_shard(s);
      delete h;
      return fail(MSB_ERR_INVALID_ARG, "msb_create: coarse cell %d is degenerate", bad_cell);
    }
  *out = h;
  return MSB_OK;
}

extern "C" int
msb_set_cells(msb_handle h, const double *corners, const double *coeff_table)
{
  if (!h || !corners)
    return fail(MSB_ERR_INVALID_ARG, "msb_set_cells: null argument");
  Shard &s = h->s;
  if (s.coeff.kind == MSB_COEFF_TABLE && !coeff_table)
    return fail(MSB_ERR_INVALID_ARG, "msb_set_cells: MSB_COEFF_TABLE needs coeff_table");
  CUDA_TRY(cudaSetDevice(s.device));
  if (s.run_pending)
    {
      // results of the run in flight belong to the old cells: wait for it, then drop them
      const cudaError_t e = cudaStreamSynchronize(s.run_stream);
      s.run_pending       = false;
      if (e != cudaSuccess)
        {
          s.valid = false;
          return fail(MSB_ERR_CUDA, "msb_set_cells: pending run failed: %s", cudaGetErrorString(e));
        }
    }
  const size_t C = (size_t)s.n_cells, NB = (size_t)s.nb, NCORN = NB * s.dim;
  if (s.coeff.kind == MSB_COEFF_TABLE)
    {
      const size_t    ncell_f = (size_t)1 << (s.dim * s.l);
      const long long bad     = first_unsymmetric_tensor(coeff_table, C * ncell_f * NB, s.dim);
      if (bad >= 0) // nothing on the device has been touched yet: the handle keeps its old cells
        return fail(MSB_ERR_INVALID_ARG, "msb_set_cells: coefficient tensor %lld of the table is not symmetric", bad);
    }
  // from here on device state is overwritten: every result of the previous batch is void, and
  // a failure leaves the handle unusable until a later msb_set_cells succeeds
  s.assembled = s.ran = s.weights_set = false;
  s.valid     = false;
  s.last_status = MSB_OK;
  CUDA_TRY(cudaMemcpyAsync(s.d_corners, corners, sizeof(double) * NCORN * C, cudaMemcpyHostToDevice, s.stream));
  int32_t bad_cell = INT_MAX;
  CUDA_TRY(launch_basis_q1(s, s.stream, &bad_cell));
  if (bad_cell != INT_MAX)
    return fail(MSB_ERR_INVALID_ARG, "msb_set_cells: coarse cell %d is degenerate", bad_cell);
  if (s.d_table)
    CUDA_TRY(cudaMemcpyAsync(s.d_table, coeff_table, sizeof(double) * C * (size_t)ncoef_table_doubles(s),
                             cudaMemcpyHostToDevice, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  s.bricks  = s.dim == 3 && all_bricks(corners, C);
  s.aligned = s.dim == 2 && all_rectangles(corners, C);
  s.valid   = true;
  return MSB_OK;
}

extern "C" int
msb_destroy(msb_handle h)
{
  if (!h)
    return MSB_OK;
  cudaSetDevice(h->s.device);
  cudaDeviceSynchronize();
  free_shard(h->s);
  delete h;
  return MSB_OK;
}

// the stage: assemble_system -> 2^dim x (condense, PCG, distribute) -> element matrices
extern "C" int
msb_run_async(msb_handle h, double tol_abs, int32_t max_iter, void *cuda_stream)
{
  if (!h)
    return fail(MSB_ERR_INVALID_ARG, "msb_run: null handle");
  if (!(tol_abs >= 0.0) || max_iter < 0)
    return fail(MSB_ERR_INVALID_ARG, "msb_run: tol_abs=%g max_iter=%d", tol_abs, max_iter);
  Shard &s = h->s;
  if (!s.valid)
    return fail(MSB_ERR_STATE, "msb_run: the last msb_set_cells failed; set valid cells first");
  CUDA_TRY(cudaSetDevice(s.device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s.stream;
  s.run_stream    = st;
  s.n_launches    = 0;
  s.ran           = false;
  const int32_t init_fail[2] = {INT_MAX, 0};
  CUDA_TRY(cudaMemcpyAsync(s.d_fail, init_fail, sizeof init_fail, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(s.ev[0], st));
  s.fused_last = fused_eligible(s);
  if (s.fused_last)
    {
      // assemble_system + 4 x (condense, PCG, distribute) + assemble_global_element_matrix of every cell
      // in ONE launch; nothing but Phi, M, b leaves the SM
      CUDA_TRY(cudaEventRecord(s.ev[1], st));
      CUDA_TRY(launch_stage_fused(s, tol_abs, max_iter, st, &s.n_launches));
      s.tier_used = s.tier;
      CUDA_TRY(cudaEventRecord(s.ev[2], st));
      CUDA_TRY(cudaEventRecord(s.ev[3], st));
      s.run_pending = true;
      s.weights_set = false;
      return MSB_OK;
    }
  CUDA_TRY(ensure_sten(s));
  if (s.dim == 3)
    CUDA_TRY(launch_assemble3(s, st, &s.n_launches));
  else
    CUDA_TRY(launch_assemble(s, st, &s.n_launches));
  s.assembled = true;
  CUDA_TRY(cudaEventRecord(s.ev[1], st));
  if (s.dim == 3)
    CUDA_TRY(launch_solve3(s, tol_abs, max_iter, st, &s.n_launches));
  else if (s.tier == MSB_TIER_SMEM && s.variant < 100)
    CUDA_TRY(launch_solve_bpx(s, tol_abs, max_iter, st, &s.n_launches)); // multilevel PCG (default)
  else if (s.tier == MSB_TIER_SMEM)
    CUDA_TRY(launch_solve_smem(s, tol_abs, max_iter, st, &s.n_launches)); // Jacobi PCG (variant >= 100)
  else
    CUDA_TRY(launch_solve_streamed(s, tol_abs, max_iter, st, &s.n_launches));
  s.tier_used = s.tier;
  CUDA_TRY(cudaEventRecord(s.ev[2], st));
  if (s.dim == 3)
    CUDA_TRY(launch_element_matrices3(s, st, &s.n_launches));
  else
    CUDA_TRY(launch_element_matrices(s, st, &s.n_launches));
  CUDA_TRY(cudaEventRecord(s.ev[3], st));
  s.run_pending = true;
  s.weights_set = false;
  return MSB_OK;
}

extern "C" int
msb_sync(msb_handle h)
{
  if (!h)
    return fail(MSB_ERR_INVALID_ARG, "msb_sync: null handle");
  Shard &s = h->s;
  if (!s.run_pending)
    return s.last_status;
  CUDA_TRY(cudaSetDevice(s.device));
  CUDA_TRY(cudaStreamSynchronize(s.run_stream));
  int32_t f[2];
  CUDA_TRY(cudaMemcpy(f, s.d_fail, sizeof f, cudaMemcpyDeviceToHost));
  s.run_pending = false;
  s.ran         = true;
  s.last_status = MSB_OK;
  if (f[0] != INT_MAX)
    s.last_status = fail(MSB_ERR_NO_CONVERGENCE,
                         "local solve (cell %d, basis %d) did not reach the tolerance within max_iter "
                         "(the reference throws SolverControl::NoConvergence here)",
                         f[0] / s.nb, f[0] % s.nb);
  return s.last_status;
}

extern "C" int
msb_run(msb_handle h, double tol_abs, int32_t max_iter)
{
  const int rc = msb_run_async(h, tol_abs, max_iter, nullptr);
  if (rc != MSB_OK)
    return rc;
  return msb_sync(h);
}

extern "C" int msb_get_bases(msb_handle h, int32_t cell0, int32_t n_cells, double *out);

// msb_run + msb_get_bases, pipelined over chunks of cells (fused one-kernel stage only): chunk k is reordered
// and copied device -> host on the staging streams while chunk k+1 is being solved on the library stream.
extern "C" int
msb_run_with_bases(msb_handle h, double tol_abs, int32_t max_iter, double *bases_out)
{
  if (!h || !bases_out)
    return fail(MSB_ERR_INVALID_ARG, "msb_run_with_bases: null argument");
  if (!(tol_abs >= 0.0) || max_iter < 0)
    return fail(MSB_ERR_INVALID_ARG, "msb_run_with_bases: tol_abs=%g max_iter=%d", tol_abs, max_iter);
  Shard &s = h->s;
  if (!s.valid)
    return fail(MSB_ERR_STATE, "msb_run_with_bases: the last msb_set_cells failed; set valid cells first");
  if (!fused_eligible(s))
    {
      const int rc = msb_run(h, tol_abs, max_iter);
      if (rc != MSB_OK && rc != MSB_ERR_NO_CONVERGENCE)
        return rc;
      const int rc2 = msb_get_bases(h, 0, s.n_cells, bases_out);
      return rc2 != MSB_OK ? rc2 : rc;
    }
  CUDA_TRY(cudaSetDevice(s.device));
  if (s.run_pending)
    CUDA_TRY(cudaStreamSynchronize(s.run_stream));
  // chunks of ~28 waves of CTAs: long enough to hide the launch, short enough that the last copy is small
  int ndev_sm = 148;
  cudaDeviceGetAttribute(&ndev_sm, cudaDevAttrMultiProcessorCount, s.device);
  const int    chunk = 28 * ndev_sm < s.n_cells ? 28 * ndev_sm : s.n_cells;
  const size_t vecs  = (size_t)chunk * s.nb;
  if (s.stage_vecs < vecs)
    {
      for (int k = 0; k < 2; ++k)
        {
          if (s.d_stage[k])
            CUDA_TRY(cudaFree(s.d_stage[k]));
          s.d_stage[k] = nullptr;
          CUDA_TRY(cudaMalloc((void **)&s.d_stage[k], sizeof(double) * vecs * (size_t)s.N));
          if (!s.stage_stream[k])
            CUDA_TRY(cudaStreamCreateWithFlags(&s.stage_stream[k], cudaStreamNonBlocking));
        }
      s.stage_vecs = vecs;
    }
  if (!s.ev_chunk[0])
    for (auto &e : s.ev_chunk)
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaStream_t st = s.stream;
  s.run_stream    = st;
  s.n_launches    = 0;
  s.ran           = false;
  s.fused_last    = true;
  const int32_t init_fail[2] = {INT_MAX, 0};
  CUDA_TRY(cudaMemcpyAsync(s.d_fail, init_fail, sizeof init_fail, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(s.ev[0], st));
  CUDA_TRY(cudaEventRecord(s.ev[1], st));
  int k = 0;
  for (int c0 = 0; c0 < s.n_cells; c0 += chunk, k ^= 1)
    {
      const int nc = s.n_cells - c0 < chunk ? s.n_cells - c0 : chunk;
      CUDA_TRY(launch_stage_fused_range(s, c0, nc, tol_abs, max_iter, st, &s.n_launches));
      CUDA_TRY(cudaEventRecord(s.ev_chunk[k], st));
      // staging buffer k was last used two chunks ago on the same staging stream: stream order protects it
      CUDA_TRY(cudaStreamWaitEvent(s.stage_stream[k], s.ev_chunk[k], 0));
      const size_t v0 = (size_t)c0 * s.nb, nv = (size_t)nc * s.nb;
      CUDA_TRY(launch_permute_batch(s, s.d_phi + v0 * (size_t)s.N, s.d_stage[k], nv, s.stage_stream[k]));
      CUDA_TRY(cudaMemcpyAsync(bases_out + v0 * (size_t)s.N, s.d_stage[k], sizeof(double) * nv * (size_t)s.N,
                               cudaMemcpyDeviceToHost, s.stage_stream[k]));
      // (the event of buffer k is re-recorded two chunks later: the copy that waits on it has been enqueued
      //  behind the wait, and cudaStreamWaitEvent captured the event's state at the time of the call)
    }
  s.tier_used = s.tier;
  CUDA_TRY(cudaEventRecord(s.ev[2], st));
  CUDA_TRY(cudaEventRecord(s.ev[3], st));
  s.run_pending = true;
  s.weights_set = false;
  const int rc = msb_sync(h);
  for (int q = 0; q < 2; ++q)
    {
      const cudaError_t e = cudaStreamSynchronize(s.stage_stream[q]);
      if (e != cudaSuccess)
        return fail(MSB_ERR_CUDA, "msb_run_with_bases: %s", cudaGetErrorString(e));
    }
  return rc;
}

static int
need_ran(msb_handle h, const char *who)
{
  if (!h)
    return fail(MSB_ERR_INVALID_ARG, "%s: null handle", who);
  if (!h->s.valid)
    return fail(MSB_ERR_STATE, "%s: the last msb_set_cells failed; set valid cells first", who);
  if (h->s.run_pending)
    {
      const int rc = msb_sync(h);
      if (rc != MSB_OK && rc != MSB_ERR_NO_CONVERGENCE)
        return rc;
    }
  if (!h->s.ran)
    return fail(MSB_ERR_STATE, "%s: msb_run has not been called", who);
  cudaError_t e = cudaSetDevice(h->s.device);
  if (e != cudaSuccess)
    return fail(MSB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return MSB_OK;
}

extern "C" int
msb_get_failure(msb_handle h, int32_t *cell, int32_t *index_basis, double *residual)
{
  int rc = need_ran(h, "msb_get_failure");
  if (rc != MSB_OK)
    return rc;
  int32_t f[2];
  CUDA_TRY(cudaMemcpy(f, h->s.d_fail, sizeof f, cudaMemcpyDeviceToHost));
  if (f[0] == INT_MAX)
    {
      if (cell)
        *cell = -1;
      if (index_basis)
        *index_basis = -1;
      if (residual)
        *residual = 0.0;
      return MSB_OK;
    }
  if (cell)
    *cell = f[0] / h->s.nb;
  if (index_basis)
    *index_basis = f[0] % h->s.nb;
  if (residual)
    CUDA_TRY(cudaMemcpy(residual, h->s.d_res + f[0], sizeof(double), cudaMemcpyDeviceToHost));
  return MSB_OK;
}

extern "C" int
msb_get_element_matrices(msb_handle h, double *M, double *b)
{
  int rc = need_ran(h, "msb_get_element_matrices");
  if (rc != MSB_OK)
    return rc;
  const size_t C = (size_t)h->s.n_cells, NB = (size_t)h->s.nb;
  if (M)
    CUDA_TRY(cudaMemcpy(M, h->s.d_M, sizeof(double) * NB * NB * C, cudaMemcpyDeviceToHost));
  if (b)
    CUDA_TRY(cudaMemcpy(b, h->s.d_b, sizeof(double) * NB * C, cudaMemcpyDeviceToHost));
  return MSB_OK;
}

extern "C" int
msb_get_iteration_counts(msb_handle h, int32_t *iters, double *residuals)
{
  int rc = need_ran(h, "msb_get_iteration_counts");
  if (rc != MSB_OK)
    return rc;
  const size_t C = (size_t)h->s.n_cells, NB = (size_t)h->s.nb;
  if (iters)
    CUDA_TRY(cudaMemcpy(iters, h->s.d_iters, sizeof(int32_t) * NB * C, cudaMemcpyDeviceToHost));
  if (residuals)
    CUDA_TRY(cudaMemcpy(residuals, h->s.d_res, sizeof(double) * NB * C, cudaMemcpyDeviceToHost));
  return MSB_OK;
}

static int
check_cell(msb_handle h, int32_t cell, int32_t ib, const char *who)
{
  if (cell < 0 || cell >= h->s.n_cells || ib < 0 || ib >= h->s.nb)
    return fail(MSB_ERR_INVALID_ARG, "%s: cell %d / basis %d out of range", who, cell, ib);
  return MSB_OK;
}

extern "C" int
msb_get_basis(msb_handle h, int32_t cell, int32_t ib, double *out)
{
  int rc = need_ran(h, "msb_get_basis");
  if (rc != MSB_OK)
    return rc;
  if ((rc = check_cell(h, cell, ib, "msb_get_basis")) != MSB_OK || !out)
    return rc != MSB_OK ? rc : fail(MSB_ERR_INVALID_ARG, "msb_get_basis: null out");
  Shard &s = h->s;
  CUDA_TRY(launch_permute(s, s.d_phi + ((size_t)cell * s.nb + ib) * s.N, s.d_tmp, true, s.stream));
  CUDA_TRY(cudaMemcpyAsync(out, s.d_tmp, sizeof(double) * s.N, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return MSB_OK;
}

extern "C" int
msb_get_dof_map(msb_handle h, uint32_t *dof_of_vertex)
{
  if (!h || !dof_of_vertex)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_dof_map: null argument");
  CUDA_TRY(cudaSetDevice(h->s.device));
  CUDA_TRY(cudaMemcpy(dof_of_vertex, h->s.d_dofmap, sizeof(uint32_t) * h->s.N, cudaMemcpyDeviceToHost));
  return MSB_OK;
}

extern "C" int
msb_get_constraints(msb_handle h, int32_t cell, int32_t ib, uint32_t *dofs, double *values)
{
  if (!h || !dofs || !values)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_constraints: null argument");
  int rc = check_cell(h, cell, ib, "msb_get_constraints");
  if (rc != MSB_OK)
    return rc;
  Shard &s = h->s;
  if (!s.valid)
    return fail(MSB_ERR_STATE, "msb_get_constraints: the last msb_set_cells failed; set valid cells first");
  CUDA_TRY(cudaSetDevice(s.device));
  const int nb    = s.dim == 2 ? 4 * s.n : s.N - (s.n - 1) * (s.n - 1) * (s.n - 1);
  uint32_t *d_dof = reinterpret_cast<uint32_t *>(s.d_tmp);       // nb uint32 <= N doubles
  double   *d_val = s.d_tmp + s.N;
  if (s.dim == 3)
    CUDA_TRY(launch_constraints3(s, cell, ib, d_dof, d_val, s.stream));
  else
    CUDA_TRY(launch_constraints(s, cell, ib, d_dof, d_val, s.stream));
  CUDA_TRY(cudaMemcpyAsync(dofs, d_dof, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaMemcpyAsync(values, d_val, sizeof(double) * nb, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return MSB_OK;
}

static int
ensure_assembled(msb_handle h)
{
  Shard &s = h->s;
  if (!s.valid)
    return fail(MSB_ERR_STATE, "the last msb_set_cells failed; set valid cells first");
  if (s.run_pending)
    {
      const int rc = msb_sync(h);
      if (rc != MSB_OK && rc != MSB_ERR_NO_CONVERGENCE)
        return rc;
    }
  if (!s.assembled)
    {
      int nl = 0;
      CUDA_TRY(ensure_sten(s));
      if (s.dim == 3)
        CUDA_TRY(launch_assemble3(s, s.stream, &nl));
      else
        CUDA_TRY(launch_assemble(s, s.stream, &nl));
      CUDA_TRY(cudaStreamSynchronize(s.stream));
      s.assembled = true;
    }
  return MSB_OK;
}

extern "C" int
msb_apply_operator(msb_handle h, int32_t cell, const double *x, double *y)
{
  if (!h || !x || !y)
    return fail(MSB_ERR_INVALID_ARG, "msb_apply_operator: null argument");
  int rc = check_cell(h, cell, 0, "msb_apply_operator");
  if (rc != MSB_OK)
    return rc;
  Shard &s = h->s;
  CUDA_TRY(cudaSetDevice(s.device));
  if ((rc = ensure_assembled(h)) != MSB_OK)
    return rc;
  double *d_in = s.d_tmp, *d_lex = s.d_tmp + s.N, *d_out = s.d_tmp + 2 * (size_t)s.N;
  CUDA_TRY(cudaMemcpyAsync(d_in, x, sizeof(double) * s.N, cudaMemcpyHostToDevice, s.stream));
  CUDA_TRY(launch_permute(s, d_in, d_lex, false, s.stream));      // dof order -> lex
  if (s.dim == 3)
    CUDA_TRY(launch_apply_operator3(s, cell, d_lex, d_out, s.stream));
  else
    CUDA_TRY(launch_apply_operator(s, cell, d_lex, d_out, s.stream));
  CUDA_TRY(launch_permute(s, d_out, d_in, true, s.stream));       // lex -> dof order
  CUDA_TRY(cudaMemcpyAsync(y, d_in, sizeof(double) * s.N, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return MSB_OK;
}

extern "C" int
msb_get_load_vector(msb_handle h, int32_t cell, double *F)
{
  if (!h || !F)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_load_vector: null argument");
  int rc = check_cell(h, cell, 0, "msb_get_load_vector");
  if (rc != MSB_OK)
    return rc;
  Shard &s = h->s;
  CUDA_TRY(cudaSetDevice(s.device));
  if ((rc = ensure_assembled(h)) != MSB_OK)
    return rc;
  CUDA_TRY(launch_permute(s, s.d_sten + ((size_t)cell * s.nst + (s.dim == 2 ? (int)ST_F : (int)ST3_F)) * s.N, s.d_tmp, true, s.stream));
  CUDA_TRY(cudaMemcpyAsync(F, s.d_tmp, sizeof(double) * s.N, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return MSB_OK;
}

extern "C" int
msb_set_global_weights(msb_handle h, const double *w)
{
  int rc = need_ran(h, "msb_set_global_weights");
  if (rc != MSB_OK)
    return rc;
  if (!w)
    return fail(MSB_ERR_INVALID_ARG, "msb_set_global_weights: null weights");
  Shard       &s = h->s;
  const size_t C = (size_t)s.n_cells;
  if (!s.d_gsol)
    CUDA_TRY(cudaMalloc((void **)&s.d_gsol, sizeof(double) * C * s.N));
  if (!s.d_w) // allocated once, with the global-solution array, on the first call
    CUDA_TRY(cudaMalloc((void **)&s.d_w, sizeof(double) * s.nb * C));
  cudaError_t e = cudaMemcpyAsync(s.d_w, w, sizeof(double) * s.nb * C, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess)
    e = launch_global_solution(s, s.d_w, s.stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess)
    return fail(MSB_ERR_CUDA, "msb_set_global_weights: %s", cudaGetErrorString(e));
  s.weights_set = true;
  return MSB_OK;
}

extern "C" int
msb_get_global_solution(msb_handle h, int32_t cell, double *out)
{
  int rc = need_ran(h, "msb_get_global_solution");
  if (rc != MSB_OK)
    return rc;
  if ((rc = check_cell(h, cell, 0, "msb_get_global_solution")) != MSB_OK)
    return rc;
  Shard &s = h->s;
  // the reference asserts is_set_global_weights (basis.tpp:425-426)
  if (!s.weights_set)
    return fail(MSB_ERR_STATE, "msb_get_global_solution: global weights must be set first");
  if (!out)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_global_solution: null out");
  CUDA_TRY(launch_permute(s, s.d_gsol + (size_t)cell * s.N, s.d_tmp, true, s.stream));
  CUDA_TRY(cudaMemcpyAsync(out, s.d_tmp, sizeof(double) * s.N, cudaMemcpyDeviceToHost, s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  return MSB_OK;
}

// n_vec consecutive device vectors (lexicographic) -> host, deal.II DoF order.  Two staging buffers on two
// streams: the device->host copy of chunk k overlaps the reordering launch of chunk k+1.
static int
bulk_to_host(Shard &s, const double *d_src, size_t n_vec, double *out, const char *who)
{
  if (n_vec == 0)
    return MSB_OK;
  if (!s.d_stage[0])
    {
      // 64 MB per staging buffer, at least one vector, at most what is asked for
      size_t vecs = ((size_t)64 << 20) / (sizeof(double) * (size_t)s.N);
      vecs        = vecs < 1 ? 1 : vecs;
      const size_t most = (size_t)s.nb * (size_t)s.n_cells;
      vecs        = vecs > most ? most : vecs;
      for (int k = 0; k < 2; ++k)
        {
          CUDA_TRY(cudaMalloc((void **)&s.d_stage[k], sizeof(double) * vecs * (size_t)s.N));
          CUDA_TRY(cudaStreamCreateWithFlags(&s.stage_stream[k], cudaStreamNonBlocking));
        }
      s.stage_vecs = vecs;
    }
  CUDA_TRY(cudaStreamSynchronize(s.stream)); // whatever produced d_src on the library stream
  int k = 0;
  for (size_t v0 = 0; v0 < n_vec; v0 += s.stage_vecs, k ^= 1)
    {
      const size_t nv = n_vec - v0 < s.stage_vecs ? n_vec - v0 : s.stage_vecs;
      // stream order protects the staging buffer: the previous copy out of it is on the same stream
      CUDA_TRY(launch_permute_batch(s, d_src + v0 * (size_t)s.N, s.d_stage[k], nv, s.stage_stream[k]));
      CUDA_TRY(cudaMemcpyAsync(out + v0 * (size_t)s.N, s.d_stage[k], sizeof(double) * nv * (size_t)s.N,
                               cudaMemcpyDeviceToHost, s.stage_stream[k]));
    }
  for (int q = 0; q < 2; ++q)
    {
      const cudaError_t e = cudaStreamSynchronize(s.stage_stream[q]);
      if (e != cudaSuccess)
        return fail(MSB_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
    }
  return MSB_OK;
}

static int
check_range(msb_handle h, int32_t cell0, int32_t n, const char *who)
{
  if (cell0 < 0 || n < 0 || (long long)cell0 + n > h->s.n_cells)
    return fail(MSB_ERR_INVALID_ARG, "%s: cells [%d, %d) outside the shard's %d cells", who, cell0, cell0 + n,
                h->s.n_cells);
  return MSB_OK;
}

extern "C" int
msb_get_bases(msb_handle h, int32_t cell0, int32_t n_cells, double *out)
{
  int rc = need_ran(h, "msb_get_bases");
  if (rc != MSB_OK)
    return rc;
  if ((rc = check_range(h, cell0, n_cells, "msb_get_bases")) != MSB_OK)
    return rc;
  if (!out && n_cells > 0)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_bases: null out");
  Shard &s = h->s;
  return bulk_to_host(s, s.d_phi + (size_t)cell0 * s.nb * s.N, (size_t)n_cells * s.nb, out, "msb_get_bases");
}

extern "C" int
msb_get_global_solutions(msb_handle h, int32_t cell0, int32_t n_cells, double *out)
{
  int rc = need_ran(h, "msb_get_global_solutions");
  if (rc != MSB_OK)
    return rc;
  if ((rc = check_range(h, cell0, n_cells, "msb_get_global_solutions")) != MSB_OK)
    return rc;
  Shard &s = h->s;
  if (!s.weights_set) // the reference asserts is_set_global_weights (basis.tpp:425-426)
    return fail(MSB_ERR_STATE, "msb_get_global_solutions: global weights must be set first");
  if (!out && n_cells > 0)
    return fail(MSB_ERR_INVALID_ARG, "msb_get_global_solutions: null out");
  return bulk_to_host(s, s.d_gsol + (size_t)cell0 * s.N, (size_t)n_cells, out, "msb_get_global_solutions");
}

extern "C" int
msb_get_device_results(msb_handle h, uint64_t *d_M, uint64_t *d_b, uint64_t *d_iters)
{
  int rc = need_ran(h, "msb_get_device_results");
  if (rc != MSB_OK)
    return rc;
  if (d_M)
    *d_M = (uint64_t)(uintptr_t)h->s.d_M;
  if (d_b)
    *d_b = (uint64_t)(uintptr_t)h->s.d_b;
  if (d_iters)
    *d_iters = (uint64_t)(uintptr_t)h->s.d_iters;
  return MSB_OK;
}

extern "C" int
msb_get_run_stats(msb_handle h, float *ms_total, float *ms_solve, int32_t *n_launches, int32_t *tier_used)
{
  int rc = need_ran(h, "msb_get_run_stats");
  if (rc != MSB_OK)
    return rc;
  Shard &s = h->s;
  if (ms_total)
    CUDA_TRY(cudaEventElapsedTime(ms_total, s.ev[0], s.ev[3]));
  if (ms_solve)
    CUDA_TRY(cudaEventElapsedTime(ms_solve, s.ev[1], s.ev[2]));
  if (n_launches)
    *n_launches = s.n_launches;
  if (tier_used)
    *tier_used = s.tier_used;
  return MSB_OK;
}

extern "C" int
msb_get_algorithmic_bytes(msb_handle h, double *bytes, double *mean_iterations)
{
  int rc = need_ran(h, "msb_get_algorithmic_bytes");
  if (rc != MSB_OK)
    return rc;
  Shard               &s = h->s;
  const size_t         n = (size_t)s.nb * s.n_cells;
  std::vector<int32_t> it(n);
  CUDA_TRY(cudaMemcpy(it.data(), s.d_iters, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  // W = N (96 k + 16) bytes per solve
  double tot = 0.0, ksum = 0.0;
  for (size_t i = 0; i < n; ++i)
    {
      tot += (double)s.N * (96.0 * it[i] + 16.0);
      ksum += it[i];
    }
  if (bytes)
    *bytes = tot;
  if (mean_iterations)
    *mean_iterations = ksum / (double)n;
  return MSB_OK;
}
