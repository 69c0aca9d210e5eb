/*
 * msfem_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restates, in plain C99, what DiffusionProblemBasis<2>::run() computes
 * (/root/reference/include/base/diffusion_problem_basis.tpp:438-474) using the
 * same algorithms the reference obtains from deal.II 9.1: first-touch DoF
 * numbering on the Morton-ordered refined cell, CSR assembly with QGauss<2>(2)
 * and MappingQ1, AffineConstraints::condense, SolverCG + PreconditionSSOR(1.6)
 * with SolverControl(1000, 1e-12), AffineConstraints::distribute and the
 * Phi^T K Phi / Phi^T F reduction.  See msfem_oracle.h for the "parity
 * unpinned" statement.  Every function cites the reference lines it follows.
 */
#include "msfem_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#  include <omp.h>
#endif

/* include/coefficients/coefficients.h:21 -- sic, not pi */
static const double PI_D = 3.14592653509793218403;
/* true pi for the BASELINE.md synthetic coefficients */
static const double PI_TRUE = 3.14159265358979323846;

/* ---------------------------------------------------------------- mesh -- */

static inline uint32_t
morton_deinterleave(uint32_t m)
{
  /* keep the even bits of m, compacted */
  uint32_t x = m & 0x55555555u;
  x          = (x | (x >> 1)) & 0x33333333u;
  x          = (x | (x >> 2)) & 0x0f0f0f0fu;
  x          = (x | (x >> 4)) & 0x00ff00ffu;
  x          = (x | (x >> 8)) & 0x0000ffffu;
  return x;
}

int
orc_n_dofs(int l)
{
  const int n = 1 << l;
  return (n + 1) * (n + 1);
}

/* basis.tpp:98 refine_global + basis.tpp:106 distribute_dofs(FE_Q(1)):
 * active cells in Morton order (child = ix_bit + 2*iy_bit), vertices 0..3 of
 * each cell numbered at first touch. */
void
orc_dof_map(int l, uint32_t *dof)
{
  const uint32_t n  = 1u << l;
  const uint32_t np = n + 1;
  for (uint32_t i = 0; i < np * np; ++i)
    dof[i] = 0xffffffffu;
  uint32_t next = 0;
  for (uint32_t m = 0; m < n * n; ++m)
    {
      const uint32_t ix = morton_deinterleave(m);
      const uint32_t iy = morton_deinterleave(m >> 1);
      for (uint32_t v = 0; v < 4; ++v)
        {
          const uint32_t jx = ix + (v & 1u), jy = iy + (v >> 1);
          if (dof[jy * np + jx] == 0xffffffffu)
            dof[jy * np + jx] = next++;
        }
    }
}

static int
cmp_u32(const void *a, const void *b)
{
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

/* basis.tpp:129-133: every DoF on the boundary of the coarse cell (boundary
 * id 0 everywhere, general_cell colorize=false) gets a constraint line. */
int
orc_boundary_dofs(int l, uint32_t *out)
{
  const uint32_t n = 1u << l, np = n + 1;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * np * np);
  orc_dof_map(l, dof);
  int cnt = 0;
  for (uint32_t jy = 0; jy < np; ++jy)
    for (uint32_t jx = 0; jx < np; ++jx)
      if (jx == 0 || jy == 0 || jx == n || jy == n)
        out[cnt++] = dof[jy * np + jx];
  qsort(out, (size_t)cnt, sizeof(uint32_t), cmp_u32);
  free(dof);
  return cnt;
}

/* fine vertex (jx,jy) of the l-times refined general_cell (basis.tpp:94-98):
 * repeated midpoint refinement of a straight-sided quad is the bilinear image
 * of the uniform grid; written so that axis-aligned dyadic cells are exact. */
static inline void
fine_vertex(const double c[8], uint32_t n, uint32_t jx, uint32_t jy, double p[2])
{
  const double s = (double)jx / (double)n, t = (double)jy / (double)n;
  for (int a = 0; a < 2; ++a)
    {
      const double v0 = c[0 + a], v1 = c[2 + a], v2 = c[4 + a], v3 = c[6 + a];
      p[a] = v0 + s * (v1 - v0) + t * (v2 - v0) + (s * t) * ((v3 - v2) - (v1 - v0));
    }
}

/* -------------------------------------------------------- coefficients -- */

/* 4x4 inverse by cofactors (stands in for FullMatrix::invert, basis_q1.tpp:46) */
static void
invert4(const double m[16], double inv[16])
{
  double a[16];
  a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] +
         m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] -
         m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] +
         m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] -
          m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] -
         m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] +
         m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] -
         m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] +
          m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] +
         m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] -
         m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] +
          m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] -
          m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] -
         m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] +
         m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] -
          m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] +
          m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const double det  = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
  const double idet = 1.0 / det;
  for (int i = 0; i < 16; ++i)
    inv[i] = a[i] * idet;
}

/* basis_q1.tpp:26-47: point_matrix rows (1, x, y, xy) at the 4 vertices;
 * coeff_matrix = point_matrix^{-1}; column i = coefficients of basis i. */
void
orc_basis_q1_coeffs(const double corners[8], double coef[16])
{
  double pm[16];
  for (int i = 0; i < 4; ++i)
    {
      const double x = corners[2 * i], y = corners[2 * i + 1];
      pm[4 * i + 0] = 1.0;
      pm[4 * i + 1] = x;
      pm[4 * i + 2] = y;
      pm[4 * i + 3] = x * y;
    }
  invert4(pm, coef);
}

/* basis_q1.tpp:86-96 */
double
orc_basis_q1_value(const double coef[16], int ib, double x, double y)
{
  return coef[0 * 4 + ib] + coef[1 * 4 + ib] * x + coef[2 * 4 + ib] * y + coef[3 * 4 + ib] * x * y;
}

static inline uint64_t
mix64(uint64_t z)
{
  z ^= z >> 33;
  z *= 0xff51afd7ed558ccdULL;
  z ^= z >> 33;
  z *= 0xc4ceb9fe1a85ec53ULL;
  z ^= z >> 33;
  return z;
}

/* Counter-based Bernoulli draw for the high-contrast inclusion lattice of
 * BASELINE.md cfg4 / SURVEY 8(d): hash of (block_ix, block_iy, seed). */
static inline int
inclusion_draw(int64_t bx, int64_t by, int32_t seed, double prob)
{
  uint64_t h = (uint64_t)bx * 0x9E3779B97F4A7C15ULL;
  h ^= mix64((uint64_t)by + 0xC2B2AE3D27D4EB4FULL * (uint64_t)(uint32_t)seed);
  h = mix64(h);
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) < prob;
}

void
orc_coeff_eval(const orc_coeff *c, double x, double y, double A[4])
{
  switch (c->kind)
    {
      case ORC_COEFF_REFERENCE:
        {
          /* matrix_coeff.hpp:45-48, matrix_coeff.tpp:17-25, :66-91 */
          const int    k     = 57;
          const double scale = 0.9999;
          const double alpha = PI_D / 3;
          const double rot[2][2] = {{cos(alpha), sin(alpha)}, {-sin(alpha), cos(alpha)}};
          const double a =
            1.0 * (1.0 - scale * (0.5 * sin(2 * PI_D * k * x) + 0.5 * sin(2 * PI_D * k * y)));
          const double v[2][2] = {{a, 0.0}, {0.0, a}};
          double       t[2][2];
          /* values = rot * values * transpose(rot) */
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
              t[i][j] = rot[i][0] * v[0][j] + rot[i][1] * v[1][j];
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
              A[2 * i + j] = t[i][0] * rot[j][0] + t[i][1] * rot[j][1];
          return;
        }
      case ORC_COEFF_PERIODIC:
        {
          const double eps = c->par[0], scale = c->par[1];
          const double a =
            1.0 - scale * (0.5 * sin(2 * PI_TRUE * x / eps) + 0.5 * sin(2 * PI_TRUE * y / eps));
          A[0] = a, A[1] = 0.0, A[2] = 0.0, A[3] = a;
          return;
        }
      case ORC_COEFF_INCLUSIONS:
        {
          const double  bs = c->par[0];
          const int64_t bx = (int64_t)floor(x / bs), by = (int64_t)floor(y / bs);
          const double  a = inclusion_draw(bx, by, c->seed, c->par[1]) ? c->par[2] : c->par[3];
          A[0] = a, A[1] = 0.0, A[2] = 0.0, A[3] = a;
          return;
        }
      case ORC_COEFF_CONSTANT:
      default:
        A[0] = c->par[0], A[1] = 0.0, A[2] = 0.0, A[3] = c->par[0];
        return;
    }
}

void
orc_constraint_values(int l, const double corners[8], int ib, double *vals)
{
  const uint32_t n = 1u << l, np = n + 1;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * np * np);
  uint32_t *bd  = (uint32_t *)malloc(sizeof(uint32_t) * 4 * n);
  double   *g   = (double *)malloc(sizeof(double) * np * np);
  double    coef[16];
  orc_dof_map(l, dof);
  orc_basis_q1_coeffs(corners, coef);
  for (uint32_t jy = 0; jy < np; ++jy)
    for (uint32_t jx = 0; jx < np; ++jx)
      {
        double p[2];
        fine_vertex(corners, n, jx, jy, p);
        g[dof[jy * np + jx]] = orc_basis_q1_value(coef, ib, p[0], p[1]);
      }
  const int nb = orc_boundary_dofs(l, bd);
  for (int i = 0; i < nb; ++i)
    vals[i] = g[bd[i]];
  free(dof), free(bd), free(g);
}

/* ------------------------------------------------------------ assembly -- */

/* sparsity (basis.tpp:137-143): 9-point pattern, diagonal first then
 * ascending column index. */
static void
build_sparsity(int l, const uint32_t *dof, uint64_t *rowptr, uint32_t *col)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np;
  uint32_t *cnt = (uint32_t *)calloc(N, sizeof(uint32_t));
  for (uint32_t jy = 0; jy < np; ++jy)
    for (uint32_t jx = 0; jx < np; ++jx)
      {
        const uint32_t wx = 1 + (jx > 0) + (jx < n), wy = 1 + (jy > 0) + (jy < n);
        cnt[dof[jy * np + jx]] = wx * wy;
      }
  rowptr[0] = 0;
  for (uint32_t r = 0; r < N; ++r)
    rowptr[r + 1] = rowptr[r] + cnt[r];
  for (uint32_t jy = 0; jy < np; ++jy)
    for (uint32_t jx = 0; jx < np; ++jx)
      {
        const uint32_t r = dof[jy * np + jx];
        uint32_t      *c = col + rowptr[r];
        uint32_t       k = 0;
        c[k++]           = r;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx)
            {
              if (dx == 0 && dy == 0)
                continue;
              const int x = (int)jx + dx, y = (int)jy + dy;
              if (x < 0 || y < 0 || x > (int)n || y > (int)n)
                continue;
              c[k++] = dof[(uint32_t)y * np + (uint32_t)x];
            }
        qsort(c + 1, k - 1, sizeof(uint32_t), cmp_u32);
      }
  free(cnt);
}

static inline double *
csr_entry(const uint64_t *rowptr, const uint32_t *col, double *val, uint32_t r, uint32_t c)
{
  for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
    if (col[k] == c)
      return val + k;
  return NULL;
}

/* basis.tpp:159-242 (assemble_system) with FEValues<2>(FE_Q(1), QGauss(2),
 * MappingQ1) restated per SURVEY A.7. */
static void
assemble(int l, const double corners[8], const orc_coeff *c, const double *table, double rhs_value,
         const uint32_t *dof, const uint64_t *rowptr, const uint32_t *col, double *val, double *F)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np;
  memset(val, 0, sizeof(double) * rowptr[N]);
  memset(F, 0, sizeof(double) * N);

  const double g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);
  const double gp[2] = {g0, g1};

  for (uint32_t m = 0; m < n * n; ++m)
    {
      const uint32_t ix = morton_deinterleave(m), iy = morton_deinterleave(m >> 1);
      double         P[4][2];
      uint32_t       ld[4];
      for (uint32_t v = 0; v < 4; ++v)
        {
          fine_vertex(corners, n, ix + (v & 1u), iy + (v >> 1), P[v]);
          ld[v] = dof[(iy + (v >> 1)) * np + ix + (v & 1u)];
        }
      double Ke[4][4] = {{0}}, Fe[4] = {0};
      for (int q = 0; q < 4; ++q)
        {
          const double xi = gp[q & 1], eta = gp[q >> 1];
          /* reference shape values / gradients, vertex order of SURVEY A.1 */
          const double Nv[4]    = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
          const double dN[4][2] = {{-(1 - eta), -(1 - xi)}, {(1 - eta), -xi}, {-eta, (1 - xi)}, {eta, xi}};
          double       J[2][2]  = {{0, 0}, {0, 0}}, xq[2] = {0, 0};
          for (int v = 0; v < 4; ++v)
            for (int a = 0; a < 2; ++a)
              {
                xq[a] += P[v][a] * Nv[v];
                J[a][0] += P[v][a] * dN[v][0];
                J[a][1] += P[v][a] * dN[v][1];
              }
          const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          const double Ji[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
          const double JxW = det * 0.25;
          double       G[4][2];
          for (int v = 0; v < 4; ++v)
            for (int a = 0; a < 2; ++a)
              G[v][a] = Ji[0][a] * dN[v][0] + Ji[1][a] * dN[v][1];

          double A[4];
          if (c->kind == ORC_COEFF_TABLE)
            memcpy(A, table + ((size_t)(iy * n + ix) * 4 + (size_t)q) * 4, sizeof A);
          else
            orc_coeff_eval(c, xq[0], xq[1], A);

          /* basis.tpp:207-222: Ke(i,j) += grad_i * A * grad_j * JxW */
          for (int i = 0; i < 4; ++i)
            {
              const double t0 = G[i][0] * A[0] + G[i][1] * A[2];
              const double t1 = G[i][0] * A[1] + G[i][1] * A[3];
              for (int j = 0; j < 4; ++j)
                Ke[i][j] += (t0 * G[j][0] + t1 * G[j][1]) * JxW;
              Fe[i] += Nv[i] * rhs_value * JxW;
            }
        }
      /* basis.tpp:225-240: scatter, no constraints */
      for (int i = 0; i < 4; ++i)
        {
          for (int j = 0; j < 4; ++j)
            *csr_entry(rowptr, col, val, ld[i], ld[j]) += Ke[i][j];
          F[ld[i]] += Fe[i];
        }
    }
}

void
orc_assemble(int l, const double corners[8], const orc_coeff *c, const double *table,
             double rhs_value, uint64_t *rowptr, uint32_t *col, double *val, double *F)
{
  const uint32_t n = 1u << l, np = n + 1;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * np * np);
  orc_dof_map(l, dof);
  build_sparsity(l, dof, rowptr, col);
  assemble(l, corners, c, table, rhs_value, dof, rowptr, col, val, F);
  free(dof);
}

void
orc_vmult(int N, const uint64_t *rowptr, const uint32_t *col, const double *val, const double *x,
          double *y)
{
  for (int r = 0; r < N; ++r)
    {
      double s = 0.0;
      for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        s += val[k] * x[col[k]];
      y[r] = s;
    }
}

/* ------------------------------------------------------------- solver -- */

/* AffineConstraints::condense(matrix, rhs) for constraint lines without
 * entries (basis.tpp:461).  is_bd marks constrained DoFs, g
 * holds their inhomogeneities. */
static void
condense(uint32_t N, const uint64_t *rowptr, const uint32_t *col, double *val, double *rhs,
         const unsigned char *is_bd, const double *g)
{
  double avg = 0.0;
  for (uint32_t r = 0; r < N; ++r)
    avg += fabs(val[rowptr[r]]);
  avg /= (double)N;
  for (uint32_t r = 0; r < N; ++r)
    {
      if (!is_bd[r])
        {
          for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
            if (is_bd[col[k]])
              {
                rhs[r] -= val[k] * g[col[k]];
                val[k] = 0.0;
              }
        }
      else
        {
          for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
            val[k] = 0.0;
          val[rowptr[r]] = avg;
          rhs[r]         = 0.0;
        }
    }
}

/* SparseMatrix::precondition_SSOR with PreconditionSSOR::initialize's
 * pos_right_of_diagonal (basis.tpp:300-301):
 * dst = om(2-om) (D+om U)^{-1} D (D+om L)^{-1} src, sweeps in DoF order. */
static void
ssor(uint32_t N, const uint64_t *rowptr, const uint32_t *col, const double *val,
     const uint64_t *right_of_diag, double om, const double *src, double *dst)
{
  for (uint32_t r = 0; r < N; ++r)
    {
      double s = 0.0;
      for (uint64_t k = rowptr[r] + 1; k < right_of_diag[r]; ++k)
        s += val[k] * dst[col[k]];
      dst[r] = (src[r] - s * om) / val[rowptr[r]];
    }
  for (uint32_t r = 0; r < N; ++r)
    dst[r] *= om * (2.0 - om) * val[rowptr[r]];
  for (uint32_t r = N; r-- > 0;)
    {
      double s = 0.0;
      for (uint64_t k = right_of_diag[r]; k < rowptr[r + 1]; ++k)
        s += val[k] * dst[col[k]];
      dst[r] = (dst[r] - s * om) / val[rowptr[r]];
    }
}

static double
dot(uint32_t N, const double *a, const double *b)
{
  double s = 0.0;
  for (uint32_t i = 0; i < N; ++i)
    s += a[i] * b[i];
  return s;
}

/* SolverCG<Vector<double>>::solve with SolverControl(max_iter, tol)
 * (basis.tpp:297-306).  x enters as zeros (basis.tpp:152).
 * Returns the iteration count, negative if max_iter was hit. */
static int
pcg(uint32_t N, const uint64_t *rowptr, const uint32_t *col, const double *val,
    const uint64_t *right_of_diag, int precond, double om, const double *b, double *x, double tol,
    int max_iter, double *g, double *h, double *d, double *res_out)
{
  /* g = A x - b with x = 0 */
  for (uint32_t i = 0; i < N; ++i)
    g[i] = -b[i];
  double res = sqrt(dot(N, g, g));
  int    it  = 0;
  if (res <= tol)
    {
      *res_out = res;
      return 0;
    }
  if (precond == ORC_PRECOND_SSOR)
    ssor(N, rowptr, col, val, right_of_diag, om, g, h);
  else
    for (uint32_t i = 0; i < N; ++i)
      h[i] = g[i] / val[rowptr[i]];
  for (uint32_t i = 0; i < N; ++i)
    d[i] = -h[i];
  double gh = dot(N, g, h);
  for (;;)
    {
      ++it;
      orc_vmult((int)N, rowptr, col, val, d, h);
      const double alpha = gh / dot(N, d, h);
      for (uint32_t i = 0; i < N; ++i)
        x[i] += alpha * d[i];
      for (uint32_t i = 0; i < N; ++i)
        g[i] += alpha * h[i];
      res = sqrt(dot(N, g, g));
      if (res <= tol)
        break;
      if (it >= max_iter || res != res)
        {
          *res_out = res;
          return -it;
        }
      if (precond == ORC_PRECOND_SSOR)
        ssor(N, rowptr, col, val, right_of_diag, om, g, h);
      else
        for (uint32_t i = 0; i < N; ++i)
          h[i] = g[i] / val[rowptr[i]];
      const double gh_old = gh;
      gh                  = dot(N, g, h);
      const double beta   = gh / gh_old;
      for (uint32_t i = 0; i < N; ++i)
        d[i] = beta * d[i] - h[i];
    }
  *res_out = res;
  return it;
}

typedef struct
{
  uint32_t      *dof, *col;
  uint64_t      *rowptr, *right_of_diag;
  double        *K, *S, *F, *rhs, *g, *w0, *w1, *w2, *phi;
  unsigned char *is_bd;
} workspace;

static void
ws_alloc(workspace *w, int l)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np;
  const uint64_t nnz = (uint64_t)(3 * n + 1) * (3 * n + 1);
  w->dof            = (uint32_t *)malloc(sizeof(uint32_t) * N);
  w->col            = (uint32_t *)malloc(sizeof(uint32_t) * nnz);
  w->rowptr         = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  w->right_of_diag  = (uint64_t *)malloc(sizeof(uint64_t) * N);
  w->K              = (double *)malloc(sizeof(double) * nnz);
  w->S              = (double *)malloc(sizeof(double) * nnz);
  w->F              = (double *)malloc(sizeof(double) * N);
  w->rhs            = (double *)malloc(sizeof(double) * N);
  w->g              = (double *)malloc(sizeof(double) * N);
  w->w0             = (double *)malloc(sizeof(double) * N);
  w->w1             = (double *)malloc(sizeof(double) * N);
  w->w2             = (double *)malloc(sizeof(double) * N);
  w->phi            = (double *)malloc(sizeof(double) * 4 * N);
  w->is_bd          = (unsigned char *)malloc(N);
  /* make_grid + setup_system part that does not depend on the cell */
  orc_dof_map(l, w->dof);
  build_sparsity(l, w->dof, w->rowptr, w->col);
  for (uint32_t r = 0; r < N; ++r)
    {
      uint64_t k = w->rowptr[r] + 1;
      while (k < w->rowptr[r + 1] && w->col[k] < r)
        ++k;
      w->right_of_diag[r] = k;
    }
  memset(w->is_bd, 0, N);
  for (uint32_t jy = 0; jy < np; ++jy)
    for (uint32_t jx = 0; jx < np; ++jx)
      if (jx == 0 || jy == 0 || jx == n || jy == n)
        w->is_bd[w->dof[jy * np + jx]] = 1;
}

static void
ws_free(workspace *w)
{
  free(w->dof), free(w->col), free(w->rowptr), free(w->right_of_diag), free(w->K), free(w->S);
  free(w->F), free(w->rhs), free(w->g), free(w->w0), free(w->w1), free(w->w2), free(w->phi);
  free(w->is_bd);
}

/* DiffusionProblemBasis<2>::run() (basis.tpp:438-474) on a prepared workspace.
 * NOTE: the reference re-does make_grid/setup_system per cell; the cell-
 * independent part lives in ws_alloc here, which only makes this CPU baseline
 * FASTER than the reference, never slower. */
static int
run_cell(workspace *w, int l, const double corners[8], const orc_coeff *c, const double *table,
         double rhs_value, double tol, int max_iter, int precond, double omega, double *phi_out,
         double *M, double *b, int32_t *iters, double *res)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np;
  const uint64_t nnz = w->rowptr[N];
  int            fail = 0;
  double         coef[16];

  assemble(l, corners, c, table, rhs_value, w->dof, w->rowptr, w->col, w->K, w->F);
  orc_basis_q1_coeffs(corners, coef);

  for (int ib = 0; ib < 4; ++ib)
    {
      /* setup_system: constraint values (basis.tpp:119-135) */
      for (uint32_t jy = 0; jy < np; ++jy)
        for (uint32_t jx = 0; jx < np; ++jx)
          {
            const uint32_t d = w->dof[jy * np + jx];
            if (w->is_bd[d])
              {
                double p[2];
                fine_vertex(corners, n, jx, jy, p);
                w->g[d] = orc_basis_q1_value(coef, ib, p[0], p[1]);
              }
            else
              w->g[d] = 0.0;
          }
      /* run(): basis.tpp:455-461 */
      memset(w->rhs, 0, sizeof(double) * N);
      memcpy(w->S, w->K, sizeof(double) * nnz);
      condense(N, w->rowptr, w->col, w->S, w->rhs, w->is_bd, w->g);
      /* solve_iterative: basis.tpp:295-308 */
      double *x = w->phi + (size_t)ib * N;
      memset(x, 0, sizeof(double) * N);
      const int it = pcg(N, w->rowptr, w->col, w->S, w->right_of_diag, precond, omega, w->rhs, x,
                         tol, max_iter, w->w0, w->w1, w->w2, &res[ib]);
      iters[ib] = it < 0 ? -it : it;
      if (it < 0 && !fail)
        fail = 1 + ib;
      /* distribute (basis.tpp:308) */
      for (uint32_t d = 0; d < N; ++d)
        if (w->is_bd[d])
          x[d] = w->g[d];
    }

  /* assemble_global_element_matrix: basis.tpp:245-285 */
  for (int i = 0; i < 4; ++i)
    {
      for (int j = 0; j < 4; ++j)
        {
          orc_vmult((int)N, w->rowptr, w->col, w->K, w->phi + (size_t)j * N, w->w0);
          M[4 * i + j] = dot(N, w->phi + (size_t)i * N, w->w0);
        }
      b[i] = dot(N, w->phi + (size_t)i * N, w->F);
    }
  if (phi_out)
    memcpy(phi_out, w->phi, sizeof(double) * 4 * N);
  return fail;
}

int
orc_run_cell(int l, const double corners[8], const orc_coeff *c, const double *table,
             double rhs_value, double tol, int max_iter, int precond, double omega, double *phi,
             double *M, double *b, int32_t *iters, double *res)
{
  workspace w;
  ws_alloc(&w, l);
  const int f =
    run_cell(&w, l, corners, c, table, rhs_value, tol, max_iter, precond, omega, phi, M, b, iters, res);
  ws_free(&w);
  return f;
}

int
orc_run_cells(int l, int n_cells, const double *corners, const orc_coeff *c, const double *table,
              double rhs_value, double tol, int max_iter, int precond, double omega, int n_threads,
              double *phi, double *M, double *b, int32_t *iters, double *res)
{
  const size_t n = (size_t)1 << l, N = (n + 1) * (n + 1);
  int          failed = 0;
  if (n_threads < 1)
    n_threads = 1;
#pragma omp parallel num_threads(n_threads) reduction(+ : failed)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), T = omp_get_num_threads();
#else
    const int t = 0, T = 1;
#endif
    workspace w;
    ws_alloc(&w, l);
    /* contiguous Morton ranges, the p4est rule */
    const int lo = (int)((long long)n_cells * t / T), hi = (int)((long long)n_cells * (t + 1) / T);
    for (int k = lo; k < hi; ++k)
      {
        const int f = run_cell(&w, l, corners + 8 * (size_t)k, c,
                               table ? table + (size_t)k * n * n * 16 : NULL, rhs_value, tol,
                               max_iter, precond, omega, phi ? phi + (size_t)k * 4 * N : NULL,
                               M + 16 * (size_t)k, b + 4 * (size_t)k, iters + 4 * (size_t)k,
                               res + 4 * (size_t)k);
        failed += (f != 0);
      }
    ws_free(&w);
  }
  return failed;
}

/* basis.tpp:352-377 */
void
orc_global_solution(int N, const double *phi, const double w[4], double *out)
{
  for (int d = 0; d < N; ++d)
    out[d] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int d = 0; d < N; ++d)
      out[d] = 1 * out[d] + w[i] * phi[(size_t)i * N + d];
}

/* ========================================================================== 3D
 * DiffusionProblemBasis<3>: same run() (basis.tpp:438-474), dim-templated code paths.
 * condense / ssor / pcg / orc_vmult above are dimension-agnostic (they see a CSR). */

static inline uint32_t
morton3_compact(uint32_t m)
{
  /* keep every third bit of m, compacted */
  uint32_t x = m & 0x09249249u;
  x          = (x | (x >> 2)) & 0x030c30c3u;
  x          = (x | (x >> 4)) & 0x0300f00fu;
  x          = (x | (x >> 8)) & 0x030000ffu;
  x          = (x | (x >> 16)) & 0x000003ffu;
  return x;
}

int
orc3_n_dofs(int l)
{
  const int n = 1 << l;
  return (n + 1) * (n + 1) * (n + 1);
}

/* refine_global + distribute_dofs(FE_Q<3>(1)): active cells in 3D Morton order (child =
 * ix_bit + 2 iy_bit + 4 iz_bit), the 8 vertices of each cell numbered at first touch */
void
orc3_dof_map(int l, uint32_t *dof)
{
  const uint32_t n = 1u << l, np = n + 1;
  for (uint32_t i = 0; i < np * np * np; ++i)
    dof[i] = 0xffffffffu;
  uint32_t next = 0;
  for (uint32_t m = 0; m < n * n * n; ++m)
    {
      const uint32_t ix = morton3_compact(m), iy = morton3_compact(m >> 1), iz = morton3_compact(m >> 2);
      for (uint32_t v = 0; v < 8; ++v)
        {
          const uint32_t jx = ix + (v & 1u), jy = iy + ((v >> 1) & 1u), jz = iz + (v >> 2);
          uint32_t      *d  = dof + (jz * np + jy) * np + jx;
          if (*d == 0xffffffffu)
            *d = next++;
        }
    }
}

int
orc3_boundary_dofs(int l, uint32_t *out)
{
  const uint32_t n = 1u << l, np = n + 1;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * np * np * np);
  orc3_dof_map(l, dof);
  int cnt = 0;
  for (uint32_t jz = 0; jz < np; ++jz)
    for (uint32_t jy = 0; jy < np; ++jy)
      for (uint32_t jx = 0; jx < np; ++jx)
        if (jx == 0 || jy == 0 || jz == 0 || jx == n || jy == n || jz == n)
          out[cnt++] = dof[(jz * np + jy) * np + jx];
  qsort(out, (size_t)cnt, sizeof(uint32_t), cmp_u32);
  free(dof);
  return cnt;
}

/* trilinear image of the uniform grid (exact for axis-aligned dyadic bricks) */
static inline void
fine_vertex3(const double c[24], uint32_t n, uint32_t jx, uint32_t jy, uint32_t jz, double p[3])
{
  const double s = (double)jx / (double)n, t = (double)jy / (double)n, u = (double)jz / (double)n;
  const double w[8] = {(1 - s) * (1 - t) * (1 - u), s * (1 - t) * (1 - u), (1 - s) * t * (1 - u),
                       s * t * (1 - u),             (1 - s) * (1 - t) * u, s * (1 - t) * u,
                       (1 - s) * t * u,             s * t * u};
  for (int a = 0; a < 3; ++a)
    {
      double v = 0.0;
      for (int k = 0; k < 8; ++k)
        v += c[3 * k + a] * w[k];
      p[a] = v;
    }
}

/* dense inverse by Gauss-Jordan with partial pivoting (stands in for FullMatrix::invert) */
static void
invert_dense(int n, const double *m, double *inv)
{
  double *a = (double *)malloc(sizeof(double) * n * 2 * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      {
        a[i * 2 * n + j]     = m[i * n + j];
        a[i * 2 * n + n + j] = i == j ? 1.0 : 0.0;
      }
  for (int c = 0; c < n; ++c)
    {
      int piv = c;
      for (int r = c + 1; r < n; ++r)
        if (fabs(a[r * 2 * n + c]) > fabs(a[piv * 2 * n + c]))
          piv = r;
      if (piv != c)
        for (int j = 0; j < 2 * n; ++j)
          {
            const double t     = a[c * 2 * n + j];
            a[c * 2 * n + j]   = a[piv * 2 * n + j];
            a[piv * 2 * n + j] = t;
          }
      const double ip = 1.0 / a[c * 2 * n + c];
      for (int j = 0; j < 2 * n; ++j)
        a[c * 2 * n + j] *= ip;
      for (int r = 0; r < n; ++r)
        if (r != c)
          {
            const double f = a[r * 2 * n + c];
            if (f != 0.0)
              for (int j = 0; j < 2 * n; ++j)
                a[r * 2 * n + j] -= f * a[c * 2 * n + j];
          }
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      inv[i * n + j] = a[i * 2 * n + n + j];
  free(a);
}

/* basis_q1.tpp:50-75: rows (1, x, y, z, xy, yz, xz, xyz) at the 8 vertices, inverted */
void
orc3_basis_q1_coeffs(const double corners[24], double coef[64])
{
  double pm[64];
  for (int i = 0; i < 8; ++i)
    {
      const double x = corners[3 * i], y = corners[3 * i + 1], z = corners[3 * i + 2];
      double      *r = pm + 8 * i;
      r[0] = 1, r[1] = x, r[2] = y, r[3] = z, r[4] = x * y, r[5] = y * z, r[6] = x * z, r[7] = x * y * z;
    }
  invert_dense(8, pm, coef);
}

/* basis_q1.tpp:99-113 */
double
orc3_basis_q1_value(const double coef[64], int ib, double x, double y, double z)
{
  return coef[0 * 8 + ib] + coef[1 * 8 + ib] * x + coef[2 * 8 + ib] * y + coef[3 * 8 + ib] * z +
         coef[4 * 8 + ib] * x * y + coef[5 * 8 + ib] * y * z + coef[6 * 8 + ib] * x * z +
         coef[7 * 8 + ib] * x * y * z;
}

void
orc3_coeff_eval(const orc_coeff *c, double x, double y, double z, double A[9])
{
  (void)z; /* the reference coefficient depends on x and y only, also in 3D (matrix_coeff.tpp:84-86) */
  if (c->kind == ORC_COEFF_REFERENCE)
    {
      /* matrix_coeff.hpp:45-48, matrix_coeff.tpp:28-41, :66-91 */
      const int    k     = 57;
      const double scale = 0.9999;
      const double al = PI_D / 3, be = PI_D / 6, ga = PI_D / 4;
      const double rot[3][3] = {
        {cos(al) * cos(ga) - sin(al) * cos(be) * sin(ga), -cos(al) * sin(ga) - sin(al) * cos(be) * cos(ga),
         sin(al) * sin(be)},
        {sin(al) * cos(ga) + cos(al) * cos(be) * sin(ga), -sin(al) * sin(ga) + cos(al) * cos(be) * cos(ga),
         -cos(al) * sin(be)},
        {sin(be) * sin(ga), sin(be) * cos(ga), cos(be)}};
      const double a =
        1.0 * (1.0 - scale * (0.5 * sin(2 * PI_D * k * x) + 0.5 * sin(2 * PI_D * k * y)));
      double t[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          t[i][j] = rot[i][j] * a; /* rot * (a I) */
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          A[3 * i + j] = t[i][0] * rot[j][0] + t[i][1] * rot[j][1] + t[i][2] * rot[j][2];
      return;
    }
  for (int i = 0; i < 9; ++i)
    A[i] = 0.0;
  A[0] = A[4] = A[8] = c->par[0];
}

void
orc3_constraint_values(int l, const double corners[24], int ib, double *vals)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np * np;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * N);
  uint32_t *bd  = (uint32_t *)malloc(sizeof(uint32_t) * N);
  double   *g   = (double *)malloc(sizeof(double) * N);
  double    coef[64];
  orc3_dof_map(l, dof);
  orc3_basis_q1_coeffs(corners, coef);
  for (uint32_t jz = 0; jz < np; ++jz)
    for (uint32_t jy = 0; jy < np; ++jy)
      for (uint32_t jx = 0; jx < np; ++jx)
        {
          double p[3];
          fine_vertex3(corners, n, jx, jy, jz, p);
          g[dof[(jz * np + jy) * np + jx]] = orc3_basis_q1_value(coef, ib, p[0], p[1], p[2]);
        }
  const int nb = orc3_boundary_dofs(l, bd);
  for (int i = 0; i < nb; ++i)
    vals[i] = g[bd[i]];
  free(dof), free(bd), free(g);
}

/* 27-point sparsity, diagonal first then ascending columns */
static uint64_t
build_sparsity3(int l, const uint32_t *dof, uint64_t *rowptr, uint32_t *col)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np * np;
  uint32_t *cnt = (uint32_t *)calloc(N, sizeof(uint32_t));
  for (uint32_t jz = 0; jz < np; ++jz)
    for (uint32_t jy = 0; jy < np; ++jy)
      for (uint32_t jx = 0; jx < np; ++jx)
        {
          const uint32_t wx = 1 + (jx > 0) + (jx < n), wy = 1 + (jy > 0) + (jy < n),
                         wz = 1 + (jz > 0) + (jz < n);
          cnt[dof[(jz * np + jy) * np + jx]] = wx * wy * wz;
        }
  rowptr[0] = 0;
  for (uint32_t r = 0; r < N; ++r)
    rowptr[r + 1] = rowptr[r] + cnt[r];
  for (uint32_t jz = 0; jz < np; ++jz)
    for (uint32_t jy = 0; jy < np; ++jy)
      for (uint32_t jx = 0; jx < np; ++jx)
        {
          const uint32_t r = dof[(jz * np + jy) * np + jx];
          uint32_t      *c = col + rowptr[r];
          uint32_t       k = 0;
          c[k++]           = r;
          for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx)
                {
                  if (dx == 0 && dy == 0 && dz == 0)
                    continue;
                  const int x = (int)jx + dx, y = (int)jy + dy, z = (int)jz + dz;
                  if (x < 0 || y < 0 || z < 0 || x > (int)n || y > (int)n || z > (int)n)
                    continue;
                  c[k++] = dof[((uint32_t)z * np + (uint32_t)y) * np + (uint32_t)x];
                }
          qsort(c + 1, k - 1, sizeof(uint32_t), cmp_u32);
        }
  free(cnt);
  return rowptr[N];
}

/* assemble_system for dim = 3: QGauss<3>(2), FE_Q<3>(1), MappingQ1 */
/* table: NULL, or the tensor values a user TensorFunction<2,3>::value_list produced at the fine quadrature
 * points (basis.tpp:202-203): [(iz n + iy) n + ix][q, x fastest][a00 a01 a02 a10 ... a22] */
static void
assemble3(int l, const double corners[24], const orc_coeff *c, const double *table, double rhs_value,
          const uint32_t *dof, const uint64_t *rowptr, const uint32_t *col, double *val, double *F)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np * np;
  memset(val, 0, sizeof(double) * rowptr[N]);
  memset(F, 0, sizeof(double) * N);
  const double gp[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
  for (uint32_t m = 0; m < n * n * n; ++m)
    {
      const uint32_t ix = morton3_compact(m), iy = morton3_compact(m >> 1), iz = morton3_compact(m >> 2);
      double         P[8][3];
      uint32_t       ld[8];
      for (uint32_t v = 0; v < 8; ++v)
        {
          const uint32_t jx = ix + (v & 1u), jy = iy + ((v >> 1) & 1u), jz = iz + (v >> 2);
          fine_vertex3(corners, n, jx, jy, jz, P[v]);
          ld[v] = dof[(jz * np + jy) * np + jx];
        }
      double Ke[8][8] = {{0}}, Fe[8] = {0};
      for (int q = 0; q < 8; ++q)
        {
          const double xi = gp[q & 1], eta = gp[(q >> 1) & 1], ze = gp[q >> 2];
          double       Nv[8], dN[8][3];
          for (int v = 0; v < 8; ++v)
            {
              const double fx = (v & 1) ? xi : 1 - xi, fy = ((v >> 1) & 1) ? eta : 1 - eta,
                           fz = (v >> 2) ? ze : 1 - ze;
              const double sx = (v & 1) ? 1.0 : -1.0, sy = ((v >> 1) & 1) ? 1.0 : -1.0,
                           sz = (v >> 2) ? 1.0 : -1.0;
              Nv[v]    = fx * fy * fz;
              dN[v][0] = sx * fy * fz;
              dN[v][1] = fx * sy * fz;
              dN[v][2] = fx * fy * sz;
            }
          double J[3][3] = {{0}}, xq[3] = {0, 0, 0};
          for (int v = 0; v < 8; ++v)
            for (int a = 0; a < 3; ++a)
              {
                xq[a] += P[v][a] * Nv[v];
                for (int b = 0; b < 3; ++b)
                  J[a][b] += P[v][a] * dN[v][b];
              }
          const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                             J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                             J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
          double Ji[3][3]; /* inverse Jacobian */
          Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
          Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
          Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
          Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
          Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
          Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
          Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
          Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
          Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
          const double JxW = det * 0.125;
          double       G[8][3];
          for (int v = 0; v < 8; ++v)
            for (int a = 0; a < 3; ++a)
              G[v][a] = Ji[0][a] * dN[v][0] + Ji[1][a] * dN[v][1] + Ji[2][a] * dN[v][2];
          double A[9];
          if (table)
            memcpy(A, table + ((((size_t)iz * n + iy) * n + ix) * 8 + (size_t)q) * 9, sizeof A);
          else
            orc3_coeff_eval(c, xq[0], xq[1], xq[2], A);
          for (int i = 0; i < 8; ++i)
            {
              double t[3];
              for (int b = 0; b < 3; ++b)
                t[b] = G[i][0] * A[b] + G[i][1] * A[3 + b] + G[i][2] * A[6 + b];
              for (int j = 0; j < 8; ++j)
                Ke[i][j] += (t[0] * G[j][0] + t[1] * G[j][1] + t[2] * G[j][2]) * JxW;
              Fe[i] += Nv[i] * rhs_value * JxW;
            }
        }
      for (int i = 0; i < 8; ++i)
        {
          for (int j = 0; j < 8; ++j)
            *csr_entry(rowptr, col, val, ld[i], ld[j]) += Ke[i][j];
          F[ld[i]] += Fe[i];
        }
    }
}

uint64_t
orc3_assemble(int l, const double corners[24], const orc_coeff *c, double rhs_value, uint64_t *rowptr,
              uint32_t *col, double *val, double *F)
{
  const uint32_t n = 1u << l, np = n + 1;
  uint32_t *dof = (uint32_t *)malloc(sizeof(uint32_t) * np * np * np);
  orc3_dof_map(l, dof);
  const uint64_t nnz = build_sparsity3(l, dof, rowptr, col);
  assemble3(l, corners, c, NULL, rhs_value, dof, rowptr, col, val, F);
  free(dof);
  return nnz;
}

static int
run_cell3(int l, const double corners[24], const orc_coeff *c, const double *table, double rhs_value, double tol,
          int max_iter, int precond, double omega, double *phi_out, double *M, double *b, int32_t *iters,
          double *res)
{
  const uint32_t n = 1u << l, np = n + 1, N = np * np * np;
  uint32_t      *dof    = (uint32_t *)malloc(sizeof(uint32_t) * N);
  uint64_t      *rowptr = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  uint32_t      *col    = (uint32_t *)malloc(sizeof(uint32_t) * 27 * (size_t)N);
  double        *K      = (double *)malloc(sizeof(double) * 27 * (size_t)N);
  double        *S      = (double *)malloc(sizeof(double) * 27 * (size_t)N);
  uint64_t      *rod    = (uint64_t *)malloc(sizeof(uint64_t) * N);
  double        *F = (double *)malloc(sizeof(double) * N), *rhs = (double *)malloc(sizeof(double) * N);
  double        *g = (double *)malloc(sizeof(double) * N), *w0 = (double *)malloc(sizeof(double) * N);
  double        *w1 = (double *)malloc(sizeof(double) * N), *w2 = (double *)malloc(sizeof(double) * N);
  double        *phi   = (double *)malloc(sizeof(double) * 8 * (size_t)N);
  unsigned char *is_bd = (unsigned char *)calloc(N, 1);
  int            fail  = 0;
  double         coef[64];

  orc3_dof_map(l, dof);
  const uint64_t nnz = build_sparsity3(l, dof, rowptr, col);
  for (uint32_t r = 0; r < N; ++r)
    {
      uint64_t k = rowptr[r] + 1;
      while (k < rowptr[r + 1] && col[k] < r)
        ++k;
      rod[r] = k;
    }
  assemble3(l, corners, c, table, rhs_value, dof, rowptr, col, K, F);
  orc3_basis_q1_coeffs(corners, coef);
  for (uint32_t jz = 0; jz < np; ++jz)
    for (uint32_t jy = 0; jy < np; ++jy)
      for (uint32_t jx = 0; jx < np; ++jx)
        if (jx == 0 || jy == 0 || jz == 0 || jx == n || jy == n || jz == n)
          is_bd[dof[(jz * np + jy) * np + jx]] = 1;

  for (int ib = 0; ib < 8; ++ib)
    {
      for (uint32_t jz = 0; jz < np; ++jz)
        for (uint32_t jy = 0; jy < np; ++jy)
          for (uint32_t jx = 0; jx < np; ++jx)
            {
              const uint32_t d = dof[(jz * np + jy) * np + jx];
              if (is_bd[d])
                {
                  double p[3];
                  fine_vertex3(corners, n, jx, jy, jz, p);
                  g[d] = orc3_basis_q1_value(coef, ib, p[0], p[1], p[2]);
                }
              else
                g[d] = 0.0;
            }
      memset(rhs, 0, sizeof(double) * N);
      memcpy(S, K, sizeof(double) * nnz);
      condense(N, rowptr, col, S, rhs, is_bd, g);
      double *x = phi + (size_t)ib * N;
      memset(x, 0, sizeof(double) * N);
      const int it =
        pcg(N, rowptr, col, S, rod, precond, omega, rhs, x, tol, max_iter, w0, w1, w2, &res[ib]);
      iters[ib] = it < 0 ? -it : it;
      if (it < 0 && !fail)
        fail = 1 + ib;
      for (uint32_t d = 0; d < N; ++d)
        if (is_bd[d])
          x[d] = g[d];
    }
  /* assemble_global_element_matrix: basis.tpp:245-285 with dofs_per_cell = 8 */
  for (int i = 0; i < 8; ++i)
    {
      for (int j = 0; j < 8; ++j)
        {
          orc_vmult((int)N, rowptr, col, K, phi + (size_t)j * N, w0);
          M[8 * i + j] = dot(N, phi + (size_t)i * N, w0);
        }
      b[i] = dot(N, phi + (size_t)i * N, F);
    }
  if (phi_out)
    memcpy(phi_out, phi, sizeof(double) * 8 * (size_t)N);
  free(dof), free(rowptr), free(col), free(K), free(S), free(rod), free(F), free(rhs), free(g);
  free(w0), free(w1), free(w2), free(phi), free(is_bd);
  return fail;
}

int
orc3_run_cells(int l, int n_cells, const double *corners, const orc_coeff *c, double rhs_value,
               double tol, int max_iter, int precond, double omega, int n_threads, double *phi, double *M,
               double *b, int32_t *iters, double *res)
{
  return orc3_run_cells_table(l, n_cells, corners, c, NULL, rhs_value, tol, max_iter, precond, omega, n_threads,
                              phi, M, b, iters, res);
}

int
orc3_run_cells_table(int l, int n_cells, const double *corners, const orc_coeff *c, const double *table,
                     double rhs_value, double tol, int max_iter, int precond, double omega, int n_threads,
                     double *phi, double *M, double *b, int32_t *iters, double *res)
{
  const size_t n = (size_t)1 << l, N = (n + 1) * (n + 1) * (n + 1);
  int          failed = 0;
  if (n_threads < 1)
    n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : failed)
  for (int k = 0; k < n_cells; ++k)
    {
      const int f = run_cell3(l, corners + 24 * (size_t)k, c, table ? table + (size_t)k * n * n * n * 72 : NULL,
                              rhs_value, tol, max_iter, precond, omega,
                              phi ? phi + (size_t)k * 8 * N : NULL, M + 64 * (size_t)k, b + 8 * (size_t)k,
                              iters + 8 * (size_t)k, res + 8 * (size_t)k);
      failed += (f != 0);
    }
  return failed;
}
