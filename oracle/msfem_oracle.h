/*
 * msfem_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's local multiscale-basis stage
 * (/root/reference/include/base/diffusion_problem_basis.tpp) together with the
 * deal.II 9.1 semantics it relies on.  deal.II is not
 * available in the authoring container and the reference holds no numerical
 * test, so this oracle is PARITY UNPINNED against the reference binary; it is
 * pinned instead against the exact invariants of SURVEY.md Appendix B and the
 * survey-time cross-check values (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may use this library.  The product (the CUDA library behind
 * include/msfem_basis.h) never links or calls it.
 */
#ifndef MSFEM_ORACLE_H
#define MSFEM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* coefficient kinds (restating include/coefficients/matrix_coeff.tpp and the
 * BASELINE.md section-4 synthetic coefficients) */
enum {
  ORC_COEFF_REFERENCE  = 0, /* MatrixCoeff<2> verbatim, PI_D typo included    */
  ORC_COEFF_PERIODIC   = 1, /* par[0]=eps, par[1]=scale; true pi, a*I         */
  ORC_COEFF_INCLUSIONS = 2, /* par[0]=block, par[1]=prob, par[2]=a_incl,
                               par[3]=a_background; seed                     */
  ORC_COEFF_CONSTANT   = 3, /* par[0]=a0                                      */
  ORC_COEFF_TABLE      = 4  /* tensor given per (cell, fine cell, q)          */
};

enum { ORC_PRECOND_SSOR = 0, ORC_PRECOND_JACOBI = 1 };

typedef struct
{
  int32_t kind;
  int32_t seed;
  double  par[6];
} orc_coeff;

/* number of fine DoFs of one coarse cell: (2^l+1)^2 */
int orc_n_dofs(int l);

/* deal.II first-touch DoF numbering (basis.tpp:106).
 * dof_of_vertex[jy*(n+1)+jx] = DoF index of fine vertex (jx,jy). */
void orc_dof_map(int l, uint32_t *dof_of_vertex);

/* sorted list of the 4n constrained (boundary) DoFs (basis.tpp:119-135). */
int orc_boundary_dofs(int l, uint32_t *out);

/* BasisQ1<2> coefficient matrix (basis_q1.tpp:26-47): coef[r*4+c], column c
 * holds the monomial coefficients (1,x,y,xy) of the basis of vertex c. */
void orc_basis_q1_coeffs(const double corners[8], double coef[16]);

/* BasisQ1<2>::value (basis_q1.tpp:86-96). */
double orc_basis_q1_value(const double coef[16], int index_basis, double x, double y);

/* diffusion tensor at a point, row-major a00,a01,a10,a11
 * (matrix_coeff.tpp:66-91 for ORC_COEFF_REFERENCE). */
void orc_coeff_eval(const orc_coeff *c, double x, double y, double A[4]);

/* inhomogeneities of constraint set `index_basis`, listed in the order of
 * orc_boundary_dofs (basis.tpp:129-133). */
void orc_constraint_values(int l, const double corners[8], int index_basis, double *vals);

/* CSR of the unconstrained fine stiffness matrix (basis.tpp:159-242), diagonal
 * first in each row then ascending columns, plus the load vector.
 * rowptr has N+1 entries, col/val have (3n+1)^2. table may be NULL. */
void orc_assemble(int l, const double corners[8], const orc_coeff *c, const double *table,
                  double rhs_value, uint64_t *rowptr, uint32_t *col, double *val, double *F);

/* y = K x with that CSR (SparseMatrix::vmult). */
void orc_vmult(int N, const uint64_t *rowptr, const uint32_t *col, const double *val,
               const double *x, double *y);

/* Full DiffusionProblemBasis<2>::run() for one coarse cell (basis.tpp:438-474).
 * phi: [4][N] in deal.II DoF order (may be NULL); M: 16 row-major; b: 4;
 * iters/res: 4 each.  Returns 0, or 1+index_basis of the first solve that hit
 * max_iter (the reference would throw SolverControl::NoConvergence). */
int orc_run_cell(int l, const double corners[8], const orc_coeff *c, const double *table,
                 double rhs_value, double tol, int max_iter, int precond, double omega,
                 double *phi, double *M, double *b, int32_t *iters, double *res);

/* The serial hot loop ms.tpp:81-87 over n_cells cells, split over n_threads
 * OpenMP threads in contiguous ranges (the p4est rule of SURVEY A.6).
 * corners: [n_cells][4][2]; table: [n_cells][n*n][4][4] or NULL;
 * phi: [n_cells][4][N] or NULL.  Returns the number of failed cells. */
int orc_run_cells(int l, int n_cells, const double *corners, const orc_coeff *c,
                  const double *table, double rhs_value, double tol, int max_iter, int precond,
                  double omega, int n_threads, double *phi, double *M, double *b, int32_t *iters,
                  double *res);

/* set_global_weights (basis.tpp:352-377): out = sum_i w[i] * phi[i]. */
void orc_global_solution(int N, const double *phi, const double w[4], double *out);

/* ------------------------------------------------------------------- 3D ----
 * The dim = 3 instantiation of the same class (diffusion_problem_basis.inst.cc:15-16;
 * BasisQ1<3> basis_q1.tpp:50-75,99-113; MatrixCoeff<3> matrix_coeff.tpp:28-41).
 * Vertices and children in deal.II's lexicographic hex order (x fastest), fine cells in
 * 3D Morton order, 2x2x2 Gauss, trilinear mapping, 2^3 = 8 bases per coarse cell.
 * Arrays are indexed [jz][jy][jx]; corners are [8][3]; coefficient tensors are 3x3 row-major.
 * Coefficient kinds: ORC_COEFF_REFERENCE, ORC_COEFF_CONSTANT. */
int    orc3_n_dofs(int l);
void   orc3_dof_map(int l, uint32_t *dof_of_vertex);
int    orc3_boundary_dofs(int l, uint32_t *out);
void   orc3_basis_q1_coeffs(const double corners[24], double coef[64]);
double orc3_basis_q1_value(const double coef[64], int index_basis, double x, double y, double z);
void   orc3_coeff_eval(const orc_coeff *c, double x, double y, double z, double A[9]);
void   orc3_constraint_values(int l, const double corners[24], int index_basis, double *vals);
/* nnz is returned; col/val must hold 27 * (n+1)^3 entries */
uint64_t orc3_assemble(int l, const double corners[24], const orc_coeff *c, double rhs_value,
                       uint64_t *rowptr, uint32_t *col, double *val, double *F);
/* phi [n_cells][8][N] or NULL, M [n_cells][64], b [n_cells][8], iters/res [n_cells][8] */
int orc3_run_cells(int l, int n_cells, const double *corners, const orc_coeff *c, double rhs_value,
                   double tol, int max_iter, int precond, double omega, int n_threads, double *phi,
                   double *M, double *b, int32_t *iters, double *res);
/* Same with tabulated tensors: table [n_cells][n^3 fine cells, (iz n + iy) n + ix][8 q-points, x fastest][9] or NULL */
int orc3_run_cells_table(int l, int n_cells, const double *corners, const orc_coeff *c, const double *table,
                         double rhs_value, double tol, int max_iter, int precond, double omega, int n_threads,
                         double *phi, double *M, double *b, int32_t *iters, double *res);

#ifdef __cplusplus
}
#endif
#endif
