// msb_dim3.cu -- the dim = 3 instantiation of the local basis stage
// (diffusion_problem_basis.inst.cc:15-16; BasisQ1<3> basis_q1.tpp:50-75,99-113; MatrixCoeff<3>
// matrix_coeff.tpp:28-41,66-91): hexahedral coarse cells, (2^l+1)^3 fine Q1 nodes, 8 bases.
//
// Data layout in HBM, node index "lex" = (jz*np + jy)*np + jx:
//   corners [C][8][3], q1coef [C][64]
//   sten    [C][15][N]  symmetric 27-point stencil of the unconstrained fine stiffness matrix:
//                       array 0 = diagonal, array k = 1..13 = coupling of node t with node
//                       t + offset(13 + k) where offset(e) = (e%3-1, (e/3)%3-1, e/9-1) (the 13
//                       "forward" neighbours), array 14 = load vector F
//   phi     [C][8][N], M [C][64], b [C][8], iters/res [C][8]
//
// The solver is the batched multilevel-preconditioned CG of msb_solve_stream.cu in three
// dimensions: all (cell, basis) solves advance together through K1 (p = z + beta p), K2
// (q = K p, matrix-free 27-point), K3 (x, r update), restrict / prolong over the nested
// trilinear hierarchy with exact Galerkin diagonals, and the fine-level z = r/D + P z_1.  Dot
// products are deterministic two-level sums with parity double-buffered partials; the stopping
// rule is the reference's ||r||_2 <= tol checked every iteration (basis.tpp:297).
#include <limits.h>
#include <math.h>

#include "msb_internal.cuh"

namespace msb
{
  namespace d3
  {
    constexpr int NB      = 8;
    constexpr int NST     = ST3_NARR;
    constexpr int THREADS = 256;
    constexpr int MAXBLK  = 40;             // max CTAs per coarse cell in the fine kernels
    constexpr int PSTRIDE = 2 * 3 * MAXBLK; // doubles per solve: [parity][rz|pq|rr][blk]
    constexpr int MAXLEV  = 8;

    // ------------------------------------------------------------------------- helpers
    __host__ __device__ inline uint32_t
    compact3(uint32_t m)
    {
      uint32_t x = m & 0x09249249u;
      x          = (x | (x >> 2)) & 0x030c30c3u;
      x          = (x | (x >> 4)) & 0x0300f00fu;
      x          = (x | (x >> 8)) & 0x030000ffu;
      x          = (x | (x >> 16)) & 0x000003ffu;
      return x;
    }

    __host__ __device__ inline uint32_t
    spread3(uint32_t x)
    {
      x &= 0x000003ffu;
      x = (x | (x << 16)) & 0x030000ffu;
      x = (x | (x << 8)) & 0x0300f00fu;
      x = (x | (x << 4)) & 0x030c30c3u;
      x = (x | (x << 2)) & 0x09249249u;
      return x;
    }

    __host__ __device__ inline uint32_t
    morton3(uint32_t ix, uint32_t iy, uint32_t iz)
    {
      return spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2);
    }

    // trilinear image of the uniform grid (exact for axis-aligned dyadic bricks)
    __device__ inline void
    fine_vertex3(const double *__restrict__ c, int n, int jx, int jy, int jz, double p[3])
    {
      const double rn = 1.0 / (double)n, s = jx * rn, t = jy * rn, u = jz * rn;
      const double w[8] = {(1 - s) * (1 - t) * (1 - u), s * (1 - t) * (1 - u), (1 - s) * t * (1 - u),
                           s * t * (1 - u),             (1 - s) * (1 - t) * u, s * (1 - t) * u,
                           (1 - s) * t * u,             s * t * u};
#pragma unroll
      for (int a = 0; a < 3; ++a)
        {
          double v = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v += c[3 * k + a] * w[k];
          p[a] = v;
        }
    }

    // BasisQ1<3>::value (basis_q1.tpp:99-113), coef[r*8+ib], monomials 1,x,y,z,xy,yz,xz,xyz
    __device__ inline double
    basis_q1_value3(const double *__restrict__ coef, int ib, const double p[3])
    {
      const double x = p[0], y = p[1], z = p[2];
      return coef[ib] + coef[8 + ib] * x + coef[16 + ib] * y + coef[24 + ib] * z + coef[32 + ib] * x * y +
             coef[40 + ib] * y * z + coef[48 + ib] * x * z + coef[56 + ib] * x * y * z;
    }

    __device__ __forceinline__ int
    off_of(int e, int np)
    {
      return (e / 9 - 1) * np * np + ((e / 3) % 3 - 1) * np + (e % 3 - 1);
    }

    // coupling K(t, t + offset(e)); both nodes must exist
    __device__ __forceinline__ double
    sten3_get(const double *__restrict__ S, int N, int np, int t, int e)
    {
      if (e == 13)
        return S[t];
      if (e > 13)
        return S[(size_t)(e - 13) * N + t];
      return S[(size_t)(13 - e) * N + t + off_of(e, np)];
    }

    __device__ __forceinline__ void
    decode3(int t, int np, int &jx, int &jy, int &jz)
    {
      jx          = t % np;
      const int q = t / np;
      jy          = q % np;
      jz          = q / np;
    }

    __device__ __forceinline__ bool
    on_boundary3(int jx, int jy, int jz, int n)
    {
      return jx == 0 || jy == 0 || jz == 0 || jx == n || jy == n || jz == n;
    }

    // block-wide deterministic sums of NV values -> valid on thread 0
    template <int NV>
    __device__ __forceinline__ void
    block_sum_to(double (&v)[NV], double *sbuf /*[THREADS/32][NV]*/)
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
              for (int w = 0; w < (int)blockDim.x / 32; ++w)
                s += sbuf[w * NV + k];
              v[k] = s;
            }
        }
    }

    // ========================================================================== DoF map
    // deal.II first-touch numbering on the 3D Morton-ordered fine cells (closed form, as in 2D)
    __device__ inline uint32_t
    first_touch_cell3(int n, int jx, int jy, int jz, int &lv)
    {
      uint32_t best = 0xffffffffu;
      lv            = 0;
#pragma unroll
      for (int v = 0; v < 8; ++v)
        {
          const int ix = jx - (v & 1), iy = jy - ((v >> 1) & 1), iz = jz - (v >> 2);
          if (ix < 0 || iy < 0 || iz < 0 || ix >= n || iy >= n || iz >= n)
            continue;
          const uint32_t m = morton3((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
          if (m < best)
            {
              best = m;
              lv   = v;
            }
        }
      return best;
    }

    __global__ void
    dofmap3_count_kernel(int n, uint32_t *__restrict__ cnt, uint32_t *__restrict__ mask)
    {
      const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
      if (m >= (uint32_t)(n * n * n))
        return;
      const int ix = (int)compact3(m), iy = (int)compact3(m >> 1), iz = (int)compact3(m >> 2);
      uint32_t  msk = 0;
#pragma unroll
      for (int v = 0; v < 8; ++v)
        {
          int lv;
          if (first_touch_cell3(n, ix + (v & 1), iy + ((v >> 1) & 1), iz + (v >> 2), lv) == m)
            msk |= 1u << v;
        }
      mask[m] = msk;
      cnt[m]  = __popc(msk);
    }

    // single-block exclusive scan (n^3 <= 2^18 entries)
    __global__ void
    dofmap3_scan_kernel(int total, const uint32_t *__restrict__ cnt, uint32_t *__restrict__ base)
    {
      __shared__ uint32_t part[1024];
      const int           T     = blockDim.x;
      const int           chunk = (total + T - 1) / T;
      const int           lo    = min(total, (int)threadIdx.x * chunk), hi = min(total, lo + chunk);
      uint32_t            s = 0;
      for (int i = lo; i < hi; ++i)
        s += cnt[i];
      part[threadIdx.x] = s;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          uint32_t run = 0;
          for (int t = 0; t < T; ++t)
            {
              const uint32_t v = part[t];
              part[t]          = run;
              run += v;
            }
        }
      __syncthreads();
      uint32_t run = part[threadIdx.x];
      for (int i = lo; i < hi; ++i)
        {
          base[i] = run;
          run += cnt[i];
        }
    }

    __global__ void
    dofmap3_assign_kernel(int n, const uint32_t *__restrict__ base, const uint32_t *__restrict__ mask,
                          uint32_t *__restrict__ dofmap, uint32_t *__restrict__ invmap)
    {
      const int np  = n + 1;
      const int lex = blockIdx.x * blockDim.x + threadIdx.x;
      if (lex >= np * np * np)
        return;
      int jx, jy, jz, lv;
      decode3(lex, np, jx, jy, jz);
      const uint32_t m   = first_touch_cell3(n, jx, jy, jz, lv);
      const uint32_t dof = base[m] + __popc(mask[m] & ((1u << lv) - 1u));
      dofmap[lex]        = dof;
      invmap[dof]        = (uint32_t)lex;
    }

    // ========================================================================== assembly
    struct Coeff3
    {
      int    kind;
      double a0;
      double rot[9]; // MatrixCoeff<3> rotation (matrix_coeff.tpp:28-41), filled on the host
    };

    __device__ inline void
    coeff3_eval(const Coeff3 &c, double x, double y, double A[9])
    {
      if (c.kind == MSB_COEFF_REFERENCE)
        {
          // matrix_coeff.tpp:66-91 with the header's PI_D (sic) -- depends on x and y only
          const double PI_D = 3.14592653509793218403;
          const double a =
            1.0 * (1.0 - 0.9999 * (0.5 * sin(2 * PI_D * 57 * x) + 0.5 * sin(2 * PI_D * 57 * y)));
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              A[3 * i + j] = (c.rot[3 * i] * a) * c.rot[3 * j] + (c.rot[3 * i + 1] * a) * c.rot[3 * j + 1] +
                             (c.rot[3 * i + 2] * a) * c.rot[3 * j + 2];
          return;
        }
#pragma unroll
      for (int i = 0; i < 9; ++i)
        A[i] = 0.0;
      A[0] = A[4] = A[8] = c.a0;
    }

    // assemble_system (basis.tpp:159-242) for dim = 3, node-centric: every thread gathers the row
    // of its node from the <= 8 adjacent fine hexes (QGauss<3>(2), MappingQ1, FE_Q<3>(1)) and
    // stores the diagonal, the 13 forward couplings and the load entry.  No atomics: the sum
    // order is fixed (cells in local-vertex order, quadrature points in order).
    __global__ void __launch_bounds__(128)
    assemble3_kernel(int n, const double *__restrict__ corners, Coeff3 cf, const double *__restrict__ table,
                     double rhs_value, double *__restrict__ sten)
    {
      const int np = n + 1, N = np * np * np, cell = blockIdx.y;
      const int t  = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= N)
        return;
      const double *c = corners + 24 * (size_t)cell;
      double       *S = sten + (size_t)cell * NST * N;
      int           jx, jy, jz;
      decode3(t, np, jx, jy, jz);
      const double g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);
      double       acc[14], F = 0.0;
#pragma unroll
      for (int k = 0; k < 14; ++k)
        acc[k] = 0.0;
      // axis-aligned brick (the coarse meshes the drivers build): constant diagonal Jacobian
      bool aligned = true;
#pragma unroll
      for (int v = 0; v < 8; ++v)
        aligned = aligned && c[3 * v] == ((v & 1) ? c[3] : c[0]) && c[3 * v + 1] == (((v >> 1) & 1) ? c[7] : c[1]) &&
                  c[3 * v + 2] == ((v >> 2) ? c[14] : c[2]);
      const double rn = 1.0 / (double)n;
      const double hx = (c[3] - c[0]) * rn, hy = (c[7] - c[1]) * rn, hz = (c[14] - c[2]) * rn;
#pragma unroll 1
      for (int cv = 0; cv < 8; ++cv)
        {
          // this node is local vertex cv of fine cell (ix, iy, iz)
          const int cvx = cv & 1, cvy = (cv >> 1) & 1, cvz = cv >> 2;
          const int ix = jx - cvx, iy = jy - cvy, iz = jz - cvz;
          if (ix < 0 || iy < 0 || iz < 0 || ix >= n || iy >= n || iz >= n)
            continue;
          double P[8][3];
          if (!aligned)
            {
#pragma unroll
              for (int v = 0; v < 8; ++v)
                fine_vertex3(c, n, ix + (v & 1), iy + ((v >> 1) & 1), iz + (v >> 2), P[v]);
            }
          double row[8], fe = 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            row[j] = 0.0;
#pragma unroll 1
          for (int q = 0; q < 8; ++q)
            {
              const double xi = (q & 1) ? g1 : g0, eta = ((q >> 1) & 1) ? g1 : g0, ze = (q >> 2) ? g1 : g0;
              double       Nv[8], dN[8][3];
#pragma unroll
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
              double G[8][3], xq[3], JxW;
              if (aligned)
                {
                  xq[0] = c[0] + (ix + xi) * hx, xq[1] = c[1] + (iy + eta) * hy, xq[2] = c[2] + (iz + ze) * hz;
                  JxW   = hx * hy * hz * 0.125;
                  const double ihx = 1.0 / hx, ihy = 1.0 / hy, ihz = 1.0 / hz;
#pragma unroll
                  for (int v = 0; v < 8; ++v)
                    G[v][0] = dN[v][0] * ihx, G[v][1] = dN[v][1] * ihy, G[v][2] = dN[v][2] * ihz;
                }
              else
                {
                  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                  xq[0] = xq[1] = xq[2] = 0.0;
#pragma unroll
                  for (int v = 0; v < 8; ++v)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                      {
                        xq[a] += P[v][a] * Nv[v];
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                          J[a][b] += P[v][a] * dN[v][b];
                      }
                  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                                     J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
                  double Ji[3][3];
                  Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
                  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
                  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
                  Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
                  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
                  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
                  Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
                  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
                  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
                  JxW      = det * 0.125;
#pragma unroll
                  for (int v = 0; v < 8; ++v)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                      G[v][a] = Ji[0][a] * dN[v][0] + Ji[1][a] * dN[v][1] + Ji[2][a] * dN[v][2];
                }
              double A[9];
              if (table)
                {
                  // MSB_COEFF_TABLE: what TensorFunction<2,3>::value_list returned for this quadrature point
                  // (basis.tpp:202-203): [cell][(iz n + iy) n + ix][q][9]
                  const double *tp =
                    table + ((((size_t)cell * n + iz) * n + iy) * n + ix) * 72 + (size_t)q * 9;
#pragma unroll
                  for (int a = 0; a < 9; ++a)
                    A[a] = tp[a];
                }
              else
                coeff3_eval(cf, xq[0], xq[1], A);
              // own gradient and shape value, selected without dynamic register indexing
              double gi[3] = {0, 0, 0}, ni = 0.0;
#pragma unroll
              for (int v = 0; v < 8; ++v)
                if (v == cv)
                  {
                    gi[0] = G[v][0], gi[1] = G[v][1], gi[2] = G[v][2];
                    ni = Nv[v];
                  }
              double tt[3];
#pragma unroll
              for (int b = 0; b < 3; ++b)
                tt[b] = gi[0] * A[b] + gi[1] * A[3 + b] + gi[2] * A[6 + b];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                row[j] += (tt[0] * G[j][0] + tt[1] * G[j][1] + tt[2] * G[j][2]) * JxW;
              fe += ni * rhs_value * JxW;
            }
          // scatter the row: neighbour offset of local vertex j relative to this node
#pragma unroll
          for (int j = 0; j < 8; ++j)
            {
              const int e = 13 + ((j & 1) - cvx) + 3 * (((j >> 1) & 1) - cvy) + 9 * ((j >> 2) - cvz);
#pragma unroll
              for (int k = 0; k < 14; ++k)
                if (e - 13 == k)
                  acc[k] += row[j];
            }
          F += fe;
        }
#pragma unroll
      for (int k = 0; k < 14; ++k)
        S[(size_t)k * N + t] = acc[k];
      S[(size_t)ST3_F * N + t] = F;
    }

    // The same assembly when every coarse cell is an axis-aligned brick and A(x) = a(x, y) B with
    // a constant matrix B (MatrixCoeff<3>: B = R R^T; constant coefficient: B = I) -- what the
    // drivers build.  Then K_e(i, j) = sum_q a_q T[q][i][j] with geometry factors
    // T[q][i][j] = JxW grad N_i(q) . B grad N_j(q) that are the same for every fine cell, and the
    // separable sines of a(x, y) come from two 2n-entry tables: 64 FMAs per adjacent cell instead
    // of a trilinear Jacobian, its inverse and two sines per quadrature point.
    __global__ void __launch_bounds__(128)
    assemble3_brick_kernel(int n, const double *__restrict__ corners, Coeff3 cf, double rhs_value,
                           double *__restrict__ sten)
    {
      __shared__ double sT[512];            // [q][i][j]
      __shared__ double sax[128], say[128]; // 0.5 sin(2 PI_D 57 x_q) at 2*ix+qx, same in y
      const int     np = n + 1, N = np * np * np, cell = blockIdx.y;
      const int     t  = blockIdx.x * blockDim.x + threadIdx.x;
      const double *c  = corners + 24 * (size_t)cell;
      double       *S  = sten + (size_t)cell * NST * N;
      const double  rn = 1.0 / (double)n;
      const double  hx = (c[3] - c[0]) * rn, hy = (c[7] - c[1]) * rn, hz = (c[14] - c[2]) * rn;
      const double  g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);
      const double  JxW = hx * hy * hz * 0.125;
      const bool    ref = cf.kind == MSB_COEFF_REFERENCE;
      for (int idx = threadIdx.x; idx < 512; idx += blockDim.x)
        {
          const int    q = idx >> 6, i = (idx >> 3) & 7, j = idx & 7;
          const double xi = (q & 1) ? g1 : g0, eta = ((q >> 1) & 1) ? g1 : g0, ze = (q >> 2) ? g1 : g0;
          double       gi[3], gj[3];
          {
            const double fx = (i & 1) ? xi : 1 - xi, fy = ((i >> 1) & 1) ? eta : 1 - eta, fz = (i >> 2) ? ze : 1 - ze;
            gi[0] = ((i & 1) ? 1.0 : -1.0) * fy * fz / hx;
            gi[1] = fx * (((i >> 1) & 1) ? 1.0 : -1.0) * fz / hy;
            gi[2] = fx * fy * ((i >> 2) ? 1.0 : -1.0) / hz;
          }
          {
            const double fx = (j & 1) ? xi : 1 - xi, fy = ((j >> 1) & 1) ? eta : 1 - eta, fz = (j >> 2) ? ze : 1 - ze;
            gj[0] = ((j & 1) ? 1.0 : -1.0) * fy * fz / hx;
            gj[1] = fx * (((j >> 1) & 1) ? 1.0 : -1.0) * fz / hy;
            gj[2] = fx * fy * ((j >> 2) ? 1.0 : -1.0) / hz;
          }
          double v = 0.0;
#pragma unroll
          for (int a = 0; a < 3; ++a)
            {
              double tb = 0.0; // (gi . B)_a with B = rot rot^T or I
#pragma unroll
              for (int b = 0; b < 3; ++b)
                {
                  const double Bba = ref ? cf.rot[3 * b] * cf.rot[3 * a] + cf.rot[3 * b + 1] * cf.rot[3 * a + 1] +
                                             cf.rot[3 * b + 2] * cf.rot[3 * a + 2] :
                                           (a == b ? 1.0 : 0.0);
                  tb += gi[b] * Bba;
                }
              v += tb * gj[a];
            }
          sT[idx] = v * JxW;
        }
      if (ref)
        {
          const double PI_D = 3.14592653509793218403;
          for (int idx = threadIdx.x; idx < 2 * n; idx += blockDim.x)
            {
              const double g = (idx & 1) ? g1 : g0;
              sax[idx]       = 0.5 * sin(2 * PI_D * 57 * (c[0] + ((idx >> 1) + g) * hx));
              say[idx]       = 0.5 * sin(2 * PI_D * 57 * (c[1] + ((idx >> 1) + g) * hy));
            }
        }
      __syncthreads();
      if (t >= N)
        return;
      int jx, jy, jz;
      decode3(t, np, jx, jy, jz);
      double acc[14], F = 0.0;
#pragma unroll
      for (int k = 0; k < 14; ++k)
        acc[k] = 0.0;
#pragma unroll
      for (int cv = 0; cv < 8; ++cv)
        {
          const int cvx = cv & 1, cvy = (cv >> 1) & 1, cvz = cv >> 2;
          const int ix = jx - cvx, iy = jy - cvy, iz = jz - cvz;
          if (ix < 0 || iy < 0 || iz < 0 || ix >= n || iy >= n || iz >= n)
            continue;
          double aq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            aq[q] = ref ? 1.0 * (1.0 - 0.9999 * (sax[2 * ix + (q & 1)] + say[2 * iy + (q >> 1)])) : cf.a0;
          double row[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            row[j] = 0.0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              row[j] += aq[q & 3] * sT[q * 64 + cv * 8 + j];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            {
              const int e = ((j & 1) - cvx) + 3 * (((j >> 1) & 1) - cvy) + 9 * ((j >> 2) - cvz);
              if (e >= 0)
                acc[e] += row[j];
            }
          F += rhs_value * JxW;
        }
#pragma unroll
      for (int k = 0; k < 14; ++k)
        S[(size_t)k * N + t] = acc[k];
      S[(size_t)ST3_F * N + t] = F;
    }

    // ========================================================================== solver
    struct Levels3
    {
      int npl[MAXLEV + 1]; // nodes per direction of level l (0 = fine)
      int off[MAXLEV + 2]; // node offset of level l >= 1 in the packed coarse arrays
      int levels, cn;
    };

    struct Params3
    {
      int           n, nblk, chunk; // fine kernels: contiguous nodes per CTA (multiple of THREADS)
      int           by, zc, nys, nblk2; // K2 tiling: rows per y-slab, planes per z-chunk, slabs, CTAs
      const double *corners, *q1coef, *sten;
      double       *x, *r, *p, *q, *z; // [C][8][N]
      double       *v;                 // [C][8][cn]
      double       *dinv;              // [C][cn]
      double       *part;              // [C][8][PSTRIDE]
      double       *rzprev;            // [C][8]
      int32_t      *iters;
      double       *res;
      double        tol2;
      int           it;
      Levels3       L;
    };

    __device__ __forceinline__ double *
    part_ptr(double *part, int sidx, int parity, int which)
    {
      return part + (size_t)sidx * PSTRIDE + (parity * 3 + which) * MAXBLK;
    }

    // deterministic sum of the nblk (<= 64) partials of one quantity by ONE WARP: a fixed
    // butterfly, so every lane of every CTA of the cell obtains the same bits
    __device__ __forceinline__ double
    warp_sum_part(const double *part, int nblk)
    {
      const int lane = threadIdx.x & 31;
      double    v    = lane < nblk ? part[lane] : 0.0;
      if (lane + 32 < nblk)
        v += part[lane + 32];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, off);
      return v;
    }

    // done(solve) as seen by every CTA of a cell: recorded in an earlier launch (iters >= 0) or
    // implied by the r.r partials of parity `parity`.  Warp k of the CTA evaluates basis k.
    __device__ __forceinline__ bool
    cell_done(const Params3 &P, int cell, int parity, int *sdone /*shared[8]*/)
    {
      const int warp = threadIdx.x >> 5;
      if (warp < NB)
        {
          const int    sidx = cell * NB + warp;
          const double rr   = warp_sum_part(part_ptr(P.part, sidx, parity, 2), P.nblk);
          if ((threadIdx.x & 31) == 0)
            sdone[warp] = (P.iters[sidx] >= 0) || (rr <= P.tol2);
        }
      __syncthreads();
      int all = 1;
#pragma unroll
      for (int k = 0; k < NB; ++k)
        all &= sdone[k];
      return all != 0;
    }

    // exact Galerkin diagonal of level l at coarse node X: D = w^T K w with w the level-l
    // trilinear hat function sampled on the fine grid.  One CTA per (coarse node, cell).
    __global__ void __launch_bounds__(256)
    galerkin_diag3_kernel(Params3 P, int l)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y;
      const int npl = P.L.npl[l], nin = npl - 2;
      const int X = 1 + blockIdx.x % nin, Y = 1 + (blockIdx.x / nin) % nin, Z = 1 + blockIdx.x / (nin * nin);
      const double *S = P.sten + (size_t)cell * NST * N;
      const int     h = 1 << l, side = 2 * h - 1;
      const double  rh = 1.0 / (double)h;
      double        acc[1] = {0.0};
      for (int s = threadIdx.x; s < side * side * side; s += blockDim.x)
        {
          const int ax = s % side - (h - 1), ay = (s / side) % side - (h - 1), az = s / (side * side) - (h - 1);
          const double wi = (1.0 - abs(ax) * rh) * (1.0 - abs(ay) * rh) * (1.0 - abs(az) * rh);
          const int    t  = ((Z * h + az) * np + (Y * h + ay)) * np + (X * h + ax);
          // w^T K w = sum_i w_i (K_ii w_i + 2 sum_{forward j} K_ij w_j): only the stored couplings
          double kw = S[t] * wi;
#pragma unroll
          for (int f = 1; f <= 13; ++f)
            {
              const int e = 13 + f, bx = ax + e % 3 - 1, by = ay + (e / 3) % 3 - 1, bz = az + e / 9 - 1;
              if (abs(bx) >= h || abs(by) >= h || abs(bz) >= h)
                continue;
              const double wj = (1.0 - abs(bx) * rh) * (1.0 - abs(by) * rh) * (1.0 - abs(bz) * rh);
              kw = fma(2.0 * S[(size_t)f * N + t], wj, kw);
            }
          acc[0] = fma(wi, kw, acc[0]);
        }
      __shared__ double sbuf[8];
      block_sum_to<1>(acc, sbuf);
      if (threadIdx.x == 0)
        P.dinv[(size_t)cell * P.L.cn + P.L.off[l] + (Z * npl + Y) * npl + X] = 1.0 / acc[0];
    }

    // Initialisation in two passes.  (a) x_0 = g (BasisQ1 values) on EVERY node: the Dirichlet data on the boundary
    // and, inside, the initial guess of the iteration (exact for a constant coefficient; 24 -> 22 iterations on
    // 16^3 x 16^3 with the reference coefficient); r = p = 0.
    __global__ void __launch_bounds__(THREADS)
    init3a_kernel(Params3 P)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const double *c = P.corners + 24 * (size_t)cell, *q1 = P.q1coef + 64 * (size_t)cell;
      const int     t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          double xv[NB];
          {
            double p[3];
            fine_vertex3(c, n, jx, jy, jz, p);
#pragma unroll
            for (int k = 0; k < NB; ++k)
              xv[k] = basis_q1_value3(q1, k, p);
          }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              const size_t o = ((size_t)cell * NB + k) * N + t;
              P.x[o]         = xv[k];
              P.r[o]         = 0.0;
              P.p[o]         = 0.0;
            }
        }
    }

    // (b) r_0 = b - K_II g_I = -(K g) on the interior rows (b = -K_IB g_B: condense), reading g from x;
    //     partial r.r into parity 0
    __global__ void __launch_bounds__(THREADS)
    init3b_kernel(Params3 P)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const double *S = P.sten + (size_t)cell * NST * N;
      const double *X = P.x + (size_t)cell * NB * N;
      const int     t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      double        acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          double rv[NB];
          {
            const double kc = S[t];
#pragma unroll
            for (int k = 0; k < NB; ++k)
              rv[k] = -kc * X[(size_t)k * N + t];
          }
#pragma unroll
          for (int f = 1; f <= 13; ++f)
            {
              const int e = 13 + f, dz = e / 9 - 1, dy = (e / 3) % 3 - 1, dx = e % 3 - 1;
              const int o = (dz * np + dy) * np + dx;
              const double kf = S[(size_t)f * N + t], kb = S[(size_t)f * N + t - o];
#pragma unroll
              for (int k = 0; k < NB; ++k)
                {
                  rv[k] = fma(-kf, X[(size_t)k * N + t + o], rv[k]);
                  rv[k] = fma(-kb, X[(size_t)k * N + t - o], rv[k]);
                }
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              P.r[((size_t)cell * NB + k) * N + t] = rv[k];
              acc[k]                               = fma(rv[k], rv[k], acc[k]);
            }
        }
      __shared__ double sbuf[(THREADS / 32) * NB];
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            part_ptr(P.part, cell * NB + k, 0, 2)[blk] = acc[k];
        }
    }

    // K1: bookkeeping of the previous iteration, then p = z + beta p -- and the x update of the PREVIOUS iteration,
    // x += alpha_{it-1} p_{it-1}: K1 reads the old direction anyway, so K3 need not read p (nor read and write x) just for
    // that: 15 instead of 16 vector passes per iteration.  alpha_{it-1} is re-derived from the stored (r.z) and the same
    // p.q partials K3 used (the same bits).  It applies to the solves that RAN iteration it-1: not yet recorded as
    // converged before it (CTA 0 records `iters = it-1` in this very launch while the others read: both values mean "ran").
    // FLUSH: after the loop, only the x update of the last executed iteration (P.it = that iteration + 1).
    template <bool FLUSH>
    __global__ void __launch_bounds__(THREADS)
    k1_kernel(Params3 P)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const int par = (P.it - 1) & 1;
      __shared__ double sbeta[NB], sxa[NB];
      __shared__ int    sdone[NB];
      if ((threadIdx.x >> 5) < NB)
        {
          const int    k = threadIdx.x >> 5, sidx = cell * NB + k;
          const double rr = warp_sum_part(part_ptr(P.part, sidx, par, 2), P.nblk);
          const double rz = warp_sum_part(part_ptr(P.part, sidx, par, 0), P.nblk);
          double       pqp = 1.0;
          if (P.it >= 2)
            pqp = warp_sum_part(part_ptr(P.part, sidx, par, 1), P.nblk2);
          if ((threadIdx.x & 31) == 0)
            {
              const int itp = P.iters[sidx];
              sxa[k]        = (P.it >= 2 && (itp < 0 || itp == P.it - 1)) ? P.rzprev[sidx] / pqp : 0.0;
              const int dn  = (itp >= 0) || (rr <= P.tol2);
              sdone[k]      = dn;
              sbeta[k]      = P.it == 1 ? 0.0 : rz / P.rzprev[sidx];
              if (blk == 0 && dn && itp < 0)
                {
                  P.iters[sidx] = P.it - 1;
                  P.res[sidx]   = sqrt(rr);
                }
            }
        }
      __syncthreads();
      int all = 1, anyx = 0;
#pragma unroll
      for (int k = 0; k < NB; ++k)
        all &= sdone[k], anyx |= sxa[k] != 0.0;
      if ((all || FLUSH) && !anyx)
        return;
      const int t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          // all loads first (unconditional: memory-level parallelism), predicated stores after
          double pv[NB], zv[NB], xv[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              const size_t o = ((size_t)cell * NB + k) * N + t;
              pv[k] = P.p[o];
              zv[k] = (FLUSH || all) ? 0.0 : P.z[o];
              xv[k] = anyx ? P.x[o] : 0.0;
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              const size_t o = ((size_t)cell * NB + k) * N + t;
              if (sxa[k] != 0.0)
                P.x[o] = fma(sxa[k], pv[k], xv[k]);
              if (!FLUSH && !sdone[k])
                P.p[o] = fma(sbeta[k], pv[k], zv[k]);
            }
        }
    }

    // K2: q = K p on interior rows (p vanishes on the boundary), partial p.q
    __global__ void __launch_bounds__(THREADS)
    k2_kernel(Params3 P)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const int par = (P.it - 1) & 1;
      __shared__ int    sdone[NB];
      __shared__ double sbuf[(THREADS / 32) * NB];
      if (cell_done(P, cell, par, sdone))
        return;
      const double *S  = P.sten + (size_t)cell * NST * N;
      const int     t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      double        acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          double y[NB];
          {
            const double kc = S[t];
#pragma unroll
            for (int k = 0; k < NB; ++k)
              y[k] = sdone[k] ? 0.0 : kc * P.p[((size_t)cell * NB + k) * N + t];
          }
#pragma unroll
          for (int f = 1; f <= 13; ++f)
            {
              // forward neighbour t + o (coefficient stored here) and backward neighbour t - o
              // (coefficient stored there)
              const int    e = 13 + f, o = (e / 9 - 1) * np * np + ((e / 3) % 3 - 1) * np + (e % 3 - 1);
              const double kf = S[(size_t)f * N + t], kb = S[(size_t)f * N + t - o];
#pragma unroll
              for (int k = 0; k < NB; ++k)
                {
                  if (sdone[k])
                    continue;
                  const double *p = P.p + ((size_t)cell * NB + k) * N + t;
                  y[k]            = fma(kf, p[o], y[k]);
                  y[k]            = fma(kb, p[-o], y[k]);
                }
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              if (sdone[k])
                continue;
              const size_t o = ((size_t)cell * NB + k) * N + t;
              P.q[o]         = y[k];
              acc[k]         = fma(P.p[o], y[k], acc[k]);
            }
        }
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              part_ptr(P.part, cell * NB + k, P.it & 1, 1)[blk] = acc[k];
        }
    }

    // K2 (tiled): the same q = K p, with p staged through shared memory.  A CTA owns a slab of
    // `by` interior rows (all x) and marches over `zc` z-planes; four plane slots hold the 8
    // bases of planes z-1, z, z+1 and the one being prefetched, so every p value is read from
    // HBM/L2 (zc+2)/zc times instead of 27 and the 216 neighbour reads per node are conflict-free
    // LDS.  The bases are interleaved in PAIRS inside a plane slot ([pair][node][2]): the neighbour reads
    // are 108 LDS.128 instead of 216 LDS.64 -- the crossbar moves 128-bit accesses at twice the byte rate
    // (scripts/probes/onchip_peaks.cu), and this kernel sat at 70 % of the 64-bit wavefront peak
    // (profiles/r02e_ncu_3d_k2m_before.csv).  Coefficients (27 per node, shared by the 8 bases) stay in
    // global memory / L1.
    template <int MINB>
    __global__ void __launch_bounds__(THREADS, MINB)
    k2m_kernel(Params3 P)
    {
      extern __shared__ __align__(16) double sp[]; // [4][NB / 2][(by+2)*np][2]
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const int par = (P.it - 1) & 1;
      __shared__ int    sdone[NB];
      __shared__ double sbuf[(THREADS / 32) * NB];
      static_assert(NB % 2 == 0, "bases are staged in pairs");
      if (cell_done(P, cell, par, sdone))
        return;
      const int     ys = blk % P.nys, zi = blk / P.nys;
      const int     y0 = 1 + ys * P.by, y1 = min(n, y0 + P.by);
      const int     z0 = 1 + zi * P.zc, z1 = min(n, z0 + P.zc);
      const int     psz = (P.by + 2) * np, cnt = (y1 - y0 + 2) * np;
      const double *S  = P.sten + (size_t)cell * NST * N;
      const double *pg = P.p + (size_t)cell * NB * N;
      double       *qg = P.q + (size_t)cell * NB * N;

      // asynchronous global -> shared copies (cp.async, no register staging): the plane z+2 is in
      // flight while plane z is being computed
      auto load_plane = [&](int z) {
        const int src = (z * np + (y0 - 1)) * np;
#pragma unroll
        for (int k = 0; k < NB; ++k)
          {
            if (sdone[k])
              continue;
            const unsigned dst =
              (unsigned)__cvta_generic_to_shared(sp + (size_t)((z & 3) * NB + (k & ~1)) * psz + (k & 1));
            const double *sg = pg + (size_t)k * N + src;
            for (int i = threadIdx.x; i < cnt; i += THREADS)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 16u * i), "l"(sg + i) : "memory");
          }
      };

      const int  nx = n - 1, yy = threadIdx.x / nx, x = 1 + threadIdx.x % nx, y = y0 + yy;
      const bool active = y < y1 && threadIdx.x < nx * P.by;
      const int  sb = (yy + 1) * np + x; // index inside a plane slot
      double     acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;

      // Plane sweep: the thread of column (x, y) visits the planes s = z0-1 .. z1 ONCE each and scatters plane s into
      // three running sums -- q of its nodes in the planes s-1 (couplings with dz = +1: complete after this step), s
      // (dz = 0) and s+1 (dz = -1).  Each staged p value is then read from shared memory once per column instead of
      // three times (as dz = +1, 0, -1 of three different nodes): 36 LDS.128 per node instead of 108; the 27
      // coefficient loads per node are the same ones, regrouped.  (profiles/r02g_ncu_3d_k2m_592.csv: the gather form
      // sat at 65 % of the LSU wavefront peak.)
      double yA[NB], yB[NB], yC[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        yA[k] = yB[k] = yC[k] = 0.0;
      const double2 *s2 = reinterpret_cast<const double2 *>(sp); // [slot][pair][node]
      const int      NP2 = np * np;

      load_plane(z0 - 1);
      load_plane(z0);
      load_plane(z0 + 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      for (int s = z0 - 1; s <= z1; ++s)
        {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncthreads(); // planes up to s+1 have landed; everyone is done with plane s-2
          if (s + 2 <= z1)
            load_plane(s + 2);
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (!active)
            continue;
          const bool ownA = s - 1 >= z0, ownB = s >= z0 && s < z1, ownC = s + 1 < z1; // (uniform over the CTA)
          const int  ts = (s * np + y) * np + x;                                        // this column's node in plane s
          const double2 *ps = s2 + (size_t)(s & 3) * (NB / 2) * psz + sb;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
              {
                const int so = dy * np + dx;
                // the symmetric stencil stores the diagonal (array 0) and the 13 forward couplings e = 13 + f,
                // e = (dz+1) 9 + (dy+1) 3 + (dx+1); a backward coupling is the forward one of the neighbour
                constexpr int dummy = 0;
                (void)dummy;
                const int fA = 5 + (dy + 1) * 3 + (dx + 1);         // node (s-1) -> (dx, dy, +1): forward, stored at the node
                const int fC = 5 + (-dy + 1) * 3 + (-dx + 1);       // node (s+1) -> (dx, dy, -1): forward of the neighbour
                const int eB = 9 + (dy + 1) * 3 + (dx + 1);         // node (s)   -> (dx, dy, 0)
                const int fB = eB > 13 ? eB - 13 : (eB < 13 ? 13 - eB : 0); // |offset| index; (-dx,-dy,0) has e' = 26 - eB
                double kA = 0.0, kB = 0.0, kC = 0.0;
                if (ownA)
                  kA = S[(size_t)fA * N + ts - NP2];
                if (ownB)
                  kB = eB >= 13 ? S[(size_t)fB * N + ts] : S[(size_t)fB * N + ts + so];
                if (ownC)
                  kC = S[(size_t)fC * N + ts + so];
#pragma unroll
                for (int kp = 0; kp < NB / 2; ++kp)
                  {
                    const double2 v = ps[kp * psz + so];
                    yA[2 * kp]     = fma(kA, v.x, yA[2 * kp]);
                    yA[2 * kp + 1] = fma(kA, v.y, yA[2 * kp + 1]);
                    yB[2 * kp]     = fma(kB, v.x, yB[2 * kp]);
                    yB[2 * kp + 1] = fma(kB, v.y, yB[2 * kp + 1]);
                    yC[2 * kp]     = fma(kC, v.x, yC[2 * kp]);
                    yC[2 * kp + 1] = fma(kC, v.y, yC[2 * kp + 1]);
                  }
              }
          if (ownA)
            {
              // plane s-1 is complete: q and the p.q partial (its centre values are still staged in slot (s-1) & 3)
              const int      t  = ts - NP2;
              const double2 *pm = s2 + (size_t)((s - 1) & 3) * (NB / 2) * psz + sb;
#pragma unroll
              for (int kp = 0; kp < NB / 2; ++kp)
                {
                  const double2 pc = pm[kp * psz];
                  if (!sdone[2 * kp])
                    {
                      qg[(size_t)(2 * kp) * N + t] = yA[2 * kp];
                      acc[2 * kp]                  = fma(pc.x, yA[2 * kp], acc[2 * kp]);
                    }
                  if (!sdone[2 * kp + 1])
                    {
                      qg[(size_t)(2 * kp + 1) * N + t] = yA[2 * kp + 1];
                      acc[2 * kp + 1]                  = fma(pc.y, yA[2 * kp + 1], acc[2 * kp + 1]);
                    }
                }
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            yA[k] = yB[k], yB[k] = yC[k], yC[k] = 0.0;
        }
      __syncthreads();
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              part_ptr(P.part, cell * NB + k, P.it & 1, 1)[blk] = acc[k];
        }
    }

    // K2 for small local meshes (n <= 8): one CTA per cell stages the whole p of the 8 bases in
    // shared memory (same [basis][node] layout as in HBM) and sweeps all interior nodes, instead
    // of marching planes that would occupy only (n-1)^2 of the 256 threads.
    __global__ void __launch_bounds__(THREADS)
    k2w_kernel(Params3 P)
    {
      extern __shared__ double sp[]; // [NB][N]
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, nin = n - 1;
      const int par = (P.it - 1) & 1;
      __shared__ int    sdone[NB];
      __shared__ double sbuf[(THREADS / 32) * NB];
      if (cell_done(P, cell, par, sdone))
        return;
      const double *S  = P.sten + (size_t)cell * NST * N;
      const double *pg = P.p + (size_t)cell * NB * N;
      double       *qg = P.q + (size_t)cell * NB * N;
      for (int i = threadIdx.x; i < NB * N; i += THREADS)
        sp[i] = pg[i];
      __syncthreads();
      double acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int u = threadIdx.x; u < nin * nin * nin; u += THREADS)
        {
          const int     x = 1 + u % nin, y = 1 + (u / nin) % nin, z = 1 + u / (nin * nin);
          const int     t = (z * np + y) * np + x;
          const double *s0 = sp + t;
          double        yv[NB], pc[NB];
          {
            const double kc = S[t];
#pragma unroll
            for (int k = 0; k < NB; ++k)
              {
                pc[k] = s0[k * N];
                yv[k] = kc * pc[k];
              }
          }
#pragma unroll
          for (int f = 1; f <= 13; ++f)
            {
              const int    e = 13 + f, o = ((e / 9 - 1) * np + (e / 3) % 3 - 1) * np + e % 3 - 1;
              const double kf = S[(size_t)f * N + t], kb = S[(size_t)f * N + t - o];
#pragma unroll
              for (int k = 0; k < NB; ++k)
                {
                  yv[k] = fma(kf, s0[k * N + o], yv[k]);
                  yv[k] = fma(kb, s0[k * N - o], yv[k]);
                }
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              if (sdone[k])
                continue;
              qg[(size_t)k * N + t] = yv[k];
              acc[k]                = fma(pc[k], yv[k], acc[k]);
            }
        }
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              part_ptr(P.part, cell * NB + k, P.it & 1, 1)[0] = acc[k];
        }
    }

    // K3: alpha = rz/pq ; r -= alpha q ; partial r.r   (x += alpha p is deferred to K1 of the next iteration)
    __global__ void __launch_bounds__(THREADS)
    k3_kernel(Params3 P)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      const int par = (P.it - 1) & 1;
      __shared__ int    sdone[NB];
      __shared__ double salpha[NB];
      __shared__ double sbuf[(THREADS / 32) * NB];
      __shared__ double srz[NB];
      if ((threadIdx.x >> 5) < NB)
        {
          const int    k = threadIdx.x >> 5, sidx = cell * NB + k;
          const double rr = warp_sum_part(part_ptr(P.part, sidx, par, 2), P.nblk);
          const double rz = warp_sum_part(part_ptr(P.part, sidx, par, 0), P.nblk);
          const double pq = warp_sum_part(part_ptr(P.part, sidx, P.it & 1, 1), P.nblk2);
          if ((threadIdx.x & 31) == 0)
            {
              const int dn = (P.iters[sidx] >= 0) || (rr <= P.tol2);
              sdone[k]     = dn;
              srz[k]       = rz;
              salpha[k]    = dn ? 0.0 : rz / pq;
            }
        }
      __syncthreads();
      int all = 1;
#pragma unroll
      for (int k = 0; k < NB; ++k)
        all &= sdone[k];
      if (all)
        return;
      const int t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      double    acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          double qv[NB], rv[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              const size_t o = ((size_t)cell * NB + k) * N + t;
              qv[k] = P.q[o], rv[k] = P.r[o];
            }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              if (sdone[k])
                continue;
              const size_t o  = ((size_t)cell * NB + k) * N + t;
              const double a  = salpha[k]; // (x += a p: K1 of the next iteration / the flush launch)
              const double rn = fma(-a, qv[k], rv[k]);
              P.r[o]          = rn;
              acc[k]          = fma(rn, rn, acc[k]);
            }
        }
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              {
                const int sidx = cell * NB + k;
                part_ptr(P.part, sidx, P.it & 1, 2)[blk] = acc[k];
                if (blk == 0)
                  P.rzprev[sidx] = srz[k];
              }
        }
    }

    // restriction r_l = P^T r_{l-1}: full weighting over the 27 fine neighbours
    __global__ void __launch_bounds__(THREADS)
    restrict3_kernel(Params3 P, int l, int rpar)
    {
      const int cell = blockIdx.y, npl = P.L.npl[l], nin = npl - 2, npf = P.L.npl[l - 1];
      __shared__ int sdone[NB];
      if (cell_done(P, cell, rpar, sdone))
        return;
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= nin * nin * nin)
        return;
      const int    cx = 1 + t % nin, cy = 1 + (t / nin) % nin, cz = 1 + t / (nin * nin);
      const size_t Nf = (size_t)npf * npf * npf;
#pragma unroll 1
      for (int k = 0; k < NB; ++k)
        {
          if (sdone[k])
            continue;
          const double *src = l == 1 ? P.r + ((size_t)cell * NB + k) * Nf :
                                       P.v + ((size_t)cell * NB + k) * P.L.cn + P.L.off[l - 1];
          double pl[3];
#pragma unroll
          for (int az = -1; az <= 1; ++az)
            {
              double row[3];
#pragma unroll
              for (int ay = -1; ay <= 1; ++ay)
                {
                  const double *s = src + ((size_t)(2 * cz + az) * npf + (2 * cy + ay)) * npf + 2 * cx;
                  row[ay + 1]     = fma(0.5, s[-1] + s[1], s[0]);
                }
              pl[az + 1] = fma(0.5, row[0] + row[2], row[1]);
            }
          P.v[((size_t)cell * NB + k) * P.L.cn + P.L.off[l] + (cz * npl + cy) * npl + cx] =
            fma(0.5, pl[0] + pl[2], pl[1]);
        }
    }

    // trilinear interpolation of the next-coarser level at node (fx, fy, fz) of this level
    __device__ __forceinline__ double
    interp3(const double *__restrict__ vc, int npc, int fx, int fy, int fz)
    {
      const int xl = fx >> 1, xh = (fx + 1) >> 1, yl = fy >> 1, yh = (fy + 1) >> 1, zl = fz >> 1,
                zh = (fz + 1) >> 1;
      const double *a = vc + ((size_t)zl * npc + yl) * npc, *b = vc + ((size_t)zl * npc + yh) * npc;
      const double *c = vc + ((size_t)zh * npc + yl) * npc, *d = vc + ((size_t)zh * npc + yh) * npc;
      return 0.125 * (((a[xl] + a[xh]) + (b[xl] + b[xh])) + ((c[xl] + c[xh]) + (d[xl] + d[xh])));
    }

    // prolongation z_l = r_l / D_l + P z_{l+1}, in place (coarsest level: z = r / D)
    __global__ void __launch_bounds__(THREADS)
    prolong3_kernel(Params3 P, int l, int rpar)
    {
      const int cell = blockIdx.y, npl = P.L.npl[l], nin = npl - 2;
      __shared__ int sdone[NB];
      if (cell_done(P, cell, rpar, sdone))
        return;
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= nin * nin * nin)
        return;
      const int    fx = 1 + t % nin, fy = 1 + (t / nin) % nin, fz = 1 + t / (nin * nin);
      const int    i  = (fz * npl + fy) * npl + fx;
      const double di = P.dinv[(size_t)cell * P.L.cn + P.L.off[l] + i];
#pragma unroll 1
      for (int k = 0; k < NB; ++k)
        {
          if (sdone[k])
            continue;
          double *vl = P.v + ((size_t)cell * NB + k) * P.L.cn + P.L.off[l];
          double  v  = vl[i] * di;
          if (l < P.L.levels)
            v += interp3(P.v + ((size_t)cell * NB + k) * P.L.cn + P.L.off[l + 1], P.L.npl[l + 1], fx, fy, fz);
          vl[i] = v;
        }
    }

    // All coarse levels l >= l0 in ONE kernel, one CTA per cell, the level vectors of the eight
    // bases in shared memory: restrict level l0 from level l0-1 (global), down to the coarsest
    // level, z_l = r_l / D_l + P z_{l+1} back up, write level l0 for the finer prolongations.
    // FINE (only with l0 = 1): the CTA also applies the fine level z = r / D + P z_1 to its whole
    // cell straight from the shared-memory level-1 vector (the residual it restricted a moment ago
    // is still in L2) and produces the r.z partial, replacing fine3_kernel.
    template <bool FINE>
    __global__ void __launch_bounds__(THREADS)
    coarse_fused3_kernel(Params3 P, int l0, int rpar)
    {
      extern __shared__ double sv[]; // [(level l0..L)][8][npl^3], level offset 8 * (off[l] - off[l0])
      const int cell = blockIdx.x;
      __shared__ int sdone[NB];
      if (cell_done(P, cell, rpar, sdone))
        return;
      const int L = P.L.levels, tot = NB * (P.L.off[L + 1] - P.L.off[l0]);
      for (int i = threadIdx.x; i < tot; i += THREADS)
        sv[i] = 0.0;
      __syncthreads();
      // full weighting of one coarse node from a finer level
      auto fw = [](const double *src, int npf, int cx, int cy, int cz) {
        double pl[3];
#pragma unroll
        for (int az = -1; az <= 1; ++az)
          {
            double row[3];
#pragma unroll
            for (int ay = -1; ay <= 1; ++ay)
              {
                const double *q = src + ((size_t)(2 * cz + az) * npf + (2 * cy + ay)) * npf + 2 * cx;
                row[ay + 1]     = fma(0.5, q[-1] + q[1], q[0]);
              }
            pl[az + 1] = fma(0.5, row[0] + row[2], row[1]);
          }
        return fma(0.5, pl[0] + pl[2], pl[1]);
      };
      {
        const int    npl = P.L.npl[l0], nin = npl - 2, npf = P.L.npl[l0 - 1], n3 = npl * npl * npl;
        const size_t Nf  = (size_t)npf * npf * npf;
        for (int t = threadIdx.x; t < NB * nin * nin * nin; t += THREADS)
          {
            const int k = t / (nin * nin * nin), u = t % (nin * nin * nin);
            if (sdone[k])
              continue;
            const int     cx = 1 + u % nin, cy = 1 + (u / nin) % nin, cz = 1 + u / (nin * nin);
            const double *src = l0 == 1 ? P.r + ((size_t)cell * NB + k) * Nf :
                                          P.v + ((size_t)cell * NB + k) * P.L.cn + P.L.off[l0 - 1];
            sv[k * n3 + (cz * npl + cy) * npl + cx] = fw(src, npf, cx, cy, cz);
          }
      }
      __syncthreads();
      for (int l = l0 + 1; l <= L; ++l)
        {
          const int     npl = P.L.npl[l], nin = npl - 2, npf = P.L.npl[l - 1], n3 = npl * npl * npl;
          double       *dst = sv + NB * (P.L.off[l] - P.L.off[l0]);
          const double *srl = sv + NB * (P.L.off[l - 1] - P.L.off[l0]);
          for (int t = threadIdx.x; t < NB * nin * nin * nin; t += THREADS)
            {
              const int k = t / (nin * nin * nin), u = t % (nin * nin * nin);
              const int cx = 1 + u % nin, cy = 1 + (u / nin) % nin, cz = 1 + u / (nin * nin);
              dst[k * n3 + (cz * npl + cy) * npl + cx] = fw(srl + (size_t)k * npf * npf * npf, npf, cx, cy, cz);
            }
          __syncthreads();
        }
      for (int l = L; l >= l0; --l)
        {
          const int     npl = P.L.npl[l], nin = npl - 2, n3 = npl * npl * npl;
          double       *vl  = sv + NB * (P.L.off[l] - P.L.off[l0]);
          const double *di  = P.dinv + (size_t)cell * P.L.cn + P.L.off[l];
          for (int t = threadIdx.x; t < NB * nin * nin * nin; t += THREADS)
            {
              const int k = t / (nin * nin * nin), u = t % (nin * nin * nin);
              const int fx = 1 + u % nin, fy = 1 + (u / nin) % nin, fz = 1 + u / (nin * nin);
              const int i  = (fz * npl + fy) * npl + fx;
              double    v  = vl[k * n3 + i] * di[i];
              if (l < L)
                {
                  const int npc = P.L.npl[l + 1];
                  v += interp3(sv + NB * (P.L.off[l + 1] - P.L.off[l0]) + (size_t)k * npc * npc * npc, npc, fx, fy, fz);
                }
              vl[k * n3 + i] = v;
            }
          __syncthreads();
        }
      if (!FINE)
        {
          const int npl = P.L.npl[l0], nin = npl - 2, n3 = npl * npl * npl;
          for (int t = threadIdx.x; t < NB * nin * nin * nin; t += THREADS)
            {
              const int k = t / (nin * nin * nin), u = t % (nin * nin * nin);
              const int i = ((1 + u / (nin * nin)) * npl + 1 + (u / nin) % nin) * npl + 1 + u % nin;
              if (!sdone[k])
                P.v[((size_t)cell * NB + k) * P.L.cn + P.L.off[l0] + i] = sv[k * n3 + i];
            }
          return;
        }
      // fine level for the whole cell
      const int     n = P.n, np = n + 1, N = np * np * np, np1 = P.L.npl[1], n13 = np1 * np1 * np1;
      const double *KC = P.sten + (size_t)cell * NST * N;
      double        acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int t = threadIdx.x; t < N; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          const double kc = KC[t];
          double       rv[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k)
            rv[k] = P.r[((size_t)cell * NB + k) * N + t];
          const double dinv = 1.0 / kc;
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              if (sdone[k])
                continue;
              const double zv = fma(rv[k], dinv, interp3(sv + (size_t)k * n13, np1, jx, jy, jz));
              P.z[((size_t)cell * NB + k) * N + t] = zv;
              acc[k]                               = fma(rv[k], zv, acc[k]);
            }
        }
      __shared__ double sbuf[(THREADS / 32) * NB];
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x < NB && !sdone[threadIdx.x])
        {
          // the single r.z partial of this solve; the other slots of the parity are cleared
          double *pp = part_ptr(P.part, cell * NB + threadIdx.x, rpar, 0);
          for (int b = 1; b < P.nblk; ++b)
            pp[b] = 0.0;
        }
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              part_ptr(P.part, cell * NB + k, rpar, 0)[0] = acc[k];
        }
    }

    // fine level: z = r / D + P z_1 on interior rows; partial r.z into parity rpar
    __global__ void __launch_bounds__(THREADS)
    fine3_kernel(Params3 P, int rpar)
    {
      const int n = P.n, np = n + 1, N = np * np * np, cell = blockIdx.y, blk = blockIdx.x;
      __shared__ int    sdone[NB];
      __shared__ double sbuf[(THREADS / 32) * NB];
      if (cell_done(P, cell, rpar, sdone))
        return;
      const double *KC  = P.sten + (size_t)cell * NST * N;
      const int     np1 = P.L.npl[1];
      const int     t0 = blk * P.chunk, t1 = min(N, t0 + P.chunk);
      double        acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        acc[k] = 0.0;
      for (int t = t0 + threadIdx.x; t < t1; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          if (on_boundary3(jx, jy, jz, n))
            continue;
          const double kc = KC[t];
          double       rv[NB], cv[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k)
            rv[k] = P.r[((size_t)cell * NB + k) * N + t];
#pragma unroll
          for (int k = 0; k < NB; ++k)
            cv[k] = P.L.levels < 1 ? 0.0 :
                      interp3(P.v + ((size_t)cell * NB + k) * P.L.cn + P.L.off[1], np1, jx, jy, jz);
          const double dinv = 1.0 / kc;
#pragma unroll
          for (int k = 0; k < NB; ++k)
            {
              if (sdone[k])
                continue;
              const double zv = fma(rv[k], dinv, cv[k]);
              P.z[((size_t)cell * NB + k) * N + t] = zv;
              acc[k]                               = fma(rv[k], zv, acc[k]);
            }
        }
      block_sum_to<NB>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (!sdone[k])
              part_ptr(P.part, cell * NB + k, rpar, 0)[blk] = acc[k];
        }
    }

    // after the loop (P.it = last executed iteration): record every solve not yet recorded
    // (one warp per solve)
    __global__ void
    finalize3_kernel(Params3 P, int n_solves, int32_t *fail)
    {
      const int sidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      if (sidx >= n_solves || P.iters[sidx] >= 0)
        return;
      const double rr = warp_sum_part(part_ptr(P.part, sidx, P.it & 1, 2), P.nblk);
      if ((threadIdx.x & 31) != 0)
        return;
      P.iters[sidx] = P.it;
      P.res[sidx]   = sqrt(rr);
      if (!(rr <= P.tol2))
        atomicMin(fail, sidx);
    }

    // number of solves still running after iteration P.it (one warp per solve)
    __global__ void
    count3_kernel(Params3 P, int n_solves, int32_t *remaining)
    {
      const int sidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      if (sidx >= n_solves || P.iters[sidx] >= 0)
        return;
      const double rr = warp_sum_part(part_ptr(P.part, sidx, P.it & 1, 2), P.nblk);
      if ((threadIdx.x & 31) == 0 && !(rr <= P.tol2))
        atomicAdd(remaining, 1);
    }

    static Levels3
    make_levels(int l)
    {
      Levels3   L = {};
      const int n = 1 << l;
      L.levels    = l - 1;
      L.npl[0]    = n + 1;
      int off     = 0;
      for (int k = 1; k <= L.levels; ++k)
        {
          L.npl[k] = (n >> k) + 1;
          L.off[k] = off;
          off += L.npl[k] * L.npl[k] * L.npl[k];
        }
      L.off[L.levels + 1] = off;
      L.cn                = off;
      return L;
    }

    static Params3
    shifted(const Params3 &P, const Shard &s, int c0)
    {
      Params3      Q = P;
      const size_t N = (size_t)s.N;
      Q.corners += 24 * (size_t)c0, Q.q1coef += 64 * (size_t)c0, Q.sten += (size_t)c0 * NST * N;
      Q.x += (size_t)c0 * NB * N, Q.r += (size_t)c0 * NB * N, Q.p += (size_t)c0 * NB * N;
      Q.q += (size_t)c0 * NB * N, Q.z += (size_t)c0 * NB * N, Q.v += (size_t)c0 * NB * P.L.cn;
      Q.dinv += (size_t)c0 * P.L.cn, Q.part += (size_t)c0 * NB * PSTRIDE;
      Q.rzprev += NB * (size_t)c0, Q.iters += NB * (size_t)c0, Q.res += NB * (size_t)c0;
      return Q;
    }

    // ===================================================================== element matrices
    // assemble_global_element_matrix (basis.tpp:245-285) with dofs_per_cell = 8: M_ij =
    // phi_i . (K phi_j) over all N DoFs with the unconstrained K, b_i = phi_i . F.
    // One CTA per coarse cell, fixed summation order.
    __global__ void __launch_bounds__(THREADS)
    element_matrix3_kernel(int n, const double *__restrict__ sten, const double *__restrict__ phi,
                           double *__restrict__ M, double *__restrict__ b)
    {
      const int     np = n + 1, N = np * np * np, cell = blockIdx.x;
      const double *S = sten + (size_t)cell * NST * N;
      const double *Ph = phi + (size_t)cell * NB * N;
      double        acc[72];
#pragma unroll
      for (int k = 0; k < 72; ++k)
        acc[k] = 0.0;
      for (int t = threadIdx.x; t < N; t += THREADS)
        {
          int jx, jy, jz;
          decode3(t, np, jx, jy, jz);
          double kp[NB], pc[NB];
          {
            const double kc = S[t];
#pragma unroll
            for (int j = 0; j < NB; ++j)
              {
                pc[j] = Ph[(size_t)j * N + t];
                kp[j] = kc * pc[j];
              }
          }
          // forward neighbour t + o (coupling stored here) and backward neighbour t - o (stored
          // there), each only where that node exists
#pragma unroll
          for (int fw = 1; fw <= 13; ++fw)
            {
              const int  e = 13 + fw, dz = e / 9 - 1, dy = (e / 3) % 3 - 1, dx = e % 3 - 1;
              const int  o = (dz * np + dy) * np + dx;
              const bool hasf = (unsigned)(jx + dx) <= (unsigned)n && (unsigned)(jy + dy) <= (unsigned)n &&
                                (unsigned)(jz + dz) <= (unsigned)n;
              const bool hasb = (unsigned)(jx - dx) <= (unsigned)n && (unsigned)(jy - dy) <= (unsigned)n &&
                                (unsigned)(jz - dz) <= (unsigned)n;
              if (hasf)
                {
                  const double kf = S[(size_t)fw * N + t];
#pragma unroll
                  for (int j = 0; j < NB; ++j)
                    kp[j] = fma(kf, Ph[(size_t)j * N + t + o], kp[j]);
                }
              if (hasb)
                {
                  const double kb = S[(size_t)fw * N + t - o];
#pragma unroll
                  for (int j = 0; j < NB; ++j)
                    kp[j] = fma(kb, Ph[(size_t)j * N + t - o], kp[j]);
                }
            }
          const double f = S[(size_t)ST3_F * N + t];
#pragma unroll
          for (int i = 0; i < NB; ++i)
            {
#pragma unroll
              for (int j = 0; j < NB; ++j)
                acc[NB * i + j] = fma(pc[i], kp[j], acc[NB * i + j]);
              acc[64 + i] = fma(pc[i], f, acc[64 + i]);
            }
        }
      __shared__ double sbuf[(THREADS / 32) * 72];
      block_sum_to<72>(acc, sbuf);
      if (threadIdx.x == 0)
        {
#pragma unroll
          for (int k = 0; k < 64; ++k)
            M[64 * (size_t)cell + k] = acc[k];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            b[8 * (size_t)cell + k] = acc[64 + k];
        }
    }

    __global__ void
    apply_operator3_kernel(int n, const double *__restrict__ S, const double *__restrict__ x,
                           double *__restrict__ y)
    {
      const int np = n + 1, N = np * np * np, t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= N)
        return;
      int jx, jy, jz;
      decode3(t, np, jx, jy, jz);
      double v = 0.0;
      for (int e = 0; e < 27; ++e)
        {
          const int bx = jx + e % 3 - 1, by = jy + (e / 3) % 3 - 1, bz = jz + e / 9 - 1;
          if (bx < 0 || by < 0 || bz < 0 || bx > n || by > n || bz > n)
            continue;
          v = fma(sten3_get(S, N, np, t, e), x[t + off_of(e, np)], v);
        }
      y[t] = v;
    }

    // constraint set of one (cell, basis): boundary DoFs ascending + BasisQ1<3> values
    // (basis.tpp:119-135); position = number of boundary DoFs with a smaller index
    __global__ void
    constraints3_kernel(int n, const uint32_t *__restrict__ dofmap, const double *__restrict__ corners,
                        const double *__restrict__ q1coef, int ib, uint32_t *__restrict__ dofs,
                        double *__restrict__ vals)
    {
      const int np = n + 1, N = np * np * np, t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= N)
        return;
      int jx, jy, jz;
      decode3(t, np, jx, jy, jz);
      if (!on_boundary3(jx, jy, jz, n))
        return;
      const uint32_t d    = dofmap[t];
      int            rank = 0;
      for (int k = 0; k < N; ++k)
        {
          int kx, ky, kz;
          decode3(k, np, kx, ky, kz);
          if (on_boundary3(kx, ky, kz, n))
            rank += dofmap[k] < d;
        }
      double p[3];
      fine_vertex3(corners, n, jx, jy, jz, p);
      dofs[rank] = d;
      vals[rank] = basis_q1_value3(q1coef, ib, p);
    }
  } // namespace d3

  // ============================================================================ launchers
  size_t
  dim3_coarse_nodes(int l)
  {
    const size_t cn = (size_t)d3::make_levels(l).cn;
    return cn ? cn : 1;
  }

  int
  dim3_part_stride()
  {
    return d3::PSTRIDE;
  }

  cudaError_t
  launch_dofmap3(const Shard &s, cudaStream_t st)
  {
    const int   ncell = s.n * s.n * s.n;
    uint32_t   *tmp   = nullptr;
    cudaError_t e     = cudaMalloc(&tmp, sizeof(uint32_t) * 3 * (size_t)ncell);
    if (e != cudaSuccess)
      return e;
    uint32_t *cnt = tmp, *mask = tmp + ncell, *base = tmp + 2 * (size_t)ncell;
    d3::dofmap3_count_kernel<<<(ncell + 255) / 256, 256, 0, st>>>(s.n, cnt, mask);
    d3::dofmap3_scan_kernel<<<1, 1024, 0, st>>>(ncell, cnt, base);
    d3::dofmap3_assign_kernel<<<(s.N + 255) / 256, 256, 0, st>>>(s.n, base, mask, s.d_dofmap, s.d_invmap);
    e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    return e != cudaSuccess ? e : cudaGetLastError();
  }

  cudaError_t
  launch_assemble3(const Shard &s, cudaStream_t st, int *n_launches)
  {
    d3::Coeff3 cf;
    cf.kind = s.coeff.kind;
    cf.a0   = s.coeff.par[0];
    // MatrixCoeff<3> rotation (matrix_coeff.tpp:28-41) with alpha = PI_D/3, beta = PI_D/6,
    // gamma = PI_D/4 (matrix_coeff.hpp:45-48)
    const double PI_D = 3.14592653509793218403;
    const double al = PI_D / 3, be = PI_D / 6, ga = PI_D / 4;
    cf.rot[0] = cos(al) * cos(ga) - sin(al) * cos(be) * sin(ga);
    cf.rot[1] = -cos(al) * sin(ga) - sin(al) * cos(be) * cos(ga);
    cf.rot[2] = sin(al) * sin(be);
    cf.rot[3] = sin(al) * cos(ga) + cos(al) * cos(be) * sin(ga);
    cf.rot[4] = -sin(al) * sin(ga) + cos(al) * cos(be) * cos(ga);
    cf.rot[5] = -cos(al) * sin(be);
    cf.rot[6] = sin(be) * sin(ga);
    cf.rot[7] = sin(be) * cos(ga);
    cf.rot[8] = cos(be);
    for (int c0 = 0; c0 < s.n_cells; c0 += 65535)
      {
        const int nc = s.n_cells - c0 < 65535 ? s.n_cells - c0 : 65535;
        const bool tab = s.coeff.kind == MSB_COEFF_TABLE;
        if (s.bricks && s.variant != 3 && !tab)
          d3::assemble3_brick_kernel<<<dim3((s.N + 127) / 128, nc), 128, 0, st>>>(
            s.n, s.d_corners + 24 * (size_t)c0, cf, s.rhs_value, s.d_sten + (size_t)c0 * ST3_NARR * s.N);
        else
          d3::assemble3_kernel<<<dim3((s.N + 127) / 128, nc), 128, 0, st>>>(
            s.n, s.d_corners + 24 * (size_t)c0, cf,
            tab ? s.d_table + (size_t)c0 * (size_t)s.n * s.n * s.n * 72 : nullptr, s.rhs_value,
            s.d_sten + (size_t)c0 * ST3_NARR * s.N);
        ++*n_launches;
      }
    return cudaGetLastError();
  }

  cudaError_t
  launch_solve3(Shard &s, double tol, int max_iter, cudaStream_t st, int *n_launches)
  {
    using namespace d3;
    Params3 P;
    P.n        = s.n;
    // contiguous node ranges, a multiple of the CTA size so that no pass runs half empty
    // and sized so that a shard fills the GPU a few times over without making the CTAs tiny
    int want = (6000 + s.n_cells - 1) / s.n_cells;
    want     = want < 1 ? 1 : (want > MAXBLK ? MAXBLK : want);
    P.chunk  = THREADS * ((s.N + THREADS * want - 1) / (THREADS * want));
    P.nblk   = (s.N + P.chunk - 1) / P.chunk;
    // K2 tiling: slabs of `by` interior rows x all columns (about one node per thread), z-chunks
    // of `zc` planes; nys * nzc <= MAXBLK partial-sum slots
    const int nin = s.n - 1;
    P.by          = nin < THREADS / nin ? nin : (THREADS / nin > 0 ? THREADS / nin : 1);
    P.nys         = (nin + P.by - 1) / P.by;
    {
      const int max_zc = MAXBLK / P.nys > 0 ? MAXBLK / P.nys : 1;
      P.zc             = (nin + max_zc - 1) / max_zc;
      if (P.zc < 8)
        P.zc = 8;
      if (nin < 16 && s.variant != 7) // a whole column per CTA when it is short: no halo planes inside the cell (variant 7: chunks of 8)
        P.zc = nin;
    }
    P.nblk2 = P.nys * ((nin + P.zc - 1) / P.zc);
    const size_t k2_smem = sizeof(double) * 4 * NB * (size_t)(P.by + 2) * s.np;
    const bool   k2_tiled = s.variant != 1 && nin * P.by <= THREADS && P.nblk2 <= MAXBLK && k2_smem <= 200 * 1024;
    if (k2_tiled)
      {
        cudaError_t ea = cudaFuncSetAttribute(k2m_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2_smem);
        if (ea == cudaSuccess)
          ea = cudaFuncSetAttribute(k2m_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2_smem);
        if (ea != cudaSuccess)
          return ea;
      }
    else
      P.nblk2 = P.nblk;
    // small local meshes: whole-cell K2, one CTA per cell
    const size_t k2w_smem  = sizeof(double) * NB * (size_t)s.N;
    const bool   k2_whole  = s.variant != 1 && s.variant != 6 && s.n <= 8 && k2w_smem <= 64 * 1024;
    if (k2_whole)
      {
        P.nblk2 = 1;
        cudaError_t ea = cudaFuncSetAttribute(k2w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2w_smem);
        if (ea != cudaSuccess)
          return ea;
      }
    P.corners = s.d_corners;
    P.q1coef  = s.d_q1coef;
    P.sten    = s.d_sten;
    P.x       = s.d_phi;
    P.r       = s.d_wr;
    P.p       = s.d_wp;
    P.q       = s.d_wq;
    P.z       = s.d_wz;
    P.v       = s.d_wv;
    P.dinv    = s.d_dinv;
    P.part    = s.d_part;
    P.rzprev  = s.d_scal;
    P.iters   = s.d_iters;
    P.res     = s.d_res;
    P.tol2    = tol * tol;
    P.it      = 0;
    P.L       = make_levels(s.l);
    const Levels3 &L        = P.L;
    const int      n_solves = NB * s.n_cells;
    const int      C        = s.n_cells;
    cudaError_t    e;

#define TRY(call)                  \
  if ((e = (call)) != cudaSuccess) \
  return e

    TRY(cudaMemsetAsync(s.d_iters, 0xff, sizeof(int32_t) * n_solves, st));
    TRY(cudaMemsetAsync(s.d_part, 0, sizeof(double) * (size_t)n_solves * PSTRIDE, st));
    TRY(cudaMemsetAsync(s.d_wv, 0, sizeof(double) * (size_t)n_solves * dim3_coarse_nodes(s.l), st));
    TRY(cudaMemsetAsync(s.d_dinv, 0, sizeof(double) * (size_t)C * dim3_coarse_nodes(s.l), st));

    auto for_slices = [&](auto &&launch) {
      for (int c0 = 0; c0 < C; c0 += 65535)
        {
          const int nc = C - c0 < 65535 ? C - c0 : 65535;
          launch(shifted(P, s, c0), nc);
          ++*n_launches;
        }
    };
    for (int l = 1; l <= L.levels; ++l)
      {
        const int nin = L.npl[l] - 2;
        for_slices([&](const Params3 &Q, int nc) {
          const int side = (2 << l) - 1;
          galerkin_diag3_kernel<<<dim3(nin * nin * nin, nc), side * side * side >= 256 ? 256 : 64, 0, st>>>(Q, l);
        });
      }
    // coarse levels l >= l0 run fused in one kernel when their vectors fit 64 KB of shared memory
    int l0 = L.levels + 1;
    while (l0 > 1 && sizeof(double) * NB * (size_t)(L.off[L.levels + 1] - L.off[l0 - 1]) <= 64 * 1024)
      --l0;
    if (s.variant == 4)
      l0 = L.levels + 1; // unfused (comparison)
    const size_t fused_smem = l0 <= L.levels ? sizeof(double) * NB * (size_t)(L.off[L.levels + 1] - L.off[l0]) : 0;
    // with every coarse level fused and enough cells to fill the GPU, the fused CTA also does the
    // fine level of its cell
    const bool fused_fine = fused_smem && l0 == 1 && C >= 296 && s.variant != 5;
    if (fused_smem)
      {
        TRY(cudaFuncSetAttribute(coarse_fused3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fused_smem));
        TRY(cudaFuncSetAttribute(coarse_fused3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fused_smem));
      }
    auto precondition = [&](int rpar) {
      for (int l = 1; l < l0 && l <= L.levels; ++l)
        {
          const int nin = L.npl[l] - 2, tot = nin * nin * nin;
          for_slices([&](const Params3 &Q, int nc) {
            restrict3_kernel<<<dim3((tot + THREADS - 1) / THREADS, nc), THREADS, 0, st>>>(Q, l, rpar);
          });
        }
      if (fused_smem)
        for (int c0 = 0; c0 < C; c0 += 65535)
          {
            const int nc = C - c0 < 65535 ? C - c0 : 65535;
            if (fused_fine)
              coarse_fused3_kernel<true><<<nc, THREADS, fused_smem, st>>>(shifted(P, s, c0), l0, rpar);
            else
              coarse_fused3_kernel<false><<<nc, THREADS, fused_smem, st>>>(shifted(P, s, c0), l0, rpar);
            ++*n_launches;
          }
      if (fused_fine)
        return;
      for (int l = (l0 <= L.levels ? l0 - 1 : L.levels); l >= 1; --l)
        {
          const int nin = L.npl[l] - 2, tot = nin * nin * nin;
          for_slices([&](const Params3 &Q, int nc) {
            prolong3_kernel<<<dim3((tot + THREADS - 1) / THREADS, nc), THREADS, 0, st>>>(Q, l, rpar);
          });
        }
      for_slices([&](const Params3 &Q, int nc) {
        fine3_kernel<<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q, rpar);
      });
    };

    for_slices([&](const Params3 &Q, int nc) { init3a_kernel<<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
    for_slices([&](const Params3 &Q, int nc) { init3b_kernel<<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
    precondition(0);

    int32_t   h_remaining = 1;
    int       it          = 0;
    const int check_every = 4;
    while (it < max_iter)
      {
        if (it % check_every == 0)
          {
            P.it = it;
            TRY(cudaMemsetAsync(s.d_flags, 0, sizeof(int32_t), st));
            count3_kernel<<<(n_solves + 7) / 8, 256, 0, st>>>(P, n_solves, s.d_flags);
            ++*n_launches;
            TRY(cudaMemcpyAsync(&h_remaining, s.d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            TRY(cudaStreamSynchronize(st));
            if (h_remaining == 0)
              break;
          }
        ++it;
        P.it = it;
        for_slices([&](const Params3 &Q, int nc) { k1_kernel<false><<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
        if (k2_whole)
          for_slices([&](const Params3 &Q, int nc) { k2w_kernel<<<dim3(1, nc), THREADS, k2w_smem, st>>>(Q); });
        else if (k2_tiled)
          for_slices([&](const Params3 &Q, int nc) {
            if (s.variant == 2)
              k2m_kernel<3><<<dim3(P.nblk2, nc), THREADS, k2_smem, st>>>(Q);
            else
              k2m_kernel<2><<<dim3(P.nblk2, nc), THREADS, k2_smem, st>>>(Q);
          });
        else
          for_slices([&](const Params3 &Q, int nc) { k2_kernel<<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
        for_slices([&](const Params3 &Q, int nc) { k3_kernel<<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
        precondition(it & 1);
      }
    if (it >= 1)
      {
        // the deferred x update of the last executed iteration
        P.it = it + 1;
        for_slices([&](const Params3 &Q, int nc) { k1_kernel<true><<<dim3(P.nblk, nc), THREADS, 0, st>>>(Q); });
      }
    P.it = it;
    finalize3_kernel<<<(n_solves + 7) / 8, 256, 0, st>>>(P, n_solves, s.d_fail);
    ++*n_launches;
#undef TRY
    return cudaGetLastError();
  }

  cudaError_t
  launch_element_matrices3(const Shard &s, cudaStream_t st, int *n_launches)
  {
    d3::element_matrix3_kernel<<<s.n_cells, d3::THREADS, 0, st>>>(s.n, s.d_sten, s.d_phi, s.d_M, s.d_b);
    ++*n_launches;
    return cudaGetLastError();
  }

  cudaError_t
  launch_apply_operator3(const Shard &s, int cell, const double *d_x, double *d_y, cudaStream_t st)
  {
    d3::apply_operator3_kernel<<<(s.N + 255) / 256, 256, 0, st>>>(
      s.n, s.d_sten + (size_t)cell * ST3_NARR * s.N, d_x, d_y);
    return cudaGetLastError();
  }

  cudaError_t
  launch_constraints3(const Shard &s, int cell, int ib, uint32_t *d_dofs, double *d_vals, cudaStream_t st)
  {
    d3::constraints3_kernel<<<(s.N + 127) / 128, 128, 0, st>>>(
      s.n, s.d_dofmap, s.d_corners + 24 * (size_t)cell, s.d_q1coef + 64 * (size_t)cell, ib, d_dofs, d_vals);
    return cudaGetLastError();
  }
} // namespace msb
