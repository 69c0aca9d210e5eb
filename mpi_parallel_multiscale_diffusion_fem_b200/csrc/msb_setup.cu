// msb_setup.cu -- everything around the solve: DoF map, stencil assembly, the
// coarse element matrices, operator application, reordering, reconstruction.
//
// Reference lines restated here (never copied):
//   DoF numbering          diffusion_problem_basis.tpp:106   
//   assemble_system        diffusion_problem_basis.tpp:159-242
//   MatrixCoeff            include/coefficients/matrix_coeff.tpp:17-25,66-91
//   constraints            diffusion_problem_basis.tpp:119-135
//   element matrix / rhs   diffusion_problem_basis.tpp:245-285
//   set_global_weights     diffusion_problem_basis.tpp:352-377
#include <math.h>

#include "msb_internal.cuh"
#include "msb_coeff.cuh"

namespace msb
{
  // ======================================================================================
  // DoF map: closed form of deal.II's first-touch numbering.  A vertex is first touched by
  // the adjacent fine cell with the smallest Morton index; its DoF index is the number of
  // vertices first touched earlier = (exclusive scan over cells of #new vertices) + rank
  // inside its cell.
  // ======================================================================================
  __device__ inline uint32_t
  first_touch_cell(int n, int jx, int jy, int &lv)
  {
    uint32_t best = 0xffffffffu;
    lv            = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        // cell for which (jx,jy) is local vertex v
        const int ix = jx - (v & 1), iy = jy - (v >> 1);
        if (ix < 0 || iy < 0 || ix >= n || iy >= n)
          continue;
        const uint32_t m = morton_encode((uint32_t)ix, (uint32_t)iy);
        if (m < best)
          {
            best = m;
            lv   = v;
          }
      }
    return best;
  }

  __global__ void
  dofmap_count_kernel(int n, uint32_t *__restrict__ cnt, uint32_t *__restrict__ mask)
  {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (uint32_t)(n * n))
      return;
    const int ix = (int)morton_compact(m), iy = (int)morton_compact(m >> 1);
    uint32_t  msk = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        int lv;
        if (first_touch_cell(n, ix + (v & 1), iy + (v >> 1), lv) == m)
          msk |= 1u << v;
      }
    mask[m] = msk;
    cnt[m]  = __popc(msk);
  }

  // single-block exclusive scan (n*n <= 2^18 entries)
  __global__ void
  dofmap_scan_kernel(int total, const uint32_t *__restrict__ cnt, uint32_t *__restrict__ base)
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
  dofmap_assign_kernel(int n, const uint32_t *__restrict__ base, const uint32_t *__restrict__ mask,
                       uint32_t *__restrict__ dofmap, uint32_t *__restrict__ invmap)
  {
    const int np  = n + 1;
    const int lex = blockIdx.x * blockDim.x + threadIdx.x;
    if (lex >= np * np)
      return;
    const int      jx = lex % np, jy = lex / np;
    int            lv;
    const uint32_t m   = first_touch_cell(n, jx, jy, lv);
    const uint32_t dof = base[m] + __popc(mask[m] & ((1u << lv) - 1u));
    dofmap[lex]        = dof;
    invmap[dof]        = (uint32_t)lex;
  }

  cudaError_t
  launch_dofmap(const Shard &s, cudaStream_t st)
  {
    const int  ncell = s.n * s.n;
    uint32_t  *tmp   = nullptr;
    cudaError_t e    = cudaMalloc(&tmp, sizeof(uint32_t) * 3 * (size_t)ncell);
    if (e != cudaSuccess)
      return e;
    uint32_t *cnt = tmp, *mask = tmp + ncell, *base = tmp + 2 * (size_t)ncell;
    dofmap_count_kernel<<<(ncell + 255) / 256, 256, 0, st>>>(s.n, cnt, mask);
    dofmap_scan_kernel<<<1, 1024, 0, st>>>(ncell, cnt, base);
    dofmap_assign_kernel<<<(s.N + 255) / 256, 256, 0, st>>>(s.n, base, mask, s.d_dofmap, s.d_invmap);
    e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    return e != cudaSuccess ? e : cudaGetLastError();
  }

  // ======================================================================================
  // BasisQ1<dim> coefficient matrices (basis_q1.tpp:26-47 for dim 2, :50-75 for dim 3) of all
  // cells: inverse of the point matrix [1, x, y, xy] resp. [1, x, y, z, xy, yz, xz, xyz] at the
  // vertices, by Gauss-Jordan with partial pivoting, one thread per coarse cell.  A singular
  // point matrix (degenerate cell) is reported through `bad` (smallest cell index).
  // ======================================================================================
  template <int DIM>
  __global__ void __launch_bounds__(128)
  basis_q1_kernel(int n_cells, const double *__restrict__ corners, double *__restrict__ q1coef,
                  int32_t *__restrict__ bad)
  {
    constexpr int NB   = 1 << DIM;
    const int     cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells)
      return;
    const double *c = corners + (size_t)cell * NB * DIM;
    double        a[NB][2 * NB];
#pragma unroll
    for (int i = 0; i < NB; ++i)
      {
        const double x = c[DIM * i], y = c[DIM * i + 1];
        a[i][0] = 1.0, a[i][1] = x, a[i][2] = y;
        if constexpr (DIM == 2)
          a[i][3] = x * y;
        else
          {
            const double z = c[DIM * i + 2];
            a[i][3] = z, a[i][4] = x * y, a[i][5] = y * z, a[i][6] = x * z, a[i][7] = x * y * z;
          }
#pragma unroll
        for (int j = 0; j < NB; ++j)
          a[i][NB + j] = i == j ? 1.0 : 0.0;
      }
    for (int col = 0; col < NB; ++col)
      {
        int piv = col;
        for (int r = col + 1; r < NB; ++r)
          if (fabs(a[r][col]) > fabs(a[piv][col]))
            piv = r;
        if (a[piv][col] == 0.0)
          {
            atomicMin(bad, cell);
            return;
          }
        if (piv != col)
          for (int j = 0; j < 2 * NB; ++j)
            {
              const double t = a[col][j];
              a[col][j]      = a[piv][j];
              a[piv][j]      = t;
            }
        const double inv = 1.0 / a[col][col];
        for (int j = 0; j < 2 * NB; ++j)
          a[col][j] *= inv;
        for (int r = 0; r < NB; ++r)
          if (r != col)
            {
              const double f = a[r][col];
              if (f != 0.0)
                for (int j = 0; j < 2 * NB; ++j)
                  a[r][j] -= f * a[col][j];
            }
      }
    double *o = q1coef + (size_t)cell * NB * NB;
    for (int i = 0; i < NB; ++i)
      for (int j = 0; j < NB; ++j)
        o[NB * i + j] = a[i][NB + j];
  }

  // d_fail[1] is used as the "degenerate cell" flag; *h_bad receives INT_MAX or the cell index
  cudaError_t
  launch_basis_q1(const Shard &s, cudaStream_t st, int32_t *h_bad)
  {
    const int32_t init = 0x7fffffff;
    cudaError_t   e    = cudaMemcpyAsync(s.d_fail + 1, &init, sizeof init, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess)
      return e;
    const int grid = (s.n_cells + 127) / 128;
    if (s.dim == 2)
      basis_q1_kernel<2><<<grid, 128, 0, st>>>(s.n_cells, s.d_corners, s.d_q1coef, s.d_fail + 1);
    else
      basis_q1_kernel<3><<<grid, 128, 0, st>>>(s.n_cells, s.d_corners, s.d_q1coef, s.d_fail + 1);
    if ((e = cudaGetLastError()) != cudaSuccess)
      return e;
    e = cudaMemcpyAsync(h_bad, s.d_fail + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess)
      return e;
    return cudaStreamSynchronize(st);
  }

  // ======================================================================================
  // Stencil assembly.  One CTA handles a strip of node rows of one coarse cell: it first
  // computes the 10 unique entries of every fine element matrix K_e (2x2 Gauss, Q1 mapping,
  // full tensor coefficient) plus the element load F_e for the R+1 rows of fine cells the
  // strip touches into shared memory, then gathers them per node into the symmetric
  // 9-point stencil.  The coefficient is evaluated once per quadrature point (+1/R halo).
  // ======================================================================================
  struct AssembleParams
  {
    int           n, R, n_cells;
    const double *corners;
    const double *table;
    double       *sten;
    double        rhs_value;
    CoeffEval     coef;
  };

  __global__ void __launch_bounds__(256)
  assemble_kernel(AssembleParams P)
  {
    extern __shared__ double ke[]; // [14][(R+1)*n] + sine tables [4n] + [4(R+1)]
    const int n = P.n, np = n + 1, N = np * np, R = P.R;
    const int cell  = blockIdx.y;
    const int jy0   = blockIdx.x * R; // first node row of the strip
    const int rows  = min(R, np - jy0);
    const int slab  = (R + 1) * n;
    const double *c = P.corners + 8 * (size_t)cell;

    const double g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);

    // ---- phase 0: on an axis-aligned coarse cell (the only kind the reference's refined
    // hyper_cube produces) a separable coefficient needs its sines only once per fine column
    // and per fine row of the strip: 4n + 4(R+1) evaluations instead of 8 per fine cell.  The
    // quadrature abscissae are computed by the same expressions as in the general path.
    double    *tsx = ke + 14 * slab, *tsy = tsx + 4 * n;
    const bool aligned = c[0] == c[4] && c[2] == c[6] && c[1] == c[3] && c[5] == c[7];
    const bool fast    = aligned && !P.table && P.coef.separable();
    if (fast)
      {
        for (int t = threadIdx.x; t < 4 * n + 4 * (rows + 1); t += blockDim.x)
          {
            const bool isx = t < 4 * n;
            const int  u = isx ? t : t - 4 * n, k = u >> 2, q = u & 3;
            const int  ix = isx ? k : 0, iy = isx ? 0 : jy0 - 1 + k;
            double     v = 0.0;
            if (isx || (iy >= 0 && iy < n))
              {
                const double xi = (q & 1) ? g1 : g0, eta = (q >> 1) ? g1 : g0;
                const double Nv[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
                double       xq = 0, yq = 0;
#pragma unroll
                for (int vv = 0; vv < 4; ++vv)
                  {
                    double px, py;
                    fine_vertex(c, n, ix + (vv & 1), iy + (vv >> 1), px, py);
                    xq += px * Nv[vv];
                    yq += py * Nv[vv];
                  }
                v = P.coef.sine_term(isx ? xq : yq);
              }
            (isx ? tsx : tsy)[u] = v;
          }
        __syncthreads();
      }

    // ---- phase 1: element matrices of cell rows jy0-1 .. jy0+rows-1
    for (int t = threadIdx.x; t < (rows + 1) * n; t += blockDim.x)
      {
        const int lr = t / n, ix = t % n;
        const int iy = jy0 - 1 + lr;
        double    K[10], Fe[4];
#pragma unroll
        for (int e = 0; e < 10; ++e)
          K[e] = 0.0;
        Fe[0] = Fe[1] = Fe[2] = Fe[3] = 0.0;
        if (iy >= 0 && iy < n && aligned)
          {
            // Axis-aligned coarse cell: the fine cells are hx x hy rectangles, the Jacobian is
            // diag(hx, hy) and  K_ij = sum_q [dNx_i dNx_j a00 hy/hx + (dNx_i dNy_j + dNy_i dNx_j) a_s
            //                               + dNy_i dNy_j a11 hx/hy] / 4
            // with the reference-cell gradients at the Gauss points folded at compile time.
            const double hx = (c[2] - c[0]) / n, hy = (c[5] - c[1]) / n;
            const double rxx = 0.25 * hy / hx, ryy = 0.25 * hx / hy;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              {
                constexpr double G0 = 0.21132486540518711775, G1 = 0.78867513459481288225;
                const double     xi = (q & 1) ? G1 : G0, eta = (q >> 1) ? G1 : G0;
                const double     dNx[4] = {-(1 - eta), (1 - eta), -eta, eta};
                const double     dNy[4] = {-(1 - xi), -xi, (1 - xi), xi};
                double           a00, a01, a10, a11;
                if (P.table)
                  {
                    const double *tp =
                      P.table + (((size_t)cell * n * n + (size_t)iy * n + ix) * 4 + q) * 4;
                    a00 = tp[0], a01 = tp[1], a10 = tp[2], a11 = tp[3];
                  }
                else if (fast)
                  P.coef.from_sines(tsx[4 * ix + q], tsy[4 * lr + q], a00, a01, a10, a11);
                else
                  P.coef(c[0] + (ix + xi) * hx, c[1] + (iy + eta) * hy, a00, a01, a10, a11);
                const double c00 = a00 * rxx, c01 = 0.125 * (a01 + a10), c11 = a11 * ryy;
                int          e   = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int j = i; j < 4; ++j)
                    {
                      K[e] = fma(dNx[i] * dNx[j], c00, K[e]);
                      K[e] = fma(dNx[i] * dNy[j] + dNy[i] * dNx[j], c01, K[e]);
                      K[e] = fma(dNy[i] * dNy[j], c11, K[e]);
                      ++e;
                    }
              }
            // sum_q N_i(q) w_q = 1/4: every vertex receives a quarter of f |cell|
            Fe[0] = Fe[1] = Fe[2] = Fe[3] = P.rhs_value * hx * hy * 0.25;
          }
        else if (iy >= 0 && iy < n)
          {
            double Px[4], Py[4];
#pragma unroll
            for (int v = 0; v < 4; ++v)
              fine_vertex(c, n, ix + (v & 1), iy + (v >> 1), Px[v], Py[v]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              {
                const double xi = (q & 1) ? g1 : g0, eta = (q >> 1) ? g1 : g0;
                const double Nv[4]  = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
                const double dNx[4] = {-(1 - eta), (1 - eta), -eta, eta};
                const double dNy[4] = {-(1 - xi), -xi, (1 - xi), xi};
                double       J00 = 0, J01 = 0, J10 = 0, J11 = 0, xq = 0, yq = 0;
#pragma unroll
                for (int v = 0; v < 4; ++v)
                  {
                    xq += Px[v] * Nv[v];
                    yq += Py[v] * Nv[v];
                    J00 += Px[v] * dNx[v];
                    J01 += Px[v] * dNy[v];
                    J10 += Py[v] * dNx[v];
                    J11 += Py[v] * dNy[v];
                  }
                const double det = J00 * J11 - J01 * J10;
                const double idet = 1.0 / det;
                const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;
                const double JxW = det * 0.25;
                double       Gx[4], Gy[4];
#pragma unroll
                for (int v = 0; v < 4; ++v)
                  {
                    Gx[v] = i00 * dNx[v] + i10 * dNy[v];
                    Gy[v] = i01 * dNx[v] + i11 * dNy[v];
                  }
                double a00, a01, a10, a11;
                if (P.table)
                  {
                    const double *tp =
                      P.table + (((size_t)cell * n * n + (size_t)iy * n + ix) * 4 + q) * 4;
                    a00 = tp[0], a01 = tp[1], a10 = tp[2], a11 = tp[3];
                  }
                else if (fast)
                  P.coef.from_sines(tsx[4 * ix + q], tsy[4 * lr + q], a00, a01, a10, a11);
                else
                  P.coef(xq, yq, a00, a01, a10, a11);
                // K is symmetrised: only the symmetric part of A enters x^T K x, and the
                // reference tensor is symmetric up to 1e-17
                const double as = 0.5 * (a01 + a10);
                int          e  = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  {
                    const double t0 = Gx[i] * a00 + Gy[i] * as, t1 = Gx[i] * as + Gy[i] * a11;
#pragma unroll
                    for (int j = i; j < 4; ++j)
                      K[e++] += (t0 * Gx[j] + t1 * Gy[j]) * JxW;
                    Fe[i] += Nv[i] * P.rhs_value * JxW;
                  }
              }
          }
#pragma unroll
        for (int e = 0; e < 10; ++e)
          ke[e * slab + t] = K[e];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          ke[(10 + e) * slab + t] = Fe[e];
      }
    __syncthreads();

    // ---- phase 2: gather per node.  K index of (i,j), i<=j: 00:0 01:1 02:2 03:3 11:4 12:5 13:6 22:7 23:8 33:9
    double *out = P.sten + (size_t)cell * ST_NARR * N;
    for (int t = threadIdx.x; t < rows * np; t += blockDim.x)
      {
        const int ly = t / np, jx = t % np;
        const int jy = jy0 + ly;
        // local slab rows: cell row jy-1 -> ly, cell row jy -> ly+1
        const bool hasW = jx > 0, hasE = jx < n, hasS = jy > 0, hasN = jy < n;
        const int  sw = ly * n + jx - 1, se = ly * n + jx, nw = (ly + 1) * n + jx - 1,
                  ne = (ly + 1) * n + jx;
        double kc = 0, kE = 0, kN = 0, kd1 = 0, kd2 = 0, f = 0;
        if (hasS && hasW)
          {
            kc += ke[9 * slab + sw];
            f += ke[13 * slab + sw];
          }
        if (hasS && hasE)
          {
            kc += ke[7 * slab + se];
            kE += ke[8 * slab + se];
            f += ke[12 * slab + se];
          }
        if (hasN && hasW)
          {
            kc += ke[4 * slab + nw];
            kN += ke[6 * slab + nw];
            f += ke[11 * slab + nw];
          }
        if (hasN && hasE)
          {
            kc += ke[0 * slab + ne];
            kE += ke[1 * slab + ne];
            kN += ke[2 * slab + ne];
            kd1 = ke[3 * slab + ne];
            kd2 = ke[5 * slab + ne];
            f += ke[10 * slab + ne];
          }
        const int lex           = jy * np + jx;
        out[ST_KC * N + lex]    = kc;
        out[ST_KE * N + lex]    = kE;
        out[ST_KN * N + lex]    = kN;
        out[ST_KD1 * N + lex]   = kd1;
        out[ST_KD2 * N + lex]   = kd2;
        out[ST_F * N + lex]     = f;
      }
  }

  cudaError_t
  launch_assemble(const Shard &s, cudaStream_t st, int *n_launches)
  {
    AssembleParams P;
    P.n         = s.n;
    P.n_cells   = s.n_cells;
    P.corners   = s.d_corners;
    P.table     = s.coeff.kind == MSB_COEFF_TABLE ? s.d_table : nullptr;
    P.sten      = s.d_sten;
    P.rhs_value = s.rhs_value;
    P.coef      = make_coeff_eval(s.coeff);
    // strip height: as many node rows as fit ~96 KB of element data, at most 16
    int R = (int)(96 * 1024 / (14 * sizeof(double) * (size_t)s.n)) - 1;
    R     = R < 1 ? 1 : (R > 16 ? 16 : R);
    P.R   = R;
    const size_t smem = sizeof(double) * (14 * (size_t)(R + 1) * s.n + 4 * (size_t)s.n + 4 * (size_t)(R + 1));
    cudaError_t  e =
      cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
      return e;
    // gridDim.y is limited to 65535: loop over slices of cells
    const int strips = (s.np + R - 1) / R;
    for (int c0 = 0; c0 < s.n_cells; c0 += 65535)
      {
        const int      nc = s.n_cells - c0 < 65535 ? s.n_cells - c0 : 65535;
        AssembleParams Q  = P;
        Q.corners += 8 * (size_t)c0;
        Q.sten += (size_t)c0 * ST_NARR * s.N;
        if (Q.table)
          Q.table += (size_t)c0 * s.n * s.n * 16;
        assemble_kernel<<<dim3(strips, nc), 256, smem, st>>>(Q);
        ++*n_launches;
      }
    return cudaGetLastError();
  }

  // ======================================================================================
  // y = K x for one vector on one cell's stencil (lexicographic order).
  // ======================================================================================
  __device__ inline double
  stencil_apply(const double *__restrict__ S, const double *__restrict__ x, int n, int jx, int jy)
  {
    const int  np = n + 1, N = np * np, i = jy * np + jx;
    const bool hasW = jx > 0, hasE = jx < n, hasS = jy > 0, hasN = jy < n;
    double     y = S[ST_KC * N + i] * x[i];
    if (hasE)
      y += S[ST_KE * N + i] * x[i + 1];
    if (hasW)
      y += S[ST_KE * N + i - 1] * x[i - 1];
    if (hasN)
      y += S[ST_KN * N + i] * x[i + np];
    if (hasS)
      y += S[ST_KN * N + i - np] * x[i - np];
    if (hasN && hasE)
      y += S[ST_KD1 * N + i] * x[i + np + 1];
    if (hasS && hasW)
      y += S[ST_KD1 * N + i - np - 1] * x[i - np - 1];
    if (hasN && hasW)
      y += S[ST_KD2 * N + i - 1] * x[i + np - 1];
    if (hasS && hasE)
      y += S[ST_KD2 * N + i - np] * x[i - np + 1];
    return y;
  }

  __global__ void
  apply_operator_kernel(int n, const double *__restrict__ S, const double *__restrict__ x,
                        double *__restrict__ y)
  {
    const int np = n + 1, lex = blockIdx.x * blockDim.x + threadIdx.x;
    if (lex < np * np)
      y[lex] = stencil_apply(S, x, n, lex % np, lex / np);
  }

  cudaError_t
  launch_apply_operator(const Shard &s, int cell, const double *d_x, double *d_y, cudaStream_t st)
  {
    apply_operator_kernel<<<(s.N + 255) / 256, 256, 0, st>>>(
      s.n, s.d_sten + (size_t)cell * ST_NARR * s.N, d_x, d_y);
    return cudaGetLastError();
  }

  // ======================================================================================
  // assemble_global_element_matrix (basis.tpp:245-285): M_ij = phi_i . (K phi_j) with the
  // UNCONSTRAINED K over all N DoFs, b_i = phi_i . F.  One CTA per coarse cell, fixed
  // summation order (deterministic).
  // ======================================================================================
  // Every thread owns one node column and a band of rows and marches up the band with a 3x3
  // register window per basis, so each phi value is loaded three times (by the three column
  // neighbours, from the same cache lines) instead of nine and each stencil coefficient once.
  // THREADS = 256 (two CTAs per SM) or 512: local meshes wider than 128 nodes leave half of a 256-thread
  // CTA idle (one band of rows), which matters when a shard has too few cells to fill the GPU (the
  // reference's default run: 64 cells of 129 x 129 nodes) -- 512 threads sweep three bands instead.
  template <int THREADS>
  __global__ void __launch_bounds__(THREADS, 512 / THREADS)
  element_matrix_kernel(int n, const double *__restrict__ sten, const double *__restrict__ phi,
                        double *__restrict__ M, double *__restrict__ b)
  {
    const int     np = n + 1, N = np * np, cell = blockIdx.x;
    const double *S = sten + (size_t)cell * ST_NARR * N;
    const double *P = phi + (size_t)cell * 4 * N;
    double        acc[20];
#pragma unroll
    for (int k = 0; k < 20; ++k)
      acc[k] = 0.0;
    // thread -> (column X, band of rows); with np > blockDim.x the columns are swept in passes
    const int nb = blockDim.x >= np ? blockDim.x / np : 1; // bands
    const int rb = (np + nb - 1) / nb;                     // rows per band
    for (int col0 = 0; col0 < np; col0 += (nb > 1 || blockDim.x >= np ? np : blockDim.x))
      {
        const int X    = blockDim.x >= np ? threadIdx.x % np : col0 + threadIdx.x;
        const int band = blockDim.x >= np ? threadIdx.x / np : 0;
        const int y0 = band * rb, y1 = min(np, y0 + rb);
        if (X < np && band < nb && y0 < y1)
          {
            const bool hasW = X > 0, hasE = X < n;
            const int  xm = hasW ? X - 1 : X, xp = hasE ? X + 1 : X;
            // window rows a = y-1, c = y, d = y+1; columns m (x-1), 0 (x), p (x+1).  Values read
            // through clamped indices are always multiplied by a zero coupling.
            double a0[4], cm[4], c0[4], cp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              {
                const double *Pj = P + (size_t)j * N;
                const int     ra = (y0 > 0 ? y0 - 1 : y0) * np, rc = y0 * np;
                a0[j] = Pj[ra + X];
                cm[j] = Pj[rc + xm], c0[j] = Pj[rc + X], cp[j] = Pj[rc + xp];
              }
            for (int y = y0; y < y1; ++y)
              {
                const int  i    = y * np + X;
                const bool hasS = y > 0, hasN = y < n;
                const int  rd = (hasN ? y + 1 : y) * np, ra = (hasS ? y - 1 : y) * np;
                // symmetric stencil row of node (X, y); couplings that leave the mesh are zero
                const double kc = S[ST_KC * N + i], f = S[ST_F * N + i];
                const double kE = hasE ? S[ST_KE * N + i] : 0.0, kW = hasW ? S[ST_KE * N + i - 1] : 0.0;
                const double kN = hasN ? S[ST_KN * N + i] : 0.0, kS = hasS ? S[ST_KN * N + i - np] : 0.0;
                const double kNE = hasN && hasE ? S[ST_KD1 * N + i] : 0.0;
                const double kSW = hasS && hasW ? S[ST_KD1 * N + i - np - 1] : 0.0;
                const double kNW = hasN && hasW ? S[ST_KD2 * N + i - 1] : 0.0;
                const double kSE = hasS && hasE ? S[ST_KD2 * N + i - np] : 0.0;
                double       kp[4], dm[4], d0[4], dp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  {
                    const double *Pj = P + (size_t)j * N;
                    dm[j] = Pj[rd + xm], d0[j] = Pj[rd + X], dp[j] = Pj[rd + xp];
                    // the two lower diagonal neighbours are re-read (L1 hits) to keep the window
                    // small enough for two resident CTAs per SM
                    const double am = Pj[ra + xm], ap = Pj[ra + xp];
                    double       t  = kc * c0[j];
                    t     = fma(kE, cp[j], t);
                    t     = fma(kW, cm[j], t);
                    t     = fma(kN, d0[j], t);
                    t     = fma(kS, a0[j], t);
                    t     = fma(kNE, dp[j], t);
                    t     = fma(kSW, am, t);
                    t     = fma(kNW, dm[j], t);
                    t     = fma(kSE, ap, t);
                    kp[j] = t;
                  }
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2)
                  {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      acc[4 * i2 + j] = fma(c0[i2], kp[j], acc[4 * i2 + j]);
                    acc[16 + i2] = fma(c0[i2], f, acc[16 + i2]);
                  }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  {
                    a0[j] = c0[j];
                    cm[j] = dm[j], c0[j] = d0[j], cp[j] = dp[j];
                  }
              }
          }
        if (blockDim.x >= np)
          break;
      }
    __shared__ double red[THREADS / 32][20];
    const int         lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 20; ++k)
      {
        double v = acc[k];
        for (int off = 16; off > 0; off >>= 1)
          v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0)
          red[warp][k] = v;
      }
    __syncthreads();
    if (threadIdx.x < 20)
      {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w)
          v += red[w][threadIdx.x];
        if (threadIdx.x < 16)
          M[16 * (size_t)cell + threadIdx.x] = v;
        else
          b[4 * (size_t)cell + threadIdx.x - 16] = v;
      }
  }

  cudaError_t
  launch_element_matrices(const Shard &s, cudaStream_t st, int *n_launches)
  {
    if (s.np > 128 && s.np <= 512 / 3 && s.n_cells < 4 * 148)
      element_matrix_kernel<512><<<s.n_cells, 512, 0, st>>>(s.n, s.d_sten, s.d_phi, s.d_M, s.d_b);
    else
      element_matrix_kernel<256><<<s.n_cells, 256, 0, st>>>(s.n, s.d_sten, s.d_phi, s.d_M, s.d_b);
    ++*n_launches;
    return cudaGetLastError();
  }

  // ======================================================================================
  // reordering between the lexicographic device layout and the deal.II DoF order
  // ======================================================================================
  __global__ void
  permute_kernel(int N, const uint32_t *__restrict__ dofmap, const double *__restrict__ src,
                 double *__restrict__ dst, int lex_to_dof)
  {
    const int lex = blockIdx.x * blockDim.x + threadIdx.x;
    if (lex >= N)
      return;
    if (lex_to_dof)
      dst[dofmap[lex]] = src[lex];
    else
      dst[lex] = src[dofmap[lex]];
  }

  cudaError_t
  launch_permute(const Shard &s, const double *d_src, double *d_dst, bool lex_to_dof, cudaStream_t st)
  {
    permute_kernel<<<(s.N + 255) / 256, 256, 0, st>>>(s.N, s.d_dofmap, d_src, d_dst, lex_to_dof ? 1 : 0);
    return cudaGetLastError();
  }

  // Bulk form (msb_get_bases / msb_get_global_solutions): n_vec consecutive vectors in one launch.
  // Reads are coalesced; the scattered writes stay inside one vector (<= 0.5 MB: L2 resident).
  __global__ void __launch_bounds__(256)
  permute_batch_kernel(int N, const uint32_t *__restrict__ dofmap, const double *__restrict__ src,
                       double *__restrict__ dst, size_t n_vec)
  {
    for (size_t v = blockIdx.y; v < n_vec; v += gridDim.y)
      {
        const double *s = src + v * (size_t)N;
        double       *d = dst + v * (size_t)N;
        for (int lex = blockIdx.x * blockDim.x + threadIdx.x; lex < N; lex += gridDim.x * blockDim.x)
          d[dofmap[lex]] = s[lex];
      }
  }

  cudaError_t
  launch_permute_batch(const Shard &s, const double *d_src, double *d_dst, size_t n_vec, cudaStream_t st)
  {
    if (n_vec == 0)
      return cudaSuccess;
    const int bx = (s.N + 255) / 256 < 16 ? (s.N + 255) / 256 : 16;
    const int by = n_vec < 32768 ? (int)n_vec : 32768;
    permute_batch_kernel<<<dim3(bx, by), 256, 0, st>>>(s.N, s.d_dofmap, d_src, d_dst, n_vec);
    return cudaGetLastError();
  }

  // ======================================================================================
  // set_global_weights (basis.tpp:352-377) for all cells: gsol = sum_i w_i phi_i.
  // Purely bandwidth bound: reads 4N, writes N doubles per cell.
  // ======================================================================================
  __global__ void __launch_bounds__(256)
  global_solution_kernel(int N, int nb, const double *__restrict__ phi, const double *__restrict__ w,
                         double *__restrict__ out)
  {
    const int     cell = blockIdx.y;
    const double *P    = phi + (size_t)cell * nb * N;
    const double *wc   = w + (size_t)nb * cell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
      {
        // same accumulation order as Vector::sadd(1, w_i, phi_i) for i = 0..2^dim-1
        double v = 0.0;
        for (int k = 0; k < nb; ++k)
          v = v + wc[k] * P[(size_t)k * N + i];
        out[(size_t)cell * N + i] = v;
      }
  }

  cudaError_t
  launch_global_solution(const Shard &s, const double *d_w, cudaStream_t st)
  {
    const int bx = (s.N + 255) / 256 < 8 ? (s.N + 255) / 256 : 8;
    for (int c0 = 0; c0 < s.n_cells; c0 += 65535)
      {
        const int nc = s.n_cells - c0 < 65535 ? s.n_cells - c0 : 65535;
        global_solution_kernel<<<dim3(bx, nc), 256, 0, st>>>(
          s.N, s.nb, s.d_phi + (size_t)c0 * s.nb * s.N, d_w + (size_t)s.nb * c0, s.d_gsol + (size_t)c0 * s.N);
      }
    return cudaGetLastError();
  }

  // ======================================================================================
  // constraint set of one (cell, basis): boundary DoFs ascending + BasisQ1 values
  // (basis.tpp:119-135).  The sorted order is produced by ranking: a boundary DoF's
  // position is the number of boundary DoFs with a smaller index.
  // ======================================================================================
  __global__ void
  constraints_kernel(int n, const uint32_t *__restrict__ dofmap, const double *__restrict__ corners,
                     const double *__restrict__ q1coef, int ib, uint32_t *__restrict__ dofs,
                     double *__restrict__ vals)
  {
    const int nb = 4 * n, np = n + 1;
    const int t  = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb)
      return;
    // boundary node t: walk the four sides
    auto node_of = [&](int k, int &jx, int &jy) {
      if (k < n)
        jx = k, jy = 0;
      else if (k < 2 * n)
        jx = n, jy = k - n;
      else if (k < 3 * n)
        jx = n - (k - 2 * n), jy = n;
      else
        jx = 0, jy = n - (k - 3 * n);
    };
    int jx, jy;
    node_of(t, jx, jy);
    const uint32_t d    = dofmap[jy * np + jx];
    int            rank = 0;
    for (int k = 0; k < nb; ++k)
      {
        int kx, ky;
        node_of(k, kx, ky);
        rank += dofmap[ky * np + kx] < d;
      }
    double px, py;
    fine_vertex(corners, n, jx, jy, px, py);
    dofs[rank] = d;
    vals[rank] = basis_q1_value(q1coef, ib, px, py);
  }

  cudaError_t
  launch_constraints(const Shard &s, int cell, int ib, uint32_t *d_dofs, double *d_vals, cudaStream_t st)
  {
    constraints_kernel<<<(4 * s.n + 127) / 128, 128, 0, st>>>(
      s.n, s.d_dofmap, s.d_corners + 8 * (size_t)cell, s.d_q1coef + 16 * (size_t)cell, ib, d_dofs, d_vals);
    return cudaGetLastError();
  }
} // namespace msb
