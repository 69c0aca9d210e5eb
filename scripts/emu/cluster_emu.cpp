// cluster_emu.cpp -- CPU emulation of solve_cluster_kernel (msb_solve_cluster.cu) for index-logic
// checks without a GPU: every CUDA thread is an OS thread, __syncthreads / cluster.sync / warp
// shuffles are std::barriers, DSMEM is ordinary memory.  The kernel source is compiled
// UNCHANGED (MSB_EMU only swaps the extern __shared__ declaration).  Development tool only:
// nothing in the library, the tests or bench.py uses it.
//
//   g++ -O2 -std=c++20 -pthread -DMSB_EMU -I/usr/local/cuda/include -I include \
//       -I mpi_parallel_multiscale_diffusion_fem_b200/csrc scripts/emu/cluster_emu.cpp -o /tmp/cluster_emu
//   /tmp/cluster_emu 5      (n_refine_local = 5, 6 or 7)
#include "emu_shims.hpp"
#include "msb_solve_cluster.cu"

using namespace msb;

// ---------------------------------------------------------------- a small problem on the host
static double
coef_a(double x, double y)
{
  return 1.0 - 0.9 * (0.5 * std::sin(2 * M_PI * x * 9) + 0.5 * std::sin(2 * M_PI * y * 7));
}

// 9-point stencil of -div(a grad u) on the unit-square coarse cell [x0,x0+H]^2, n x n Q1 cells
static void
assemble(int n, double x0, double y0, double H, std::vector<double> &S)
{
  const int np = n + 1, N = np * np;
  S.assign((size_t)ST_NARR * N, 0.0);
  const double h = H / n, gp[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  for (int iy = 0; iy < n; ++iy)
    for (int ix = 0; ix < n; ++ix)
      {
        double Ke[4][4] = {};
        for (int qy = 0; qy < 2; ++qy)
          for (int qx = 0; qx < 2; ++qx)
            {
              const double s = gp[qx], t = gp[qy];
              const double a = coef_a(x0 + (ix + s) * h, y0 + (iy + t) * h);
              // reference gradients of the 4 bilinear shape functions (vertex order x fastest)
              const double gx[4] = {-(1 - t), (1 - t), -t, t}, gy[4] = {-(1 - s), -s, (1 - s), s};
              for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j)
                  Ke[i][j] += 0.25 * a * (gx[i] * gx[j] + gy[i] * gy[j]); // h^2 * (1/h)^2 * w(1/4)
            }
        const int v[4] = {iy * np + ix, iy * np + ix + 1, (iy + 1) * np + ix, (iy + 1) * np + ix + 1};
        for (int i = 0; i < 4; ++i)
          S[ST_KC * N + v[i]] += Ke[i][i];
        S[ST_KE * N + v[0]] += Ke[0][1];
        S[ST_KE * N + v[2]] += Ke[2][3];
        S[ST_KN * N + v[0]] += Ke[0][2];
        S[ST_KN * N + v[1]] += Ke[1][3];
        S[ST_KD1 * N + v[0]] += Ke[0][3];
        S[ST_KD2 * N + v[0]] += Ke[1][2]; // (jx+1,jy)-(jx,jy+1), stored at (jx,jy)
      }
}

static double
sget(const std::vector<double> &S, int np, int x, int y, int ex, int ey)
{
  return bpx::sten_get(S.data(), np, np * np, x, y, ex, ey);
}

int
main(int argc, char **argv)
{
  const int  l  = argc > 1 ? atoi(argv[1]) : 5;
  const bool tm = argc > 2 ? atoi(argv[2]) != 0 : true; // 1: tensor-memory kernel (4 bases per pass)
  const int n = 1 << l, np = n + 1, N = np * np;
  const int CS = n / 16, T = 4 * n;
  const double H = 1.0 / 8, X0 = 0.25, Y0 = 0.5;
  std::vector<double> S;
  assemble(n, X0, Y0, H, S);
  double corners[8] = {X0, Y0, X0 + H, Y0, X0, Y0 + H, X0 + H, Y0 + H};
  // BasisQ1 coefficients (basis_q1.tpp:35-46): columns of the inverse of [1 x y xy] at the vertices
  double q1[16];
  {
    double A[4][8];
    for (int v = 0; v < 4; ++v)
      {
        const double x = corners[2 * v], y = corners[2 * v + 1];
        A[v][0] = 1, A[v][1] = x, A[v][2] = y, A[v][3] = x * y;
        for (int j = 0; j < 4; ++j)
          A[v][4 + j] = v == j;
      }
    for (int cidx = 0; cidx < 4; ++cidx)
      {
        int pv = cidx;
        for (int r = cidx; r < 4; ++r)
          if (std::fabs(A[r][cidx]) > std::fabs(A[pv][cidx]))
            pv = r;
        for (int j = 0; j < 8; ++j)
          std::swap(A[cidx][j], A[pv][j]);
        const double d = A[cidx][cidx];
        for (int j = 0; j < 8; ++j)
          A[cidx][j] /= d;
        for (int r = 0; r < 4; ++r)
          if (r != cidx)
            {
              const double f = A[r][cidx];
              for (int j = 0; j < 8; ++j)
                A[r][j] -= f * A[cidx][j];
            }
      }
    for (int r = 0; r < 4; ++r)
      for (int ib = 0; ib < 4; ++ib)
        q1[r * 4 + ib] = A[r][4 + ib]; // coef[r*4+ib] = inverse(r, ib)
  }
  // Galerkin diagonals level by level (what stream_galerkin_kernel does)
  const int levels = l - 1;
  std::vector<int> npl(levels + 2), off(levels + 2);
  npl[0] = np;
  int cn = 0;
  for (int k = 1; k <= levels; ++k)
    npl[k] = (n >> k) + 1, off[k] = cn, cn += npl[k] * npl[k];
  std::vector<double> dinv(cn, 0.0), Sf = S;
  for (int k = 1; k <= levels; ++k)
    {
      const int npc = npl[k], nin = npc - 2, npf = npl[k - 1], Nc = npc * npc;
      std::vector<double> Sc((size_t)ST_NARR * Nc, 0.0);
      for (int Yc = 1; Yc <= nin; ++Yc)
        for (int Xc = 1; Xc <= nin; ++Xc)
          {
            double a[5];
            bpx::galerkin_row(Sf.data(), npf, npf * npf, Xc, Yc, a);
            const int i = Yc * npc + Xc;
            Sc[ST_KC * Nc + i] = a[0];
            if (Xc < nin)
              Sc[ST_KE * Nc + i] = a[1];
            if (Yc < nin)
              Sc[ST_KN * Nc + i] = a[2];
            if (Xc < nin && Yc < nin)
              Sc[ST_KD1 * Nc + i] = a[3];
            if (Xc > 1 && Yc < nin)
              Sc[ST_KD2 * Nc + i - 1] = a[4];
            dinv[off[k] + i] = 1.0 / a[0];
          }
      // NOTE: Sf has ST_NARR arrays of stride Nf; galerkin_row is told Nf
      Sf = Sc;
    }

  std::vector<double> phi((size_t)4 * N, -777.0), res(4, -1);
  std::vector<int32_t> iters(4, -5);
  int32_t fail[2] = {INT_MAX, 0};
  clus::Params P;
  P.corners = corners, P.q1coef = q1, P.sten = S.data(), P.dinv = dinv.data(), P.phi = phi.data();
  P.iters = iters.data(), P.res = res.data(), P.fail = fail, P.tol2 = 1e-24, P.max_iter = 500, P.cn = cn, P.cell0 = 0, P.split = 0;

  size_t smem_doubles =
    tm ? (l == 5 ? clus::Lay<5, 4, true>::total : l == 6 ? clus::Lay<6, 4, true>::total : clus::Lay<7, 4, true>::total) :
         (l == 5 ? clus::Lay<5, 2, false>::total : l == 6 ? clus::Lay<6, 2, false>::total : clus::Lay<7, 2, false>::total);
  printf("l=%d n=%d cluster=%d threads=%d tmem=%d smem=%zu bytes\n", l, n, CS, T, (int)tm, smem_doubles * 8);
  emu::Cluster cl;
  cl.bar = std::make_unique<std::barrier<>>(CS * T);
  cl.ctas.resize(CS);
  for (auto &cta : cl.ctas)
    {
      cta.bar = std::make_unique<std::barrier<>>(T);
      cta.smem.assign(smem_doubles, NAN); // poison: the kernel must initialise what it reads
      for (int w = 0; w < T / 32; ++w)
        cta.warps.emplace_back(new emu::Warp);
    }
  std::vector<std::thread> th;
  for (int rank = 0; rank < CS; ++rank)
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, rank, t] {
        emu::t_cluster = &cl, emu::t_rank = rank;
        threadIdx.x = t, blockIdx.x = rank, blockDim.x = T, gridDim.x = CS;
        if (tm)
          {
            if (l == 5)
              clus::solve_cluster_kernel<5, 4, true>(P);
            else if (l == 6)
              clus::solve_cluster_kernel<6, 4, true>(P);
            else
              clus::solve_cluster_kernel<7, 4, true>(P);
          }
        else
          {
            if (l == 5)
              clus::solve_cluster_kernel<5, 2, false>(P);
            else if (l == 6)
              clus::solve_cluster_kernel<6, 2, false>(P);
            else
              clus::solve_cluster_kernel<7, 2, false>(P);
          }
      });
  for (auto &t : th)
    t.join();

  // ---- checks: residual of the condensed system, boundary data, partition of unity
  double worst = 0, pu = 0;
  for (int ib = 0; ib < 4; ++ib)
    {
      const double *x = phi.data() + (size_t)ib * N;
      double        rr = 0;
      for (int y = 1; y < n; ++y)
        for (int xx = 1; xx < n; ++xx)
          {
            double s = 0;
            for (int ey = -1; ey <= 1; ++ey)
              for (int ex = -1; ex <= 1; ++ex)
                s += sget(S, np, xx, y, ex, ey) * x[(y + ey) * np + xx + ex];
            rr += s * s;
          }
      printf("basis %d: iters %d  reported res %.3e  true ||K x||_interior %.3e  fail %d\n", ib, iters[ib],
             res[ib], std::sqrt(rr), fail[0]);
      worst = std::fmax(worst, std::sqrt(rr));
    }
  for (int i = 0; i < N; ++i)
    pu = std::fmax(pu, std::fabs(phi[i] + phi[N + i] + phi[2 * N + i] + phi[3 * N + i] - 1.0));
  printf("partition of unity defect %.3e, worst residual %.3e\n", pu, worst);

  // ---- the same PCG written plainly on full arrays (iteration counts must agree)
  for (int ib = 0; ib < 4; ++ib)
    {
      std::vector<double> x(N, 0.0), r(N, 0.0), z(N, 0.0), p(N, 0.0), q(N, 0.0);
      auto g = [&](int bx, int by) {
        double px, py;
        const double rn = 1.0 / n, s = bx * rn, t = by * rn;
        px = corners[0] + s * (corners[2] - corners[0]) + t * (corners[4] - corners[0]);
        py = corners[1] + s * (corners[3] - corners[1]) + t * (corners[5] - corners[1]);
        return q1[0 * 4 + ib] + q1[1 * 4 + ib] * px + q1[2 * 4 + ib] * py + q1[3 * 4 + ib] * px * py;
      };
      // initial guess: the coarse Q1 shape function on every node, r_0 = -(K g) on interior rows (as the kernel)
      for (int y = 0; y <= n; ++y)
        for (int xx = 0; xx <= n; ++xx)
          x[y * np + xx] = g(xx, y);
      for (int y = 1; y < n; ++y)
        for (int xx = 1; xx < n; ++xx)
          for (int ey = -1; ey <= 1; ++ey)
            for (int ex = -1; ex <= 1; ++ex)
              r[y * np + xx] -= sget(S, np, xx, y, ex, ey) * x[(y + ey) * np + xx + ex];
      auto dot = [&](const std::vector<double> &a, const std::vector<double> &b) {
        double s = 0;
        for (int y = 1; y < n; ++y)
          for (int xx = 1; xx < n; ++xx)
            s += a[y * np + xx] * b[y * np + xx];
        return s;
      };
      auto precond = [&]() {
        std::vector<std::vector<double>> v(levels + 1);
        v[0] = r;
        for (int k = 1; k <= levels; ++k)
          {
            v[k].assign(npl[k] * npl[k], 0.0);
            for (int cy = 1; cy < npl[k] - 1; ++cy)
              for (int cx = 1; cx < npl[k] - 1; ++cx)
                v[k][cy * npl[k] + cx] = clus::restrict_node(v[k - 1].data() + 2 * cy * npl[k - 1] + 2 * cx, npl[k - 1]);
          }
        for (int k = levels; k >= 0; --k)
          for (int fy = 1; fy < npl[k] - 1; ++fy)
            for (int fx = 1; fx < npl[k] - 1; ++fx)
              {
                const int    i  = fy * npl[k] + fx;
                const double di = k == 0 ? 1.0 / S[ST_KC * N + i] : dinv[off[k] + i];
                double       val = v[k][i] * di;
                if (k < levels)
                  {
                    const int     npc = npl[k + 1];
                    const double *vc  = v[k + 1].data();
                    const int     xl = fx >> 1, xh = (fx + 1) >> 1, yl = fy >> 1, yh = (fy + 1) >> 1;
                    val += 0.25 * ((vc[yl * npc + xl] + vc[yl * npc + xh]) + (vc[yh * npc + xl] + vc[yh * npc + xh]));
                  }
                v[k][i] = val;
              }
        z = v[0];
      };
      int    it = 0;
      double rz = 1, rr = dot(r, r);
      while (rr > 1e-24 && it < 500)
        {
          precond();
          const double rzn = dot(r, z), beta = it == 0 ? 0.0 : rzn / rz;
          rz = rzn;
          ++it;
          for (int i = 0; i < N; ++i)
            p[i] = z[i] + beta * p[i];
          for (int y = 1; y < n; ++y)
            for (int xx = 1; xx < n; ++xx)
              {
                double s = 0;
                for (int ey = -1; ey <= 1; ++ey)
                  for (int ex = -1; ex <= 1; ++ex)
                    {
                      const int bx = xx + ex, by = y + ey;
                      if (!(bx == 0 || by == 0 || bx == n || by == n))
                        s += sget(S, np, xx, y, ex, ey) * p[by * np + bx];
                    }
                q[y * np + xx] = s;
              }
          const double alpha = rz / dot(p, q);
          for (int y = 1; y < n; ++y)
            for (int xx = 1; xx < n; ++xx)
              x[y * np + xx] += alpha * p[y * np + xx], r[y * np + xx] -= alpha * q[y * np + xx];
          rr = dot(r, r);
        }
      double d = 0, nr = 0;
      for (int i = 0; i < N; ++i)
        d += (x[i] - phi[(size_t)ib * N + i]) * (x[i] - phi[(size_t)ib * N + i]), nr += x[i] * x[i];
      printf("plain PCG basis %d: iters %d (kernel %d)  rel diff %.3e\n", ib, it, iters[ib], std::sqrt(d / nr));
    }
  return worst < 1e-10 && pu < 1e-9 ? 0 : 1;
}
