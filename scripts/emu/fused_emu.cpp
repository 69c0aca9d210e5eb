// fused_emu.cpp -- CPU emulation of solve_fused_kernel (msb_solve_fused.cu): the whole basis stage of one
// coarse cell (on-chip stencil assembly, Galerkin hierarchy, 2 x 2 multilevel PCG solves, element matrix
// epilogue) executed by 512 OS threads, checked against a plain host implementation of the reference's
// sequence (assemble K and F, condense, solve K_II phi_I = -K_IB g_B, M = Phi^T K Phi over ALL DoFs,
// b = Phi^T F: diffusion_problem_basis.tpp:159-285, 450-465).  The kernel source is compiled UNCHANGED
// (MSB_EMU swaps shared / tensor memory and the PTX).  Development and test tool only: nothing in the
// library, bench.py or the GPU tests uses it.
//
//   g++ -O1 -std=c++20 -pthread -DMSB_EMU -I/usr/local/cuda/include -I include \
//       -I mpi_parallel_multiscale_diffusion_fem_b200/csrc scripts/emu/fused_emu.cpp -o /tmp/fused_emu
//   /tmp/fused_emu [kind 0|1|2|3] [max_iter] [flavour]          (-DEMU_NL=5 for the 32 x 32 instantiation)
#include "emu_shims.hpp"
#include "msb_solve_fused.cu"

using namespace msb;

int
main(int argc, char **argv)
{
  const int kind     = argc > 1 ? atoi(argv[1]) : 1;
  const int max_iter = argc > 2 ? atoi(argv[2]) : 500;
  const int flavor   = argc > 3 ? atoi(argv[3]) : 0;
#ifndef EMU_NL
#  define EMU_NL 6 // -DEMU_NL=5: the 32 x 32 instantiation (128 threads)
#endif
  constexpr int NL = EMU_NL, n = 1 << NL, np = n + 1, N = np * np;
  // a coarse cell of the target configuration (256 x 256 coarse mesh), or a 2:1 rectangle for kind 3
  const double H = 1.0 / 256, X0 = 37 * H, Y0 = 101 * H, HY = kind == 3 ? 0.5 * H : H;
  double       corners[8] = {X0, Y0, X0 + H, Y0, X0, Y0 + HY, X0 + H, Y0 + HY};
  msb_coeff_desc d{};
  d.kind = kind;
  if (kind == MSB_COEFF_PERIODIC)
    d.par[0] = 1.0 / 64, d.par[1] = 0.9999;
  if (kind == MSB_COEFF_INCLUSIONS)
    d.par[0] = 1.0 / 2048, d.par[1] = 0.2, d.par[2] = 1e4, d.par[3] = 1.0, d.seed = 1234;
  if (kind == MSB_COEFF_CONSTANT)
    d.par[0] = 2.5;
  const CoeffEval coef = make_coeff_eval(d);
  const double    f    = 3.5;

  // ---- host reference: element loop with the full tensor, 2x2 Gauss
  const double        hx = H / n, hy = HY / n, gp[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  std::vector<double> S((size_t)ST_NARR * N, 0.0);
  for (int iy = 0; iy < n; ++iy)
    for (int ix = 0; ix < n; ++ix)
      {
        double Ke[4][4] = {}, Fe[4] = {};
        for (int q = 0; q < 4; ++q)
          {
            const double s = gp[q & 1], t = gp[q >> 1];
            double       a00, a01, a10, a11;
            coef(X0 + (ix + s) * hx, Y0 + (iy + t) * hy, a00, a01, a10, a11);
            const double gx[4] = {-(1 - t) / hx, (1 - t) / hx, -t / hx, t / hx};
            const double gy[4] = {-(1 - s) / hy, -s / hy, (1 - s) / hy, s / hy};
            const double Nv[4] = {(1 - s) * (1 - t), s * (1 - t), (1 - s) * t, s * t};
            const double JxW   = 0.25 * hx * hy;
            for (int i = 0; i < 4; ++i)
              {
                for (int j = 0; j < 4; ++j)
                  Ke[i][j] += (gx[i] * (a00 * gx[j] + a01 * gy[j]) + gy[i] * (a10 * gx[j] + a11 * gy[j])) * JxW;
                Fe[i] += Nv[i] * f * JxW;
              }
          }
        const int v[4] = {iy * np + ix, iy * np + ix + 1, (iy + 1) * np + ix, (iy + 1) * np + ix + 1};
        for (int i = 0; i < 4; ++i)
          {
            S[ST_KC * N + v[i]] += Ke[i][i];
            S[ST_F * N + v[i]] += Fe[i];
          }
        S[ST_KE * N + v[0]] += Ke[0][1];
        S[ST_KE * N + v[2]] += Ke[2][3];
        S[ST_KN * N + v[0]] += Ke[0][2];
        S[ST_KN * N + v[1]] += Ke[1][3];
        S[ST_KD1 * N + v[0]] += Ke[0][3];
        S[ST_KD2 * N + v[0]] += Ke[1][2];
      }
  auto K = [&](int x, int y, int ex, int ey) -> double {
    const int bx = x + ex, by = y + ey;
    if (bx < 0 || bx > n || by < 0 || by > n)
      return 0.0;
    return bpx::sten_get(S.data(), np, N, x, y, ex, ey);
  };
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
    for (int c = 0; c < 4; ++c)
      {
        int pv = c;
        for (int r = c; r < 4; ++r)
          if (std::fabs(A[r][c]) > std::fabs(A[pv][c]))
            pv = r;
        for (int j = 0; j < 8; ++j)
          std::swap(A[c][j], A[pv][j]);
        const double dd = A[c][c];
        for (int j = 0; j < 8; ++j)
          A[c][j] /= dd;
        for (int r = 0; r < 4; ++r)
          if (r != c)
            {
              const double ff = A[r][c];
              for (int j = 0; j < 8; ++j)
                A[r][j] -= ff * A[c][j];
            }
      }
    for (int r = 0; r < 4; ++r)
      for (int ib = 0; ib < 4; ++ib)
        q1[r * 4 + ib] = A[r][4 + ib];
  }
  auto g = [&](int ib, int bx, int by) {
    double px, py;
    fine_vertex(corners, n, bx, by, px, py);
    return basis_q1_value(q1, ib, px, py);
  };
  // Jacobi-preconditioned CG on the condensed systems, far below the kernel's tolerance
  std::vector<double> ref((size_t)4 * N, 0.0);
  for (int ib = 0; ib < 4; ++ib)
    {
      double *x = ref.data() + (size_t)ib * N;
      for (int y = 0; y <= n; ++y)
        for (int xx = 0; xx <= n; ++xx)
          if (xx == 0 || y == 0 || xx == n || y == n)
            x[y * np + xx] = g(ib, xx, y);
      std::vector<double> r(N, 0.0), z(N, 0.0), p(N, 0.0), qv(N, 0.0);
      auto apply = [&](const std::vector<double> &u, std::vector<double> &out, bool interior_only) {
        for (int y = 1; y < n; ++y)
          for (int xx = 1; xx < n; ++xx)
            {
              double s = 0;
              for (int ey = -1; ey <= 1; ++ey)
                for (int ex = -1; ex <= 1; ++ex)
                  {
                    const int  bx = xx + ex, by = y + ey;
                    const bool bd = bx == 0 || by == 0 || bx == n || by == n;
                    if (!(interior_only && bd))
                      s += K(xx, y, ex, ey) * u[by * np + bx];
                  }
              out[y * np + xx] = s;
            }
      };
      std::vector<double> xv(x, x + N), Kg(N, 0.0);
      apply(xv, Kg, false); // K [0; g]: only boundary columns contribute
      for (int i = 0; i < N; ++i)
        r[i] = -Kg[i];
      auto dot = [&](const std::vector<double> &a, const std::vector<double> &b) {
        double s = 0;
        for (int y = 1; y < n; ++y)
          for (int xx = 1; xx < n; ++xx)
            s += a[y * np + xx] * b[y * np + xx];
        return s;
      };
      std::vector<double> xi(N, 0.0);
      double              rz = 0;
      for (int it = 0; it < 5000 && dot(r, r) > 1e-30; ++it)
        {
          for (int i = 0; i < N; ++i)
            z[i] = r[i] / S[ST_KC * N + i];
          const double rzn = dot(r, z), beta = it == 0 ? 0.0 : rzn / rz;
          rz = rzn;
          for (int i = 0; i < N; ++i)
            p[i] = z[i] + beta * p[i];
          apply(p, qv, true);
          const double alpha = rz / dot(p, qv);
          for (int y = 1; y < n; ++y)
            for (int xx = 1; xx < n; ++xx)
              xi[y * np + xx] += alpha * p[y * np + xx], r[y * np + xx] -= alpha * qv[y * np + xx];
        }
      for (int y = 1; y < n; ++y)
        for (int xx = 1; xx < n; ++xx)
          x[y * np + xx] = xi[y * np + xx];
    }
  // M = Phi^T K Phi over all DoFs, b = Phi^T F
  double Mref[16] = {}, bref[4] = {};
  for (int j = 0; j < 4; ++j)
    {
      const double       *pj = ref.data() + (size_t)j * N;
      std::vector<double> Kp(N, 0.0);
      for (int y = 0; y <= n; ++y)
        for (int xx = 0; xx <= n; ++xx)
          {
            double s = 0;
            for (int ey = -1; ey <= 1; ++ey)
              for (int ex = -1; ex <= 1; ++ex)
                if (xx + ex >= 0 && xx + ex <= n && y + ey >= 0 && y + ey <= n)
                  s += K(xx, y, ex, ey) * pj[(y + ey) * np + xx + ex];
            Kp[y * np + xx] = s;
          }
      for (int i = 0; i < 4; ++i)
        for (int a = 0; a < N; ++a)
          Mref[4 * i + j] += ref[(size_t)i * N + a] * Kp[a];
      for (int a = 0; a < N; ++a)
        bref[j] += pj[a] * S[ST_F * N + a];
    }

  // ---- the kernel, 512 (NL = 6) or 128 (NL = 5) emulated threads
  std::vector<double>  phi((size_t)4 * N, -777.0), M(16, -777.0), b(4, -777.0), res(4, -1);
  std::vector<int32_t> iters(4, -5);
  int32_t              fail[2] = {INT_MAX, 0};
  FusedParams          P;
  P.corners = corners, P.q1coef = q1, P.phi = phi.data(), P.M = M.data(), P.b = b.data();
  P.iters = iters.data(), P.res = res.data(), P.fail = fail, P.fail_base = 0, P.tol2 = 1e-24, P.max_iter = max_iter, P.n_cells = 1;
  P.rhs_value = f, P.coef = coef, P.flavor = flavor, P.split = 0;
  constexpr int T = fused::Cfg<NL>::THREADS;
  emu::Cluster  cl;
  cl.bar = std::make_unique<std::barrier<>>(T);
  cl.ctas.resize(1);
  cl.ctas[0].bar = std::make_unique<std::barrier<>>(T);
  cl.ctas[0].smem.assign(fused::Cfg<NL>::smem_bytes / 8 + 8, NAN); // poison: the kernel must initialise what it reads
  for (int w = 0; w < T / 32; ++w)
    cl.ctas[0].warps.emplace_back(new emu::Warp);
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      emu::t_cluster = &cl, emu::t_rank = 0;
      threadIdx.x = t, blockIdx.x = 0, blockDim.x = T, gridDim.x = 1;
      if (flavor == 1)
        fused::solve_fused_kernel<NL, 1>(P);
      else if (flavor == 2)
        fused::solve_fused_kernel<NL, 2>(P);
      else
        fused::solve_fused_kernel<NL, 0>(P);
    });
  for (auto &t : th)
    t.join();

  int    bad = 0;
  double worst_phi = 0, pu = 0;
  for (int ib = 0; ib < 4; ++ib)
    {
      double dd = 0, nr = 0;
      for (int i = 0; i < N; ++i)
        dd += (ref[(size_t)ib * N + i] - phi[(size_t)ib * N + i]) * (ref[(size_t)ib * N + i] - phi[(size_t)ib * N + i]),
          nr += ref[(size_t)ib * N + i] * ref[(size_t)ib * N + i];
      printf("basis %d: iters %d  reported res %.3e  rel diff vs host solve %.3e  fail %d\n", ib, iters[ib], res[ib],
             std::sqrt(dd / nr), fail[0]);
      worst_phi = std::fmax(worst_phi, std::sqrt(dd / nr));
    }
  for (int i = 0; i < N; ++i)
    pu = std::fmax(pu, std::fabs(phi[i] + phi[N + i] + phi[2 * N + i] + phi[3 * N + i] - 1.0));
  double dm = 0, nm = 0, db = 0, nb = 0;
  for (int i = 0; i < 16; ++i)
    dm = std::fmax(dm, std::fabs(M[i] - Mref[i])), nm = std::fmax(nm, std::fabs(Mref[i]));
  for (int i = 0; i < 4; ++i)
    db = std::fmax(db, std::fabs(b[i] - bref[i])), nb = std::fmax(nb, std::fabs(bref[i]));
  printf("partition of unity defect %.3e, worst basis diff %.3e\n", pu, worst_phi);
  printf("M max diff %.3e (scale %.3e)  b max diff %.3e (scale %.3e)\n", dm, nm, db, nb);
  printf("M row 0: %.15g %.15g %.15g %.15g   ref %.15g %.15g %.15g %.15g\n", M[0], M[1], M[2], M[3], Mref[0], Mref[1],
         Mref[2], Mref[3]);
  if (max_iter >= 100)
    bad = !(worst_phi < 1e-9 && pu < 1e-9 && dm < 1e-9 * nm && db < 1e-10 * nb && fail[0] == INT_MAX);
  else
    bad = !(fail[0] == 0 && iters[0] == max_iter);
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
