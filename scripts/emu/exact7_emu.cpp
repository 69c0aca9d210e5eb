// exact7_emu.cpp -- host check of bpx::exact7_build (msb_bpx_common.cuh): the banded L D L^T inverse of the
// 7x7-level Galerkin operator, run by 128 emulated CUDA threads, against A * A^-1 = I.
//   g++ -O1 -std=c++20 -pthread -DMSB_EMU -I/usr/local/cuda/include -I include \
//       -I mpi_parallel_multiscale_diffusion_fem_b200/csrc scripts/emu/exact7_emu.cpp -o /tmp/exact7_emu
#include "emu_shims.hpp"
#include "msb_bpx_common.cuh"

using namespace msb;

int
main(int argc, char **argv)
{
  const double contrast = argc > 1 ? atof(argv[1]) : 1e4;
  // a 9-point SPD operator on 9 x 9 nodes: Q1 stiffness of a piecewise-constant coefficient with jumps
  const int np = 9, N = 81, n = 8;
  std::vector<double> G((size_t)ST_NARR * N, 0.0);
  unsigned seed = 12345;
  for (int iy = 0; iy < n; ++iy)
    for (int ix = 0; ix < n; ++ix)
      {
        seed = seed * 1664525u + 1013904223u;
        const double a = (seed >> 28) < 4 ? contrast : 1.0;
        // Q1 stiffness of a square cell: diag 2/3, edge -1/6, diagonal -1/3
        const int v[4] = {iy * np + ix, iy * np + ix + 1, (iy + 1) * np + ix, (iy + 1) * np + ix + 1};
        for (int i = 0; i < 4; ++i)
          G[ST_KC * N + v[i]] += a * 2.0 / 3.0;
        G[ST_KE * N + v[0]] += a * -1.0 / 6.0, G[ST_KE * N + v[2]] += a * -1.0 / 6.0;
        G[ST_KN * N + v[0]] += a * -1.0 / 6.0, G[ST_KN * N + v[1]] += a * -1.0 / 6.0;
        G[ST_KD1 * N + v[0]] += a * -1.0 / 3.0;
        G[ST_KD2 * N + v[0]] += a * -1.0 / 3.0;
      }
  const int T = 128;
  emu::Cluster cl;
  cl.bar = std::make_unique<std::barrier<>>(T);
  cl.ctas.resize(1);
  cl.ctas[0].bar = std::make_unique<std::barrier<>>(T);
  cl.ctas[0].smem.assign(49 * 49 + bpx::EXACT7_SCRATCH, NAN);
  for (int w = 0; w < T / 32; ++w)
    cl.ctas[0].warps.emplace_back(new emu::Warp);
  double *sGi = cl.ctas[0].smem.data(), *sBand = sGi + 49 * 49;
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      emu::t_cluster = &cl, emu::t_rank = 0;
      threadIdx.x = t, blockDim.x = T;
      bpx::exact7_build<T>(G.data(), sGi, sBand, t);
    });
  for (auto &t : th)
    t.join();
  // A * Ainv against the identity (A = interior 7 x 7 block of the stencil operator)
  double worst = 0, asym = 0, amax = 0;
  for (int i = 0; i < 49; ++i)
    for (int j = 0; j < 49; ++j)
      {
        const int X = 1 + i % 7, Y = 1 + i / 7;
        double    s = 0;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx)
            {
              const int bx = X + dx, by = Y + dy;
              if (bx >= 1 && bx <= 7 && by >= 1 && by <= 7)
                s += bpx::sten_get(G.data(), np, N, X, Y, dx, dy) * sGi[((by - 1) * 7 + bx - 1) * 49 + j];
            }
        worst = std::fmax(worst, std::fabs(s - (i == j)));
        asym  = std::fmax(asym, std::fabs(sGi[i * 49 + j] - sGi[j * 49 + i]));
        amax  = std::fmax(amax, std::fabs(sGi[i * 49 + j]));
      }
  printf("contrast %g: max |A Ainv - I| = %.3e, asymmetry %.3e (max entry %.3e)\n", contrast, worst, asym, amax);
  // the packed lower-triangle variant (exact7_build<T, true> + Exact7::fetch_tri) must deliver the same entries
  std::vector<double> full(sGi, sGi + 49 * 49);
  cl.ctas[0].smem.assign(49 * 49 + bpx::EXACT7_SCRATCH, NAN);
  sGi = cl.ctas[0].smem.data(), sBand = sGi + 49 * 49;
  th.clear();
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      emu::t_cluster = &cl, emu::t_rank = 0;
      threadIdx.x = t, blockDim.x = T;
      bpx::exact7_build<T, true>(G.data(), sGi, sBand, t);
    });
  for (auto &t : th)
    t.join();
  double dtri = 0;
  for (int r = 0; r < 49; ++r)
    for (int j = 0; j < 49; ++j)
      dtri = std::fmax(dtri, std::fabs(bpx::Exact7<T>::fetch_tri(sGi, r, j / 8, j % 8) - full[j * 49 + r]));
  printf("packed variant: max difference to the full inverse %.3e\n", dtri);
  return worst < 1e-10 && asym < 1e-12 * amax && dtri <= 1e-12 * amax ? 0 : 1;
}
