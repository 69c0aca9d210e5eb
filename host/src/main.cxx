// main.cxx -- driver with the shape of the reference's source/main.cxx:16-84: the same
// hard-wired default problem (n_refine = 3, n_refine_local = 7, 2D; main.cxx:23-25), the
// same top-level exception handling (return 1), running the MsFEM problem whose basis
// stage executes on the B200.  With --truth the two standard-FEM runs of the reference
// (main.cxx:29-35: DiffusionProblem<2> at n_refine and at n_refine + n_refine_local) are run as
// well and the MsFEM-vs-fine-FEM difference is reported.
//
//   msfem_main [--n-refine R] [--n-refine-local L] [--coeff reference|periodic|inclusions]
//              [--dump coarse_solution.txt] [--output] [--device D] [--gpus P] [--truth]
//              [--dim 3]   the 3D block of the reference's main (main.cxx:42-55, disabled there by
//                          is_2d = true): the multiscale problem on the unit cube
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>

#include "msfem/diffusion_problem.hpp"
#include "msfem/diffusion_problem_ms.hpp"

// main.cxx:29-35 / :44-50: the standard problem on the coarse and on the fine mesh, then the MsFEM
// reconstruction u_ms = sum_i w_i phi_i on every coarse cell against the fine FEM solution at the same
// vertices (relative l2 over all (cell, vertex) pairs) and the coarse standard FEM at the coarse vertices
template <int dim>
static void
truth_runs(DiffusionProblem::DiffusionProblemMultiscale<dim> &ms, unsigned n_refine, unsigned n_refine_local,
           const Coefficients::TensorCoefficient<dim> *c, int device, bool output, double &ms_vs_fine,
           double &coarse_vs_fine)
{
  DiffusionProblem::DiffusionProblem<dim> coarse(n_refine, device);
  coarse.set_coefficient(c);
  coarse.set_output(output);
  coarse.run();
  DiffusionProblem::DiffusionProblem<dim> fine(n_refine + n_refine_local, device);
  fine.set_coefficient(c);
  fine.set_output(output);
  fine.run();

  const unsigned        nl = 1u << n_refine_local, npl = nl + 1, nc = 1u << n_refine, npz = dim == 3 ? npl : 1;
  std::vector<uint32_t> ldof;
  std::vector<double>   ums;
  double                num = 0.0, den = 0.0, num_c = 0.0, den_c = 0.0;
  for (auto &kv : ms.get_cell_basis_map())
    {
      const std::uint64_t m      = kv.first.morton();
      unsigned            idx[3] = {0, 0, 0};
      for (unsigned b = 0; b < n_refine; ++b)
        for (unsigned a = 0; a < (unsigned)dim; ++a)
          idx[a] |= unsigned((m >> (dim * b + a)) & 1u) << b;
      if (ldof.empty())
        kv.second.get_dof_map(ldof);
      kv.second.get_global_solution(ums);
      for (unsigned jz = 0; jz < npz; ++jz)
        for (unsigned jy = 0; jy < npl; ++jy)
          for (unsigned jx = 0; jx < npl; ++jx)
            {
              const double uf = fine.value_at_vertex(idx[0] * nl + jx, idx[1] * nl + jy, idx[2] * nl + jz);
              const double d  = ums[ldof[(std::size_t(jz) * npl + jy) * npl + jx]] - uf;
              num += d * d, den += uf * uf;
            }
    }
  for (unsigned jz = 0; jz <= (dim == 3 ? nc : 0); ++jz)
    for (unsigned jy = 0; jy <= nc; ++jy)
      for (unsigned jx = 0; jx <= nc; ++jx)
        {
          const double uf = fine.value_at_vertex(jx * nl, jy * nl, jz * nl);
          const double d  = coarse.value_at_vertex(jx, jy, jz) - uf;
          num_c += d * d, den_c += uf * uf;
        }
  ms_vs_fine     = std::sqrt(num / den);
  coarse_vs_fine = std::sqrt(num_c / den_c);
  std::cout << "   MsFEM reconstruction vs fine standard FEM (" << (nc * nl) << "^" << dim
            << " cells), relative l2 over all fine vertices: " << ms_vs_fine << std::endl
            << "   coarse standard FEM vs fine standard FEM, relative l2 over the coarse vertices: "
            << coarse_vs_fine << std::endl;
}

int
main(int argc, char *argv[])
{
  try
    {
      unsigned int n_refine = 3, n_refine_local = 7;
      std::string  coeff = "reference", dump;
      bool         output = false, truth = false;
      int          device = 0, gpus = 1, dim = 2;
      for (int i = 1; i < argc; ++i)
        {
          const std::string a = argv[i];
          if (a == "--n-refine" && i + 1 < argc)
            n_refine = std::atoi(argv[++i]);
          else if (a == "--n-refine-local" && i + 1 < argc)
            n_refine_local = std::atoi(argv[++i]);
          else if (a == "--coeff" && i + 1 < argc)
            coeff = argv[++i];
          else if (a == "--dump" && i + 1 < argc)
            dump = argv[++i];
          else if (a == "--device" && i + 1 < argc)
            device = std::atoi(argv[++i]);
          else if (a == "--gpus" && i + 1 < argc)
            gpus = std::atoi(argv[++i]);
          else if (a == "--output")
            output = true;
          else if (a == "--truth")
            truth = true;
          else if (a == "--dim" && i + 1 < argc)
            dim = std::atoi(argv[++i]);
          else
            throw std::runtime_error("unknown argument " + a);
        }
      if (dim != 2 && dim != 3)
        throw std::runtime_error("--dim must be 2 or 3");

      std::unique_ptr<Coefficients::TensorCoefficient<2>> c;
      if (coeff == "reference")
        c.reset(new Coefficients::MatrixCoeff<2>());
      else if (coeff == "periodic")
        c.reset(new Coefficients::PeriodicCoeff<2>(1.0 / 64));
      else if (coeff == "inclusions")
        c.reset(new Coefficients::InclusionCoeff<2>(std::ldexp(1.0, -11), 0.2, 1e4, 1.0, 1234));
      else
        throw std::runtime_error("unknown coefficient " + coeff);

      // create the CUDA contexts and load the kernels before any timed section (a one-off
      // process start-up cost of ~1 s per device that has no counterpart in the reference)
      for (int g = 0; g < gpus; ++g)
        {
          msb_config warm{};
          warm.abi_version = MSB_ABI_VERSION, warm.dim = 2, warm.n_refine_local = 1, warm.n_cells = 1;
          warm.device_id = device + g, warm.rhs_value = 2.0, warm.coeff.kind = MSB_COEFF_CONSTANT;
          warm.coeff.par[0]           = 1.0;
          const double unit_square[8] = {0, 0, 1, 0, 0, 1, 1, 1};
          msb_handle   h              = nullptr;
          if (msb_create(&warm, unit_square, nullptr, &h) != MSB_OK)
            throw std::runtime_error(std::string("device start-up failed: ") + msb_last_error());
          msb_destroy(h);
        }

      if (dim == 3)
        {
          // main.cxx:52-54 (MatrixCoeff<3> is the only coefficient the reference has in 3D)
          if (coeff != "reference")
            throw std::runtime_error("--dim 3 supports --coeff reference only");
          DiffusionProblem::DiffusionProblemMultiscale<3> diffusion_ms_problem_3d(n_refine, n_refine_local, device,
                                                                                  gpus);
          diffusion_ms_problem_3d.set_output(output);
          diffusion_ms_problem_3d.run();
          double ms_vs_fine = -1.0, coarse_vs_fine = -1.0;
          if (truth) // main.cxx:44-50
            truth_runs<3>(diffusion_ms_problem_3d, n_refine, n_refine_local, nullptr, device, output, ms_vs_fine,
                          coarse_vs_fine);
          if (!dump.empty())
            {
              std::ofstream f(dump.c_str());
              f << std::setprecision(17);
              const auto &u = diffusion_ms_problem_3d.get_solution();
              f << u.size() << "\n";
              for (double v : u)
                f << v << "\n";
              f << "basis_seconds " << diffusion_ms_problem_3d.basis_seconds() << "\n";
              if (truth)
                f << "ms_vs_fine_rel_l2 " << ms_vs_fine << "\ncoarse_vs_fine_rel_l2 " << coarse_vs_fine << "\n";
            }
          return 0;
        }

      DiffusionProblem::DiffusionProblemMultiscale<2> diffusion_ms_problem_2d(n_refine, n_refine_local, device, gpus);
      diffusion_ms_problem_2d.set_coefficient(c.get());
      diffusion_ms_problem_2d.set_output(output);
      diffusion_ms_problem_2d.run();

      double ms_vs_fine = -1.0, coarse_vs_fine = -1.0;
      if (truth)
        {
          truth_runs<2>(diffusion_ms_problem_2d, n_refine, n_refine_local, c.get(), device, output, ms_vs_fine,
                        coarse_vs_fine);
        }

      if (!dump.empty())
        {
          std::ofstream f(dump.c_str());
          f << std::setprecision(17);
          const auto &u = diffusion_ms_problem_2d.get_solution();
          f << u.size() << "\n";
          for (double v : u)
            f << v << "\n";
          f << "basis_seconds " << diffusion_ms_problem_2d.basis_seconds() << "\n";
          if (truth)
            f << "ms_vs_fine_rel_l2 " << ms_vs_fine << "\ncoarse_vs_fine_rel_l2 " << coarse_vs_fine << "\n";
        }
    }
  catch (std::exception &exc)
    {
      std::cerr << std::endl
                << "----------------------------------------------------" << std::endl
                << "Exception on processing: " << std::endl
                << exc.what() << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  catch (...)
    {
      std::cerr << std::endl
                << "----------------------------------------------------" << std::endl
                << "Unknown exception!" << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  return 0;
}
