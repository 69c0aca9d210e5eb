// main.cxx -- driver with the shape of the reference's source/main.cxx:16-84: the same
// hard-wired default problem (n_refine = 3, n_refine_local = 7, 2D; main.cxx:23-25), the
// same top-level exception handling (return 1), running the MsFEM problem whose basis
// stage executes on the B200.  The two standard-FEM "truth" runs of the reference
// (main.cxx:29-35) are out of scope (component 4).
//
//   msfem_main [--n-refine R] [--n-refine-local L] [--coeff reference|periodic|inclusions]
//              [--dump coarse_solution.txt] [--output] [--device D] [--gpus P]
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>

#include "msfem/diffusion_problem_ms.hpp"

int
main(int argc, char *argv[])
{
  try
    {
      unsigned int n_refine = 3, n_refine_local = 7;
      std::string  coeff = "reference", dump;
      bool         output = false;
      int          device = 0, gpus = 1;
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
          else
            throw std::runtime_error("unknown argument " + a);
        }

      std::unique_ptr<Coefficients::TensorCoefficient<2>> c;
      if (coeff == "reference")
        c.reset(new Coefficients::MatrixCoeff<2>());
      else if (coeff == "periodic")
        c.reset(new Coefficients::PeriodicCoeff<2>(1.0 / 64));
      else if (coeff == "inclusions")
        c.reset(new Coefficients::InclusionCoeff<2>(std::ldexp(1.0, -11), 0.2, 1e4, 1.0, 1234));
      else
        throw std::runtime_error("unknown coefficient " + coeff);

      DiffusionProblem::DiffusionProblemMultiscale<2> diffusion_ms_problem_2d(n_refine, n_refine_local, device, gpus);
      diffusion_ms_problem_2d.set_coefficient(c.get());
      diffusion_ms_problem_2d.set_output(output);
      diffusion_ms_problem_2d.run();

      if (!dump.empty())
        {
          std::ofstream f(dump.c_str());
          f << std::setprecision(17);
          const auto &u = diffusion_ms_problem_2d.get_solution();
          f << u.size() << "\n";
          for (double v : u)
            f << v << "\n";
          f << "basis_seconds " << diffusion_ms_problem_2d.basis_seconds() << "\n";
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
