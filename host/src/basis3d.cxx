// basis3d.cxx -- the dim = 3 basis stage through the C++ mirror of the reference interface:
// builds the (2^R)^3 coarse mesh of the unit cube in CellId (3D Morton) order, creates one
// DiffusionProblemBasis<3> per coarse cell as ms.tpp:50-73 does, and runs them all in one GPU
// batch (the replacement of ms.tpp:81-87).  The reference instantiates the class for dim 3
// (diffusion_problem_basis.inst.cc:15-16) but its main() only drives 2D; this program is the
// 3D counterpart used by the tests.
//
//   msfem_basis3d [--n-refine R] [--n-refine-local L] [--dump file] [--output] [--device D]
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>

#include "msfem/diffusion_problem_basis.hpp"

int
main(int argc, char *argv[])
{
  try
    {
      unsigned    n_refine = 1, n_refine_local = 3;
      std::string dump;
      bool        output = false;
      int         device = 0;
      for (int i = 1; i < argc; ++i)
        {
          const std::string a = argv[i];
          if (a == "--n-refine" && i + 1 < argc)
            n_refine = std::atoi(argv[++i]);
          else if (a == "--n-refine-local" && i + 1 < argc)
            n_refine_local = std::atoi(argv[++i]);
          else if (a == "--dump" && i + 1 < argc)
            dump = argv[++i];
          else if (a == "--device" && i + 1 < argc)
            device = std::atoi(argv[++i]);
          else if (a == "--output")
            output = true;
          else
            throw std::runtime_error("unknown argument " + a);
        }
      using Basis = DiffusionProblem::DiffusionProblemBasis<3>;
      const unsigned            nc = 1u << n_refine;
      const double              H  = 1.0 / nc;
      std::map<msfem::CellId, Basis> cell_basis_map;
      for (std::uint64_t m = 0; m < std::uint64_t(nc) * nc * nc; ++m)
        {
          unsigned idx[3] = {0, 0, 0};
          for (unsigned bit = 0; bit < n_refine; ++bit)
            for (unsigned a = 0; a < 3; ++a)
              idx[a] |= unsigned((m >> (3 * bit + a)) & 1u) << bit;
          msfem::CoarseCell<3> cell;
          for (unsigned v = 0; v < 8; ++v)
            cell.vertices[v] =
              msfem::Point<3>((idx[0] + (v & 1)) * H, (idx[1] + ((v >> 1) & 1)) * H, (idx[2] + (v >> 2)) * H);
          cell.cell_id = msfem::CellId(n_refine, m, 3);
          for (unsigned f = 0; f < 6; ++f)
            cell.boundary_id[f] = 255;
          cell_basis_map.emplace(cell.id(), Basis(n_refine_local, cell, 0));
        }
      if (output)
        for (auto &kv : cell_basis_map)
          kv.second.set_output_flag(true);
      Basis::run_all(cell_basis_map, Basis::default_coefficient(), DiffusionProblem::BasisSolverControl(),
                     device);

      double   checksum = 0.0;
      unsigned max_it   = 0;
      for (auto &kv : cell_basis_map)
        {
          const auto &M = kv.second.get_global_element_matrix();
          for (unsigned i = 0; i < 8; ++i)
            {
              checksum += M(i, i);
              max_it = std::max(max_it, kv.second.last_step(i));
            }
        }
      std::cout << "3D basis stage: " << cell_basis_map.size() << " coarse cells, " << 8 * cell_basis_map.size()
                << " local solves, max fine CG iterations " << max_it << ", trace sum " << std::setprecision(15)
                << checksum << std::endl;
      if (!dump.empty())
        {
          std::ofstream out(dump.c_str());
          out << std::setprecision(17);
          for (auto &kv : cell_basis_map)
            {
              out << kv.first.to_string();
              const auto &M = kv.second.get_global_element_matrix();
              const auto &b = kv.second.get_global_element_rhs();
              for (unsigned i = 0; i < 8; ++i)
                for (unsigned j = 0; j < 8; ++j)
                  out << " " << M(i, j);
              for (unsigned i = 0; i < 8; ++i)
                out << " " << b(i);
              out << "\n";
            }
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
  return 0;
}
