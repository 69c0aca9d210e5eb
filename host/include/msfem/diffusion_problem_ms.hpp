// diffusion_problem_ms.hpp -- the caller of the basis stage, shaped like the reference's
// DiffusionProblemMultiscale<dim> (/root/reference/include/base/diffusion_problem_ms.{hpp,tpp})
// so that the GPU path can be exercised end to end on a machine without deal.II, MPI,
// p4est and Trilinos.  Method names and the order of run() follow ms.tpp:445-491.
//
// What is NOT the reference here (and is out of the accelerated path):
// the coarse mesh is the structured (2^r)^dim refinement of hyper_cube(0,1,colorize) held
// as plain arrays in Morton / CellId order (ms.tpp:97-99), the coarse matrix is a host CSR,
// and the coarse system is solved by a host Jacobi-PCG to the reference's tolerance
// (SolverControl(n_dofs, 1e-12), ms.tpp:264) instead of Trilinos AMG-CG.  One process
// drives one GPU; with several GPUs each rank owns a contiguous Morton range (ms.tpp:52).
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "msfem/coefficients.hpp"
#include "msfem/diffusion_problem_basis.hpp"

namespace DiffusionProblem
{
  using namespace msfem;

  template <int dim>
  class DiffusionProblemMultiscale
  {
    static_assert(dim == 2 || dim == 3, "the reference instantiates dim 2 and 3");
    static constexpr unsigned NB = 1u << dim; // vertices (= multiscale bases) per coarse cell

  public:
    DiffusionProblemMultiscale(unsigned int n_refine, unsigned int n_refine_local, int device_id = 0,
                               int n_devices = 1)
      : n_refine(n_refine)
      , n_refine_local(n_refine_local)
      , device_id(device_id)
      , n_devices(n_devices < 1 ? 1 : n_devices)
    {}

    void set_coefficient(const Coefficients::TensorCoefficient<dim> *c) { matrix_coeff = c; }
    void set_output(bool coarse_and_fine) { write_output = coarse_and_fine; }

    // ms.tpp:445-491
    void run()
    {
      std::cout << std::endl
                << "===========================================" << std::endl
                << "Solving >> MULTISCALE << problem in " << dim << "D." << std::endl;
      std::cout << "Running with the B200 basis stage on " << n_devices << " GPU(s)..." << std::endl;
      make_grid();
      setup_system();
      timed("basis initialization and computation", [&] { initialize_and_compute_basis(); });
      timed("global multiscale assembly", [&] { assemble_system(); });
      timed("global iterative solver", [&] { solve_iterative(); });
      send_global_weights_to_cell();
      if (write_output)
        {
          timed("coarse output vtu", [&] { output_global_coarse(); });
          timed("fine output vtu", [&] { output_global_fine(); });
        }
      print_summary();
      std::cout << std::endl << "===========================================" << std::endl;
    }

    const std::vector<double> &get_solution() const { return solution; }
    unsigned int                n_dofs() const { return n_coarse_dofs; }
    const std::vector<unsigned> &get_dof_map() const { return dof_of_vertex; }
    double basis_seconds() const { return timings.count("basis initialization and computation") ?
                                            timings.at("basis initialization and computation") : 0.0; }
    std::map<CellId, DiffusionProblemBasis<dim>> &get_cell_basis_map() { return cell_basis_map; }

  private:
    template <class F>
    void timed(const std::string &name, F f)
    {
      const auto t0 = std::chrono::steady_clock::now();
      f();
      timings[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    // coordinate index `axis` of the coarse cell with Morton index m (bits interleaved x, y[, z])
    unsigned cell_index(std::uint64_t m, unsigned axis) const
    {
      unsigned v = 0;
      for (unsigned b = 0; b < n_refine; ++b)
        v |= unsigned((m >> (dim * b + axis)) & 1u) << b;
      return v;
    }
    std::size_t n_cells_total() const
    {
      std::size_t t = 1;
      for (int a = 0; a < dim; ++a)
        t *= nc;
      return t;
    }
    std::size_t n_vertices_total() const
    {
      std::size_t t = 1;
      for (int a = 0; a < dim; ++a)
        t *= nc + 1;
      return t;
    }
    // lexicographic index of vertex v of cell m in the (nc+1)^dim vertex grid
    std::size_t vertex_lex(std::uint64_t m, unsigned v) const
    {
      std::size_t lex = 0, stride = 1;
      for (unsigned a = 0; a < (unsigned)dim; ++a)
        {
          lex += (cell_index(m, a) + ((v >> a) & 1u)) * stride;
          stride *= nc + 1;
        }
      return lex;
    }

    // ms.tpp:91-103: hyper_cube(0,1,colorize=true), refine_global(n_refine); active cells in
    // Morton order (= CellId order = p4est order)
    void make_grid()
    {
      nc = 1u << n_refine;
      H  = 1.0 / nc;
      cells.resize(n_cells_total());
      for (std::uint64_t m = 0; m < cells.size(); ++m)
        {
          CoarseCell<dim> &c = cells[m];
          for (unsigned v = 0; v < NB; ++v)
            for (unsigned a = 0; a < (unsigned)dim; ++a)
              c.vertices[v](a) = (cell_index(m, a) + ((v >> a) & 1u)) * H;
          c.cell_id = CellId(n_refine, m, dim);
          for (unsigned a = 0; a < (unsigned)dim; ++a)
            {
              c.boundary_id[2 * a]     = cell_index(m, a) == 0 ? 2 * a : 255;          // x_a = 0
              c.boundary_id[2 * a + 1] = cell_index(m, a) == nc - 1 ? 2 * a + 1 : 255; // x_a = 1
            }
        }
      std::cout << "Number of active global cells: " << cells.size() << std::endl;
    }

    // ms.tpp:106-156: first-touch DoF numbering, Dirichlet values on the even boundary ids
    // (x_a = 0 for every axis a, ms.tpp:129-139)
    void setup_system()
    {
      const unsigned np = nc + 1;
      dof_of_vertex.assign(n_vertices_total(), ~0u);
      unsigned next = 0;
      for (std::uint64_t m = 0; m < cells.size(); ++m)
        for (unsigned v = 0; v < NB; ++v)
          {
            unsigned &d = dof_of_vertex[vertex_lex(m, v)];
            if (d == ~0u)
              d = next++;
          }
      n_coarse_dofs = next;
      is_constrained.assign(n_coarse_dofs, 0);
      constraint_value.assign(n_coarse_dofs, 0.0);
      const Coefficients::DirichletBC<dim> dirichlet_bc;
      for (std::size_t lex = 0; lex < dof_of_vertex.size(); ++lex)
        {
          Point<dim>  p;
          bool        on_dirichlet = false;
          std::size_t rest         = lex;
          for (unsigned a = 0; a < (unsigned)dim; ++a)
            {
              const unsigned j = rest % np;
              rest /= np;
              p(a) = j * H;
              on_dirichlet |= j == 0;
            }
          if (on_dirichlet)
            {
              const unsigned d    = dof_of_vertex[lex];
              is_constrained[d]   = 1;
              constraint_value[d] = dirichlet_bc.value(p);
            }
        }
      solution.assign(n_coarse_dofs, 0.0);
    }

    // ms.tpp:40-88 -- THE stage.  Construction loop as in the reference; the serial run()
    // loop (ms.tpp:81-87) is replaced by one batched call.
    void initialize_and_compute_basis()
    {
      for (std::size_t m = 0; m < cells.size(); ++m)
        {
          DiffusionProblemBasis<dim> current_cell_problem(n_refine_local, cells[m], /*subdomain*/ 0, 0);
          cell_basis_map.insert(std::make_pair(cells[m].id(), current_cell_problem));
        }
      const Coefficients::TensorCoefficient<dim> &coeff =
        matrix_coeff ? *matrix_coeff : DiffusionProblemBasis<dim>::default_coefficient();
      if (n_devices == 1)
        {
          DiffusionProblemBasis<dim>::run_all(cell_basis_map, coeff, BasisSolverControl(), device_id);
          return;
        }
      // one host thread and one GPU batch per device, contiguous Morton ranges
      // [C g / P, C (g+1) / P) -- what "mpirun -n P" does to the reference (ms.tpp:52)
      const std::size_t               C = cell_basis_map.size();
      std::vector<std::thread>        workers;
      std::vector<std::exception_ptr> errors(n_devices);
      auto                            it = cell_basis_map.begin();
      std::size_t                     pos = 0;
      for (int g = 0; g < n_devices; ++g)
        {
          const std::size_t hi    = C * (g + 1) / n_devices;
          auto              first = it;
          while (pos < hi)
            ++it, ++pos;
          auto last = it;
          workers.emplace_back([=, &coeff, &errors]() {
            try
              {
                DiffusionProblemBasis<dim>::run_range(first, last, coeff, BasisSolverControl(), device_id + g);
              }
            catch (...)
              {
                errors[g] = std::current_exception();
              }
          });
        }
      for (auto &w : workers)
        w.join();
      for (auto &e : errors)
        if (e)
          std::rethrow_exception(e);
    }

    void cell_dofs(std::size_t m, unsigned (&ld)[NB]) const
    {
      for (unsigned v = 0; v < NB; ++v)
        ld[v] = dof_of_vertex[vertex_lex(m, v)];
    }

    // ms.tpp:159-255: element matrices from the basis objects, Neumann face terms on the odd
    // boundary ids (QGauss<dim-1>(2), standard Q1 face shape values), scatter into the coarse system
    void assemble_system()
    {
      rows.assign(n_coarse_dofs, {});
      system_rhs.assign(n_coarse_dofs, 0.0);
      const Coefficients::NeumannBC<dim> neumann_bc;
      const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
      const unsigned n_face_q = 1u << (dim - 1);
      double         JxW      = 1.0;
      for (int a = 1; a < dim; ++a)
        JxW *= 0.5 * H;
      std::size_t m = 0;
      for (auto &kv : cell_basis_map)
        {
          const FullMatrix<double> &cell_matrix = kv.second.get_global_element_matrix();
          Vector<double>            cell_rhs    = kv.second.get_global_element_rhs();
          const CoarseCell<dim>    &c           = cells[m];
          for (unsigned axis = 0; axis < (unsigned)dim; ++axis)
            {
              const unsigned face = 2 * axis + 1; // x_axis = 1 side of the cell
              if (c.boundary_id[face] != face)
                continue;
              for (unsigned q = 0; q < n_face_q; ++q)
                {
                  // quadrature point: tensor Gauss point in the tangential axes
                  double     t[3] = {0, 0, 0};
                  Point<dim> xq;
                  unsigned   bit = 0;
                  for (unsigned b = 0; b < (unsigned)dim; ++b)
                    {
                      t[b]  = b == axis ? 1.0 : g[(q >> bit++) & 1u];
                      xq(b) = c.vertex(0)(b) + t[b] * H;
                    }
                  const double val = neumann_bc.value(xq);
                  for (unsigned v = 0; v < NB; ++v)
                    {
                      if (!((v >> axis) & 1u))
                        continue; // only the vertices of this face carry a face shape function
                      double shape = 1.0;
                      for (unsigned b = 0; b < (unsigned)dim; ++b)
                        if (b != axis)
                          shape *= ((v >> b) & 1u) ? t[b] : 1.0 - t[b];
                      cell_rhs(v) += val * shape * JxW;
                    }
                }
            }
          unsigned ld[NB];
          cell_dofs(m, ld);
          for (unsigned i = 0; i < NB; ++i)
            {
              for (unsigned j = 0; j < NB; ++j)
                rows[ld[i]][ld[j]] += cell_matrix(i, j);
              system_rhs[ld[i]] += cell_rhs(i);
            }
          ++m;
        }
    }

    // ms.tpp:258-296: CG to ||r||_2 <= 1e-12 on the system with the Dirichlet DoFs eliminated,
    // then constraints.distribute
    void solve_iterative()
    {
      const unsigned      n = n_coarse_dofs;
      std::vector<double> b(n, 0.0), x(n, 0.0), r(n), z(n), p(n), q(n), dinv(n, 1.0);
      for (unsigned i = 0; i < n; ++i)
        {
          if (is_constrained[i])
            continue;
          double bi = system_rhs[i];
          for (auto &e : rows[i])
            if (is_constrained[e.first])
              bi -= e.second * constraint_value[e.first];
          b[i]    = bi;
          dinv[i] = 1.0 / rows[i][i];
        }
      auto vmult = [&](const std::vector<double> &in, std::vector<double> &out) {
        for (unsigned i = 0; i < n; ++i)
          {
            double s = 0.0;
            if (!is_constrained[i])
              for (auto &e : rows[i])
                if (!is_constrained[e.first])
                  s += e.second * in[e.first];
            out[i] = s;
          }
      };
      auto dot = [&](const std::vector<double> &a, const std::vector<double> &c) {
        double s = 0.0;
        for (unsigned i = 0; i < n; ++i)
          s += a[i] * c[i];
        return s;
      };
      r = b;
      for (unsigned i = 0; i < n; ++i)
        z[i] = dinv[i] * r[i];
      p          = z;
      double   rz = dot(r, z);
      unsigned it = 0;
      while (std::sqrt(dot(r, r)) > 1e-12 && it < 10 * n)
        {
          ++it;
          vmult(p, q);
          const double alpha = rz / dot(p, q);
          for (unsigned i = 0; i < n; ++i)
            x[i] += alpha * p[i], r[i] -= alpha * q[i];
          for (unsigned i = 0; i < n; ++i)
            z[i] = dinv[i] * r[i];
          const double rz_new = dot(r, z), beta = rz_new / rz;
          rz                  = rz_new;
          for (unsigned i = 0; i < n; ++i)
            p[i] = z[i] + beta * p[i];
        }
      std::cout << "   Global problem solved in " << it << " iterations." << std::endl;
      for (unsigned i = 0; i < n; ++i)
        solution[i] = is_constrained[i] ? constraint_value[i] : x[i];
    }

    // ms.tpp:299-325
    void send_global_weights_to_cell()
    {
      std::size_t m = 0;
      for (auto &kv : cell_basis_map)
        {
          unsigned ld[NB];
          cell_dofs(m++, ld);
          std::vector<double> extracted_weights(NB);
          for (unsigned i = 0; i < NB; ++i)
            extracted_weights[i] = solution[ld[i]];
          kv.second.set_global_weights(extracted_weights);
        }
    }

    // ms.tpp:328-379 (one rank: a single VTU)
    void output_global_coarse() const
    {
      const unsigned    np = nc + 1, npz = dim == 3 ? np : 1, ncz = dim == 3 ? nc : 1;
      const std::string d  = dim == 2 ? "2d" : "3d";
      std::ofstream     out("solution-ms_coarse-" + d + "_refinements-" + std::to_string(n_refine) + ".0000.vtu");
      out.precision(17);
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" "
             "byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n<Piece NumberOfPoints=\""
          << n_vertices_total() << "\" NumberOfCells=\"" << n_cells_total()
          << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
      for (unsigned jz = 0; jz < npz; ++jz)
        for (unsigned jy = 0; jy < np; ++jy)
          for (unsigned jx = 0; jx < np; ++jx)
            out << jx * H << " " << jy * H << " " << jz * H << "\n";
      out << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" "
             "format=\"ascii\">\n";
      for (unsigned iz = 0; iz < ncz; ++iz)
        for (unsigned iy = 0; iy < nc; ++iy)
          for (unsigned ix = 0; ix < nc; ++ix)
            {
              const std::size_t b = (std::size_t(iz) * np + iy) * np + ix, up = std::size_t(np) * np;
              out << b << " " << b + 1 << " " << b + np + 1 << " " << b + np;
              if (dim == 3)
                out << " " << b + up << " " << b + up + 1 << " " << b + up + np + 1 << " " << b + up + np;
              out << "\n";
            }
      out << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
      for (std::size_t k = 1; k <= n_cells_total(); ++k)
        out << NB * k << "\n";
      out << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
      for (std::size_t k = 0; k < n_cells_total(); ++k)
        out << (dim == 2 ? "9\n" : "12\n");
      out << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n<DataArray type=\"Float64\" "
             "Name=\"u\" format=\"ascii\">\n";
      for (std::size_t i = 0; i < n_vertices_total(); ++i)
        out << solution[dof_of_vertex[i]] << "\n";
      out << "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
    }

    // ms.tpp:382-423: one VTU per coarse cell + a PVTU record naming them
    void output_global_fine()
    {
      std::vector<std::string> filenames;
      // one bulk device->host transfer per GPU batch (msb_get_global_solutions) instead of one
      // launch + synchronisation per coarse cell; repeated calls on the same batch are no-ops
      for (auto &kv : cell_basis_map)
        kv.second.prefetch_global_solutions();
      for (auto &kv : cell_basis_map)
        {
          kv.second.output_global_solution_in_cell();
          filenames.push_back(kv.second.get_filename_global());
        }
      std::ofstream out(dim == 2 ? "solution-ms_fine-2d.pvtu" : "solution-ms_fine-3d.pvtu");
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" "
             "byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n<PPointData "
             "Scalars=\"scalars\">\n<PDataArray type=\"Float64\" Name=\"solution\" format=\"ascii\"/>\n"
             "</PPointData>\n<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n"
             "</PPoints>\n";
      for (auto &f : filenames)
        out << "<Piece Source=\"" << f << "\"/>\n";
      out << "</PUnstructuredGrid>\n</VTKFile>\n";
    }

    void print_summary() const
    {
      std::cout << "\n+---------------------------------------------+------------+\n"
                << "| Section                                     | wall time  |\n"
                << "+---------------------------------------------+------------+\n";
      for (auto &kv : timings)
        {
          std::string name = kv.first;
          name.resize(43, ' ');
          std::cout << "| " << name << " | " << kv.second << "s |\n";
        }
      std::cout << "+---------------------------------------------+------------+\n";
    }

    unsigned int n_refine, n_refine_local;
    int          device_id, n_devices;
    unsigned     nc = 0, n_coarse_dofs = 0;
    double       H = 1.0;
    bool         write_output = false;
    const Coefficients::TensorCoefficient<dim> *matrix_coeff = nullptr;

    std::vector<CoarseCell<dim>>                 cells;
    std::vector<unsigned>                        dof_of_vertex;
    std::vector<char>                            is_constrained;
    std::vector<double>                          constraint_value, system_rhs, solution;
    std::vector<std::map<unsigned, double>>      rows;
    std::map<CellId, DiffusionProblemBasis<dim>> cell_basis_map;
    std::map<std::string, double>                timings;
  };
} // namespace DiffusionProblem
