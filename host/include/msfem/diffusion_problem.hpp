// diffusion_problem.hpp -- the standard-FEM "truth" run of the reference
// (/root/reference/include/base/diffusion_problem.hpp / .tpp, driven from main.cxx:29-35):
// Q1 elements on the n_refine times refined unit square / cube with the same data as the multiscale
// problem -- tensor coefficient MatrixCoeff, f = RightHandSide, Dirichlet data DirichletBC on the
// even boundary ids (x_a = 0; diffusion_problem.tpp:71-80), Neumann data NeumannBC on the odd ids
// (x_a = 1; :185-214, QGauss<dim-1>(2)), CG to SolverControl(n_dofs, 1e-12) (:250).
//
// It is NOT on the accelerated path ("no performance relevance"); it exists to
// give the MsFEM-vs-fine-FEM error number.  The operator is never assembled on the host: the unit
// square is handed to the basis-stage library as ONE "coarse cell" refined n_refine times, whose
// assembly kernel produces the matrix-free fine operator and load vector; K x is then
// msb_apply_operator (device), while the boundary conditions, the Jacobi-preconditioned CG and
// the reductions run on the host.  The reference uses Trilinos AMG-CG here; the solution of the
// linear system is the same to the solver tolerance.
#pragma once

#include <chrono>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "msfem/diffusion_problem_basis.hpp"

namespace DiffusionProblem
{
  using namespace msfem;

  template <int dim>
  class DiffusionProblem
  {
  public:
    DiffusionProblem() = delete;
    explicit DiffusionProblem(unsigned int n_refine, int device_id = 0)
      : n_refine(n_refine)
      , device_id(device_id)
    {
      static_assert(dim == 2 || dim == 3, "the reference instantiates dim 2 and 3");
    }
    ~DiffusionProblem()
    {
      if (handle)
        msb_destroy(handle);
    }
    DiffusionProblem(const DiffusionProblem &)            = delete;
    DiffusionProblem &operator=(const DiffusionProblem &) = delete;

    void set_coefficient(const Coefficients::TensorCoefficient<dim> *c) { matrix_coeff = c; }
    void set_output(bool flag) { write_output = flag; }
    void set_verbose(bool flag) { verbose = flag; }

    // diffusion_problem.tpp:335-377
    void run()
    {
      if (verbose)
        std::cout << std::endl
                  << "===========================================" << std::endl
                  << "Solving >> STANDARD << problem in " << dim << "D." << std::endl;
      make_grid();
      setup_system();
      if (verbose)
        std::cout << "   Number of active cells:       " << n_cells() << std::endl
                  << "   Number of degrees of freedom: " << n_dofs() << std::endl;
      assemble_system();
      solve_iterative();
      if (write_output)
        output_results();
      if (verbose)
        std::cout << std::endl << "===========================================" << std::endl;
    }

    std::size_t  n_dofs() const { return dim == 2 ? std::size_t(np) * np : std::size_t(np) * np * np; }
    std::size_t  n_cells() const { return dim == 2 ? std::size_t(n) * n : std::size_t(n) * n * n; }
    unsigned int last_step() const { return n_iterations; }
    // nodal value at vertex (jx, jy[, jz]) of the refined mesh
    double value_at_vertex(unsigned jx, unsigned jy, unsigned jz = 0) const
    {
      return solution[dof_of_vertex[(std::size_t(jz) * np + jy) * np + jx]];
    }
    // the solution in deal.II DoF order and the vertex -> DoF map
    const std::vector<double>   &get_solution() const { return solution; }
    const std::vector<uint32_t> &get_dof_map() const { return dof_of_vertex; }

  private:
    // hyper_cube(0, 1, colorize = true) + refine_global(n_refine), diffusion_problem.tpp:43-45
    void make_grid()
    {
      if (n_refine < 1 || n_refine > (dim == 2 ? 9u : 6u))
        throw BasisStageError(MSB_ERR_UNSUPPORTED, "DiffusionProblem: the standard-FEM run supports n_refine <= 9 "
                                                   "in 2D (512 x 512 cells) and <= 6 in 3D (64^3 cells)");
      n  = 1u << n_refine;
      np = n + 1;
      h  = 1.0 / n;
    }

    void setup_system()
    {
      static const Coefficients::MatrixCoeff<dim> default_coeff;
      const Coefficients::TensorCoefficient<dim> &coeff = matrix_coeff ? *matrix_coeff : default_coeff;
      msb_config cfg{};
      cfg.abi_version    = MSB_ABI_VERSION;
      cfg.dim            = dim;
      cfg.n_refine_local = (int32_t)n_refine;
      cfg.n_cells        = 1;
      cfg.device_id      = device_id;
      cfg.tier           = MSB_TIER_AUTO;
      cfg.rhs_value      = Coefficients::RightHandSide<dim>().value(Point<dim>());
      cfg.coeff          = coeff.device_descriptor();
      if (cfg.coeff.kind == MSB_COEFF_TABLE)
        throw BasisStageError(MSB_ERR_UNSUPPORTED, "DiffusionProblem: tabulated coefficients are not wired here");
      double unit_cell[24]; // hyper_cube(0, 1) as one cell, deal.II vertex order
      for (unsigned v = 0; v < (1u << dim); ++v)
        for (unsigned a = 0; a < (unsigned)dim; ++a)
          unit_cell[dim * v + a] = (v >> a) & 1u;
      internal::check(msb_create(&cfg, unit_cell, nullptr, &handle));
      dof_of_vertex.resize(n_dofs());
      internal::check(msb_get_dof_map(handle, dof_of_vertex.data()));

      // Dirichlet values on the even boundary ids (x_a = 0), end points included
      const Coefficients::DirichletBC<dim> dirichlet_bc;
      is_constrained.assign(n_dofs(), 0);
      solution.assign(n_dofs(), 0.0);
      for (std::size_t lex = 0; lex < n_dofs(); ++lex)
        {
          Point<dim>  p;
          bool        on_dirichlet = false;
          std::size_t rest         = lex;
          for (unsigned a = 0; a < (unsigned)dim; ++a)
            {
              const unsigned j = rest % np;
              rest /= np;
              p(a) = j * h;
              on_dirichlet |= j == 0;
            }
          if (on_dirichlet)
            {
              const uint32_t d  = dof_of_vertex[lex];
              is_constrained[d] = 1;
              solution[d]       = dirichlet_bc.value(p);
            }
        }
    }

    // cell integrals come from the device (load vector); the Neumann face integrals on the odd ids
    // (x_a = 1) use QGauss<dim-1>(2) and the Q1 face shape values as the reference does
    // (diffusion_problem.tpp:185-214)
    void assemble_system()
    {
      system_rhs.resize(n_dofs());
      internal::check(msb_get_load_vector(handle, 0, system_rhs.data()));
      const Coefficients::NeumannBC<dim> neumann_bc;
      const double   g[2]     = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
      const unsigned n_face_q = 1u << (dim - 1);
      double         JxW      = 1.0;
      std::size_t    n_faces  = 1; // fine faces per side of the domain
      for (int a = 1; a < dim; ++a)
        JxW *= 0.5 * h, n_faces *= n;
      for (unsigned axis = 0; axis < (unsigned)dim; ++axis)
        for (std::size_t fc = 0; fc < n_faces; ++fc)
          {
            // tangential cell indices of this boundary face
            unsigned    ti[3] = {0, 0, 0};
            std::size_t rest  = fc;
            for (unsigned b = 0; b < (unsigned)dim; ++b)
              if (b != axis)
                ti[b] = rest % n, rest /= n;
            for (unsigned q = 0; q < n_face_q; ++q)
              {
                double     t[3] = {0, 0, 0};
                Point<dim> xq;
                unsigned   bit = 0;
                for (unsigned b = 0; b < (unsigned)dim; ++b)
                  {
                    t[b]  = b == axis ? 0.0 : g[(q >> bit++) & 1u];
                    xq(b) = b == axis ? 1.0 : (ti[b] + t[b]) * h;
                  }
                const double val = neumann_bc.value(xq) * JxW;
                for (unsigned v = 0; v < n_face_q; ++v) // the 2^(dim-1) vertices of the face
                  {
                    double      shape = 1.0;
                    std::size_t lex = 0, stride = 1;
                    unsigned    vb = 0;
                    for (unsigned b = 0; b < (unsigned)dim; ++b)
                      {
                        unsigned j;
                        if (b == axis)
                          j = n;
                        else
                          {
                            const unsigned up = (v >> vb++) & 1u;
                            shape *= up ? t[b] : 1.0 - t[b];
                            j = ti[b] + up;
                          }
                        lex += j * stride;
                        stride *= np;
                      }
                    system_rhs[dof_of_vertex[lex]] += val * shape;
                  }
              }
          }
    }

    void apply(const std::vector<double> &x, std::vector<double> &y) const
    {
      y.resize(x.size());
      internal::check(msb_apply_operator(handle, 0, x.data(), y.data()));
    }

    // CG on the unconstrained DoFs, A_ff u_f = b_f - A_fc g_c, Jacobi preconditioner, stop at
    // ||r||_2 <= 1e-12 or n_dofs steps (SolverControl(n_dofs, 1e-12), diffusion_problem.tpp:250)
    void solve_iterative()
    {
      const std::size_t   N = n_dofs();
      std::vector<double> diag(N, 0.0), probe(N), tmp(N);
      // diagonal by 2^dim-colour probing: in a 9- / 27-point stencil no two nodes of one parity
      // class (jx%2, jy%2[, jz%2]) are coupled
      const unsigned npz = dim == 3 ? np : 1;
      for (unsigned c = 0; c < (1u << dim); ++c)
        {
          std::fill(probe.begin(), probe.end(), 0.0);
          for (unsigned jz = (c >> 2) & 1u; jz < npz; jz += 2)
            for (unsigned jy = (c >> 1) & 1u; jy < np; jy += 2)
              for (unsigned jx = c & 1u; jx < np; jx += 2)
                probe[dof_of_vertex[(std::size_t(jz) * np + jy) * np + jx]] = 1.0;
          apply(probe, tmp);
          for (unsigned jz = (c >> 2) & 1u; jz < npz; jz += 2)
            for (unsigned jy = (c >> 1) & 1u; jy < np; jy += 2)
              for (unsigned jx = c & 1u; jx < np; jx += 2)
                {
                  const uint32_t d = dof_of_vertex[(std::size_t(jz) * np + jy) * np + jx];
                  diag[d]          = tmp[d];
                }
        }
      std::vector<double> r(N), z(N), p(N, 0.0), q(N);
      apply(solution, tmp); // K (0 + g_c)
      double rr = 0.0;
      for (std::size_t i = 0; i < N; ++i)
        {
          r[i] = is_constrained[i] ? 0.0 : system_rhs[i] - tmp[i];
          rr += r[i] * r[i];
        }
      const double   tol      = 1e-12;
      const unsigned max_step = (unsigned)N;
      double         rz_old   = 1.0;
      n_iterations            = 0;
      while (std::sqrt(rr) > tol && n_iterations < max_step)
        {
          double rz = 0.0;
          for (std::size_t i = 0; i < N; ++i)
            {
              z[i] = is_constrained[i] ? 0.0 : r[i] / diag[i];
              rz += r[i] * z[i];
            }
          const double beta = n_iterations == 0 ? 0.0 : rz / rz_old;
          for (std::size_t i = 0; i < N; ++i)
            p[i] = z[i] + beta * p[i];
          apply(p, q);
          double pq = 0.0;
          for (std::size_t i = 0; i < N; ++i)
            {
              if (is_constrained[i])
                q[i] = 0.0;
              pq += p[i] * q[i];
            }
          const double alpha = rz / pq;
          rr                 = 0.0;
          for (std::size_t i = 0; i < N; ++i)
            {
              solution[i] += alpha * p[i];
              r[i] -= alpha * q[i];
              rr += r[i] * r[i];
            }
          rz_old = rz;
          ++n_iterations;
        }
      if (std::sqrt(rr) > tol)
        throw NoConvergence("standard problem", 0, std::sqrt(rr), max_step);
      if (verbose)
        std::cout << "   Solved in " << n_iterations << " iterations." << std::endl;
    }

    // output_results (diffusion_problem.tpp:281-332): one piece, the reference's file names
    void output_results() const
    {
      std::ostringstream base;
      base << (dim == 2 ? "solution-std_2d" : "solution-std_3d") << "_refinements-" << n_refine;
      const std::string piece = base.str() + ".0000.vtu";
      std::ofstream     out(piece.c_str());
      out << std::setprecision(17);
      const unsigned npz = dim == 3 ? np : 1, nz = dim == 3 ? n : 1;
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" "
             "byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n<Piece NumberOfPoints=\""
          << n_dofs() << "\" NumberOfCells=\"" << n_cells()
          << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
      for (unsigned jz = 0; jz < npz; ++jz)
        for (unsigned jy = 0; jy < np; ++jy)
          for (unsigned jx = 0; jx < np; ++jx)
            out << jx * h << " " << jy * h << " " << jz * h << "\n";
      out << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
      for (unsigned iz = 0; iz < nz; ++iz)
        for (unsigned iy = 0; iy < n; ++iy)
          for (unsigned ix = 0; ix < n; ++ix)
            {
              const std::size_t b = (std::size_t(iz) * np + iy) * np + ix, up = std::size_t(np) * np;
              out << b << " " << b + 1 << " " << b + np + 1 << " " << b + np;
              if (dim == 3)
                out << " " << b + up << " " << b + up + 1 << " " << b + up + np + 1 << " " << b + up + np;
              out << "\n";
            }
      out << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
      for (std::size_t k = 1; k <= n_cells(); ++k)
        out << (1u << dim) * k << "\n";
      out << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
      for (std::size_t k = 0; k < n_cells(); ++k)
        out << (dim == 2 ? "9\n" : "12\n");
      out << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n<DataArray type=\"Float64\" Name=\"u\" "
             "format=\"ascii\">\n";
      for (std::size_t lex = 0; lex < n_dofs(); ++lex)
        out << solution[dof_of_vertex[lex]] << "\n";
      out << "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
      std::ofstream master((base.str() + ".pvtu").c_str());
      master << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" "
                "byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n<PPointData Scalars=\"scalars\">\n"
                "<PDataArray type=\"Float64\" Name=\"u\" format=\"ascii\"/>\n</PPointData>\n<PPoints>\n"
                "<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n<Piece Source=\""
             << piece << "\"/>\n</PUnstructuredGrid>\n</VTKFile>\n";
    }

    unsigned int n_refine;
    int          device_id;
    unsigned     n = 0, np = 0;
    double       h = 1.0;
    bool         write_output = false, verbose = true;
    unsigned     n_iterations = 0;

    const Coefficients::TensorCoefficient<dim> *matrix_coeff = nullptr;
    msb_handle                                  handle       = nullptr;
    std::vector<uint32_t>                       dof_of_vertex;
    std::vector<unsigned char>                  is_constrained;
    std::vector<double>                         solution, system_rhs;
  };
} // namespace DiffusionProblem
