// diffusion_problem_basis.hpp -- host C++ mirror of the reference's per-cell basis object
// (/root/reference/include/base/diffusion_problem_basis.hpp:68-281) on top of the C ABI
// (include/msfem_basis.h).  Same public names, argument meaning and error behaviour:
//
//   ctor(n_refine_local, global_cell, local_subdomain, mpi_communicator)   basis.tpp:18-50
//   copy ctor (objects are copied into the std::map while still empty)     basis.tpp:53-87
//   run()                                                                  basis.tpp:438-474
//   get_global_element_matrix() / get_global_element_rhs()                 basis.tpp:320-333
//   set_global_weights(weights)                                            basis.tpp:352-377
//   output_global_solution_in_cell()                                       basis.tpp:421-435
//   get_filename_global() / set_output_flag(flag)                          basis.tpp:336-349
//
// A non-converged local solve throws DiffusionProblem::NoConvergence out of run() exactly
// where the reference's SolverCG throws SolverControl::NoConvergence (basis.tpp:303).
//
// What is new is run_all(): the reference runs its objects one by one (ms.tpp:81-87);
// run_all() hands the whole std::map to ONE GPU batch.  run() on a single object is a batch
// of one and exists for interface parity, not for speed.
#pragma once

#include <algorithm>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "msfem/coefficients.hpp"
#include "msfem/shims.hpp"
#include "msfem_basis.h"

namespace DiffusionProblem
{
  using namespace msfem;

  class NoConvergence : public std::runtime_error
  {
  public:
    NoConvergence(const std::string &cell, unsigned basis, double residual, unsigned last_step)
      : std::runtime_error("Iterative method reported convergence failure in step " +
                           std::to_string(last_step) + " (cell " + cell + ", basis " +
                           std::to_string(basis) + "). The residual in the last step was " +
                           std::to_string(residual) + ".")
      , last_step(last_step)
      , last_residual(residual)
    {}
    unsigned last_step;
    double   last_residual;
  };

  class BasisStageError : public std::runtime_error
  {
  public:
    BasisStageError(int code, const std::string &what)
      : std::runtime_error("msfem_basis error " + std::to_string(code) + ": " + what)
      , code(code)
    {}
    int code;
  };

  // solver parameters of solve_iterative (basis.tpp:297: SolverControl(1000, 1e-12)).  The GPU
  // iterates CG with a multilevel preconditioner instead of SSOR (whose sweeps are sequential);
  // it needs fewer iterations than the reference, but the cap stays a parameter.  The stopping
  // rule is unchanged: absolute l2 norm of the unpreconditioned residual, every iteration.
  struct BasisSolverControl
  {
    double   tolerance = 1e-12;
    unsigned max_steps = 5000;
  };

  namespace internal
  {
    inline void
    check(int rc)
    {
      if (rc != MSB_OK)
        throw BasisStageError(rc, msb_last_error());
    }

    // one GPU batch shared by the basis objects it serves
    template <int dim>
    class Batch
    {
    public:
      Batch(unsigned n_refine_local, const std::vector<double> &corners, const msb_coeff_desc &coeff,
            const std::vector<double> &table, double rhs_value, int device_id)
        : n_cells(corners.size() / (dim * (1 << dim)))
        , n_refine_local(n_refine_local)
        , weights((std::size_t(1) << dim) * n_cells, 0.0)
      {
        msb_config cfg{};
        cfg.abi_version    = MSB_ABI_VERSION;
        cfg.dim            = dim;
        cfg.n_refine_local = (int32_t)n_refine_local;
        cfg.n_cells        = (int32_t)n_cells;
        cfg.device_id      = device_id;
        cfg.tier           = MSB_TIER_AUTO;
        cfg.rhs_value      = rhs_value;
        cfg.coeff          = coeff;
        check(msb_create(&cfg, corners.data(), table.empty() ? nullptr : table.data(), &handle));
      }
      ~Batch() { msb_destroy(handle); }
      Batch(const Batch &)            = delete;
      Batch &operator=(const Batch &) = delete;

      void upload_weights()
      {
        if (weights_dirty)
          {
            check(msb_set_global_weights(handle, weights.data()));
            weights_dirty = false;
            gsol_valid    = false;
          }
      }

      // global_solution of EVERY cell of the batch with one bulk call (msb_get_global_solutions)
      // instead of one reordering launch + synchronisation per cell: what the walk over all cells
      // in output_global_fine (ms.tpp:386-393) needs
      void fetch_global_solutions(std::size_t n_dofs)
      {
        upload_weights();
        if (!gsol_valid)
          {
            gsol_cache.resize(n_cells * n_dofs);
            check(msb_get_global_solutions(handle, 0, (int32_t)n_cells, gsol_cache.data()));
            gsol_valid = true;
          }
      }

      msb_handle          handle = nullptr;
      std::size_t         n_cells;
      unsigned            n_refine_local;
      std::vector<double> weights;
      bool                weights_dirty = true;
      std::vector<double> gsol_cache; // [n_cells][N], deal.II DoF order
      bool                gsol_valid = false;
    };
  } // namespace internal

  template <int dim>
  class DiffusionProblemBasis
  {
  public:
    DiffusionProblemBasis() = delete;
    static constexpr unsigned NB = 1u << dim; // GeometryInfo<dim>::vertices_per_cell

#ifdef MSFEM_WITH_DEALII
    // The reference's constructor, argument for argument (diffusion_problem_basis.hpp:79-83), so that the
    // construction loop diffusion_problem_ms.tpp:54-66 compiles against this class unchanged.
    DiffusionProblemBasis(unsigned int                                               n_refine_local,
                          typename dealii::Triangulation<dim>::active_cell_iterator &global_cell,
                          unsigned int local_subdomain, MPI_Comm mpi_communicator)
      : DiffusionProblemBasis(n_refine_local, CoarseCell<dim>::from(global_cell), local_subdomain, mpi_communicator)
    {}
#endif
    DiffusionProblemBasis(unsigned int n_refine_local, const CoarseCell<dim> &global_cell,
                          unsigned int local_subdomain, MPI_Comm_shim mpi_communicator = MPI_Comm_shim())
      : mpi_communicator(mpi_communicator)
      , corner_points(1 << dim)
      , filename_global("")
      , global_element_matrix(1 << dim, 1 << dim)
      , is_built_global_element_matrix(false)
      , global_element_rhs(1 << dim)
      , global_weights(1 << dim, 0)
      , is_set_global_weights(false)
      , n_refine_local(n_refine_local)
      , global_cell_id(global_cell.id())
      , local_subdomain(local_subdomain)
      , output_flag(false)
      , verbose(false)
      , last_steps(1 << dim, 0)
    {
      static_assert(dim == 2 || dim == 3, "the reference instantiates dim 2 and 3");
      for (unsigned int v = 0; v < (1u << dim); ++v)
        corner_points[v] = global_cell.vertex(v);
    }

    // the reference copies objects into the map while they are empty (ms.tpp:65-66); a copy
    // shares the GPU batch of its source, it never duplicates device memory
    DiffusionProblemBasis(const DiffusionProblemBasis<dim> &) = default;

    // run() of a single object: a batch of one cell
    void run()
    {
      std::map<CellId, DiffusionProblemBasis<dim> *> one;
      one[global_cell_id] = this;
      run_pointers(one, default_coefficient(), BasisSolverControl(), 0);
    }

    // THE replacement of the hot loop ms.tpp:81-87: every object of the map in one GPU batch
    static void run_all(std::map<CellId, DiffusionProblemBasis<dim>> &cell_basis_map,
                        const Coefficients::TensorCoefficient<dim>   &matrix_coeff,
                        const BasisSolverControl &control = BasisSolverControl(), int device_id = 0)
    {
      std::map<CellId, DiffusionProblemBasis<dim> *> ptrs;
      for (auto &kv : cell_basis_map)
        ptrs[kv.first] = &kv.second;
      run_pointers(ptrs, matrix_coeff, control, device_id);
    }
    static void run_all(std::map<CellId, DiffusionProblemBasis<dim>> &cell_basis_map)
    {
      run_all(cell_basis_map, default_coefficient());
    }

    // The same for the objects [first, last) of the map only, on one device: with several GPUs
    // every device gets a contiguous CellId (= Morton) range, the ownership rule the reference
    // inherits from p4est (ms.tpp:52).  Safe to call concurrently from one host thread per device.
    static void run_range(typename std::map<CellId, DiffusionProblemBasis<dim>>::iterator first,
                          typename std::map<CellId, DiffusionProblemBasis<dim>>::iterator last,
                          const Coefficients::TensorCoefficient<dim> &matrix_coeff,
                          const BasisSolverControl &control, int device_id)
    {
      std::map<CellId, DiffusionProblemBasis<dim> *> ptrs;
      for (auto it = first; it != last; ++it)
        ptrs[it->first] = &it->second;
      run_pointers(ptrs, matrix_coeff, control, device_id);
    }

    void output_global_solution_in_cell() const
    {
      if (!is_set_global_weights)
        throw std::logic_error("Global weights must be set first."); // Assert, basis.tpp:425-426
      std::vector<double> sol;
      get_global_solution(sol);
      write_vtu(filename_global, sol, "solution");
    }

    const FullMatrix<double> &get_global_element_matrix() const { return global_element_matrix; }
    const Vector<double>     &get_global_element_rhs() const { return global_element_rhs; }
    const std::string        &get_filename_global() { return filename_global; }
    void                      set_output_flag(bool flag) { output_flag = flag; }

    void set_global_weights(const std::vector<double> &weights)
    {
      global_weights = weights;
      if (batch)
        {
          for (unsigned i = 0; i < (1u << dim); ++i)
            batch->weights[NB * index_in_batch + i] = weights[i];
          batch->weights_dirty = true;
        }
      is_set_global_weights = true;
    }

    // --- additions (not in the reference's public interface) -------------------------------
    // solver_control.last_step() of solve `index_basis` (printed at basis.tpp:310-316)
    unsigned last_step(unsigned index_basis) const { return last_steps.at(index_basis); }
    // solution_vector[index_basis] (basis.hpp:216), fetched from the device on demand
    void get_basis(unsigned index_basis, std::vector<double> &out) const
    {
      require_batch();
      out.resize(n_dofs());
      internal::check(msb_get_basis(batch->handle, (int32_t)index_in_batch, (int32_t)index_basis, out.data()));
    }
    void get_global_solution(std::vector<double> &out) const
    {
      require_batch();
      batch->upload_weights();
      out.resize(n_dofs());
      if (batch->gsol_valid) // prefetched in bulk (prefetch_global_solutions)
        std::copy(batch->gsol_cache.begin() + index_in_batch * n_dofs(),
                  batch->gsol_cache.begin() + (index_in_batch + 1) * n_dofs(), out.begin());
      else
        internal::check(msb_get_global_solution(batch->handle, (int32_t)index_in_batch, out.data()));
    }
    // Bring the global solutions of all cells of this object's batch to the host in one call; the
    // per-object getters / output_global_solution_in_cell() then read the host copy.
    void prefetch_global_solutions() const
    {
      require_batch();
      batch->fetch_global_solutions(n_dofs());
    }
    // the 2^dim bases of the cells [first, first + count) of this object's batch, [count][2^dim][N]
    void get_bases_of_batch(std::size_t first, std::size_t count, std::vector<double> &out) const
    {
      require_batch();
      out.resize(count * NB * n_dofs());
      internal::check(msb_get_bases(batch->handle, (int32_t)first, (int32_t)count, out.data()));
    }
    std::size_t batch_size() const { return batch ? batch->n_cells : 0; }
    std::size_t n_dofs() const
    {
      const std::size_t np = (std::size_t(1) << n_refine_local) + 1;
      return dim == 2 ? np * np : np * np * np;
    }
    void set_verbose(bool v) { verbose = v; }
    // vertex (lexicographic) -> deal.II DoF index of the local mesh (same for every cell)
    void get_dof_map(std::vector<uint32_t> &out) const
    {
      require_batch();
      out.resize(n_dofs());
      internal::check(msb_get_dof_map(batch->handle, out.data()));
    }

  private:
  public:
    static const Coefficients::TensorCoefficient<dim> &default_coefficient()
    {
      // assemble_system hard-wires Coefficients::MatrixCoeff<dim> (basis.tpp:184)
      static const Coefficients::MatrixCoeff<dim> c;
      return c;
    }

  private:

    void require_batch() const
    {
      if (!batch)
        throw std::logic_error("DiffusionProblemBasis: run() has not been called");
    }

    // set_filename_global (basis.tpp:380-389)
    void set_filename_global()
    {
      std::ostringstream s;
      s << (dim == 2 ? "solution-ms_fine-2d" : "solution-ms_fine-3d") << "." << std::setw(5)
        << std::setfill('0') << local_subdomain << ".cell-" << global_cell_id.to_string() << ".vtu";
      filename_global += s.str();
    }

    static void run_pointers(std::map<CellId, DiffusionProblemBasis<dim> *> &objs,
                             const Coefficients::TensorCoefficient<dim>     &coeff,
                             const BasisSolverControl &control, int device_id)
    {
      if (objs.empty())
        return;
      const unsigned l = objs.begin()->second->n_refine_local;
      const unsigned n = 1u << l;
      std::vector<double> corners;
      corners.reserve(objs.size() * NB * dim);
      for (auto &kv : objs)
        for (unsigned v = 0; v < (1u << dim); ++v)
          for (int d = 0; d < dim; ++d)
            corners.push_back(kv.second->corner_points[v](d));

      const msb_coeff_desc desc = coeff.device_descriptor();
      std::vector<double>  table;
      if (desc.kind == MSB_COEFF_TABLE)
        {
          // a coefficient class the device has no formula for: evaluate its value_list at the
          // fine quadrature points (what assemble_system does per cell, basis.tpp:202-203);
          // dim 2: [cell][iy n + ix][4 q][2x2], dim 3: [cell][(iz n + iy) n + ix][8 q][3x3], q with x fastest
          const double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
          std::vector<Point<dim>>     pts(NB);
          std::vector<Tensor<2, dim>> vals(NB);
          const unsigned              nz = dim == 3 ? n : 1;
          table.reserve(objs.size() * std::size_t(n) * n * nz * NB * dim * dim);
          for (auto &kv : objs)
            {
              const auto &c = kv.second->corner_points;
              for (unsigned iz = 0; iz < nz; ++iz)
                for (unsigned iy = 0; iy < n; ++iy)
                  for (unsigned ix = 0; ix < n; ++ix)
                    {
                      for (unsigned q = 0; q < NB; ++q)
                        {
                          // multilinear image of the reference point: sum_v c_v N_v(s, t, u)
                          const double s = (ix + g[q & 1]) / n, t = (iy + g[(q >> 1) & 1]) / n,
                                       u = dim == 3 ? (iz + g[q >> 2]) / n : 0.0;
                          for (int d = 0; d < dim; ++d)
                            {
                              double x = 0.0;
                              for (unsigned v = 0; v < NB; ++v)
                                x += c[v](d) * ((v & 1) ? s : 1 - s) * (((v >> 1) & 1) ? t : 1 - t) *
                                     (dim == 3 ? ((v >> 2) ? u : 1 - u) : 1.0);
                              pts[q](d) = x;
                            }
                        }
                      coeff.value_list(pts, vals);
                      for (unsigned q = 0; q < NB; ++q)
                        for (int i = 0; i < dim; ++i)
                          for (int j = 0; j < dim; ++j)
                            table.push_back(vals[q][i][j]);
                    }
            }
        }

      const Coefficients::RightHandSide<dim> right_hand_side; // basis.tpp:190
      auto batch = std::make_shared<internal::Batch<dim>>(l, corners, desc, table,
                                                          right_hand_side.value(Point<dim>()), device_id);
      const int rc = msb_run(batch->handle, control.tolerance, (int32_t)control.max_steps);
      if (rc != MSB_OK && rc != MSB_ERR_NO_CONVERGENCE)
        throw BasisStageError(rc, msb_last_error());

      const std::size_t    C = objs.size();
      std::vector<double>  M(NB * NB * C), b(NB * C), res(NB * C);
      std::vector<int32_t> its(NB * C);
      internal::check(msb_get_element_matrices(batch->handle, M.data(), b.data()));
      internal::check(msb_get_iteration_counts(batch->handle, its.data(), res.data()));

      std::size_t k = 0;
      for (auto &kv : objs)
        {
          DiffusionProblemBasis<dim> &o = *kv.second;
          o.batch          = batch;
          o.index_in_batch = k;
          for (unsigned i = 0; i < NB; ++i)
            {
              for (unsigned j = 0; j < NB; ++j)
                o.global_element_matrix(i, j) = M[NB * NB * k + NB * i + j];
              // the reference accumulates b with += and never resets it (basis.tpp:280)
              o.global_element_rhs(i) += b[NB * k + i];
              o.last_steps[i] = (unsigned)its[NB * k + i];
            }
          o.is_built_global_element_matrix = true;
          if (o.filename_global.empty())
            o.set_filename_global();
          if (o.verbose)
            for (unsigned i = 0; i < NB; ++i)
              std::cout << "   (cell   " << o.global_cell_id.to_string() << ") (basis   " << i << ")   "
                        << o.last_steps[i] << " fine CG iterations needed to obtain convergence."
                        << std::endl;
          ++k;
        }
      if (rc == MSB_ERR_NO_CONVERGENCE)
        {
          int32_t cell = -1, ib = -1;
          double  r    = 0.0;
          msb_get_failure(batch->handle, &cell, &ib, &r);
          auto it = objs.begin();
          std::advance(it, cell);
          throw NoConvergence(it->first.to_string(), (unsigned)ib, r, control.max_steps);
        }
      if (objs.begin()->second->output_flag)
        for (auto &kv : objs)
          kv.second->output_basis();
    }

    // output_basis (basis.tpp:392-418): the 2^dim bases of this cell as one VTU
    void output_basis() const
    {
      std::ostringstream name;
      name << (dim == 2 ? "2d-" : "3d-") << "basis." << std::setw(5) << std::setfill('0') << 0
           << ".cell-" << global_cell_id.to_string() << ".vtu";
      std::vector<std::vector<double>> phi(NB);
      for (unsigned i = 0; i < NB; ++i)
        get_basis(i, phi[i]);
      write_vtu_multi(name.str(), phi, "basis_");
    }

    // ---- minimal VTU writer (unstructured quads; DoF order = deal.II first-touch) ---------
    void write_vtu(const std::string &fn, const std::vector<double> &data, const std::string &name) const
    {
      write_vtu_multi(fn, {data}, name, /*numbered=*/false);
    }
    void write_vtu_multi(const std::string &fn, const std::vector<std::vector<double>> &fields,
                         const std::string &prefix, bool numbered = true) const
    {
      const unsigned        n = 1u << n_refine_local, np = n + 1, npz = dim == 3 ? np : 1, nz = dim == 3 ? n : 1;
      const std::size_t     n_pts = n_dofs(), n_cells = std::size_t(n) * n * nz;
      std::vector<uint32_t> dof(n_pts);
      internal::check(msb_get_dof_map(batch->handle, dof.data()));
      std::ofstream out(fn.c_str());
      out << std::setprecision(17);
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" "
             "byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n<Piece NumberOfPoints=\""
          << n_pts << "\" NumberOfCells=\"" << n_cells << "\">\n<Points>\n<DataArray type=\"Float64\" "
             "NumberOfComponents=\"3\" format=\"ascii\">\n";
      const auto &c = corner_points;
      for (unsigned jz = 0; jz < npz; ++jz)
        for (unsigned jy = 0; jy < np; ++jy)
          for (unsigned jx = 0; jx < np; ++jx)
            {
              // multilinear image of the uniform grid: sum_v c_v N_v(s, t, u)
              const double s = double(jx) / n, t = double(jy) / n, u = dim == 3 ? double(jz) / n : 0.0;
              for (int d = 0; d < 3; ++d)
                {
                  double x = 0.0;
                  if (d < dim)
                    for (unsigned v = 0; v < NB; ++v)
                      x += c[v](d) * ((v & 1) ? s : 1 - s) * (((v >> 1) & 1) ? t : 1 - t) *
                           (dim == 3 ? ((v >> 2) ? u : 1 - u) : 1.0);
                  out << x << (d < 2 ? " " : "\n");
                }
            }
      out << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" "
             "format=\"ascii\">\n";
      for (unsigned iz = 0; iz < nz; ++iz)
        for (unsigned iy = 0; iy < n; ++iy)
          for (unsigned ix = 0; ix < n; ++ix)
            {
              // VTK_QUAD / VTK_HEXAHEDRON vertex order (counter-clockwise, bottom then top)
              const std::size_t b = (std::size_t(iz) * np + iy) * np + ix, up = std::size_t(np) * np;
              out << b << " " << b + 1 << " " << b + np + 1 << " " << b + np;
              if (dim == 3)
                out << " " << b + up << " " << b + up + 1 << " " << b + up + np + 1 << " " << b + up + np;
              out << "\n";
            }
      out << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
      for (std::size_t k = 1; k <= n_cells; ++k)
        out << NB * k << "\n";
      out << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
      for (std::size_t k = 0; k < n_cells; ++k)
        out << (dim == 2 ? "9\n" : "12\n");
      out << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n";
      for (std::size_t f = 0; f < fields.size(); ++f)
        {
          out << "<DataArray type=\"Float64\" Name=\"" << prefix;
          if (numbered)
            out << f;
          out << "\" format=\"ascii\">\n";
          for (std::size_t lex = 0; lex < n_pts; ++lex)
            out << fields[f][dof[lex]] << "\n";
          out << "</DataArray>\n";
        }
      out << "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
    }

    MPI_Comm_shim           mpi_communicator;
    std::vector<Point<dim>> corner_points;
    std::string             filename_global;
    FullMatrix<double>      global_element_matrix;
    bool                    is_built_global_element_matrix;
    Vector<double>          global_element_rhs;
    std::vector<double>     global_weights;
    bool                    is_set_global_weights;
    unsigned int            n_refine_local;
    CellId                  global_cell_id;
    unsigned int            local_subdomain;
    bool                    output_flag;
    bool                    verbose;
    std::vector<unsigned>   last_steps;

    std::shared_ptr<internal::Batch<dim>> batch;
    std::size_t                           index_in_batch = 0;
  };
} // namespace DiffusionProblem
