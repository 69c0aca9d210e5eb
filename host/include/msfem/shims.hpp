// shims.hpp -- the handful of deal.II value types the reference's basis interface exposes,
// re-stated without deal.II so that the host side of the B200 basis stage compiles on a
// machine that has neither deal.II nor MPI.  Only what
// /root/reference/include/base/diffusion_problem_basis.hpp:72-138 and
// /root/reference/include/coefficients/*.hpp touch is provided.  On a machine with deal.II,
// define MSFEM_WITH_DEALII and the real types are used instead.
#pragma once

#include <array>

#ifdef MSFEM_WITH_DEALII
#  include <deal.II/base/point.h>
#  include <deal.II/base/tensor.h>
#  include <deal.II/grid/cell_id.h>
#  include <deal.II/grid/tria.h>
#  include <deal.II/lac/full_matrix.h>
#  include <deal.II/lac/vector.h>
#  include <mpi.h>
namespace msfem
{
  using dealii::CellId;
  using dealii::FullMatrix;
  using dealii::Point;
  using dealii::Tensor;
  using dealii::Vector;
  using MPI_Comm_shim = MPI_Comm;
} // namespace msfem
#else
#  define MSFEM_SHIM_NAMESPACE msfem
#  include "msfem/shim_types.hpp"
#  undef MSFEM_SHIM_NAMESPACE
namespace msfem
{
  // MPI is not needed on one box; the reference stores the communicator without using it
  // (basis.tpp:24)
  using MPI_Comm_shim = int;
} // namespace msfem
#endif

namespace msfem
{
  // What the reference reads off Triangulation<dim>::active_cell_iterator when it builds a
  // basis object (basis.tpp:36,44-49; basis_q1.tpp:35-44): the vertices and the CellId.
  template <int dim>
  struct CoarseCell
  {
    std::array<Point<dim>, (1 << dim)> vertices;
    CellId                             cell_id;
    unsigned                           boundary_id[2 * dim]; // 255 = interior face

    const Point<dim> &vertex(unsigned i) const { return vertices[i]; }
    CellId            id() const { return cell_id; }

#ifdef MSFEM_WITH_DEALII
    // from the reference's own argument type (basis.hpp:79-83, caller ms.tpp:54-66)
    static CoarseCell
    from(const typename dealii::Triangulation<dim>::active_cell_iterator &cell)
    {
      CoarseCell c;
      for (unsigned v = 0; v < (1u << dim); ++v)
        c.vertices[v] = cell->vertex(v);
      c.cell_id = cell->id();
      for (unsigned f = 0; f < 2 * dim; ++f)
        c.boundary_id[f] = cell->face(f)->at_boundary() ? (unsigned)cell->face(f)->boundary_id() : 255u;
      return c;
    }
#endif
  };
} // namespace msfem
