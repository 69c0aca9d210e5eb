#pragma once
// stub (tests/fake_dealii/README.md): just the iterator surface the host mirror uses
#include <deal.II/base/point.h>

#include <vector>
namespace dealii
{
  template <int dim>
  class Triangulation
  {
  public:
    struct Face
    {
      bool         boundary = false;
      unsigned int bid      = 0;
      bool         at_boundary() const { return boundary; }
      unsigned int boundary_id() const { return bid; }
      const Face  *operator->() const { return this; }
    };
    struct Cell
    {
      Point<dim>  v[1 << dim];
      CellId      cid;
      Face        faces[2 * dim];
      const Point<dim> &vertex(unsigned i) const { return v[i]; }
      CellId            id() const { return cid; }
      Face              face(unsigned f) const { return faces[f]; }
    };
    struct active_cell_iterator
    {
      Cell       *c = nullptr;
      Cell       *operator->() const { return c; }
    };
  };
} // namespace dealii
