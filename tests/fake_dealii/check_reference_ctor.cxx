// Compile check of the MSFEM_WITH_DEALII branch of the host mirror: the construction loop of the reference
// (diffusion_problem_ms.tpp:54-66) written against the mirror, with the reference's argument types.
#include "msfem/diffusion_problem_basis.hpp"

#include <map>

template <int dim>
void
construct_like_the_reference(dealii::Triangulation<dim> &, typename dealii::Triangulation<dim>::active_cell_iterator cell,
                             unsigned n_refine_local, MPI_Comm mpi_communicator)
{
  using namespace DiffusionProblem;
  std::map<dealii::CellId, DiffusionProblemBasis<dim>> cell_basis_map;
  DiffusionProblemBasis<dim> current_cell_problem(n_refine_local, cell, /*local_subdomain*/ 0u, mpi_communicator);
  cell_basis_map.emplace(cell->id(), current_cell_problem);
  DiffusionProblemBasis<dim>::run_all(cell_basis_map);
  const dealii::FullMatrix<double> &M = cell_basis_map.begin()->second.get_global_element_matrix();
  (void)M;
}
template void construct_like_the_reference<2>(dealii::Triangulation<2> &, dealii::Triangulation<2>::active_cell_iterator,
                                              unsigned, MPI_Comm);
template void construct_like_the_reference<3>(dealii::Triangulation<3> &, dealii::Triangulation<3>::active_cell_iterator,
                                              unsigned, MPI_Comm);
