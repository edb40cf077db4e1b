// nsx_dealii_adapter.h -- hands the outputs of NSSolverStationary::setup() / NSSolver::setup() (deal.II + Trilinos objects) to the
// device library behind include/nsx.h.  This is the piece a maintainer of HliasGit/navier_stokes_solver adds to make libnsx.so a
// drop-in for assemble_system / solve_system: everything deal.II does in setup() stays (mesh generation / GridIn, partitioning,
// distribute_dofs + component_wise, IndexSets, sparsity patterns); only its RESULTS cross the boundary, once.
//
// Reference members read here (lab_new/src/NSSolverStationary.{hpp,cpp}; NSSolver.{hpp,cpp} has the same members):
//   dof_handler, fe, mesh                         NSSolverStationary.hpp:406-428, cell loop NSSolverStationary.cpp:356-361, 528
//   block_owned_dofs / block_relevant_dofs        NSSolverStationary.cpp:237-242
//   jacobian_matrix, pressure_mass (patterns)     NSSolverStationary.cpp:276-305
//   boundary ids 7 / 6 / 10 (Dirichlet), 8, 10    NSSolverStationary.cpp:503-508, 540-572, 842-843
//
// NOT BUILT in the image this repository is developed in: deal.II, Trilinos and MPI are absent there, so this
// file is compiled only by adapters/dealii/CMakeLists.txt where find_package(deal.II) succeeds, and it has never been run.  The
// stand-in that produces the same arrays without deal.II -- and that the tests exercise -- is include/nsx_host.h.
#pragma once
#include <deal.II/base/index_set.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>
#include <deal.II/fe/fe_system.h>
#include <deal.II/lac/trilinos_block_sparse_matrix.h>
#include <deal.II/numerics/vector_tools.h>

#include <map>
#include <vector>

#include "nsx.h"

namespace nsx_dealii {

// What setup() has built, by reference (no copies of deal.II objects)
template <int dim>
struct SetupView {
  const dealii::DoFHandler<dim> &dof_handler;
  const dealii::FESystem<dim> &fe;
  const std::vector<dealii::IndexSet> &block_owned_dofs;     // [0] velocity, [1] pressure, block-local indices
  const std::vector<dealii::IndexSet> &block_relevant_dofs;
  const dealii::TrilinosWrappers::BlockSparseMatrix &jacobian_matrix;
  const dealii::TrilinosWrappers::BlockSparseMatrix &pressure_mass;
  bool simplex;                                              // FE_SimplexP (mesh from file) or FE_Q (generated mesh)
  MPI_Comm comm;
};

// Creates the context of this rank and fills it; `inlet` is the function imposed on boundary id 7 in the one non-homogeneous
// assembly (InletVelocity, NSSolverStationary.hpp:60-111).  Throws std::runtime_error with nsx_last_error() on failure.
template <int dim>
nsx_ctx *hand_over(const SetupView<dim> &s, const dealii::Function<dim> &inlet, int device_id);

// ghosted block vector <-> the library's owned layout [velocity | pressure] (output(), NSSolverStationary.cpp:765-800)
void download_solution(nsx_ctx *ctx, dealii::TrilinosWrappers::MPI::BlockVector &solution_owned);

}  // namespace nsx_dealii
