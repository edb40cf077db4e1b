// dump_reference_fixture.cc -- writes what would PIN the oracle to the real reference: the assembled Jacobian blocks, the residual
// and the converged increment of HliasGit/navier_stokes_solver's own NSSolverStationary on the generated 16x6 mesh, as plain text
// under tests/golden/reference_16x6/.  It needs deal.II + Trilinos + MPI (adapters/dealii/CMakeLists.txt, -DREFERENCE_SRC=...);
// none of them exists in the image this repository is developed in, so this program has never been run and DESIGN.md section 2
// keeps the oracle's status at "parity unpinned".  A maintainer with the reference's toolchain runs
//     mpirun -n 1 ./dump_reference_fixture tests/golden/reference_16x6
// commits the files, and tests/test_oracle.py::test_reference_fixture (skipped while the directory is absent) compares
// the oracle's J, r and delta with them entry by entry (1e-12 / 1e-12 / 1e-8).
//
// The reference keeps its members `protected` and its matrices are only reachable from inside the class, hence the subclass.
#include <deal.II/lac/trilinos_block_sparse_matrix.h>

#include <fstream>
#include <iomanip>
#include <string>

#include "NSSolverStationary.hpp"

struct Dumper : public NSSolverStationary {
  using NSSolverStationary::NSSolverStationary;
  static void write_block(const std::string &path, const dealii::TrilinosWrappers::SparseMatrix &A) {
    std::ofstream f(path);
    f << std::setprecision(17);
    for (auto r = A.local_range().first; r < A.local_range().second; ++r)
      for (auto it = A.begin(r); it != A.end(r); ++it) f << r << ' ' << it->column() << ' ' << it->value() << '\n';
  }
  static void write_vector(const std::string &path, const dealii::TrilinosWrappers::MPI::BlockVector &v) {
    std::ofstream f(path);
    f << std::setprecision(17);
    for (unsigned b = 0; b < v.n_blocks(); ++b)
      for (auto i : v.block(b).locally_owned_elements()) f << b << ' ' << i << ' ' << v.block(b)[i] << '\n';
  }
  void dump(const std::string &dir) {
    setup();
    nu = 1.0 / 10.0;
    assemble_system(/*global_first_iter*/ true, /*computing_stokes*/ true);   // NSSolverStationary.cpp:685-690, first call of the run
    write_block(dir + "/F.txt", jacobian_matrix.block(0, 0));
    write_block(dir + "/Bt.txt", jacobian_matrix.block(0, 1));
    write_block(dir + "/B.txt", jacobian_matrix.block(1, 0));
    write_block(dir + "/Mp.txt", pressure_mass.block(1, 1));
    write_vector(dir + "/residual.txt", residual_vector);
    const int its = solve_system();                                            // FGMRES + blockDiagonal, tol 1e-12
    write_vector(dir + "/delta.txt", delta_owned);
    std::ofstream(dir + "/iterations.txt") << its << '\n';
  }
};

int main(int argc, char *argv[]) {
  dealii::Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);
  const std::string dir = argc > 1 ? argv[1] : "tests/golden/reference_16x6";
  const std::string no_mesh_file;
  // ctor arguments as testStationary.cpp:119-130: mesh file, degrees (3, 2), Re, solver 1 (FGMRES), tol, preconditioner 0, 16 x 6, generated mesh
  Dumper problem(no_mesh_file, 3, 2, 100.0, 1, 1e-12, 0, 16, 6, false);
  problem.dump(dir);
  return 0;
}
