// StationaryNSSolver -- the stationary solver's executable (reference: lab_new/src/testStationary.cpp:19-138).
#include "cli.hpp"
#include "ns_stationary.hpp"

int main(int argc, char *argv[]) {
  const bool root = app::Ranks().rank == 0;
  app::Options o;
  switch (app::parse_command_line(argc, argv, /*unsteady=*/false, root, o)) {
    case app::Parse::ExitOk: return 0;
    case app::Parse::ExitError: return 1;
    case app::Parse::Run: break;
  }
  if (root) app::print_banner(o, false);

  app::NSSolverStationary problem(o.mesh_file_name, o.degree_velocity, o.degree_pressure, o.mesh_size_x, o.mesh_size_y, o.solver_type, o.tolerance,
                                  o.preconditioner, o.Re, o.read_mesh_from_file);
  if (std::getenv("NSX_NO_OUTPUT")) problem.write_output = false;

  // run sequence of testStationary.cpp:131-136; as there, nothing is caught: SolverControl::NoConvergence and
  // std::invalid_argument end the program
  problem.setup();
  problem.solve_newton();
  problem.output();
  problem.compute_lift_drag();
  problem.print_lift_coeff();
  problem.print_drag_coeff();
  return 0;
}
