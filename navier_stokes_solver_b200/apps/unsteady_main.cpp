// NSSolver -- the time-dependent solver's executable (reference: lab_new/src/test.cpp:21-155).
#include "cli.hpp"
#include "ns_unsteady.hpp"

int main(int argc, char *argv[]) {
  const bool root = app::Ranks().rank == 0;
  app::Options o;
  switch (app::parse_command_line(argc, argv, /*unsteady=*/true, root, o)) {
    case app::Parse::ExitOk: return 0;
    case app::Parse::ExitError: return 1;
    case app::Parse::Run: break;
  }
  if (root) app::print_banner(o, true);

  app::NSSolver problem(o.mesh_file_name, o.degree_velocity, o.degree_pressure, o.time_span, o.time_step, o.mesh_size_x, o.mesh_size_y, o.solver_type,
                        o.tolerance, o.preconditioner, o.Re, o.read_mesh_from_file);
  if (std::getenv("NSX_NO_OUTPUT")) problem.write_output = false;
  if (const char *ms = std::getenv("NSX_MAX_TIME_STEPS")) problem.max_time_steps = (unsigned)std::atoi(ms);

  problem.setup();   // run sequence of test.cpp:151-152
  problem.solve();
  return 0;
}
