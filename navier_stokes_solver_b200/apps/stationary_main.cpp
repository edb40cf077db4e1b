// StationaryNSSolver -- command line of the stationary solver (reference: lab_new/src/testStationary.cpp:19-138).
// Same flags, defaults, messages and exit codes; the short -M consumes one argument because the option string
// declares it so ("M:m:r:s:t:p:h", testStationary.cpp:46) while --read-mesh-from-file takes none.
#include <getopt.h>

#include <cstdlib>
#include <cstring>
#include <iostream>

#include "ns_stationary.hpp"

static void print_help() {
  std::cout << "Usage: ./NSSolver [options]\n\n"
            << "Options:\n"
            << "  -M, --read-mesh-from-file  Read mesh from file instead or generate it inside the program\n"
            << "  -m, --mesh-size X,Y       Set mesh size (two integers separated by a comma)\n"
            << "  -r, --reynolds N         Set Reynolds number (floating point value)\n"
            << "  -s, --solver N            Select solver (valid values: 0: GMRES, 1: FGMRES, 2: Bicgstab)\n"
            << "  -t, --tolerance D         Set tolerance (floating point value)\n"
            << "  -p, --preconditioner N    Select preconditioner (valid values: 0: blockDiagonal, 1: blockTriangular, 2: aSIMPLE)\n"
            << "  -h, --help                Display this help message\n";
}

int main(int argc, char *argv[]) {
  const bool root = app::Ranks().rank == 0;
  bool read_mesh_from_file = false;
  unsigned int degree_velocity = 3, degree_pressure = 2;
  double Re = 100.0;
  int mesh_size_x = 100, mesh_size_y = 100;
  int solver_type = 1;
  double tolerance = 1e-6;
  int preconditioner = 0;

  static struct option long_options[] = {{"read-mesh-from-file", no_argument, 0, 'M'}, {"mesh-size", required_argument, 0, 'm'},
                                         {"reynolds", required_argument, 0, 'r'},      {"solver", required_argument, 0, 's'},
                                         {"tolerance", required_argument, 0, 't'},     {"preconditioner", required_argument, 0, 'p'},
                                         {"help", no_argument, 0, 'h'},                {0, 0, 0, 0}};
  int opt;
  while ((opt = getopt_long(argc, argv, "M:m:r:s:t:p:h", long_options, NULL)) != -1) {
    switch (opt) {
      case 'M': read_mesh_from_file = true; degree_velocity = 2; degree_pressure = 1; break;
      case 'm': {
        char *comma = strchr(optarg, ',');
        if (comma) { *comma = '\0'; mesh_size_x = std::atoi(optarg); mesh_size_y = std::atoi(comma + 1); }
        else { if (root) std::cerr << "Error: mesh-size requires two values separated by comma\n"; return 1; }
        break;
      }
      case 'r': Re = std::atof(optarg); break;
      case 's': solver_type = std::atoi(optarg); break;
      case 't': tolerance = std::atof(optarg); break;
      case 'p': preconditioner = std::atoi(optarg); break;
      case 'h': if (root) print_help(); return 0;
      default: if (root) print_help(); return 1;
    }
  }
  if (tolerance <= 0) { if (root) std::cerr << "Error: tolerance must be positive\n"; return 1; }

  if (root) {
    std::cout << "--------- CONFIGURATION PARAMETERS --------- \n";
    std::cout << "Mesh size: " << mesh_size_x << "x" << mesh_size_y << "\n";
    std::cout << "Reynolds number: " << Re << "\n";
    std::cout << "Solver type: ";
    if (solver_type == 0) std::cout << "GMRES\n";
    else if (solver_type == 1) std::cout << "FGMRES\n";
    else if (solver_type == 2) std::cout << "Bicgstab\n";
    std::cout << "Tolerance: " << tolerance << "\n";
    std::cout << "Preconditioner: ";
    if (preconditioner == 0) std::cout << "blockDiagonal\n";
    else if (preconditioner == 1) std::cout << "blockTriangular\n";
    else if (preconditioner == 2) std::cout << "aSIMPLE\n";
    std::cout << "-----------------------------------------------\n";
  }

  // the reference hard-codes this path (testStationary.cpp:127); NSX_MESH_FILE points somewhere else without a rebuild
  const char *mesh_env = std::getenv("NSX_MESH_FILE");
  const std::string mesh_file_name = mesh_env ? mesh_env : "/home/users/gdaneri/navier_stokes_solver/lab_new/mesh/new_mesh.msh";

  app::NSSolverStationary problem(mesh_file_name, degree_velocity, degree_pressure, mesh_size_x, mesh_size_y, solver_type, tolerance, preconditioner, Re,
                                  read_mesh_from_file);
  if (std::getenv("NSX_NO_OUTPUT")) problem.write_output = false;

  // as in the reference nothing is caught: SolverControl::NoConvergence / std::invalid_argument end the program
  problem.setup();
  problem.solve_newton();
  problem.output();
  problem.compute_lift_drag();
  problem.print_lift_coeff();
  problem.print_drag_coeff();
  return 0;
}
