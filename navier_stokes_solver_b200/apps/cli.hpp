// cli.hpp -- the command line shared by the two executables.
//
// Drop-in for the option handling of the reference's mains (lab_new/src/testStationary.cpp:22-124, test.cpp:24-145):
// same flags, defaults, texts and exit codes.  One table-driven parser serves both binaries; the unsteady one adds -T.
// Quirk kept on purpose: the short option string declares an argument for -M ("M:m:r:s:t:p:h"), the long option
// --read-mesh-from-file takes none, so `-M -m 100,70` swallows the -m.
#pragma once
#include <getopt.h>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

namespace app {

struct Options {
  bool read_mesh_from_file = false;
  unsigned degree_velocity = 3, degree_pressure = 2;
  double Re = 100.0;
  int mesh_size_x = 100, mesh_size_y = 100;
  int solver_type = 1;
  double tolerance = 1e-6;
  int preconditioner = 0;
  double time_span = 1.0, time_step = 0.01;  // unsteady only
  // the reference hard-codes this path (testStationary.cpp:127, test.cpp:147); NSX_MESH_FILE points elsewhere without a rebuild
  std::string mesh_file_name = "/home/users/gdaneri/navier_stokes_solver/lab_new/mesh/new_mesh.msh";
};

enum class Parse { Run, ExitOk, ExitError };

inline void print_help(bool unsteady) {
  std::cout << "Usage: ./NSSolver [options]\n\n"
            << "Options:\n";
  if (unsteady) std::cout << "  -T, --time-span and time-step T,D\n";
  std::cout << "  -M, --read-mesh-from-file  Read mesh from file instead or generate it inside the program\n"
            << "  -m, --mesh-size X,Y       Set mesh size (two integers separated by a comma)\n"
            << "  -r, --reynolds N         Set Reynolds number (floating point value)\n"
            << "  -s, --solver N            Select solver (valid values: 0: GMRES, 1: FGMRES, 2: Bicgstab)\n"
            << "  -t, --tolerance D         Set tolerance (floating point value)\n"
            << "  -p, --preconditioner N    Select preconditioner (valid values: 0: blockDiagonal, 1: blockTriangular, 2: aSIMPLE)\n"
            << "  -h, --help                Display this help message\n";
}

// "X,Y" -> two numbers; false when the comma is missing
template <class T, class Conv>
inline bool split_pair(char *arg, T &a, T &b, Conv conv) {
  char *comma = std::strchr(arg, ',');
  if (!comma) return false;
  *comma = '\0';
  a = conv(arg);
  b = conv(comma + 1);
  return true;
}

inline Parse parse_command_line(int argc, char *argv[], bool unsteady, bool root, Options &o) {
  static struct option stationary_long[] = {{"read-mesh-from-file", no_argument, 0, 'M'}, {"mesh-size", required_argument, 0, 'm'},
                                            {"reynolds", required_argument, 0, 'r'},      {"solver", required_argument, 0, 's'},
                                            {"tolerance", required_argument, 0, 't'},     {"preconditioner", required_argument, 0, 'p'},
                                            {"help", no_argument, 0, 'h'},                {0, 0, 0, 0}};
  static struct option unsteady_long[] = {{"timespan-step", required_argument, 0, 'T'}, {"read-mesh-from-file", no_argument, 0, 'M'},
                                          {"mesh-size", required_argument, 0, 'm'},     {"reynolds", required_argument, 0, 'r'},
                                          {"solver", required_argument, 0, 's'},        {"tolerance", required_argument, 0, 't'},
                                          {"preconditioner", required_argument, 0, 'p'}, {"help", no_argument, 0, 'h'},
                                          {0, 0, 0, 0}};
  const char *short_opts = unsteady ? "T:M:m:r:s:t:p:h" : "M:m:r:s:t:p:h";
  int opt;
  while ((opt = getopt_long(argc, argv, short_opts, unsteady ? unsteady_long : stationary_long, nullptr)) != -1) {
    switch (opt) {
      case 'T':
        if (!split_pair(optarg, o.time_span, o.time_step, [](const char *s) { return std::atof(s); })) {
          if (root) std::cerr << "Error: timespan-step requires two values separated by comma\n";
          return Parse::ExitError;
        }
        break;
      case 'M': o.read_mesh_from_file = true; o.degree_velocity = 2; o.degree_pressure = 1; break;
      case 'm':
        if (!split_pair(optarg, o.mesh_size_x, o.mesh_size_y, [](const char *s) { return std::atoi(s); })) {
          if (root) std::cerr << "Error: mesh-size requires two values separated by comma\n";
          return Parse::ExitError;
        }
        break;
      case 'r': o.Re = std::atof(optarg); break;
      case 's': o.solver_type = std::atoi(optarg); break;
      case 't': o.tolerance = std::atof(optarg); break;
      case 'p': o.preconditioner = std::atoi(optarg); break;
      case 'h': if (root) print_help(unsteady); return Parse::ExitOk;
      default: if (root) print_help(unsteady); return Parse::ExitError;
    }
  }
  if (unsteady) {
    if (o.time_step <= 0 || o.time_span <= 0 || o.tolerance <= 0) {
      if (root) std::cerr << "Error: time_step, time_span, and tolerance must be positive\n";
      return Parse::ExitError;
    }
  } else if (o.tolerance <= 0) {
    if (root) std::cerr << "Error: tolerance must be positive\n";
    return Parse::ExitError;
  }
  if (const char *mesh_env = std::getenv("NSX_MESH_FILE")) o.mesh_file_name = mesh_env;
  return Parse::Run;
}

inline void print_banner(const Options &o, bool unsteady) {
  static const char *solver_names[] = {"GMRES\n", "FGMRES\n", "Bicgstab\n"};
  static const char *prec_names[] = {"blockDiagonal\n", "blockTriangular\n", "aSIMPLE\n"};
  std::cout << "--------- CONFIGURATION PARAMETERS --------- \n";
  if (unsteady) std::cout << "Time span: " << o.time_span << "\n" << "Time step: " << o.time_step << "\n";
  std::cout << "Mesh size: " << o.mesh_size_x << "x" << o.mesh_size_y << "\n";
  std::cout << "Reynolds number: " << o.Re << "\n";
  std::cout << "Solver type: ";
  if (o.solver_type >= 0 && o.solver_type <= 2) std::cout << solver_names[o.solver_type];   // other values print nothing, as the reference
  std::cout << "Tolerance: " << o.tolerance << "\n";
  std::cout << "Preconditioner: ";
  if (o.preconditioner >= 0 && o.preconditioner <= 2) std::cout << prec_names[o.preconditioner];
  std::cout << "-----------------------------------------------\n";
}

}  // namespace app
