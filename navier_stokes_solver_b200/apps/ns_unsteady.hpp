// ns_unsteady.hpp -- host class of the time-dependent solver: the reference's entry points with the device
// library behind them.
//
// Mirrors NSSolver (lab_new/src/NSSolver.hpp:388-551, NSSolver.cpp:3-975): backward Euler time loop, a Newton
// solve with its own Reynolds continuation 1, 11, 21, ... inside every time step, the first assembly of a
// step in the massless Stokes branch, the inlet imposed in the very first assembly only, line search
// accepting on `<=`.  All matrix / vector arithmetic goes through include/nsx.h.
#pragma once
#include <iomanip>

#include "nsx_app.hpp"

namespace app {

class NSSolver {
 public:
  struct InletVelocity {  // NSSolver.hpp:57-90
    const double U_m = 0.3, H = 0.41;
    double value_x(double y) const { return 4 * U_m * y * (H - y) / (H * H); }
  };

  NSSolver(const std::string &mesh_file_name_, unsigned degree_velocity_, unsigned degree_pressure_, double T_, double deltat_, int mesh_size_x_,
           int mesh_size_y_, int solver_type_, double tolerance_, int preconditioner_, double Re_, bool read_mesh_from_file_)
      : mesh_file_name(mesh_file_name_), degree_velocity(degree_velocity_), degree_pressure(degree_pressure_), T(T_), deltat(deltat_),
        mesh_size_x(mesh_size_x_), mesh_size_y(mesh_size_y_), solver_type(solver_type_), tolerance(tolerance_), preconditioner_type(preconditioner_),
        Re(Re_), read_mesh_from_file(read_mesh_from_file_), pcout(prob.ranks.rank == 0) {}

  // NSSolver.cpp:3-311
  void setup() {
    pcout << "Initializing the mesh" << std::endl;
    if (read_mesh_from_file) pcout << "Mesh file name = " << mesh_file_name << std::endl;
    prob.make_mesh(read_mesh_from_file, mesh_file_name, mesh_size_x, mesh_size_y);
    if (read_mesh_from_file) pcout << "Here1" << std::endl;
    pcout << "  Number of elements = " << prob.ginfo(NSX_DI_NCELLS) << std::endl;
    if (!read_mesh_from_file) pcout << "Mesh written to mesh.msh" << std::endl;
    pcout << "-----------------------------------------------" << std::endl;
    pcout << "Initializing the finite element space" << std::endl;
    pcout << "  Velocity degree:           = " << degree_velocity << std::endl;
    pcout << "  Pressure degree:           = " << degree_pressure << std::endl;
    pcout << "  DoFs per cell              = " << prob.ginfo(NSX_DI_DOFS_PER_CELL) << std::endl;
    pcout << "  Quadrature points per cell = " << prob.ginfo(NSX_DI_NQ) << std::endl;
    pcout << "  Quadrature points per face = " << prob.ginfo(NSX_DI_NQF) << std::endl;
    pcout << "-----------------------------------------------" << std::endl;
    pcout << "Initializing the DoF handler" << std::endl;
    const int64_t n_u = prob.ginfo(NSX_DI_N_U), n_p = prob.ginfo(NSX_DI_N_P);
    pcout << "  Number of DoFs: " << std::endl;
    pcout << "    velocity = " << n_u << std::endl;
    pcout << "    pressure = " << n_p << std::endl;
    pcout << "    total    = " << n_u + n_p << std::endl;
    pcout << "-----------------------------------------------" << std::endl;
    pcout << "Initializing the linear system" << std::endl;
    pcout << "  Initializing the sparsity pattern" << std::endl;
    pcout << "  Initializing the matrices" << std::endl;
    pcout << "  Initializing the system right-hand side" << std::endl;
    pcout << "  Initializing the solution vector" << std::endl;
    prob.to_device(inlet_velocity.U_m);
  }

  // NSSolver.cpp:313-599: the inlet goes in only when `first_iter && apply_first` (:573)
  void assemble_system(bool first_iter) {
    check(prob.ctx, nsx_assemble(prob.ctx, first_iter ? NSX_MODE_UNSTEADY_FIRST : NSX_MODE_UNSTEADY_NEWTON, (first_iter && apply_first) ? 1 : 0, nu, deltat,
                                 p_out, &last_residual_norm), "nsx_assemble");
  }

  // line-search assembly (NSSolver.cpp:733-736): residual only by default, see ns_stationary.hpp
  void assemble_for_line_search() {
    static const bool full = [] { const char *e = std::getenv("NSX_FULL_LINESEARCH_ASSEMBLY"); return e && e[0] == '1'; }();
    if (full) { assemble_system(false); return; }
    check(prob.ctx, nsx_assemble_residual(prob.ctx, NSX_MODE_UNSTEADY_NEWTON, nu, deltat, p_out, &last_residual_norm), "nsx_assemble_residual");
  }

  // NSSolver.cpp:601-672
  int solve_system() {
    int it = 0;
    double res = 0;
    check(prob.ctx, nsx_solve(prob.ctx, NSX_UNSTEADY, solver_type, preconditioner_type, tolerance, 100000, 0.5, &it, &res), "nsx_solve");
    pcout << "   " << it << " iterations" << std::endl;   // NSSolver.cpp:670 (the stationary binary says "solver iterations")
    return it;
  }

  // NSSolver.cpp:674-754
  void solve_newton() {
    pcout << "===============================================" << std::endl;
    const unsigned int n_max_iters = 10;
    const double residual_tolerance = 1e-9;
    const double target_Re = Re;
    bool first_iter = true;
    pcout << "Target Re = " << target_Re << std::endl;
    for (double current_Re = 1.0; current_Re <= target_Re; current_Re += 10.0) {
      pcout << "===============================================" << std::endl;
      nu = 1.0 / current_Re;
      pcout << "Solving for Re = " << get_reynolds() << std::endl;
      unsigned int n_iter = 0;
      double residual_norm = residual_tolerance + 1, prev_residual = 0;
      int GMRES_iter = 0;
      while (n_iter < n_max_iters && residual_norm > residual_tolerance) {
        if (first_iter) { first_iter = false; assemble_system(n_iter == 0); }
        else assemble_system(false);
        residual_norm = last_residual_norm;
        prev_residual = n_iter == 0 ? residual_norm + 1 : prev_residual;
        pcout << "Newton iteration " << n_iter << "/" << n_max_iters << " - ||r|| = " << std::scientific << std::setprecision(6) << residual_norm
              << std::flush;
        if (residual_norm > residual_tolerance) {
          GMRES_iter = solve_system();
          krylov_iterations.push_back(GMRES_iter);
          if (GMRES_iter == 0) break;
          check(prob.ctx, nsx_save_eval_point(prob.ctx), "nsx_save_eval_point");
          for (double alpha = 1; alpha > 1e-12; alpha *= 0.1) {
            check(prob.ctx, nsx_update(prob.ctx, alpha), "nsx_update");
            assemble_for_line_search();
            residual_norm = last_residual_norm;
            pcout << "  Evaluating alpha=" << alpha << ", ||r||=" << residual_norm << std::endl;
            if (residual_norm <= prev_residual) break;
          }
          prev_residual = residual_norm;
        } else {
          pcout << " < tolerance" << std::endl;
          break;
        }
        ++n_iter;
      }
    }
    pcout << "===============================================" << std::endl;
  }

  // NSSolver.cpp:799-837
  void solve() {
    pcout << "===============================================" << std::endl;
    time = 0.0;
    output(0);
    pcout << "-----------------------------------------------" << std::endl;
    unsigned int time_step = 0;
    while (time < T - 0.5 * deltat) {
      time += deltat;
      ++time_step;
      check(prob.ctx, nsx_copy_old(prob.ctx), "nsx_copy_old");
      pcout << "n = " << std::setw(3) << time_step << ", t = " << std::setw(5) << std::fixed << time << std::endl;
      solve_newton();
      apply_first = false;
      output(time_step);
      compute_lift_drag();
      print_lift_coeff();
      print_drag_coeff();
      pcout << std::endl;
      if (max_time_steps && time_step >= max_time_steps) break;
    }
  }

  // NSSolver.cpp:761-797
  void output(const unsigned int &time_step) const {
    pcout << "===============================================" << std::endl;
    const std::string output_file_name = "output-stokes";   // printed name; the files are output_<step> (NSSolver.cpp:788-793)
    if (write_output) prob.write_vtu("output", time_step);
    pcout << "Output written to " << output_file_name << std::endl;
    pcout << "===============================================" << std::endl;
  }

  // NSSolver.cpp:839-938
  void compute_lift_drag() {
    pcout << "===============================================" << std::endl;
    pcout << "Computing lift and drag forces" << std::endl;
    int64_t n_faces = 0;
    prob.arr<int32_t>(NSX_DA_CYL_CELL, &n_faces);
    for (int64_t k = 0; k < n_faces; ++k) pcout << "Debug " << std::endl;   // NSSolver.cpp:883
    check(prob.ctx, nsx_lift_drag(prob.ctx, nu, &drag_force, &lift_force), "nsx_lift_drag");
    pcout << "Lift force: " << lift_force << std::endl;
    pcout << "Drag force: " << drag_force << std::endl;
    lift_history.push_back(lift_force); drag_history.push_back(drag_force);
  }
  double get_avg_inlet_velocity() const { return 2 * inlet_velocity.value_x(0.41 / 2.0) / 3; }
  double get_reynolds() const { return get_avg_inlet_velocity() * 0.1 / nu; }
  void compute_lift_coeff() { const double U_avg = get_avg_inlet_velocity(); lift_coeff = 2 * lift_force / (U_avg * U_avg * 0.1); }
  void compute_drag_coeff() { const double U_avg = get_avg_inlet_velocity(); drag_coeff = 2 * drag_force / (U_avg * U_avg * 0.1); }
  void print_lift_coeff() {
    pcout << "===============================================" << std::endl;
    compute_lift_coeff();
    pcout << "Lift coefficient: " << lift_coeff << std::endl;
    print_more_digits(pcout, "lift coefficient", lift_coeff);
  }
  void print_drag_coeff() {
    pcout << "===============================================" << std::endl;
    compute_drag_coeff();
    pcout << "Drag coefficient: " << drag_coeff << std::endl;
    print_more_digits(pcout, "drag coefficient", drag_coeff);
  }

  Problem prob;
  bool write_output = true;
  unsigned max_time_steps = 0;   // 0: run the whole span (test hook: NSX_MAX_TIME_STEPS)
  std::vector<int> krylov_iterations;
  std::vector<double> lift_history, drag_history;
  double lift_force = 0, drag_force = 0, lift_coeff = 0, drag_coeff = 0, last_residual_norm = 0;

 protected:
  std::string mesh_file_name;
  unsigned degree_velocity, degree_pressure;
  double T, deltat;
  int mesh_size_x, mesh_size_y, solver_type;
  double tolerance;
  int preconditioner_type;
  double Re;
  bool read_mesh_from_file;
  double nu = 1.0;
  const double p_out = 1.0;
  double time = 0.0;
  bool apply_first = true;
  InletVelocity inlet_velocity;
  mutable Pcout pcout;
};

}  // namespace app
