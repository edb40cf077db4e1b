// ns_stationary.hpp -- host class of the stationary solver: the reference's entry points with the device
// library behind them.
//
// Mirrors NSSolverStationary (lab_new/src/NSSolverStationary.hpp:339-376, .cpp:3-933): same public members
// (setup, assemble_system, solve_system, solve_newton, output, compute_lift_drag, print_*), same control
// flow of the Reynolds continuation / inlet ladder / Newton / line search (quirks of SURVEY.md appendix B
// included: the inlet is imposed once, Stokes mode lasts for the whole first Reynolds stage, delta is never
// zeroed between solves).  Every piece of arithmetic on matrices and vectors is a call into include/nsx.h.
#pragma once
#include <iomanip>

#include "nsx_app.hpp"

namespace app {

class NSSolverStationary {
 public:
  // inlet profile amplitude ladder (NSSolverStationary.hpp:60-108)
  struct InletVelocity {
    double u = 0.1;
    const double U_m = 1.0, H = 0.41;
    double value_x(double y) const { return 4 * u * y * (H - y) / (H * H); }
    double getVelocity() const { return u; }
    bool incrementVelocity(double re) {
      if (u == U_m) return true;
      u += 0.15;
      if (re == 0.0) u = 0.01;
      if (u > U_m) u = U_m;
      return false;
    }
  };

  NSSolverStationary(const std::string &mesh_file_name_, unsigned degree_velocity_, unsigned degree_pressure_, int mesh_size_x_, int mesh_size_y_,
                     int solver_type_, double tolerance_, int preconditioner_, double Re_, bool read_mesh_from_file_)
      : mesh_file_name(mesh_file_name_), degree_velocity(degree_velocity_), degree_pressure(degree_pressure_), mesh_size_x(mesh_size_x_),
        mesh_size_y(mesh_size_y_), solver_type(solver_type_), tolerance(tolerance_), preconditioner_type(preconditioner_), Re(Re_),
        read_mesh_from_file(read_mesh_from_file_), pcout(prob.ranks.rank == 0) {}

  // NSSolverStationary.cpp:3-315 (deal.II in the reference; include/nsx_host.h here), then the hand-over to the GPU
  void setup() {
    if (!read_mesh_from_file) {
      pcout << "Initializing the mesh" << std::endl;
      prob.make_mesh(false, mesh_file_name, mesh_size_x, mesh_size_y);
      pcout << "  Number of elements = " << prob.ginfo(NSX_DI_NCELLS) << std::endl;
      pcout << "Mesh written to mesh.msh" << std::endl;  // the mesh file itself is deal.II's business (GridOut)
    } else {
      pcout << "Initializing the mesh" << std::endl;
      pcout << "Mesh file name = " << mesh_file_name << std::endl;
      prob.make_mesh(true, mesh_file_name, 0, 0);
      pcout << "Here1" << std::endl;
      pcout << "  Number of elements = " << prob.ginfo(NSX_DI_NCELLS) << std::endl;
    }
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
    prob.to_device(inlet_velocity.getVelocity());
  }

  // NSSolverStationary.cpp:317-577; leaves ||residual_vector||_2 in last_residual_norm (the l2_norm() of :698, :729)
  void assemble_system(bool global_first_iter, bool computing_stokes) {
    const int mode = (global_first_iter || computing_stokes) ? NSX_MODE_STOKES : NSX_MODE_NEWTON;
    check(prob.ctx, nsx_assemble(prob.ctx, mode, global_first_iter ? 1 : 0, nu, 0.0, p_out, &last_residual_norm), "nsx_assemble");
  }

  // The line search's assembly (NSSolverStationary.cpp:724-727): the reference re-assembles everything for a norm; the matrices of
  // those assemblies are never read (the loop assembles again before the next solve), so by default only the residual is
  // formed -- bit-identical ||r||.  NSX_FULL_LINESEARCH_ASSEMBLY=1 restores the full assembly.
  void assemble_for_line_search(bool computing_stokes) {
    static const bool full = [] { const char *e = std::getenv("NSX_FULL_LINESEARCH_ASSEMBLY"); return e && e[0] == '1'; }();
    if (full) { assemble_system(false, computing_stokes); return; }
    check(prob.ctx, nsx_assemble_residual(prob.ctx, computing_stokes ? NSX_MODE_STOKES : NSX_MODE_NEWTON, nu, 0.0, p_out, &last_residual_norm), "nsx_assemble_residual");
  }

  // NSSolverStationary.cpp:579-647
  int solve_system() {
    int it = 0;
    double res = 0;
    check(prob.ctx, nsx_solve(prob.ctx, NSX_STATIONARY, solver_type, preconditioner_type, tolerance, 20000, 0.5, &it, &res), "nsx_solve");
    pcout << "   " << it << " solver iterations" << std::endl;
    return it;
  }

  // NSSolverStationary.cpp:649-758
  void solve_newton() {
    pcout << "===============================================" << std::endl;
    const unsigned int n_max_iters = 15;
    const double residual_tolerance = 1e-9;
    const double target_Re = Re;
    bool global_first_iter = true, computing_stokes = true, inlet_reached = false;
    pcout << "Target Re = " << target_Re << std::endl;
    for (double current_Re = 10.0; current_Re <= target_Re; current_Re += 20.0) {
      pcout << "===============================================" << std::endl;
      nu = 1.0 / current_Re;
      inlet_reached = false;
      pcout << "Solving for nu = " << nu << ", Re = " << get_reynolds() << std::endl;
      while (!inlet_reached) {
        pcout << "Solving for inlet velocity: " << inlet_velocity.getVelocity() << std::endl;
        if (global_first_iter) pcout << "Solving Stokes adding BCs" << std::endl;
        else if (computing_stokes) pcout << "Solving Stokes without adding BCs" << std::endl;
        else pcout << "Solving NS" << std::endl;
        unsigned int n_iter = 0;
        double residual_norm = residual_tolerance + 1, prev_residual = 0;
        int GMRES_iter = 0;
        while (n_iter < n_max_iters && residual_norm > residual_tolerance) {
          if (global_first_iter) { global_first_iter = false; assemble_system(true, true); }
          else assemble_system(false, computing_stokes);
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
              assemble_for_line_search(computing_stokes);
              residual_norm = last_residual_norm;
              pcout << "  Evaluating alpha=" << alpha << ", ||r||=" << residual_norm << std::endl;
              if (residual_norm < prev_residual) break;
            }
            prev_residual = residual_norm;
          } else {
            pcout << " < tolerance" << std::endl;
            output();
            break;
          }
          output();
          ++n_iter;
        }
        inlet_reached = inlet_velocity.incrementVelocity(get_reynolds());
        if (inlet_reached) computing_stokes = false;
      }
      output();
    }
    pcout << "===============================================" << std::endl;
  }

  // NSSolverStationary.cpp:765-800
  void output() const {
    pcout << "===============================================" << std::endl;
    const std::string output_file_name = "output-stokes";
    if (write_output) prob.write_vtu(output_file_name, 0);
    pcout << "Output written to " << output_file_name << std::endl;
    pcout << "===============================================" << std::endl;
  }

  // NSSolverStationary.cpp:802-897
  void compute_lift_drag() {
    pcout << "===============================================" << std::endl;
    pcout << "Computing lift and drag forces" << std::endl;
    int64_t n_faces = 0;
    prob.arr<int32_t>(NSX_DA_CYL_CELL, &n_faces);
    // the reference prints this line once per face quadrature point (NSSolverStationary.cpp:855); rank 0 only sees its own
    for (int64_t k = 0; k < n_faces * prob.info(NSX_DI_NQF); ++k) pcout << "Computing drag and lift forces" << std::endl;
    check(prob.ctx, nsx_lift_drag(prob.ctx, nu, &drag_force, &lift_force), "nsx_lift_drag");
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
  std::vector<int> krylov_iterations;
  double lift_force = 0, drag_force = 0, lift_coeff = 0, drag_coeff = 0, last_residual_norm = 0;

 protected:
  std::string mesh_file_name;
  unsigned degree_velocity, degree_pressure;
  int mesh_size_x, mesh_size_y, solver_type;
  double tolerance;
  int preconditioner_type;
  double Re;
  bool read_mesh_from_file;
  double nu = 1.0;              // NSSolverStationary.hpp:392
  const double p_out = 1.0;     // NSSolverStationary.hpp:398
  InletVelocity inlet_velocity;
  mutable Pcout pcout;
};

}  // namespace app
