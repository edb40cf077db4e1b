// nsx_over_oracle.cpp -- TEST INFRASTRUCTURE ONLY: the subset of the device C ABI (include/nsx.h) that the executables
// call, implemented over the CPU oracle (oracle/liboracle.so).  Linked into CPU-only copies of StationaryNSSolver / NSSolver
// by tests/test_apps_host_logic.py so that the host-side Newton / continuation / time loops of navier_stokes_solver_b200/apps
// can be compared, print for print, with the oracle's own restatement of those loops -- without a GPU.  Nothing in the
// product links this file; libnsx.so has no CPU path.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nsx.h"

struct Problem;  // oracle handle
extern "C" {
Problem *orc_create(int elem, int ncells, const double *cell_vertices, const uint32_t *cell_dofs, int64_t n_u, int64_t n_p);
void orc_destroy(Problem *P);
const char *orc_last_error();
int orc_set_pattern(Problem *P, int block, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int32_t *col);
int orc_set_faces(Problem *P, int kind, int n, const int *cell, const int *face);
int orc_set_dirichlet(Problem *P, int n, const uint32_t *dof, const double *inlet_value);
int orc_set_ranks(Problem *P, int nranks, const int64_t *owned_u, const int64_t *owned_p);
double *orc_vec(Problem *P, int which);
int orc_assemble(Problem *P, int mode, int apply_inlet, double nu, double dt, double p_out, double *residual_l2);
int orc_solve(Problem *P, int flavour, int solver, int prec, double tol, int max_it, double alpha, int *iters, double *final_res, int64_t *inner);
int orc_lift_drag(Problem *P, double nu, double *drag, double *lift);
}

struct nsx_ctx {
  Problem *P = nullptr;
  int64_t n = 0;
  std::string err;
};

static int map_rc(nsx_ctx *c, int rc) {
  if (rc == 0) return NSX_OK;
  c->err = orc_last_error();
  return rc == 1 ? NSX_E_NOCONV : rc == 2 ? NSX_E_BADARG : NSX_E_STATE;
}

extern "C" {

int nsx_create(int, int nranks, int, void *, nsx_ctx **out) {
  if (!out || nranks != 1) return NSX_E_BADARG;
  *out = new nsx_ctx;
  return NSX_OK;
}
int nsx_destroy(nsx_ctx *c) { if (c) { if (c->P) orc_destroy(c->P); delete c; } return NSX_OK; }
const char *nsx_last_error(const nsx_ctx *c) { return c ? c->err.c_str() : "null context"; }
int nsx_set_option(nsx_ctx *, int, int64_t) { return NSX_OK; }   // kernel / ordering choices have no meaning for the oracle
int nsx_set_discretisation(nsx_ctx *c, int elem, int64_t n_cells, const double *cell_vertices, const uint32_t *cell_dofs, int64_t n_u, int64_t n_p) {
  c->P = orc_create(elem, (int)n_cells, cell_vertices, cell_dofs, n_u, n_p);
  c->n = n_u + n_p;
  return c->P ? NSX_OK : NSX_E_STATE;
}
int nsx_set_pattern(nsx_ctx *c, int block, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int32_t *col) {
  return map_rc(c, orc_set_pattern(c->P, block, nrows, ncols, rowptr, col));
}
int nsx_set_faces(nsx_ctx *c, int kind, int64_t n, const int32_t *cell, const int32_t *face) { return map_rc(c, orc_set_faces(c->P, kind, (int)n, cell, face)); }
int nsx_set_dirichlet(nsx_ctx *c, int64_t n, const uint32_t *dof, const double *inlet_value) { return map_rc(c, orc_set_dirichlet(c->P, (int)n, dof, inlet_value)); }
int nsx_set_ranks(nsx_ctx *c, int nranks, const int64_t *owned_u, const int64_t *owned_p) { return map_rc(c, orc_set_ranks(c->P, nranks, owned_u, owned_p)); }
int nsx_set_partition(nsx_ctx *, int64_t, int64_t) { return NSX_E_STATE; }
int nsx_set_halo(nsx_ctx *, int, int, const int32_t *, const int64_t *, const int32_t *, const int64_t *) { return NSX_E_STATE; }
int nsx_comm_unique_id(void *) { return NSX_E_COMM; }
int nsx_comm_init(nsx_ctx *, const void *) { return NSX_E_COMM; }
int nsx_finalize_setup(nsx_ctx *) { return NSX_OK; }
int nsx_vec_download(nsx_ctx *c, int which, double *host) {
  const double *v = orc_vec(c->P, which);
  if (!v || !host) return NSX_E_BADARG;
  std::memcpy(host, v, (size_t)c->n * sizeof(double));
  return NSX_OK;
}
int nsx_halo_exchange(nsx_ctx *, int) { return NSX_OK; }
int nsx_vec_download_ghosts(nsx_ctx *, int, double *, double *) { return NSX_OK; }
int nsx_assemble(nsx_ctx *c, int mode, int apply_inlet, double nu, double dt, double p_out, double *residual_l2) {
  return map_rc(c, orc_assemble(c->P, mode, apply_inlet, nu, dt, p_out, residual_l2));
}
int nsx_assemble_residual(nsx_ctx *c, int mode, double nu, double dt, double p_out, double *residual_l2) {
  return map_rc(c, orc_assemble(c->P, mode, 0, nu, dt, p_out, residual_l2));   // the oracle has no short cut: same residual
}
int nsx_solve(nsx_ctx *c, int flavour, int solver, int prec, double tol, int max_it, double alpha, int *iterations, double *final_residual) {
  int64_t inner[3];
  return map_rc(c, orc_solve(c->P, flavour, solver, prec, tol, max_it, alpha, iterations, final_residual, inner));
}
int nsx_save_eval_point(nsx_ctx *c) {
  std::memcpy(orc_vec(c->P, 4), orc_vec(c->P, 0), (size_t)c->n * sizeof(double));
  return NSX_OK;
}
int nsx_update(nsx_ctx *c, double alpha) {
  double *s = orc_vec(c->P, 0);
  const double *e = orc_vec(c->P, 4), *d = orc_vec(c->P, 2);
  for (int64_t i = 0; i < c->n; ++i) s[i] = e[i] + alpha * d[i];
  return NSX_OK;
}
int nsx_copy_old(nsx_ctx *c) {
  std::memcpy(orc_vec(c->P, 1), orc_vec(c->P, 0), (size_t)c->n * sizeof(double));
  return NSX_OK;
}
int nsx_lift_drag(nsx_ctx *c, double nu, double *drag_force, double *lift_force) { return map_rc(c, orc_lift_drag(c->P, nu, drag_force, lift_force)); }

}  // extern "C"
