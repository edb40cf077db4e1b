// krylov.cu -- outer and inner Krylov solvers and the three block preconditioners, device-resident.
//
// Replaces, for solve_system (NSSolverStationary.cpp:579-647, NSSolver.cpp:601-672):
//   deal.II SolverGMRES / SolverFGMRES / SolverBicgstab on block vectors (K1-K3 of SURVEY.md 8a),
//   the inner SolverFGMRES / SolverCG on single blocks (I4), and the preconditioner classes
//   PreconditionBlockDiagonal / BlockTriangular / aSIMPLE in both flavours
//   (NSSolverStationary.hpp:115-335, NSSolver.hpp:138-384).
// The iteration logic of the OUTER solvers (what deal.II's solvers do step by step, default AdditionalData) runs on the
// host; every vector lives on the device and all arithmetic on vectors is a CUDA kernel.  The INNER FGMRES solves on F --
// a couple of hundred thousand iterations per Newton step of the README run -- keep their recurrences on the device
// (solver_fgmres_dev): Hessenberg column, Givens rotations, residual estimate and SolverControl::check run in k_fg_step (or in
// the last CTA of the norm kernel), the host polls a mapped record instead of synchronising the stream, and the next
// iteration's sweep is queued behind the verdict and skips itself if the solve is over.
// Orthogonalisation of GMRES / FGMRES (NSX_OPT_ORTHO): 0 = deal.II's modified Gram-Schmidt as a chain of fused
// add_and_dot kernels that hand their coefficients on in device memory; 1 = two passes of batched classical
// Gram-Schmidt (all dot products of a pass in one launch, all updates + the norm in another); 2 (default) = as 1 for
// the outer solver, and for the inner FGMRES solves one pass plus a second only after heavy cancellation.  Every
// variant costs one host read per Krylov iteration, not one per basis vector.
#include <cmath>
#include <functional>
#include <limits>

#include "device.cuh"

namespace nsx {

namespace {

typedef std::function<void(double *dst, const double *src)> DOp;

enum State { ITERATE, SUCCESS, FAILURE };
struct Control {  // deal.II SolverControl
  int max_steps; double tol; int last_step = 0; double last_value = 0;
  Control(int m, double t) : max_steps(m), tol(t) {}
  State check(int step, double value) {
    last_step = step; last_value = value;
    if (value <= tol) return SUCCESS;
    if (step >= max_steps || std::isnan(value)) return FAILURE;
    return ITERATE;
  }
};

// workspace of one solver instance: vectors of one length, allocated on first use, kept in the context
struct Work {
  Ctx &c;
  std::vector<DevBuf<double>> &pool;
  int64_t n;
  Work(Ctx &ctx, std::vector<DevBuf<double>> &p, int64_t n_) : c(ctx), pool(p), n(n_) {}
  double *vec(size_t i) {
    if (pool.size() <= i) pool.resize(i + 1);
    if (pool[i].n != (size_t)c.nvec) pool[i].alloc(c.nvec);  // room for the ghost tail of an SpMV input
    return pool[i].p;
  }
};

// deal.II Householder::least_squares on the (m+1) x m Hessenberg matrix, H[col][row]
double hessenberg_least_squares(const std::vector<std::vector<double>> &H, int m, double beta, std::vector<double> &y) {
  const int rows = m + 1;
  std::vector<double> A((size_t)rows * m), b(rows, 0.0), v(rows);
  for (int j = 0; j < m; ++j) for (int i = 0; i < rows; ++i) A[(size_t)i * m + j] = H[j][i];
  b[0] = beta;
  for (int j = 0; j < m; ++j) {
    double sigma = 0;
    for (int i = j; i < rows; ++i) sigma += A[(size_t)i * m + j] * A[(size_t)i * m + j];
    if (sigma == 0) continue;
    double s = std::sqrt(sigma);
    if (A[(size_t)j * m + j] > 0) s = -s;
    std::fill(v.begin(), v.end(), 0.0);
    for (int i = j; i < rows; ++i) v[i] = A[(size_t)i * m + j];
    v[j] -= s;
    double vv = 0;
    for (int i = j; i < rows; ++i) vv += v[i] * v[i];
    if (vv == 0) continue;
    for (int cc = j; cc < m; ++cc) {
      double t = 0;
      for (int i = j; i < rows; ++i) t += v[i] * A[(size_t)i * m + cc];
      t = 2 * t / vv;
      for (int i = j; i < rows; ++i) A[(size_t)i * m + cc] -= t * v[i];
    }
    double t = 0;
    for (int i = j; i < rows; ++i) t += v[i] * b[i];
    t = 2 * t / vv;
    for (int i = j; i < rows; ++i) b[i] -= t * v[i];
  }
  y.assign(m, 0.0);
  for (int i = m - 1; i >= 0; --i) {
    double s = b[i];
    for (int cc = i + 1; cc < m; ++cc) s -= A[(size_t)i * m + cc] * y[cc];
    y[i] = s / A[(size_t)i * m + i];
  }
  return std::fabs(b[m]);
}

// SolverCG::solve
void solver_cg(Ctx &c, Control &ctl, const DOp &A, double *x, const double *b, const DOp &M, Work &W) {
  const int64_t n = W.n;
  double *g = W.vec(0), *d = W.vec(1), *h = W.vec(2);
  int it = 0;
  if (vec_dot(c, x, x, n) != 0.0) { A(g, x); vec_axpy(c, g, -1.0, b, n); }
  else vec_equ(c, g, -1.0, b, n);
  double res = vec_norm(c, g, n);
  State conv = ctl.check(0, res);
  if (conv != ITERATE) { if (conv != SUCCESS) throw NoConvergence(it, res, "CG (start)"); return; }
  M(h, g);
  vec_equ(c, d, -1.0, h, n);
  double gh = vec_dot(c, g, h, n);
  while (conv == ITERATE) {
    it++;
    A(h, d);
    double alpha = vec_dot(c, d, h, n);
    alpha = gh / alpha;
    vec_axpy(c, x, alpha, d, n);
    res = std::sqrt(std::fabs(vec_add_and_dot(c, g, alpha, h, g, n)));
    conv = ctl.check(it, res);
    if (conv != ITERATE) break;
    M(h, g);
    double beta = gh;
    gh = vec_dot(c, g, h, n);
    beta = gh / beta;
    vec_sadd(c, d, beta, -1.0, h, n);
  }
  if (conv != SUCCESS) throw NoConvergence(it, res, "CG");
}

// SolverFGMRES::solve (max_basis_size = 30)
void solver_fgmres(Ctx &c, Control &ctl, const DOp &A, double *x, const double *b, const DOp &M, Work &W, int slot0) {
  const int basis_size = 30;
  const int64_t n = W.n;
  std::vector<std::vector<double>> H(basis_size, std::vector<double>(basis_size + 1, 0.0));
  std::vector<double> y;
  std::vector<char> z_used(basis_size, 0);
  double hs[80];
  int accumulated_iterations = 0;
  double res = -std::numeric_limits<double>::max();
  double *aux = W.vec(0);
  auto v = [&](int j) { return W.vec(1 + j); };
  auto z = [&](int j) { return W.vec(1 + basis_size + j); };
  State state = ITERATE;
  do {
    A(aux, x);
    vec_sadd(c, aux, -1.0, 1.0, b, n);
    const double beta = vec_norm(c, aux, n);
    res = beta;
    state = ctl.check(accumulated_iterations, res);
    if (state == SUCCESS) break;
    for (auto &col : H) std::fill(col.begin(), col.end(), 0.0);
    double a = beta;
    y.clear();
    for (int j = 0; j < basis_size; ++j) {
      if (std::isfinite(a)) vec_equ(c, v(j), 1.0 / a, aux, n);
      else vec_set(c, v(j), 0.0, n);
      if (!z_used[j]) { vec_set(c, z(j), 0.0, n); z_used[j] = 1; }
      M(z(j), v(j));
      A(aux, z(j));
      if (c.ortho == 0) {  // modified Gram-Schmidt chain, the arithmetic of deal.II's SolverFGMRES
        vec_dot_dev(c, slot0, aux, v(0), n);
        for (int i = 1; i <= j; ++i) vec_add_and_dot_dev(c, slot0 + i, aux, -1.0, slot_ptr(c, slot0 + i - 1), v(i - 1), v(i), n);
        vec_add_and_dot_dev(c, slot0 + j + 1, aux, -1.0, slot_ptr(c, slot0 + j), v(j), aux, n);
        read_slots(c, slot0, j + 2, hs);
        for (int i = 0; i <= j; ++i) H[j][i] = hs[i];
        H[j][j + 1] = a = std::sqrt(hs[j + 1]);
      } else if (c.ortho == 2 && slot0 != 0) {
        // inner solves (tolerances 1e-1 ... 1e-5 relative): one pass of batched classical Gram-Schmidt, and a second one
        // only when the first removed more than 99 % of the vector's square norm (|w'|^2 < 0.01 (|w'|^2 + sum h^2)), i.e.
        // when cancellation would cost more than a digit of orthogonality.  deal.II's SolverFGMRES itself orthogonalises
        // once (modified Gram-Schmidt, no second pass).
        VecList V;
        for (int i = 0; i <= j; ++i) V.v[i] = v(i);
        vec_multi_dot_dev(c, slot0, V, j + 1, aux, n);
        vec_multi_axpy_norm_dev(c, slot0 + 64, V, j + 1, slot0, aux, n);
        read_slots(c, slot0, 65, hs);
        double h2 = 0;
        for (int i = 0; i <= j; ++i) { H[j][i] = hs[i]; h2 += hs[i] * hs[i]; }
        double nrm2 = hs[64];
        if (nrm2 < 0.01 * (nrm2 + h2)) {
          vec_multi_dot_dev(c, slot0 + 32, V, j + 1, aux, n);
          vec_multi_axpy_norm_dev(c, slot0 + 65, V, j + 1, slot0 + 32, aux, n);
          read_slots(c, slot0 + 32, 34, hs + 32);
          for (int i = 0; i <= j; ++i) H[j][i] += hs[32 + i];
          nrm2 = hs[65];
        }
        H[j][j + 1] = a = std::sqrt(nrm2);
      } else {  // two passes of batched classical Gram-Schmidt: 4 launches and one host read per column
        VecList V;
        for (int i = 0; i <= j; ++i) V.v[i] = v(i);
        vec_multi_dot_dev(c, slot0, V, j + 1, aux, n);
        vec_multi_axpy_norm_dev(c, slot0 + 64, V, j + 1, slot0, aux, n);
        vec_multi_dot_dev(c, slot0 + 32, V, j + 1, aux, n);
        vec_multi_axpy_norm_dev(c, slot0 + 65, V, j + 1, slot0 + 32, aux, n);
        read_slots(c, slot0, 66, hs);
        for (int i = 0; i <= j; ++i) H[j][i] = hs[i] + hs[32 + i];
        H[j][j + 1] = a = std::sqrt(hs[65]);
      }
      if (j > 0) {
        res = hessenberg_least_squares(H, j, beta, y);
        state = ctl.check(++accumulated_iterations, res);
        if (state != ITERATE) break;
      }
    }
    for (size_t j = 0; j < y.size(); ++j) vec_axpy(c, x, y[j], z((int)j), n);
  } while (state == ITERATE);
  if (state != SUCCESS) throw NoConvergence(accumulated_iterations, res, slot0 ? "inner FGMRES" : "FGMRES");
}

// Inner preconditioner of a device-driven FGMRES solve: a block-local sweep (fusable with the scaling of the next basis
// vector and launchable behind the device-side verdict), or any other operator
struct InnerM {
  TriPlan *plan = nullptr;   // block-local SGS / ILU(0) plan, or null
  bool sgs = false;
  bool node_layout = false;  // the solve runs on vectors in the node layout (decouple.cu)
  DOp op;                    // used when plan is null or not block-local
  bool fusable() const { return plan && plan->nblk > 0; }
};

// SolverFGMRES::solve for the inner solves (max_basis_size = 30), recurrences on the device.  The host only launches: the
// Hessenberg update, the residual estimate and SolverControl::check run in k_fg_step, whose verdict lands in a mapped host
// record that the host polls (no stream synchronisation).  With a block-local sweep as the preconditioner the next
// iteration's first kernel (scaling fused into the sweep) is already queued behind that verdict and skips itself if the
// solve is over, so the GPU never waits for the host.
void solver_fgmres_dev(Ctx &c, Control &ctl, const DOp &A, double *x, const double *b, const InnerM &M, Work &W, int slot0) {
  const int basis_size = 30;
  const int64_t n = W.n;
  FgDev *st = fg_state(c);
  double *aux = W.vec(0);
  auto v = [&](int j) { return W.vec(1 + j); };
  auto z = [&](int j) { return W.vec(1 + basis_size + j); };
  for (int j = 0; j < basis_size; ++j) { v(j); z(j); }   // allocate before the first pointer is baked into a launch
  aux = W.vec(0);
  const double *slots = slot_ptr(c, slot0);
  const int mode1 = c.ortho == 0 ? 0 : (c.ortho == 2 ? 1 : 3);
  int it = 0;
  double res = 0;
  int gate = 0;
  auto launch_M = [&](int j) {   // v_j = aux / a ; z_j = M^-1 v_j, both behind the device-side verdict
    if (M.fusable()) {
      if (M.sgs && M.plan->bl_sgs != 1) bl_refresh(c, *M.plan, true);
      bl_sweep(c, *M.plan, M.sgs, z(j), aux, &st->a, v(j), &st->gate, M.node_layout);
    } else {
      vec_scale_to_dev(c, v(j), aux, &st->a, &st->gate, n);
      M.op(z(j), v(j));
    }
  };
  do {
    A(aux, x);
    vec_sadd(c, aux, -1.0, 1.0, b, n);
    vec_dot_dev(c, slot0 + 70, aux, aux, n);
    fg_begin(c, slot_ptr(c, slot0 + 70), ctl.tol, ctl.max_steps, it);
    const bool spec = M.fusable();
    if (spec) launch_M(0);
    FgRec r = fg_wait(c);
    it = r.it; res = r.res; gate = r.gate;
    if (gate == 2) break;
    int j = 0;
    for (; j < basis_size; ++j) {
      if (!spec) launch_M(j);
      A(aux, z(j));
      bool stepped = false;
      VecList V;
      for (int i = 0; i <= j; ++i) V.v[i] = v(i);
      if (mode1 == 0) {
        vec_dot_dev(c, slot0, aux, v(0), n);
        for (int i = 1; i <= j; ++i) vec_add_and_dot_dev(c, slot0 + i, aux, -1.0, slot_ptr(c, slot0 + i - 1), v(i - 1), v(i), n);
        vec_add_and_dot_dev(c, slot0 + j + 1, aux, -1.0, slot_ptr(c, slot0 + j), v(j), aux, n);
      } else if (mode1 == 1 && c.comm) {
        // partitioned: ONE reduction per iteration -- the vector itself rides along as the last "basis vector", so that |w|^2 arrives
        // with the coefficients and the norm after the update follows from Pythagoras (k_fg_step mode 4)
        VecList V2 = V;
        V2.v[j + 1] = aux;
        vec_multi_dot_dev(c, slot0, V2, j + 2, aux, n);
        vec_multi_axpy_norm_dev(c, slot0 + 64, V, j + 1, slot0, aux, n, false);
        fg_step(c, slots, j, 4);
        stepped = true;
      } else if (mode1 == 1) {
        vec_multi_dot_dev(c, slot0, V, j + 1, aux, n);
        stepped = vec_multi_axpy_norm_fg(c, slot0 + 64, V, j + 1, slot0, aux, n, slots, j, 1);   // step folded into the norm kernel's last CTA
      } else {
        vec_multi_dot_dev(c, slot0, V, j + 1, aux, n);
        vec_multi_axpy_norm_dev(c, slot0 + 64, V, j + 1, slot0, aux, n);
        vec_multi_dot_dev(c, slot0 + 32, V, j + 1, aux, n);
        stepped = vec_multi_axpy_norm_fg(c, slot0 + 65, V, j + 1, slot0 + 32, aux, n, slots, j, 3);
      }
      if (!stepped) fg_step(c, slots, j, mode1);
      if (spec && j + 1 < basis_size) launch_M(j + 1);
      r = fg_wait(c);
      if (r.gate == 1) {   // heavy cancellation in the first pass: orthogonalise again, then finish the step
        vec_multi_dot_dev(c, slot0 + 32, V, j + 1, aux, n);
        if (!vec_multi_axpy_norm_fg(c, slot0 + 65, V, j + 1, slot0 + 32, aux, n, slots, j, 2)) fg_step(c, slots, j, 2);
        if (spec && j + 1 < basis_size) launch_M(j + 1);   // the launch queued before saw gate 1 and skipped itself
        r = fg_wait(c);
      }
      it = r.it; res = r.res; gate = r.gate;
      if (gate != 0) break;
    }
    // x += sum_{m < ny} y_m z_m, count and coefficients on the device
    VecList Z;
    for (int i = 0; i < basis_size; ++i) Z.v[i] = z(i);
    vec_multi_add_dev(c, x, Z, st->y, &st->ny, n);
  } while (gate == 0);
  ctl.last_step = it; ctl.last_value = res;
  if (gate != 2) throw NoConvergence(it, res, "inner FGMRES");
}

// SolverGMRES::solve (max_n_tmp_vectors = 30, left preconditioning, preconditioned residual,
// modified Gram-Schmidt with the re-orthogonalisation test every 5th step)
void solver_gmres(Ctx &c, Control &ctl, const DOp &A, double *x, const double *b, const DOp &M, Work &W, int slot0) {
  const int n_tmp = 30;
  const int64_t n = W.n;
  std::vector<std::vector<double>> H(n_tmp - 1, std::vector<double>(n_tmp, 0.0));
  std::vector<double> gamma(n_tmp), ci(n_tmp - 1), si(n_tmp - 1), h(n_tmp - 1);
  std::vector<char> used(n_tmp, 0);
  double hs[80];
  int accumulated_iterations = 0, dim = 0;
  State state = ITERATE;
  double last_res = -std::numeric_limits<double>::max();
  auto tmp = [&](int i) {
    double *p = W.vec(i);
    if (!used[i]) { vec_set(c, p, 0.0, n); used[i] = 1; }
    return p;
  };
  double *vfirst = tmp(0);
  double *p = tmp(n_tmp - 1);
  bool re_orthogonalize = false;
  do {
    std::fill(h.begin(), h.end(), 0.0);
    A(p, x);
    vec_sadd(c, p, -1.0, 1.0, b, n);
    M(vfirst, p);
    double rho = vec_norm(c, vfirst, n);
    last_res = rho;
    state = ctl.check(accumulated_iterations, rho);
    if (state != ITERATE) break;
    gamma[0] = rho;
    vec_scale(c, vfirst, 1.0 / rho, n);
    for (int inner = 0; inner < n_tmp - 2 && state == ITERATE; ++inner) {
      ++accumulated_iterations;
      double *vv = tmp(inner + 1);
      A(p, tmp(inner));
      M(vv, p);
      dim = inner + 1;
      double norm_vv = 0;
      if (c.ortho == 0) {
        double norm_vv_start = 0;
        const bool consider = (re_orthogonalize == false) && (accumulated_iterations % 5 == 0);
        if (consider) norm_vv_start = vec_norm(c, vv, n);
        vec_dot_dev(c, slot0, vv, tmp(0), n);
        for (int i = 1; i < dim; ++i) vec_add_and_dot_dev(c, slot0 + i, vv, -1.0, slot_ptr(c, slot0 + i - 1), tmp(i - 1), tmp(i), n);
        vec_add_and_dot_dev(c, slot0 + dim, vv, -1.0, slot_ptr(c, slot0 + dim - 1), tmp(dim - 1), vv, n);
        read_slots(c, slot0, dim + 1, hs);
        for (int i = 0; i < dim; ++i) h[i] = hs[i];
        norm_vv = std::sqrt(hs[dim]);
        bool done = false;
        if (consider) {
          if (norm_vv > 10. * norm_vv_start * std::sqrt(std::numeric_limits<double>::epsilon())) done = true;
          else re_orthogonalize = true;
        }
        if (!done && re_orthogonalize) {
          vec_dot_dev(c, slot0, vv, tmp(0), n);
          for (int i = 1; i < dim; ++i) vec_add_and_dot_dev(c, slot0 + i, vv, -1.0, slot_ptr(c, slot0 + i - 1), tmp(i - 1), tmp(i), n);
          vec_add_and_dot_dev(c, slot0 + dim, vv, -1.0, slot_ptr(c, slot0 + dim - 1), tmp(dim - 1), vv, n);
          read_slots(c, slot0, dim + 1, hs);
          for (int i = 0; i < dim; ++i) h[i] += hs[i];
          norm_vv = std::sqrt(hs[dim]);
        }
      } else {
        VecList V;
        for (int i = 0; i < dim; ++i) V.v[i] = tmp(i);
        vec_multi_dot_dev(c, slot0, V, dim, vv, n);
        vec_multi_axpy_norm_dev(c, slot0 + 64, V, dim, slot0, vv, n);
        vec_multi_dot_dev(c, slot0 + 32, V, dim, vv, n);
        vec_multi_axpy_norm_dev(c, slot0 + 65, V, dim, slot0 + 32, vv, n);
        read_slots(c, slot0, 66, hs);
        for (int i = 0; i < dim; ++i) h[i] = hs[i] + hs[32 + i];
        norm_vv = std::sqrt(hs[65]);
      }
      const double s = norm_vv;
      h[inner + 1] = s;
      if (std::isfinite(1. / s)) vec_scale(c, vv, 1. / s, n);
      for (int i = 0; i < inner; ++i) {
        const double sn = si[i], cs = ci[i], dummy = h[i];
        h[i] = cs * dummy + sn * h[i + 1];
        h[i + 1] = -sn * dummy + cs * h[i + 1];
      }
      const double r = 1. / std::sqrt(h[inner] * h[inner] + h[inner + 1] * h[inner + 1]);
      si[inner] = h[inner + 1] * r;
      ci[inner] = h[inner] * r;
      h[inner] = ci[inner] * h[inner] + si[inner] * h[inner + 1];
      gamma[inner + 1] = -si[inner] * gamma[inner];
      gamma[inner] *= ci[inner];
      for (int i = 0; i < dim; ++i) H[inner][i] = h[i];
      rho = std::fabs(gamma[dim]);
      last_res = rho;
      state = ctl.check(accumulated_iterations, rho);
    }
    std::vector<double> yv(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int cc = i + 1; cc < dim; ++cc) s -= H[cc][i] * yv[cc];
      yv[i] = s / H[i][i];
    }
    for (int i = 0; i < dim; ++i) vec_axpy(c, x, yv[i], tmp(i), n);
  } while (state == ITERATE);
  if (state != SUCCESS) throw NoConvergence(accumulated_iterations, last_res, "GMRES");
}

// SolverBicgstab::solve (exact_residual = true; breakdown threshold of deal.II >= 9.4)
void solver_bicgstab(Ctx &c, Control &ctl, const DOp &A, double *x, const double *b, const DOp &M, Work &W) {
  const int64_t n = W.n;
  const double breakdown_tol = std::numeric_limits<double>::min();
  double *r = W.vec(0), *rbar = W.vec(1), *p = W.vec(2), *y = W.vec(3), *z = W.vec(4), *t = W.vec(5), *v = W.vec(6);
  for (double *q : {y, z, v}) vec_set(c, q, 0.0, n);
  int step = 0;
  double res = 0;
  State state = ITERATE;
  bool breakdown;
  do {
    breakdown = false;
    A(r, x);
    vec_sadd(c, r, -1.0, 1.0, b, n);
    res = vec_norm(c, r, n);
    {
      const State st = ctl.check(step, res);
      if (st == SUCCESS) { state = SUCCESS; break; }
      if (st == FAILURE) { state = FAILURE; break; }
    }
    state = ITERATE;
    double alpha = 1, omega = 1, rho = 1, rhobar, beta;
    vec_copy(c, rbar, r, n);
    bool startup = true;
    do {
      ++step;
      rhobar = vec_dot(c, r, rbar, n);
      if (std::fabs(rhobar) < breakdown_tol) { breakdown = true; break; }
      beta = rhobar * alpha / (rho * omega);
      rho = rhobar;
      if (startup) { vec_copy(c, p, r, n); startup = false; }
      else { vec_sadd(c, p, beta, 1.0, r, n); vec_axpy(c, p, -beta * omega, v, n); }
      M(y, p);
      A(v, y);
      rhobar = vec_dot(c, rbar, v, n);
      if (std::fabs(rhobar) < breakdown_tol) { breakdown = true; break; }
      alpha = rho / rhobar;
      res = std::sqrt(vec_add_and_dot(c, r, -alpha, v, r, n));
      if (ctl.check(step, res) == SUCCESS) { vec_axpy(c, x, alpha, y, n); state = SUCCESS; break; }
      M(z, r);
      A(t, z);
      rhobar = vec_dot(c, t, r, n);
      const double t_squared = vec_dot(c, t, t, n);
      if (t_squared < breakdown_tol) { breakdown = true; break; }
      omega = rhobar / t_squared;
      vec_axpy(c, x, alpha, y, n);
      vec_axpy(c, x, omega, z, n);
      vec_axpy(c, r, -omega, t, n);
      A(t, x);
      vec_axpy(c, t, -1.0, b, n);
      res = vec_norm(c, t, n);
      state = ctl.check(step, res);
    } while (state == ITERATE);
  } while (breakdown);
  if (state != SUCCESS) throw NoConvergence(step, res, "BiCGStab");
}

// ---------------------------------------------------------------------------------------------
// the six preconditioner variants
// ---------------------------------------------------------------------------------------------
struct Preconditioner {
  Ctx &c;
  int flavour, type;
  double alpha;
  TriPlan *F = nullptr, *Mp = nullptr, *S = nullptr;
  DOp opF, opM, opS, opKn;
  int view = 0;

  Preconditioner(Ctx &ctx, int fl, int ty, double al) : c(ctx), flavour(fl), type(ty), alpha(al) {}

  // the initialize() calls of solve_system: every inner preconditioner is rebuilt from the
  // current matrix values (NSSolverStationary.cpp:583-585, 601-604, 620-626)
  void initialize() {
    // F products of the inner solves: on the same-component entries only while the cross-component ones are exact zeros
    // (view 1), or on the scalar matrix over the velocity nodes when moreover F = K (x) I_2 (view 2) -- decouple.cu
    view = effective_view(c);
    const bool dec = view >= 1;
    if (dec) opF = [this](double *y, const double *x) { spmv(c, c.Fd, x, y); };
    else opF = [this](double *y, const double *x) { spmv(c, c.F, x, y); };
    opKn = [this](double *y, const double *x) { spmv(c, c.Kn, x, y); };   // node layout in, node layout out (view 2)
    // ILU(0) of K (x) I_2 is ILU(0)(K) (x) I_2 (fill on the zero cross couplings stays zero): the node plan serves ILU too;
    // the same-component plan (view 1) does not -- general cross positions receive fill
    const int ilu_variant = view == 2 ? 2 : 0;
    opM = [this](double *y, const double *x) { spmv(c, c.Mp, x, y); };
    if (c.n_ug || c.n_pg) {
      // partitioned: S holds the rank-local block only (what ILU(0) needs); the product is B (diag(F)^-1 (Bt x)) with the two
      // ghost imports of the SpMVs -- the same sum as the reference's assembled S x, in another order
      c.tmp_s.alloc(c.nvec);
      opS = [this](double *y, const double *x) {
        spmv(c, c.Bt, x, c.tmp_s.p);
        vec_mul(c, c.tmp_s.p, c.Dinv.p, c.n_u);
        spmv(c, c.B, c.tmp_s.p, y);
      };
    } else {
      opS = [this](double *y, const double *x) { spmv(c, c.S, x, y); };
    }
    // numeric builds, skipped while NSX_OPT_PRECOND_LAG lets the data of an earlier solve stand in (a plan that never held values, or
    // whose values are older than the lag, is always rebuilt)
    auto reuse = [this](long long built_at) { return c.precond_lag > 0 && built_at >= 0 && c.solve_seq - built_at <= c.precond_lag; };
    auto sgs_values = [&](TriPlan &P, const DevCSR &A) {
      if (!P.factored && reuse(P.built_at)) return;   // (a plan holding LU factors of another preconditioner type does not count)
      tri_refresh_values(c, P, A); P.built_at = c.solve_seq; c.stat_precond_builds++;
    };
    auto ilu = [&](TriPlan &P, const DevCSR &A) {
      if (P.factored && reuse(P.built_at)) return;
      ilu0_factor(c, P, A); P.built_at = c.solve_seq; c.stat_precond_builds++;
    };
    if (type == 0) {
      // Gauss-Seidel sweeps skip exact zeros too (plan variant 1); ILU(0) needs the full pattern (fill lands on those entries)
      F = &tri_plan(c, NSX_BLOCK_F, flavour == NSX_STATIONARY ? view : ilu_variant); Mp = &tri_plan(c, NSX_BLOCK_MP);
      if (flavour == NSX_STATIONARY) { sgs_values(*F, c.F); sgs_values(*Mp, c.Mp); }
      else { ilu(*F, c.F); ilu(*Mp, c.Mp); }
    } else if (type == 1) {
      Mp = &tri_plan(c, NSX_BLOCK_MP);
      if (flavour == NSX_STATIONARY) {
        if (!(c.amg && reuse(c.amg_built_at))) { amg_setup(c, c.F); c.amg_built_at = c.solve_seq; c.stat_precond_builds++; }
      } else { F = &tri_plan(c, NSX_BLOCK_F, ilu_variant); ilu(*F, c.F); }
      ilu(*Mp, c.Mp);
    } else {
      // S and diag(F)^-1 belong together (vmult uses both): one decision for the pair and for the factors of S
      const bool keep_S = c.S_symbolic && reuse(c.schur_built_at);
      if (!keep_S) { schur_complement(c); c.schur_built_at = c.solve_seq; c.stat_precond_builds++; }
      // The unsteady aSIMPLE applies ONE ILU(0) sweep pair per block and iteration (NSSolver.hpp:294-350), so its outer iteration
      // count follows the quality of the factorisation: unless the caller chose an order, the blocks keep Ifpack's natural order
      // inside (ordering 3; measured on the reference's mesh: 81 k outer iterations against 107 k with the multicolour order)
      const int ord = (flavour == NSX_UNSTEADY && c.ordering_auto && c.ordering == 2) ? 3 : -1;
      F = &tri_plan(c, NSX_BLOCK_F, ilu_variant, ord); S = &tri_plan(c, NSX_BLOCK_S, 0, ord);
      ilu(*F, c.F);
      if (!(keep_S && S->factored && S->built_at >= 0)) { ilu0_factor(c, *S, c.S); S->built_at = c.solve_seq; c.stat_precond_builds++; }
      c.delta_p.alloc(c.nvec);
      c.delta_p.zero(c.stream);
    }
    c.tmp_u.alloc(c.nvec);
    c.tmp_p.alloc(c.nvec);
  }

  // inner FGMRES on F: recurrences on the device unless NSX_OPT_HOST_INNER asks for the host-driven solver
  void inner_fgmres(Control &ctl, double *x, const double *b, TriPlan *plan, bool sgs, const DOp &op, Work &W, int slot0) {
    if (c.host_inner) { solver_fgmres(c, ctl, opF, x, b, op, W, slot0); return; }
    InnerM M;
    M.plan = plan; M.sgs = sgs; M.op = op;
    if (view == 2 && plan && plan->node && plan->nblk) {
      // F = K (x) I_2: the whole solve runs in the node layout -- one matrix value per node pair in the SpMV and in the sweeps --
      // and only its right-hand side, initial guess and result are permuted
      c.node_b.alloc(c.nvec); c.node_x.alloc(c.nvec);
      vec_to_node_layout(c, c.node_b.p, b);
      vec_to_node_layout(c, c.node_x.p, x);
      M.node_layout = true;
      try { solver_fgmres_dev(c, ctl, opKn, c.node_x.p, c.node_b.p, M, W, slot0); }
      catch (const NoConvergence &) { vec_from_node_layout(c, x, c.node_x.p); throw; }
      vec_from_node_layout(c, x, c.node_x.p);
      return;
    }
    solver_fgmres_dev(c, ctl, opF, x, b, M, W, slot0);
  }

  void vmult(double *dst, const double *src) {
    const int64_t nu = c.n_u, np = c.n_p;
    const double *su = src, *sp = src + nu;
    double *du = dst, *dp = dst + nu;
    Work WF(c, c.work_inner_u, nu), WP(c, c.work_inner_p, np);
    const int slot_inner = 80;
    c.stat_applies++;
    if (flavour == NSX_STATIONARY && type == 0) {  // NSSolverStationary.hpp:132-153
      Control cu(100001, 1e-1 * vec_norm(c, su, nu));
      inner_fgmres(cu, du, su, F, true, [&](double *y, const double *x) { sgs_apply(c, *F, y, x); }, WF, slot_inner);
      c.stat_inner_F += cu.last_step;
      Control cp(100000, 1e-1 * vec_norm(c, sp, np));
      solver_cg(c, cp, opM, dp, sp, [&](double *y, const double *x) { sgs_apply(c, *Mp, y, x); }, WP);
      c.stat_inner_S += cp.last_step;
    } else if (flavour == NSX_UNSTEADY && type == 0) {  // NSSolver.hpp:155-176
      Control cu(1000, 1e-1);
      inner_fgmres(cu, du, su, F, false, [&](double *y, const double *x) { ilu0_apply(c, *F, y, x); }, WF, slot_inner);
      c.stat_inner_F += cu.last_step;
      Control cp(1000, 1e-1);
      solver_cg(c, cp, opM, dp, sp, [&](double *y, const double *x) { ilu0_apply(c, *Mp, y, x); }, WP);
      c.stat_inner_S += cp.last_step;
    } else if (type == 1) {  // NSSolverStationary.hpp:190-218, NSSolver.hpp:212-237
      const bool st = flavour == NSX_STATIONARY;
      Control cu(st ? 10000001 : 2000001, (st ? 1e-2 : 1e-4) * vec_norm(c, su, nu));
      if (st) inner_fgmres(cu, du, su, nullptr, false, [&](double *y, const double *x) { amg_apply(c, y, x); }, WF, slot_inner);
      else inner_fgmres(cu, du, su, F, false, [&](double *y, const double *x) { ilu0_apply(c, *F, y, x); }, WF, slot_inner);
      c.stat_inner_F += cu.last_step;
      double *tmp = c.tmp_p.p;
      spmv(c, c.B, du, tmp);
      vec_sadd(c, tmp, -1.0, 1.0, sp, np);
      Control cp(st ? 100000 : 2000000, (st ? 1e-2 : 1e-5) * vec_norm(c, sp, np));
      solver_cg(c, cp, opM, dp, tmp, [&](double *y, const double *x) { ilu0_apply(c, *Mp, y, x); }, WP);
      c.stat_inner_S += cp.last_step;
    } else if (flavour == NSX_STATIONARY) {  // aSIMPLE, NSSolverStationary.hpp:282-311
      Control cu(100000, 1e-1 * vec_norm(c, su, nu));
      inner_fgmres(cu, du, su, F, false, [&](double *y, const double *x) { ilu0_apply(c, *F, y, x); }, WF, slot_inner);
      c.stat_inner_F += cu.last_step;
      double *tp = c.tmp_p.p, *tu = c.tmp_u.p, *delta_p = c.delta_p.p;
      spmv(c, c.B, du, tp);
      vec_sadd(c, tp, -1.0, 1.0, sp, np);
      Control cs(100000, 1e-1 * vec_norm(c, tp, np));
      solver_cg(c, cs, opS, delta_p, tp, [&](double *y, const double *x) { ilu0_apply(c, *S, y, x); }, WP);
      c.stat_inner_S += cs.last_step;
      vec_scale(c, delta_p, alpha, np);
      spmv(c, c.Bt, delta_p, tu);
      vec_mul(c, tu, c.Dinv.p, nu);
      vec_axpy(c, du, -1.0, tu, nu);
      vec_copy(c, dp, delta_p, np);
    } else {  // aSIMPLE, NSSolver.hpp:294-350
      double *tp = c.tmp_p.p, *tu = c.tmp_u.p;
      ilu0_apply(c, *F, du, su);
      vec_copy(c, tp, sp, np);
      spmv(c, c.B, du, tp, true);
      ilu0_apply(c, *S, dp, tp);
      vec_mul(c, du, c.Dvec.p, nu);
      vec_scale(c, dp, 1.0 / alpha, np);
      spmv(c, c.Bt, dp, tu);
      vec_axpy(c, du, -1.0, tu, nu);
      vec_mul(c, du, c.Dinv.p, nu);
    }
  }
};

}  // namespace

int solve_system(Ctx &c, int flavour, int solver, int prec, double tol, int max_it, double alpha, double *final_res) {
  if (prec < 0 || prec > 2)
    throw std::invalid_argument("Invalid preconditioner type. Use 0: blockDiagonal, 1: blockTriangular, 2: aSIMPLE.");
  c.stat_inner_F = c.stat_inner_S = c.stat_applies = 0;
  c.solve_seq++;
  Control ctl(max_it, tol);
  Preconditioner pc(c, flavour, prec, alpha);
  pc.initialize();
  const int64_t n = c.n;
  DOp A = [&c](double *y, const double *x) { block_spmv(c, x, y); };
  DOp M = [&pc](double *y, const double *x) { pc.vmult(y, x); };
  Work W(c, c.work_outer, n);
  double *x = c.vec[NSX_VEC_DELTA].p;
  const double *b = c.vec[NSX_VEC_RESIDUAL].p;
  if (solver == 0) solver_gmres(c, ctl, A, x, b, M, W, 0);
  else if (solver == 1) solver_fgmres(c, ctl, A, x, b, M, W, 0);
  else if (solver == 2) solver_bicgstab(c, ctl, A, x, b, M, W);
  if (final_res) *final_res = ctl.last_value;
  return ctl.last_step;
}

void precond_apply_once(Ctx &c, int flavour, int prec, double alpha, const double *src, double *dst) {
  if (prec < 0 || prec > 2) throw std::invalid_argument("Invalid preconditioner type.");
  Preconditioner pc(c, flavour, prec, alpha);
  pc.initialize();
  pc.vmult(dst, src);
}

}  // namespace nsx
