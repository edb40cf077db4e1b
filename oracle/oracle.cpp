// oracle.cpp -- CPU restatement of the reference's assembly + linear-algebra hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product (libnsx.so) never
// links or calls it.
//
// *** PARITY UNPINNED. ***  The reference (HliasGit/navier_stokes_solver) cannot be built here
// (it needs deal.II >= 9.3.1 with Trilinos, MPI, METIS and Boost >= 1.72 -- none are in the
// image) and holds no golden vector or test for this path.  This file
// therefore restates
//   * the reference's own source, line by line where it spells out arithmetic
//     (lab_new/src/NSSolverStationary.cpp:317-577, 579-758, 802-919; NSSolver.cpp:313-599,
//      601-837; NSSolverStationary.hpp:115-335; NSSolver.hpp:138-384), and
//   * the published algorithms of the third-party pieces it calls (deal.II 9.3-9.5 SolverCG,
//     SolverGMRES, SolverFGMRES, SolverBicgstab, MatrixTools::apply_boundary_values; Trilinos
//     Ifpack point relaxation "symmetric Gauss-Seidel", Ifpack ILU(0), ML smoothed aggregation),
//     from knowledge of those libraries, not from their source.
// The only reference artefact that pins anything is the 154 244-DoF count of the 100x70 mesh
// (performance_analysis.ipynb, cell 2), which tests/test_host_setup.py checks.
#include "oracle.hpp"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>

namespace orc {

// ---------------------------------------------------------------------------------------------
// reference cell: nodes, shape functions, quadrature (deal.II FE_Q / FE_SimplexP conventions)
// ---------------------------------------------------------------------------------------------
static const double SQ5 = 2.23606797749978969641;
static const double GLL3[4] = {0.0, (1.0 - 1.0 / SQ5) / 2.0, (1.0 + 1.0 / SQ5) / 2.0, 1.0};
static const double EQ2[3] = {0.0, 0.5, 1.0};

static double lag(const double *x, int n, int k, double t) {
  double v = 1;
  for (int m = 0; m < n; ++m) if (m != k) v *= (t - x[m]) / (x[k] - x[m]);
  return v;
}
static double dlag(const double *x, int n, int k, double t) {
  double s = 0;
  for (int j = 0; j < n; ++j) {
    if (j == k) continue;
    double p = 1.0 / (x[k] - x[j]);
    for (int m = 0; m < n; ++m) if (m != k && m != j) p *= (t - x[m]) / (x[k] - x[m]);
    s += p;
  }
  return s;
}
// tensor-product node (i,j) of the hierarchically numbered FE_Q(p) node `a`
static void q_ij(int p, int a, int &i, int &j) {
  if (a < 4) { i = (a % 2) * p; j = (a / 2) * p; return; }
  a -= 4;
  const int m = p - 1;
  if (a < 4 * m) {
    const int line = a / m, k = 1 + a % m;
    switch (line) {
      case 0: i = 0; j = k; break;
      case 1: i = p; j = k; break;
      case 2: i = k; j = 0; break;
      default: i = k; j = p; break;
    }
    return;
  }
  a -= 4 * m;
  i = 1 + a % m; j = 1 + a / m;
}
// scalar shape function `a` of the velocity (which=0) or pressure (which=1) element at (x,y)
static void shape(const RefCell &rc, int which, int a, double x, double y, double &v, double g[2]) {
  if (rc.elem == 0) {
    const int p = which == 0 ? 3 : 2;
    const double *nd = which == 0 ? GLL3 : EQ2;
    int i, j;
    q_ij(p, a, i, j);
    const double vx = lag(nd, p + 1, i, x), vy = lag(nd, p + 1, j, y);
    v = vx * vy; g[0] = dlag(nd, p + 1, i, x) * vy; g[1] = vx * dlag(nd, p + 1, j, y);
  } else {
    const double L[3] = {1 - x - y, x, y};
    const double dL[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
    if (which == 1) { v = L[a]; g[0] = dL[a][0]; g[1] = dL[a][1]; return; }
    if (a < 3) {
      v = L[a] * (2 * L[a] - 1);
      g[0] = (4 * L[a] - 1) * dL[a][0]; g[1] = (4 * L[a] - 1) * dL[a][1];
    } else {
      const int s = a - 3, e = (s + 1) % 3;
      v = 4 * L[s] * L[e];
      g[0] = 4 * (dL[s][0] * L[e] + L[s] * dL[e][0]);
      g[1] = 4 * (dL[s][1] * L[e] + L[s] * dL[e][1]);
    }
  }
}

static void init_refcell(RefCell &rc, int elem) {
  std::memset(&rc, 0, sizeof(rc));
  rc.elem = elem;
  int n = 0;
  if (elem == 0) {
    rc.nvpc = 4; rc.ndofs = 41; rc.nq = 16; rc.nqf = 4; rc.nfaces = 4; rc.nvn = 16; rc.npn = 9;
    for (int v = 0; v < 4; ++v) for (int c = 0; c < 3; ++c) { rc.comp[n] = c; rc.node[n++] = v; }
    for (int l = 0; l < 4; ++l) {
      for (int c = 0; c < 2; ++c) for (int k = 0; k < 2; ++k) { rc.comp[n] = c; rc.node[n++] = 4 + 2 * l + k; }
      rc.comp[n] = 2; rc.node[n++] = 4 + l;
    }
    for (int c = 0; c < 2; ++c) for (int k = 0; k < 4; ++k) { rc.comp[n] = c; rc.node[n++] = 12 + k; }
    rc.comp[n] = 2; rc.node[n++] = 8;
    // 4-point Gauss-Legendre on [0,1]
    const double r = std::sqrt(6.0 / 5.0);
    const double xa = std::sqrt(3.0 / 7.0 - 2.0 / 7.0 * r) / 2, xb = std::sqrt(3.0 / 7.0 + 2.0 / 7.0 * r) / 2;
    const double gx[4] = {0.5 - xb, 0.5 - xa, 0.5 + xa, 0.5 + xb};
    const double wa = (18 + std::sqrt(30.0)) / 72, wb = (18 - std::sqrt(30.0)) / 72;
    const double gw[4] = {wb, wa, wa, wb};
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) {
      rc.qp[4 * j + i][0] = gx[i]; rc.qp[4 * j + i][1] = gx[j]; rc.qw[4 * j + i] = gw[i] * gw[j];
    }
    for (int i = 0; i < 4; ++i) { rc.qfp[i] = gx[i]; rc.qfw[i] = gw[i]; }
  } else {
    rc.nvpc = 3; rc.ndofs = 15; rc.nq = 7; rc.nqf = 3; rc.nfaces = 3; rc.nvn = 6; rc.npn = 3;
    for (int v = 0; v < 3; ++v) for (int c = 0; c < 3; ++c) { rc.comp[n] = c; rc.node[n++] = v; }
    for (int l = 0; l < 3; ++l) for (int c = 0; c < 2; ++c) { rc.comp[n] = c; rc.node[n++] = 3 + l; }
    const double s = std::sqrt(15.0);
    const double a = (6 - s) / 21, b = (6 + s) / 21;
    const double pts[7][2] = {{1 / 3., 1 / 3.}, {1 - 2 * a, a}, {a, 1 - 2 * a}, {a, a}, {1 - 2 * b, b}, {b, 1 - 2 * b}, {b, b}};
    const double w[7] = {9 / 80., (155 - s) / 2400, (155 - s) / 2400, (155 - s) / 2400, (155 + s) / 2400, (155 + s) / 2400, (155 + s) / 2400};
    for (int q = 0; q < 7; ++q) { rc.qp[q][0] = pts[q][0]; rc.qp[q][1] = pts[q][1]; rc.qw[q] = w[q]; }
    const double h = std::sqrt(3.0 / 5.0) / 2;
    rc.qfp[0] = 0.5 - h; rc.qfp[1] = 0.5; rc.qfp[2] = 0.5 + h;
    rc.qfw[0] = 5 / 18.; rc.qfw[1] = 4 / 9.; rc.qfw[2] = 5 / 18.;
  }
}

// ---------------------------------------------------------------------------------------------
// FEValues-like per-cell data
// ---------------------------------------------------------------------------------------------
struct CellValues {
  int n, nq;
  double JxW[16];
  double val[41][16][2];      // fe_values[velocity].value(i,q)
  double grad[41][16][2][2];  // fe_values[velocity].gradient(i,q)[k][l] = d_l phi_k
  double div[41][16];         // fe_values[velocity].divergence(i,q)
  double pv[41][16];          // fe_values[pressure].value(i,q)
  double normal[16][2];       // face only
};

static void geometry_at(const RefCell &rc, const double *xv, double x, double y, double J[2][2]) {
  if (rc.elem == 0) {
    const double dN[4][2] = {{-(1 - y), -(1 - x)}, {(1 - y), -x}, {-y, (1 - x)}, {y, x}};
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) {
      double s = 0;
      for (int v = 0; v < 4; ++v) s += xv[2 * v + r] * dN[v][c];
      J[r][c] = s;
    }
  } else {
    J[0][0] = xv[2] - xv[0]; J[0][1] = xv[4] - xv[0];
    J[1][0] = xv[3] - xv[1]; J[1][1] = xv[5] - xv[1];
  }
}

static void fill_point(const RefCell &rc, const double *xv, double x, double y, int q, CellValues &cv, double J[2][2]) {
  geometry_at(rc, xv, x, y, J);
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double Ji[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
  for (int i = 0; i < rc.ndofs; ++i) {
    const int c = rc.comp[i];
    double v, g[2];
    shape(rc, c == 2 ? 1 : 0, rc.node[i], x, y, v, g);
    // physical gradient = J^{-T} g_ref
    const double gx = Ji[0][0] * g[0] + Ji[1][0] * g[1], gy = Ji[0][1] * g[0] + Ji[1][1] * g[1];
    cv.val[i][q][0] = cv.val[i][q][1] = 0;
    for (int k = 0; k < 2; ++k) for (int l = 0; l < 2; ++l) cv.grad[i][q][k][l] = 0;
    cv.div[i][q] = 0; cv.pv[i][q] = 0;
    if (c < 2) {
      cv.val[i][q][c] = v;
      cv.grad[i][q][c][0] = gx; cv.grad[i][q][c][1] = gy;
      cv.div[i][q] = cv.grad[i][q][0][0] + cv.grad[i][q][1][1];
    } else
      cv.pv[i][q] = v;
  }
}

static void reinit_cell(const RefCell &rc, const double *xv, CellValues &cv) {
  cv.n = rc.ndofs; cv.nq = rc.nq;
  for (int q = 0; q < rc.nq; ++q) {
    double J[2][2];
    fill_point(rc, xv, rc.qp[q][0], rc.qp[q][1], q, cv, J);
    cv.JxW[q] = rc.qw[q] * std::fabs(J[0][0] * J[1][1] - J[0][1] * J[1][0]);
  }
}

static void face_ref_point(const RefCell &rc, int f, double t, double &x, double &y) {
  if (rc.elem == 0) {
    switch (f) {
      case 0: x = 0; y = t; break;
      case 1: x = 1; y = t; break;
      case 2: x = t; y = 0; break;
      default: x = t; y = 1; break;
    }
  } else {
    switch (f) {
      case 0: x = t; y = 0; break;
      case 1: x = 1 - t; y = t; break;
      default: x = 0; y = 1 - t; break;
    }
  }
}

static void reinit_face(const RefCell &rc, const double *xv, int f, CellValues &cv) {
  cv.n = rc.ndofs; cv.nq = rc.nqf;
  // end points of the face and a vertex off the face (to orient the normal outwards)
  int a, b, o;
  if (rc.elem == 0) {
    const int fv[4][3] = {{0, 2, 1}, {1, 3, 0}, {0, 1, 2}, {2, 3, 0}};
    a = fv[f][0]; b = fv[f][1]; o = fv[f][2];
  } else {
    a = f; b = (f + 1) % 3; o = (f + 2) % 3;
  }
  const double tx = xv[2 * b] - xv[2 * a], ty = xv[2 * b + 1] - xv[2 * a + 1];
  const double len = std::sqrt(tx * tx + ty * ty);
  double nx = ty / len, ny = -tx / len;
  if (nx * (xv[2 * o] - xv[2 * a]) + ny * (xv[2 * o + 1] - xv[2 * a + 1]) > 0) { nx = -nx; ny = -ny; }
  for (int q = 0; q < rc.nqf; ++q) {
    double x, y, J[2][2];
    face_ref_point(rc, f, rc.qfp[q], x, y);
    fill_point(rc, xv, x, y, q, cv, J);
    cv.JxW[q] = rc.qfw[q] * len;
    cv.normal[q][0] = nx; cv.normal[q][1] = ny;
  }
}

static inline double sp2(const double a[2][2], const double b[2][2]) {
  return a[0][0] * b[0][0] + a[0][1] * b[0][1] + a[1][0] * b[1][0] + a[1][1] * b[1][1];
}
static inline double sp1(const double a[2], const double b[2]) { return a[0] * b[0] + a[1] * b[1]; }

// ---------------------------------------------------------------------------------------------
// assembly (NSSolverStationary.cpp:317-537, NSSolver.cpp:313-562)
// ---------------------------------------------------------------------------------------------
enum Mode { STOKES = 0, NEWTON = 1, UNSTEADY_FIRST = 2, UNSTEADY_NEWTON = 3 };

struct CellOut {
  double cm[41][41], cpm[41][41], rhs[41];
};

static void cell_assemble(const Problem &P, int cell, Mode mode, double nu, double delta_t, CellOut &out) {
  const RefCell &rc = P.rc;
  const int n = rc.ndofs, nq = rc.nq;
  const double *xv = &P.cell_vertices[(size_t)cell * rc.nvpc * 2];
  const uint32_t *dofs = &P.cell_dofs[(size_t)cell * n];
  static thread_local CellValues fe_values;
  reinit_cell(rc, xv, fe_values);
  for (int i = 0; i < n; ++i) {
    out.rhs[i] = 0;
    for (int j = 0; j < n; ++j) out.cm[i][j] = out.cpm[i][j] = 0;
  }
  // get_function_values / get_function_gradients of the ghosted solution
  double velocity_loc[16][2], velocity_old_loc[16][2], velocity_gradient_loc[16][2][2], pressure_loc[16];
  for (int q = 0; q < nq; ++q) {
    velocity_loc[q][0] = velocity_loc[q][1] = 0; velocity_old_loc[q][0] = velocity_old_loc[q][1] = 0;
    for (int k = 0; k < 2; ++k) for (int l = 0; l < 2; ++l) velocity_gradient_loc[q][k][l] = 0;
    pressure_loc[q] = 0;
    for (int i = 0; i < n; ++i) {
      const double s = P.solution[dofs[i]];
      for (int k = 0; k < 2; ++k) {
        velocity_loc[q][k] += s * fe_values.val[i][q][k];
        for (int l = 0; l < 2; ++l) velocity_gradient_loc[q][k][l] += s * fe_values.grad[i][q][k][l];
      }
      pressure_loc[q] += s * fe_values.pv[i][q];
      if (mode >= UNSTEADY_FIRST) {
        const double so = P.solution_old[dofs[i]];
        for (int k = 0; k < 2; ++k) velocity_old_loc[q][k] += so * fe_values.val[i][q][k];
      }
    }
  }
  double nonlinear_term[2];
  for (int q = 0; q < nq; ++q) {
    const double JxW = fe_values.JxW[q];
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        if (mode == STOKES) {  // S.cpp:383-406
          out.cm[i][j] += nu * sp2(fe_values.grad[i][q], fe_values.grad[j][q]) * JxW;
          out.cm[i][j] -= fe_values.div[i][q] * fe_values.pv[j][q] * JxW;
          out.cm[i][j] -= fe_values.div[j][q] * fe_values.pv[i][q] * JxW;
          out.cpm[i][j] += fe_values.pv[i][q] * fe_values.pv[j][q] / nu * JxW;
        } else if (mode == UNSTEADY_FIRST) {  // U.cpp:381-409
          out.cm[i][j] += nu * sp2(fe_values.grad[i][q], fe_values.grad[j][q]) * JxW;
          out.cm[i][j] -= fe_values.div[i][q] * fe_values.pv[j][q] * JxW;
          const double du[2] = {velocity_loc[q][0] - velocity_old_loc[q][0], velocity_loc[q][1] - velocity_old_loc[q][1]};
          out.cm[i][j] += sp1(du, fe_values.val[i][q]) / delta_t * JxW;
          out.cm[i][j] -= fe_values.div[j][q] * fe_values.pv[i][q] * JxW;
          out.cpm[i][j] += fe_values.pv[i][q] * fe_values.pv[j][q] / nu * JxW;
        } else {  // S.cpp:408-452, U.cpp:411-469
          for (int k = 0; k < 2; ++k) {
            nonlinear_term[k] = 0.0;
            for (int l = 0; l < 2; ++l) {
              nonlinear_term[k] += velocity_loc[q][l] * fe_values.grad[j][q][k][l];
              nonlinear_term[k] += fe_values.val[j][q][l] * velocity_gradient_loc[q][k][l];
            }
          }
          out.cm[i][j] += sp1(nonlinear_term, fe_values.val[i][q]) * JxW;
          if (mode == UNSTEADY_NEWTON)
            out.cm[i][j] += sp1(fe_values.val[j][q], fe_values.val[i][q]) / delta_t * JxW;
          out.cm[i][j] += nu * sp2(fe_values.grad[j][q], fe_values.grad[i][q]) * JxW;
          out.cm[i][j] -= fe_values.pv[j][q] * fe_values.div[i][q] * JxW;
          out.cm[i][j] += fe_values.pv[i][q] * fe_values.div[j][q] * JxW;
          out.cpm[i][j] += fe_values.pv[i][q] * fe_values.pv[j][q] / nu * JxW;
        }
      }
      if (mode == STOKES || mode == UNSTEADY_FIRST) continue;
      // -R(u,v): S.cpp:460-493, U.cpp:477-519
      if (mode == UNSTEADY_NEWTON) {
        const double du[2] = {velocity_loc[q][0] - velocity_old_loc[q][0], velocity_loc[q][1] - velocity_old_loc[q][1]};
        out.rhs[i] -= sp1(du, fe_values.val[i][q]) / delta_t * JxW;
      }
      out.rhs[i] -= nu * sp2(velocity_gradient_loc[q], fe_values.grad[i][q]) * JxW;
      for (int k = 0; k < 2; ++k) {
        nonlinear_term[k] = 0.0;
        for (int l = 0; l < 2; ++l) nonlinear_term[k] += velocity_loc[q][l] * velocity_gradient_loc[q][k][l];
      }
      out.rhs[i] -= sp1(nonlinear_term, fe_values.val[i][q]) * JxW;
      out.rhs[i] += pressure_loc[q] * fe_values.div[i][q] * JxW;
      const double velocity_divergence_loc = velocity_gradient_loc[q][0][0] + velocity_gradient_loc[q][1][1];
      out.rhs[i] += velocity_divergence_loc * fe_values.pv[i][q] * JxW;
    }
  }
}

static void scatter_row(CSR &A, int64_t row, int32_t col, double v) {
  const int32_t *b = &A.col[A.rowptr[row]], *e = &A.col[A.rowptr[row + 1]];
  const int32_t *it = std::lower_bound(b, e, col);
  if (it == e || *it != col) throw std::logic_error("oracle: entry outside the sparsity pattern");
  A.val[A.rowptr[row] + (it - b)] += v;
}

static void assemble(Problem &P, Mode mode, bool apply_inlet, double nu, double delta_t, double p_out);
static void apply_boundary_values(Problem &P, bool apply_inlet);

static void assemble_cells(Problem &P, Mode mode, double nu, double delta_t, double p_out) {
  const RefCell &rc = P.rc;
  const int n = rc.ndofs;
  const int64_t nu_ = P.n_u;
  std::fill(P.F.val.begin(), P.F.val.end(), 0.0);
  std::fill(P.Bt.val.begin(), P.Bt.val.end(), 0.0);
  std::fill(P.B.val.begin(), P.B.val.end(), 0.0);
  std::fill(P.Mp.val.begin(), P.Mp.val.end(), 0.0);
  std::fill(P.residual.begin(), P.residual.end(), 0.0);
  // outlet faces per cell (S.cpp:503-526)
  std::multimap<int, int> outlet;
  for (size_t k = 0; k < P.outlet_cell.size(); ++k) outlet.emplace(P.outlet_cell[k], P.outlet_face[k]);

  const int CH = 256;
  std::vector<CellOut> buf(CH);
  for (int c0 = 0; c0 < P.ncells; c0 += CH) {
    const int c1 = std::min(P.ncells, c0 + CH);
#pragma omp parallel for schedule(dynamic, 4)
    for (int c = c0; c < c1; ++c) cell_assemble(P, c, mode, nu, delta_t, buf[c - c0]);
    for (int c = c0; c < c1; ++c) {
      CellOut &o = buf[c - c0];
      auto rng = outlet.equal_range(c);
      for (auto it = rng.first; it != rng.second; ++it) {
        CellValues fv;
        reinit_face(rc, &P.cell_vertices[(size_t)c * rc.nvpc * 2], it->second, fv);
        for (int q = 0; q < rc.nqf; ++q)
          for (int i = 0; i < n; ++i)
            o.rhs[i] -= p_out * sp1(fv.normal[q], fv.val[i][q]) * fv.JxW[q];
      }
      const uint32_t *dofs = &P.cell_dofs[(size_t)c * n];
      for (int i = 0; i < n; ++i) {
        const bool ip = rc.comp[i] == 2;
        for (int j = 0; j < n; ++j) {
          const bool jp = rc.comp[j] == 2;
          const double v = o.cm[i][j];
          if (v != 0.0) {  // elide_zero_values
            if (!ip && !jp) scatter_row(P.F, dofs[i], (int32_t)dofs[j], v);
            else if (!ip && jp) scatter_row(P.Bt, dofs[i], (int32_t)(dofs[j] - nu_), v);
            else if (ip && !jp) scatter_row(P.B, dofs[i] - nu_, (int32_t)dofs[j], v);
            else throw std::logic_error("oracle: non-zero p-p entry in the Jacobian");
          }
          const double w = o.cpm[i][j];
          if (w != 0.0) scatter_row(P.Mp, dofs[i] - nu_, (int32_t)(dofs[j] - nu_), w);
        }
        P.residual[dofs[i]] += o.rhs[i];
      }
    }
  }
}

// MatrixTools::apply_boundary_values(bv, jacobian_matrix, delta_owned, residual_vector, false)
// for Trilinos block matrices (S.cpp:574-575): per diagonal block, constrained rows are cleared
// but a non-zero diagonal is preserved (else it is set to the first non-zero diagonal entry of
// the local range); solution[i] = v; rhs[i] = v * A_ii; the off-diagonal block rows are cleared.
// Only velocity dofs are constrained, so block (1,1) is untouched.
static void apply_boundary_values(Problem &P, bool apply_inlet) {
  CSR &F = P.F;
  for (int r = 0; r < P.nranks; ++r) {
    const int64_t lo = P.owned_u[r], hi = P.owned_u[r + 1];
    double first_nonzero_diag = 1.0;
    for (int64_t i = lo; i < hi; ++i) {
      double dgl = 0;
      for (int64_t k = F.rowptr[i]; k < F.rowptr[i + 1]; ++k) if (F.col[k] == i) dgl = F.val[k];
      if (dgl != 0) { first_nonzero_diag = std::fabs(dgl); break; }
    }
    for (size_t b = 0; b < P.bc_dof.size(); ++b) {
      const int64_t i = P.bc_dof[b];
      if (i < lo || i >= hi) continue;
      const double v = apply_inlet ? P.bc_val[b] : 0.0;
      double dgl = 0;
      for (int64_t k = F.rowptr[i]; k < F.rowptr[i + 1]; ++k) {
        if (F.col[k] == i) {
          if (F.val[k] == 0.0) F.val[k] = first_nonzero_diag;
          dgl = F.val[k];
        } else
          F.val[k] = 0.0;
      }
      for (int64_t k = P.Bt.rowptr[i]; k < P.Bt.rowptr[i + 1]; ++k) P.Bt.val[k] = 0.0;
      P.delta[i] = v;
      P.residual[i] = v * dgl;
    }
  }
}

static void assemble(Problem &P, Mode mode, bool apply_inlet, double nu, double delta_t, double p_out) {
  assemble_cells(P, mode, nu, delta_t, p_out);
  apply_boundary_values(P, apply_inlet);
}

// ---------------------------------------------------------------------------------------------
// vectors and sparse kernels
// ---------------------------------------------------------------------------------------------
static double dot(const Vec &a, const Vec &b) {
  const int64_t n = (int64_t)a.size(), BS = 4096, nb = (n + BS - 1) / BS;
  std::vector<double> part(nb);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nb; ++k) {
    double s = 0;
    const int64_t e = std::min(n, (k + 1) * BS);
    for (int64_t i = k * BS; i < e; ++i) s += a[i] * b[i];
    part[k] = s;
  }
  double s = 0;
  for (double p : part) s += p;
  return s;
}
static double norm2(const Vec &a) { return std::sqrt(dot(a, a)); }
static void axpy(Vec &y, double a, const Vec &x) {
  const int64_t n = (int64_t)y.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
// y = s*y + a*x
static void sadd(Vec &y, double s, double a, const Vec &x) {
  const int64_t n = (int64_t)y.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = s * y[i] + a * x[i];
}
static void equ(Vec &y, double a, const Vec &x) {
  const int64_t n = (int64_t)x.size();
  y.resize(n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = a * x[i];
}
// w += a*x ; return w . v     (deal.II Vector::add_and_dot)
static double add_and_dot(Vec &w, double a, const Vec &x, const Vec &v) {
  axpy(w, a, x);
  return dot(w, v);
}
static bool all_zero(const Vec &x) {
  for (double v : x) if (v != 0) return false;
  return true;
}

static void spmv(const CSR &A, const double *x, double *y, bool add = false) {
#pragma omp parallel for schedule(static, 256)
  for (int64_t i = 0; i < A.nrows; ++i) {
    double s = 0;
    for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) s += A.val[k] * x[A.col[k]];
    y[i] = add ? y[i] + s : s;
  }
}
// jacobian_matrix.vmult: y_u = F x_u + Bt x_p ; y_p = B x_u
static void block_spmv(const Problem &P, const Vec &x, Vec &y) {
  y.resize(x.size());
  spmv(P.F, x.data(), y.data());
  spmv(P.Bt, x.data() + P.n_u, y.data(), true);
  spmv(P.B, x.data(), y.data() + P.n_u);
}

// ---------------------------------------------------------------------------------------------
// SolverControl and the deal.II Krylov solvers (9.3-9.5 behaviour, default AdditionalData)
// ---------------------------------------------------------------------------------------------
struct NoConvergence : std::runtime_error {
  int last_step; double last_residual;
  NoConvergence(int s, double r) : std::runtime_error("SolverControl::NoConvergence"), last_step(s), last_residual(r) {}
};
enum State { ITERATE, SUCCESS, FAILURE };
struct Control {
  int max_steps; double tol; int last_step = 0; double last_value = 0;
  Control(int m, double t) : max_steps(m), tol(t) {}
  State check(int step, double value) {
    last_step = step; last_value = value;
    if (value <= tol) return SUCCESS;
    if (step >= max_steps || std::isnan(value)) return FAILURE;
    return ITERATE;
  }
};

// SolverCG::solve
static void solver_cg(Control &ctl, const Op &A, Vec &x, const Vec &b, const Op &M) {
  const size_t n = b.size();
  Vec g(n), d(n), h(n);
  int it = 0;
  if (!all_zero(x)) { A(g, x); axpy(g, -1.0, b); } else equ(g, -1.0, b);
  double res = norm2(g);
  State conv = ctl.check(0, res);
  if (conv != ITERATE) { if (conv != SUCCESS) throw NoConvergence(it, res); return; }
  M(h, g);
  equ(d, -1.0, h);
  double gh = dot(g, h);
  while (conv == ITERATE) {
    it++;
    A(h, d);
    double alpha = dot(d, h);
    alpha = gh / alpha;
    axpy(x, alpha, d);
    res = std::sqrt(std::fabs(add_and_dot(g, alpha, h, g)));
    conv = ctl.check(it, res);
    if (conv != ITERATE) break;
    M(h, g);
    double beta = gh;
    gh = dot(g, h);
    beta = gh / beta;
    sadd(d, beta, -1.0, h);
  }
  if (conv != SUCCESS) throw NoConvergence(it, res);
}

// minimise || rhs - H y || for the (m+1) x m upper Hessenberg H (column-major Hm[j][i]) by
// Householder QR; returns the residual norm (deal.II Householder::least_squares)
static double hessenberg_least_squares(const std::vector<std::vector<double>> &H, int m, double beta, std::vector<double> &y) {
  const int rows = m + 1;
  std::vector<double> A((size_t)rows * m), b(rows, 0.0);
  for (int j = 0; j < m; ++j) for (int i = 0; i < rows; ++i) A[(size_t)i * m + j] = H[j][i];
  b[0] = beta;
  for (int j = 0; j < m; ++j) {
    double sigma = 0;
    for (int i = j; i < rows; ++i) sigma += A[(size_t)i * m + j] * A[(size_t)i * m + j];
    if (sigma == 0) continue;
    double s = std::sqrt(sigma);
    if (A[(size_t)j * m + j] > 0) s = -s;
    std::vector<double> v(rows, 0.0);
    for (int i = j; i < rows; ++i) v[i] = A[(size_t)i * m + j];
    v[j] -= s;
    double vv = 0;
    for (int i = j; i < rows; ++i) vv += v[i] * v[i];
    if (vv == 0) continue;
    for (int c = j; c < m; ++c) {
      double t = 0;
      for (int i = j; i < rows; ++i) t += v[i] * A[(size_t)i * m + c];
      t = 2 * t / vv;
      for (int i = j; i < rows; ++i) A[(size_t)i * m + c] -= t * v[i];
    }
    double t = 0;
    for (int i = j; i < rows; ++i) t += v[i] * b[i];
    t = 2 * t / vv;
    for (int i = j; i < rows; ++i) b[i] -= t * v[i];
  }
  y.assign(m, 0.0);
  for (int i = m - 1; i >= 0; --i) {
    double s = b[i];
    for (int c = i + 1; c < m; ++c) s -= A[(size_t)i * m + c] * y[c];
    y[i] = s / A[(size_t)i * m + i];
  }
  return std::fabs(b[m]);
}

// SolverFGMRES::solve (max_basis_size = 30)
static void solver_fgmres(Control &ctl, const Op &A, Vec &x, const Vec &b, const Op &M) {
  const int basis_size = 30;
  const size_t n = b.size();
  std::vector<Vec> v(basis_size), z(basis_size);
  std::vector<std::vector<double>> H(basis_size, std::vector<double>(basis_size + 1, 0.0));
  std::vector<double> y;
  int accumulated_iterations = 0;
  double res = -std::numeric_limits<double>::max();
  Vec aux(n);
  State state = ITERATE;
  do {
    A(aux, x);
    sadd(aux, -1.0, 1.0, b);
    const double beta = norm2(aux);
    res = beta;
    state = ctl.check(accumulated_iterations, res);
    if (state == SUCCESS) break;
    for (auto &c : H) std::fill(c.begin(), c.end(), 0.0);
    double a = beta;
    y.clear();
    for (int j = 0; j < basis_size; ++j) {
      if (std::isfinite(a)) equ(v[j], 1.0 / a, aux);
      else v[j].assign(n, 0.0);
      if (z[j].empty()) z[j].assign(n, 0.0);  // TmpVectors: zero on first use, kept across restarts
      M(z[j], v[j]);
      A(aux, z[j]);
      H[j][0] = dot(aux, v[0]);
      for (int i = 1; i <= j; ++i) H[j][i] = add_and_dot(aux, -H[j][i - 1], v[i - 1], v[i]);
      H[j][j + 1] = a = std::sqrt(add_and_dot(aux, -H[j][j], v[j], aux));
      if (j > 0) {
        res = hessenberg_least_squares(H, j, beta, y);
        state = ctl.check(++accumulated_iterations, res);
        if (state != ITERATE) break;
      }
    }
    for (size_t j = 0; j < y.size(); ++j) axpy(x, y[j], z[j]);
  } while (state == ITERATE);
  if (state != SUCCESS) throw NoConvergence(accumulated_iterations, res);
}

// SolverGMRES::solve (max_n_tmp_vectors = 30, left preconditioning, default residual)
static void solver_gmres(Control &ctl, const Op &A, Vec &x, const Vec &b, const Op &M) {
  const int n_tmp = 30;
  const size_t n = b.size();
  std::vector<Vec> tmp(n_tmp);
  std::vector<std::vector<double>> H(n_tmp - 1, std::vector<double>(n_tmp, 0.0));  // H[col][row]
  std::vector<double> gamma(n_tmp), ci(n_tmp - 1), si(n_tmp - 1), h(n_tmp - 1);
  int accumulated_iterations = 0, dim = 0;
  State state = ITERATE;
  double last_res = -std::numeric_limits<double>::max();
  Vec &vfirst = tmp[0];
  Vec &p = tmp[n_tmp - 1];
  bool re_orthogonalize = false;
  do {
    std::fill(h.begin(), h.end(), 0.0);
    p.resize(n);
    if (vfirst.empty()) vfirst.assign(n, 0.0);  // TmpVectors: zero on first use, kept across restarts
    A(p, x);
    sadd(p, -1.0, 1.0, b);
    M(vfirst, p);
    double rho = norm2(vfirst);
    last_res = rho;
    state = ctl.check(accumulated_iterations, rho);
    if (state != ITERATE) break;
    gamma[0] = rho;
    equ(vfirst, 1.0 / rho, Vec(vfirst));
    for (int inner = 0; inner < n_tmp - 2 && state == ITERATE; ++inner) {
      ++accumulated_iterations;
      Vec &vv = tmp[inner + 1];
      if (vv.empty()) vv.assign(n, 0.0);
      A(p, tmp[inner]);
      M(vv, p);
      dim = inner + 1;
      // modified Gram-Schmidt with Kelley's re-orthogonalisation test every 5th iteration
      double norm_vv_start = 0;
      const bool consider = (re_orthogonalize == false) && (accumulated_iterations % 5 == 0);
      if (consider) norm_vv_start = norm2(vv);
      h[0] = dot(vv, tmp[0]);
      for (int i = 1; i < dim; ++i) h[i] = add_and_dot(vv, -h[i - 1], tmp[i - 1], tmp[i]);
      double norm_vv = std::sqrt(add_and_dot(vv, -h[dim - 1], tmp[dim - 1], vv));
      bool done = false;
      if (consider) {
        if (norm_vv > 10. * norm_vv_start * std::sqrt(std::numeric_limits<double>::epsilon())) done = true;
        else re_orthogonalize = true;
      }
      if (!done && re_orthogonalize) {
        double htmp = dot(vv, tmp[0]);
        h[0] += htmp;
        for (int i = 1; i < dim; ++i) {
          htmp = add_and_dot(vv, -htmp, tmp[i - 1], tmp[i]);
          h[i] += htmp;
        }
        norm_vv = std::sqrt(add_and_dot(vv, -htmp, tmp[dim - 1], vv));
      }
      const double s = norm_vv;
      h[inner + 1] = s;
      if (std::isfinite(1. / s)) equ(vv, 1. / s, Vec(vv));
      // Givens rotations
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
    // back substitution H1 y = gamma
    std::vector<double> yv(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int c = i + 1; c < dim; ++c) s -= H[c][i] * yv[c];
      yv[i] = s / H[i][i];
    }
    for (int i = 0; i < dim; ++i) axpy(x, yv[i], tmp[i]);
  } while (state == ITERATE);
  if (state != SUCCESS) throw NoConvergence(accumulated_iterations, last_res);
}

// SolverBicgstab::solve (exact_residual = true).  Breakdown threshold: deal.II <= 9.3 used 1e-10,
// which restarts for ever once |r.rbar| < 1e-10 (any run with ||r|| < 1e-5 at a restart); later
// releases use numeric_limits<double>::min().  The reference does not pin a version; the later
// default is restated here, and a restart that can no longer make progress fails instead of
// spinning.
static void solver_bicgstab(Control &ctl, const Op &A, Vec &x, const Vec &b, const Op &M) {
  const size_t n = b.size();
  const double breakdown_tol = std::numeric_limits<double>::min();
  Vec r(n), rbar(n), p(n), y(n), z(n), t(n), v(n);
  int step = 0;
  double res = 0;
  State state = ITERATE;
  bool breakdown;
  do {
    breakdown = false;
    // start()
    A(r, x);
    sadd(r, -1.0, 1.0, b);
    res = norm2(r);
    {
      const State st = ctl.check(step, res);
      if (st == SUCCESS) { state = SUCCESS; break; }
      if (st == FAILURE) { state = FAILURE; break; }
    }
    // iterate()
    state = ITERATE;
    double alpha = 1, omega = 1, rho = 1, rhobar, beta;
    rbar = r;
    bool startup = true;
    do {
      ++step;
      rhobar = dot(r, rbar);
      if (std::fabs(rhobar) < breakdown_tol) { breakdown = true; break; }
      beta = rhobar * alpha / (rho * omega);
      rho = rhobar;
      if (startup) { p = r; startup = false; }
      else { sadd(p, beta, 1.0, r); axpy(p, -beta * omega, v); }
      M(y, p);  // y, z keep their previous content (initial guess of inner solves)
      A(v, y);
      rhobar = dot(rbar, v);
      if (std::fabs(rhobar) < breakdown_tol) { breakdown = true; break; }
      alpha = rho / rhobar;
      res = std::sqrt(add_and_dot(r, -alpha, v, r));
      if (ctl.check(step, res) == SUCCESS) { axpy(x, alpha, y); state = SUCCESS; break; }
      M(z, r);
      A(t, z);
      rhobar = dot(t, r);
      const double t_squared = dot(t, t);
      if (t_squared < breakdown_tol) { breakdown = true; break; }
      omega = rhobar / dot(t, t);
      axpy(x, alpha, y);
      axpy(x, omega, z);
      axpy(r, -omega, t);
      A(t, x);
      axpy(t, -1.0, b);
      res = norm2(t);
      state = ctl.check(step, res);
    } while (state == ITERATE);
  } while (breakdown);
  if (state != SUCCESS) throw NoConvergence(step, res);
}

// ---------------------------------------------------------------------------------------------
// rank-local inner preconditioners (Ifpack, overlap 0)
// ---------------------------------------------------------------------------------------------
struct LocalBlocks {
  std::vector<int64_t> off;  // nranks+1 row offsets; couplings across blocks are dropped
  // Optional general form (the sub-blocks and elimination orders the product's sweeps may use): block b eliminates the
  // rows ord[ord_off[b] .. ord_off[b+1]) in that sequence; blk[i] = block of row i; couplings between different blocks
  // are dropped.  Empty = the contiguous ranges `off` in natural order (Ifpack's own behaviour).
  const std::vector<int32_t> *ord = nullptr, *blk = nullptr;
  const std::vector<int64_t> *ord_off = nullptr;
  bool general() const { return ord && !ord->empty(); }
};

// Ifpack point relaxation, symmetric Gauss-Seidel, 1 sweep, omega 1, zero starting solution
static void sgs_apply(const CSR &A, const LocalBlocks &lb, Vec &y, const Vec &x) {
  y.assign(x.size(), 0.0);
  if (lb.general()) {
    const std::vector<int32_t> &ord = *lb.ord, &blk = *lb.blk;
    const int nb = (int)lb.ord_off->size() - 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nb; ++b) {
      const int64_t t0 = (*lb.ord_off)[b], t1 = (*lb.ord_off)[b + 1];
      for (int pass = 0; pass < 2; ++pass)
        for (int64_t t = pass == 0 ? t0 : t1 - 1; pass == 0 ? t < t1 : t >= t0; pass == 0 ? ++t : --t) {
          const int64_t i = ord[t];
          double dtemp = 0, dgl = 0;
          for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            const int64_t c = A.col[k];
            if (c >= A.nrows || blk[c] != b) continue;
            if (c == i) dgl = A.val[k];
            dtemp += A.val[k] * y[c];
          }
          y[i] += (x[i] - dtemp) / dgl;
        }
    }
    return;
  }
  const int nb = (int)lb.off.size() - 1;
#pragma omp parallel for schedule(static, 1)
  for (int b = 0; b < nb; ++b) {
    const int64_t lo = lb.off[b], hi = lb.off[b + 1];
    for (int64_t i = lo; i < hi; ++i) {
      double dtemp = 0, dgl = 0;
      for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
        const int64_t c = A.col[k];
        if (c < lo || c >= hi) continue;
        if (c == i) dgl = A.val[k];
        dtemp += A.val[k] * y[c];
      }
      y[i] += (x[i] - dtemp) / dgl;
    }
    for (int64_t i = hi - 1; i >= lo; --i) {
      double dtemp = 0, dgl = 0;
      for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
        const int64_t c = A.col[k];
        if (c < lo || c >= hi) continue;
        if (c == i) dgl = A.val[k];
        dtemp += A.val[k] * y[c];
      }
      y[i] += (x[i] - dtemp) / dgl;
    }
  }
}

// Ifpack ILU, level-of-fill 0, absolute threshold 0, relative threshold 1
struct ILU0 {
  const CSR *A = nullptr;
  LocalBlocks lb;
  std::vector<double> lu;       // same pattern as A; entries coupling different blocks unused
  std::vector<int64_t> diag;    // position of the diagonal in each row
  // general blocks / elimination orders: position of every row in its block's sequence, and per row its in-block
  // entries sorted by that position (the order in which IKJ elimination visits them)
  std::vector<int32_t> rank;
  std::vector<int64_t> srt_ptr, srt;
  void compute_general() {
    const CSR &A_ = *A;
    const int64_t n = A_.nrows;
    const std::vector<int32_t> &ord = *lb.ord, &blk = *lb.blk;
    rank.assign(n, 0);
    for (int64_t t = 0; t < n; ++t) rank[ord[t]] = (int32_t)t;
    srt_ptr.assign(n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
      int64_t cnt = 0;
      for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k) cnt += A_.col[k] < n && blk[A_.col[k]] == blk[i];
      srt_ptr[i + 1] = srt_ptr[i] + cnt;
    }
    srt.resize(srt_ptr[n]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      int64_t o = srt_ptr[i];
      for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k)
        if (A_.col[k] < n && blk[A_.col[k]] == blk[i]) srt[o++] = k;
      std::sort(srt.begin() + srt_ptr[i], srt.begin() + srt_ptr[i + 1], [&](int64_t a, int64_t b) { return rank[A_.col[a]] < rank[A_.col[b]]; });
    }
    const int nb = (int)lb.ord_off->size() - 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nb; ++b) {
      for (int64_t t = (*lb.ord_off)[b]; t < (*lb.ord_off)[b + 1]; ++t) {
        const int64_t i = ord[t];
        for (int64_t q = srt_ptr[i]; q < srt_ptr[i + 1]; ++q) {
          const int64_t k = srt[q];
          const int64_t kk = A_.col[k];
          if (rank[kk] >= rank[i]) break;
          const double lik = lu[k] / lu[diag[kk]];
          lu[k] = lik;
          // a_ij -= l_ik u_kj for the j behind k in row k that are also in row i (both lists ascend in rank)
          int64_t qi = q + 1;
          for (int64_t qk = srt_ptr[kk]; qk < srt_ptr[kk + 1]; ++qk) {
            const int64_t m = srt[qk];
            const int32_t rj = rank[A_.col[m]];
            if (rj <= rank[kk]) continue;
            while (qi < srt_ptr[i + 1] && rank[A_.col[srt[qi]]] < rj) ++qi;
            if (qi < srt_ptr[i + 1] && A_.col[srt[qi]] == A_.col[m]) lu[srt[qi]] -= lik * lu[m];
          }
        }
      }
    }
  }
  void compute(const CSR &A_, const LocalBlocks &lb_) {
    A = &A_; lb = lb_;
    lu = A_.val;
    const int64_t n = A_.nrows;
    diag.assign(n, -1);
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k) if (A_.col[k] == i) diag[i] = k;
    if (lb.general()) { compute_general(); return; }
    const int nb = (int)lb.off.size() - 1;
#pragma omp parallel for schedule(static, 1)
    for (int b = 0; b < nb; ++b) {
      const int64_t lo = lb.off[b], hi = lb.off[b + 1];
      std::vector<int64_t> pos(hi - lo, -1);
      for (int64_t i = lo; i < hi; ++i) {
        for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k) {
          const int64_t c = A_.col[k];
          if (c >= lo && c < hi) pos[c - lo] = k;
        }
        for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k) {
          const int64_t kk = A_.col[k];
          if (kk < lo || kk >= hi) continue;
          if (kk >= i) break;
          const double lik = lu[k] / lu[diag[kk]];
          lu[k] = lik;
          for (int64_t m = diag[kk] + 1; m < A_.rowptr[kk + 1]; ++m) {
            const int64_t c = A_.col[m];
            if (c >= hi) break;
            const int64_t p = pos[c - lo];
            if (p >= 0) lu[p] -= lik * lu[m];
          }
        }
        for (int64_t k = A_.rowptr[i]; k < A_.rowptr[i + 1]; ++k) {
          const int64_t c = A_.col[k];
          if (c >= lo && c < hi) pos[c - lo] = -1;
        }
      }
    }
  }
  void apply(Vec &y, const Vec &x) const {
    y.resize(x.size());
    if (lb.general()) {
      const std::vector<int32_t> &ord = *lb.ord;
      const int nb = (int)lb.ord_off->size() - 1;
#pragma omp parallel for schedule(dynamic, 1)
      for (int b = 0; b < nb; ++b) {
        const int64_t t0 = (*lb.ord_off)[b], t1 = (*lb.ord_off)[b + 1];
        for (int64_t t = t0; t < t1; ++t) {
          const int64_t i = ord[t];
          double s = x[i];
          for (int64_t q = srt_ptr[i]; q < srt_ptr[i + 1]; ++q) {
            const int64_t k = srt[q];
            if (rank[A->col[k]] >= rank[i]) break;
            s -= lu[k] * y[A->col[k]];
          }
          y[i] = s;
        }
        for (int64_t t = t1 - 1; t >= t0; --t) {
          const int64_t i = ord[t];
          double s = y[i];
          for (int64_t q = srt_ptr[i]; q < srt_ptr[i + 1]; ++q) {
            const int64_t k = srt[q];
            if (rank[A->col[k]] > rank[i]) s -= lu[k] * y[A->col[k]];
          }
          y[i] = s / lu[diag[i]];
        }
      }
      return;
    }
    const int nb = (int)lb.off.size() - 1;
#pragma omp parallel for schedule(static, 1)
    for (int b = 0; b < nb; ++b) {
      const int64_t lo = lb.off[b], hi = lb.off[b + 1];
      for (int64_t i = lo; i < hi; ++i) {
        double s = x[i];
        for (int64_t k = A->rowptr[i]; k < diag[i]; ++k) {
          const int64_t c = A->col[k];
          if (c >= lo) s -= lu[k] * y[c];
        }
        y[i] = s;
      }
      for (int64_t i = hi - 1; i >= lo; --i) {
        double s = y[i];
        for (int64_t k = diag[i] + 1; k < A->rowptr[i + 1]; ++k) {
          const int64_t c = A->col[k];
          if (c < hi) s -= lu[k] * y[c];
        }
        y[i] = s / lu[diag[i]];
      }
    }
  }
};

}  // namespace orc

#include "oracle_amg.inc"
#include "oracle_solve.inc"
