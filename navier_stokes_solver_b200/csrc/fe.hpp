// Reference-cell tables for the two Taylor-Hood pairs of the hot path.
//
//   elem 0: Q3/Q2 on quadrilaterals  (FESystem(FE_Q(3)^2, FE_Q(2)), QGauss<2>(4), QGauss<1>(4))
//           reference: lab_new/src/NSSolverStationary.cpp:118-138
//   elem 1: P2/P1 on triangles       (FESystem(FE_SimplexP(2)^2, FE_SimplexP(1)),
//           QGaussSimplex<2>(3), QGaussSimplex<1>(3))
//           reference: lab_new/src/NSSolverStationary.cpp:184-204
//
// The numbering conventions (cell-local dof order, node order, quadrature order) restate what
// deal.II does for these elements; deal.II itself is not available in
// this image, so the conventions are pinned by tests/test_fe_tables.py instead.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define NSX_HD __host__ __device__
#else
#define NSX_HD
#endif

namespace nsx {

constexpr int MAX_DOFS = 41;  // dofs per cell (Q3/Q2)
constexpr int MAX_VN = 16;    // scalar velocity nodes per cell
constexpr int MAX_PN = 9;     // scalar pressure nodes per cell
constexpr int MAX_Q = 16;     // cell quadrature points
constexpr int MAX_QF = 4;     // face quadrature points
constexpr int MAX_FACES = 4;

struct FETables {
  int elem;    // 0 = Q3/Q2 quads, 1 = P2/P1 triangles
  int nvpc;    // vertices per cell
  int nvn;     // scalar velocity nodes
  int npn;     // scalar pressure nodes
  int ndofs;   // dofs per cell
  int nq;      // cell quadrature points
  int nqf;     // face quadrature points
  int nfaces;  // faces per cell
  int dof_comp[MAX_DOFS];  // 0 = u_x, 1 = u_y, 2 = p
  int dof_node[MAX_DOFS];  // scalar node index inside the component's scalar element
  double qp[MAX_Q][2];     // reference quadrature points
  double qw[MAX_Q];        // reference weights
  double Nv[MAX_VN][MAX_Q];
  double dNv[MAX_VN][MAX_Q][2];  // reference-cell gradients
  double Np[MAX_PN][MAX_Q];
  double qwf[MAX_QF];
  double Nvf[MAX_FACES][MAX_VN][MAX_QF];
  double dNvf[MAX_FACES][MAX_VN][MAX_QF][2];
  double Npf[MAX_FACES][MAX_PN][MAX_QF];
  double qpf[MAX_QF];  // face quadrature parameter t in [0,1] along the face direction
};

namespace detail {

// Lagrange basis k on 1-D nodes x[0..n), value and derivative at t.
inline void lagrange1d(const double *x, int n, int k, double t, double &v, double &d) {
  v = 1.0;
  for (int m = 0; m < n; ++m)
    if (m != k) v *= (t - x[m]) / (x[k] - x[m]);
  d = 0.0;
  for (int j = 0; j < n; ++j) {
    if (j == k) continue;
    double term = 1.0 / (x[k] - x[j]);
    for (int m = 0; m < n; ++m)
      if (m != k && m != j) term *= (t - x[m]) / (x[k] - x[m]);
    d += term;
  }
}

// FE_Q(p) hierarchical node numbering on the unit square -> (ix, iy) 1-D node indices.
// Order: 4 vertices (lexicographic), line 0 (x=0), line 1 (x=1), line 2 (y=0), line 3 (y=1),
// each line's interior nodes along its direction, then the interior nodes x-fastest.
inline int feq_node_ij(int p, int node, int &ix, int &iy) {
  const int m = p - 1;  // interior nodes per line
  if (node < 4) { ix = (node & 1) ? p : 0; iy = (node & 2) ? p : 0; return 0; }
  node -= 4;
  if (node < 4 * m) {
    const int l = node / m, k = node % m + 1;
    if (l == 0) { ix = 0; iy = k; }
    else if (l == 1) { ix = p; iy = k; }
    else if (l == 2) { ix = k; iy = 0; }
    else { ix = k; iy = p; }
    return 0;
  }
  node -= 4 * m;
  ix = node % m + 1; iy = node / m + 1;
  return 0;
}

inline void feq_eval(int p, const double *nodes1d, int node, double x, double y, double &v,
                     double &dx, double &dy) {
  int ix, iy;
  feq_node_ij(p, node, ix, iy);
  double vx, dvx, vy, dvy;
  lagrange1d(nodes1d, p + 1, ix, x, vx, dvx);
  lagrange1d(nodes1d, p + 1, iy, y, vy, dvy);
  v = vx * vy; dx = dvx * vy; dy = vx * dvy;
}

// P2 on the reference triangle (0,0),(1,0),(0,1): 3 vertex functions then the midpoints of
// lines (v0v1), (v1v2), (v2v0).
inline void p2_eval(int node, double x, double y, double &v, double &dx, double &dy) {
  const double l[3] = {1.0 - x - y, x, y};
  const double gl[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
  if (node < 3) {
    v = l[node] * (2.0 * l[node] - 1.0);
    dx = (4.0 * l[node] - 1.0) * gl[node][0];
    dy = (4.0 * l[node] - 1.0) * gl[node][1];
  } else {
    const int a = node - 3, b = (node - 3 + 1) % 3;
    v = 4.0 * l[a] * l[b];
    dx = 4.0 * (gl[a][0] * l[b] + l[a] * gl[b][0]);
    dy = 4.0 * (gl[a][1] * l[b] + l[a] * gl[b][1]);
  }
}
inline void p1_eval(int node, double x, double y, double &v) {
  const double l[3] = {1.0 - x - y, x, y};
  v = l[node];
}

inline void gauss01(int n, double *x, double *w) {
  if (n == 3) {
    const double a = 0.5 * std::sqrt(3.0 / 5.0);
    x[0] = 0.5 - a; x[1] = 0.5; x[2] = 0.5 + a;
    w[0] = 5.0 / 18.0; w[1] = 4.0 / 9.0; w[2] = 5.0 / 18.0;
  } else {  // n == 4
    const double s = std::sqrt(6.0 / 5.0);
    const double a = 0.5 * std::sqrt(3.0 / 7.0 - 2.0 / 7.0 * s);
    const double b = 0.5 * std::sqrt(3.0 / 7.0 + 2.0 / 7.0 * s);
    const double wa = (18.0 + std::sqrt(30.0)) / 72.0, wb = (18.0 - std::sqrt(30.0)) / 72.0;
    x[0] = 0.5 - b; x[1] = 0.5 - a; x[2] = 0.5 + a; x[3] = 0.5 + b;
    w[0] = wb; w[1] = wa; w[2] = wa; w[3] = wb;
  }
}

}  // namespace detail

// Reference face quadrature point q (parameter t in [0,1]) of face f mapped into the cell.
NSX_HD inline void face_point(int elem, int f, double t, double &x, double &y) {
  if (elem == 0) {
    if (f == 0) { x = 0.0; y = t; }
    else if (f == 1) { x = 1.0; y = t; }
    else if (f == 2) { x = t; y = 0.0; }
    else { x = t; y = 1.0; }
  } else {
    if (f == 0) { x = t; y = 0.0; }
    else if (f == 1) { x = 1.0 - t; y = t; }
    else { x = 0.0; y = 1.0 - t; }
  }
}

inline void build_fe_tables(int elem, FETables &T) {
  std::memset(&T, 0, sizeof(T));
  T.elem = elem;
  if (elem == 0) {
    T.nvpc = 4; T.nvn = 16; T.npn = 9; T.ndofs = 41; T.nq = 16; T.nqf = 4; T.nfaces = 4;
    // cell-local system dofs: per vertex [ux uy p]; per line [ux ux uy uy p]; cell [ux*4 uy*4 p]
    int i = 0;
    for (int v = 0; v < 4; ++v) {
      T.dof_comp[i] = 0; T.dof_node[i++] = v;
      T.dof_comp[i] = 1; T.dof_node[i++] = v;
      T.dof_comp[i] = 2; T.dof_node[i++] = v;
    }
    for (int l = 0; l < 4; ++l) {
      for (int c = 0; c < 2; ++c)
        for (int k = 0; k < 2; ++k) { T.dof_comp[i] = c; T.dof_node[i++] = 4 + 2 * l + k; }
      T.dof_comp[i] = 2; T.dof_node[i++] = 4 + l;
    }
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < 4; ++k) { T.dof_comp[i] = c; T.dof_node[i++] = 12 + k; }
    T.dof_comp[i] = 2; T.dof_node[i++] = 8;

    const double s5 = std::sqrt(5.0);
    const double gl3[4] = {0.0, 0.5 * (1.0 - 1.0 / s5), 0.5 * (1.0 + 1.0 / s5), 1.0};
    const double eq2[3] = {0.0, 0.5, 1.0};
    double gx[4], gw[4];
    detail::gauss01(4, gx, gw);
    for (int qy = 0; qy < 4; ++qy)
      for (int qx = 0; qx < 4; ++qx) {
        const int q = qy * 4 + qx;
        T.qp[q][0] = gx[qx]; T.qp[q][1] = gx[qy]; T.qw[q] = gw[qx] * gw[qy];
      }
    for (int q = 0; q < T.nq; ++q) {
      for (int a = 0; a < 16; ++a)
        detail::feq_eval(3, gl3, a, T.qp[q][0], T.qp[q][1], T.Nv[a][q], T.dNv[a][q][0], T.dNv[a][q][1]);
      for (int m = 0; m < 9; ++m) {
        double dx, dy;
        detail::feq_eval(2, eq2, m, T.qp[q][0], T.qp[q][1], T.Np[m][q], dx, dy);
      }
    }
    for (int qf = 0; qf < 4; ++qf) { T.qwf[qf] = gw[qf]; T.qpf[qf] = gx[qf]; }
    for (int f = 0; f < 4; ++f)
      for (int qf = 0; qf < 4; ++qf) {
        double x, y;
        face_point(0, f, gx[qf], x, y);
        for (int a = 0; a < 16; ++a)
          detail::feq_eval(3, gl3, a, x, y, T.Nvf[f][a][qf], T.dNvf[f][a][qf][0], T.dNvf[f][a][qf][1]);
        for (int m = 0; m < 9; ++m) {
          double dx, dy;
          detail::feq_eval(2, eq2, m, x, y, T.Npf[f][m][qf], dx, dy);
        }
      }
  } else {
    T.nvpc = 3; T.nvn = 6; T.npn = 3; T.ndofs = 15; T.nq = 7; T.nqf = 3; T.nfaces = 3;
    int i = 0;
    for (int v = 0; v < 3; ++v) {
      T.dof_comp[i] = 0; T.dof_node[i++] = v;
      T.dof_comp[i] = 1; T.dof_node[i++] = v;
      T.dof_comp[i] = 2; T.dof_node[i++] = v;
    }
    for (int l = 0; l < 3; ++l) {
      T.dof_comp[i] = 0; T.dof_node[i++] = 3 + l;
      T.dof_comp[i] = 1; T.dof_node[i++] = 3 + l;
    }
    // Radon's 7-point degree-5 rule on the reference triangle (area 1/2).
    const double s15 = std::sqrt(15.0);
    const double a1 = (6.0 - s15) / 21.0, a2 = (6.0 + s15) / 21.0;
    const double w0 = 0.5 * 9.0 / 40.0, w1 = 0.5 * (155.0 - s15) / 1200.0, w2 = 0.5 * (155.0 + s15) / 1200.0;
    const double P[7][2] = {{1.0 / 3.0, 1.0 / 3.0},
                            {1.0 - 2.0 * a1, a1}, {a1, 1.0 - 2.0 * a1}, {a1, a1},
                            {1.0 - 2.0 * a2, a2}, {a2, 1.0 - 2.0 * a2}, {a2, a2}};
    const double W[7] = {w0, w1, w1, w1, w2, w2, w2};
    for (int q = 0; q < 7; ++q) { T.qp[q][0] = P[q][0]; T.qp[q][1] = P[q][1]; T.qw[q] = W[q]; }
    for (int q = 0; q < 7; ++q) {
      for (int a = 0; a < 6; ++a)
        detail::p2_eval(a, P[q][0], P[q][1], T.Nv[a][q], T.dNv[a][q][0], T.dNv[a][q][1]);
      for (int m = 0; m < 3; ++m) detail::p1_eval(m, P[q][0], P[q][1], T.Np[m][q]);
    }
    double gx[3], gw[3];
    detail::gauss01(3, gx, gw);
    for (int qf = 0; qf < 3; ++qf) { T.qwf[qf] = gw[qf]; T.qpf[qf] = gx[qf]; }
    for (int f = 0; f < 3; ++f)
      for (int qf = 0; qf < 3; ++qf) {
        double x, y;
        face_point(1, f, gx[qf], x, y);
        for (int a = 0; a < 6; ++a)
          detail::p2_eval(a, x, y, T.Nvf[f][a][qf], T.dNvf[f][a][qf][0], T.dNvf[f][a][qf][1]);
        for (int m = 0; m < 3; ++m) detail::p1_eval(m, x, y, T.Npf[f][m][qf]);
      }
  }
}

// Local vertex numbers of the two end points of face f, in the face's direction.
inline void face_vertices(int elem, int f, int &a, int &b) {
  if (elem == 0) {
    const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
    a = fv[f][0]; b = fv[f][1];
  } else {
    a = f; b = (f + 1) % 3;
  }
}

}  // namespace nsx
