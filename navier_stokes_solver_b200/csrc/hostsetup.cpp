#include "hostsetup.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace nsx {

namespace {

inline uint64_t edge_key(int a, int b) {
  const uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
  return ((uint64_t)lo << 32) | hi;
}

struct EdgeTable {
  // line id per (cell, local line); number of cells per line; first (cell,face) per line
  std::vector<int> cell_line;
  std::vector<int> line_ncells;
  int nlines = 0;
};

// Global line numbering by first appearance in the cell walk.
EdgeTable build_edges(const Mesh &m) {
  EdgeTable E;
  const int nc = m.ncells(), nf = m.nvpc;  // faces per cell == vertices per cell in 2-D
  E.cell_line.assign((size_t)nc * nf, -1);
  std::unordered_map<uint64_t, int> ids;
  ids.reserve((size_t)nc * 2 + 16);
  for (int c = 0; c < nc; ++c)
    for (int f = 0; f < nf; ++f) {
      int a, b;
      face_vertices(m.elem, f, a, b);
      const uint64_t k = edge_key(m.cells[(size_t)c * nf + a], m.cells[(size_t)c * nf + b]);
      auto it = ids.find(k);
      int id;
      if (it == ids.end()) {
        id = E.nlines++;
        ids.emplace(k, id);
        E.line_ncells.push_back(0);
      } else
        id = it->second;
      E.cell_line[(size_t)c * nf + f] = id;
      E.line_ncells[id]++;
    }
  return E;
}

}  // namespace

void generate_mesh(int nx, int ny, bool triangles, Mesh &m) {
  if (nx <= 0 || ny <= 0) throw std::invalid_argument("mesh size must be positive");
  const double x0 = 0.0, y0 = 0.0, x1 = 2.2, y1 = 0.41;
  const double dx = (x1 - x0) / nx, dy = (y1 - y0) / ny;
  const double ccx = x0 + 0.2, ccy = (y0 + y1) / 2.0, rad = 0.05;

  std::vector<double> fullv((size_t)(nx + 1) * (ny + 1) * 2);
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      fullv[2 * ((size_t)j * (nx + 1) + i)] = x0 + i * dx;
      fullv[2 * ((size_t)j * (nx + 1) + i) + 1] = y0 + j * dy;
    }
  std::vector<int> qcells, qmat;
  std::vector<char> used((size_t)(nx + 1) * (ny + 1), 0);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int v[4] = {j * (nx + 1) + i, j * (nx + 1) + i + 1, (j + 1) * (nx + 1) + i, (j + 1) * (nx + 1) + i + 1};
      double cx = 0, cy = 0;
      for (int k = 0; k < 4; ++k) { cx += fullv[2 * (size_t)v[k]]; cy += fullv[2 * (size_t)v[k] + 1]; }
      cx /= 4.0; cy /= 4.0;
      const double dist = std::sqrt((cx - ccx) * (cx - ccx) + (cy - ccy) * (cy - ccy));
      if (dist < rad) continue;
      const double d03 = std::hypot(fullv[2 * (size_t)v[3]] - fullv[2 * (size_t)v[0]], fullv[2 * (size_t)v[3] + 1] - fullv[2 * (size_t)v[0] + 1]);
      const double d12 = std::hypot(fullv[2 * (size_t)v[2]] - fullv[2 * (size_t)v[1]], fullv[2 * (size_t)v[2] + 1] - fullv[2 * (size_t)v[1] + 1]);
      const double diam = std::max(d03, d12);
      const int mat = (dist < rad + diam / 2 && dist > rad - diam / 2) ? 10 : 0;
      for (int k = 0; k < 4; ++k) { qcells.push_back(v[k]); used[v[k]] = 1; }
      qmat.push_back(mat);
    }
  // delete_unused_vertices: compact, order preserved
  std::vector<int> newid(used.size(), -1);
  m.vx.clear();
  int nv = 0;
  for (size_t v = 0; v < used.size(); ++v)
    if (used[v]) {
      newid[v] = nv++;
      m.vx.push_back(fullv[2 * v]);
      m.vx.push_back(fullv[2 * v + 1]);
    }
  for (auto &v : qcells) v = newid[v];

  m.cells.clear(); m.material.clear();
  if (!triangles) {
    m.elem = 0; m.nvpc = 4;
    m.cells = qcells; m.material = qmat;
  } else {
    m.elem = 1; m.nvpc = 3;
    for (size_t c = 0; c < qmat.size(); ++c) {
      const int *v = &qcells[4 * c];
      const int t0[3] = {v[0], v[1], v[3]}, t1[3] = {v[0], v[3], v[2]};
      m.cells.insert(m.cells.end(), t0, t0 + 3); m.material.push_back(qmat[c]);
      m.cells.insert(m.cells.end(), t1, t1 + 3); m.material.push_back(qmat[c]);
    }
  }
  // boundary ids (NSSolverStationary.cpp:77-95)
  EdgeTable E = build_edges(m);
  m.bfaces.clear();
  const int nc = m.ncells(), nf = m.nvpc;
  for (int c = 0; c < nc; ++c)
    for (int f = 0; f < nf; ++f) {
      if (E.line_ncells[E.cell_line[(size_t)c * nf + f]] != 1) continue;
      int a, b;
      face_vertices(m.elem, f, a, b);
      const int va = m.cells[(size_t)c * nf + a], vb = m.cells[(size_t)c * nf + b];
      const double fcx = (m.vx[2 * (size_t)va] + m.vx[2 * (size_t)vb]) / 2.0;
      int bid;
      if (std::fabs(fcx - x0) < 1e-12) bid = 7;
      else if (std::fabs(fcx - x1) < 1e-12) bid = 8;
      else if (m.material[c] == 10) bid = 10;
      else bid = 6;
      m.bfaces.push_back({c, f, bid});
    }
  m.cell_rank.assign(nc, 0);
}

void read_gmsh2(const std::string &path, Mesh &m) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open mesh file " + path);
  m.elem = 1; m.nvpc = 3;
  m.vx.clear(); m.cells.clear(); m.material.clear(); m.bfaces.clear();
  std::string line;
  std::unordered_map<long, int> node_id;
  std::map<uint64_t, int> line_bid;
  while (std::getline(in, line)) {
    if (line.rfind("$MeshFormat", 0) == 0) {
      double ver; int ft, ds;
      in >> ver >> ft >> ds;
      if (ver < 2.0 || ver >= 3.0 || ft != 0) throw std::runtime_error("only Gmsh 2.x ASCII meshes are supported");
    } else if (line.rfind("$Nodes", 0) == 0) {
      long n; in >> n;
      m.vx.reserve(2 * n);
      for (long i = 0; i < n; ++i) {
        long tag; double x, y, z;
        in >> tag >> x >> y >> z;
        node_id[tag] = (int)i;
        m.vx.push_back(x); m.vx.push_back(y);
      }
    } else if (line.rfind("$Elements", 0) == 0) {
      long n; in >> n;
      std::getline(in, line);
      for (long i = 0; i < n; ++i) {
        std::getline(in, line);
        std::istringstream ss(line);
        long id; int type, ntags;
        ss >> id >> type >> ntags;
        int phys = 0;
        for (int t = 0; t < ntags; ++t) { int tag; ss >> tag; if (t == 0) phys = tag; }
        if (type == 1) {
          long a, b; ss >> a >> b;
          line_bid[edge_key(node_id.at(a), node_id.at(b))] = phys;
        } else if (type == 2) {
          long a, b, c; ss >> a >> b >> c;
          int v[3] = {node_id.at(a), node_id.at(b), node_id.at(c)};
          const double ax = m.vx[2 * (size_t)v[1]] - m.vx[2 * (size_t)v[0]], ay = m.vx[2 * (size_t)v[1] + 1] - m.vx[2 * (size_t)v[0] + 1];
          const double bx = m.vx[2 * (size_t)v[2]] - m.vx[2 * (size_t)v[0]], by = m.vx[2 * (size_t)v[2] + 1] - m.vx[2 * (size_t)v[0] + 1];
          if (ax * by - ay * bx < 0) std::swap(v[1], v[2]);  // keep positive measure
          m.cells.insert(m.cells.end(), v, v + 3);
          m.material.push_back(phys);
        }  // points (15) and everything else are ignored
      }
    }
  }
  if (m.cells.empty()) throw std::runtime_error("mesh file holds no triangles: " + path);
  EdgeTable E = build_edges(m);
  const int nc = m.ncells();
  for (int c = 0; c < nc; ++c)
    for (int f = 0; f < 3; ++f) {
      if (E.line_ncells[E.cell_line[(size_t)c * 3 + f]] != 1) continue;
      int a, b;
      face_vertices(1, f, a, b);
      auto it = line_bid.find(edge_key(m.cells[(size_t)c * 3 + a], m.cells[(size_t)c * 3 + b]));
      m.bfaces.push_back({c, f, it == line_bid.end() ? 0 : it->second});
    }
  m.cell_rank.assign(nc, 0);
}

void partition_strips(Mesh &m, int nranks) {
  const int nc = m.ncells();
  m.cell_rank.assign(nc, 0);
  if (nranks <= 1) return;
  std::vector<int> order(nc);
  std::iota(order.begin(), order.end(), 0);
  std::vector<double> cx(nc);
  for (int c = 0; c < nc; ++c) {
    double s = 0;
    for (int k = 0; k < m.nvpc; ++k) s += m.vx[2 * (size_t)m.cells[(size_t)c * m.nvpc + k]];
    cx[c] = s / m.nvpc;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cx[a] < cx[b]; });
  for (int k = 0; k < nc; ++k) m.cell_rank[order[k]] = (int)(((int64_t)k * nranks) / nc);
}

namespace {

// Sparsity of one block from a cell -> dof table (ids of the pressure block offset by `p_off`); rows at and
// beyond `nrows` (ghost rows of a local view) are left out.
void build_block_pattern(const FETables &T, const uint32_t *cell_dofs, int64_t nc, int64_t p_off, bool row_is_p, bool col_is_p,
                         int64_t nrows, int64_t ncols, CSRPattern &P) {
  const int nd = T.ndofs;
  const uint32_t roff = row_is_p ? (uint32_t)p_off : 0u, coff = col_is_p ? (uint32_t)p_off : 0u;
  P.nrows = nrows; P.ncols = ncols;
  // row -> cells adjacency (CSR)
  std::vector<int64_t> rc_ptr(nrows + 1, 0);
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < nd; ++i)
      if ((T.dof_comp[i] == 2) == row_is_p) {
        const int64_t r = (int64_t)cell_dofs[(size_t)c * nd + i] - roff;
        if (r < nrows) rc_ptr[r + 1]++;
      }
  for (int64_t r = 0; r < nrows; ++r) rc_ptr[r + 1] += rc_ptr[r];
  std::vector<int32_t> rc(rc_ptr[nrows]);
  {
    std::vector<int64_t> fill(rc_ptr.begin(), rc_ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int i = 0; i < nd; ++i)
        if ((T.dof_comp[i] == 2) == row_is_p) {
          const int64_t r = (int64_t)cell_dofs[(size_t)c * nd + i] - roff;
          if (r < nrows) rc[fill[r]++] = (int32_t)c;
        }
  }
  std::vector<int> lcols;
  for (int j = 0; j < nd; ++j)
    if ((T.dof_comp[j] == 2) == col_is_p) lcols.push_back(j);
  P.rowptr.assign(nrows + 1, 0);
  P.col.clear();
  // pass 1: counts, pass 2: fill
  for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel
    {
      std::vector<int32_t> tmp;
#pragma omp for schedule(static)
      for (int64_t r = 0; r < nrows; ++r) {
        tmp.clear();
        for (int64_t k = rc_ptr[r]; k < rc_ptr[r + 1]; ++k) {
          const uint32_t *cd = &cell_dofs[(size_t)rc[k] * nd];
          for (int j : lcols) tmp.push_back((int32_t)(cd[j] - coff));
        }
        std::sort(tmp.begin(), tmp.end());
        const size_t n = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
        if (pass == 0) P.rowptr[r + 1] = (int64_t)n;
        else std::copy(tmp.begin(), tmp.begin() + n, P.col.begin() + P.rowptr[r]);
      }
    }
    if (pass == 0) {
      for (int64_t r = 0; r < nrows; ++r) P.rowptr[r + 1] += P.rowptr[r];
      P.col.resize(P.rowptr[nrows]);
    }
  }
}

}  // namespace

void build_discretisation(Discretisation &d) {
  Mesh &m = d.mesh;
  build_fe_tables(m.elem, d.fe);
  const FETables &T = d.fe;
  const int nc = m.ncells(), nv = m.nverts(), nf = m.nvpc, nd = T.ndofs;
  d.nranks = 1;
  for (int r : m.cell_rank) d.nranks = std::max(d.nranks, r + 1);
  const int nranks = d.nranks;

  EdgeTable E = build_edges(m);
  if (m.elem == 0) {
    // Q3 line dofs are ordered along the line; the generated meshes have every line pointing
    // from its lower to its higher vertex index in every cell, so no flips are needed.
    for (int c = 0; c < nc; ++c)
      for (int f = 0; f < 4; ++f) {
        int a, b;
        face_vertices(0, f, a, b);
        if (m.cells[(size_t)c * 4 + a] > m.cells[(size_t)c * 4 + b])
          throw std::runtime_error("quad mesh with non-standard line orientation is not supported");
      }
  }
  // object owners = lowest subdomain id touching the object
  std::vector<int> vowner(nv, nranks), lowner(E.nlines, nranks);
  for (int c = 0; c < nc; ++c) {
    const int r = m.cell_rank[c];
    for (int k = 0; k < nf; ++k) {
      int &vo = vowner[m.cells[(size_t)c * nf + k]]; vo = std::min(vo, r);
      int &lo = lowner[E.cell_line[(size_t)c * nf + k]]; lo = std::min(lo, r);
    }
  }
  // dofs per object: vertex [ux uy p]; line Q3/Q2 [ux ux uy uy p], P2/P1 [ux uy]; quad [ux*4 uy*4 p]
  const int dpv = 3, dpl = (m.elem == 0) ? 5 : 2, dpq = (m.elem == 0) ? 9 : 0;
  std::vector<int64_t> vfirst(nv, -1), lfirst(E.nlines, -1), qfirst(nc, -1);
  int64_t next = 0;
  for (int r = 0; r < nranks; ++r)
    for (int c = 0; c < nc; ++c) {
      if (m.cell_rank[c] != r) continue;
      for (int k = 0; k < nf; ++k) {
        const int v = m.cells[(size_t)c * nf + k];
        if (vfirst[v] < 0 && vowner[v] == r) { vfirst[v] = next; next += dpv; }
      }
      for (int k = 0; k < nf; ++k) {
        const int l = E.cell_line[(size_t)c * nf + k];
        if (lfirst[l] < 0 && lowner[l] == r) { lfirst[l] = next; next += dpl; }
      }
      if (dpq) { qfirst[c] = next; next += dpq; }
    }
  const int64_t ntot = next;
  // component of each old index
  std::vector<uint8_t> is_p(ntot, 0);
  for (int v = 0; v < nv; ++v) if (vfirst[v] >= 0) is_p[vfirst[v] + 2] = 1;
  for (int l = 0; l < E.nlines; ++l) if (lfirst[l] >= 0 && m.elem == 0) is_p[lfirst[l] + 4] = 1;
  if (dpq) for (int c = 0; c < nc; ++c) is_p[qfirst[c] + 8] = 1;
  // component_wise({0,0,1}): stable by old index inside each block
  std::vector<uint32_t> renum(ntot);
  int64_t cu = 0, cp = 0;
  for (int64_t i = 0; i < ntot; ++i) if (!is_p[i]) renum[i] = (uint32_t)cu++;
  d.n_u = cu;
  for (int64_t i = 0; i < ntot; ++i) if (is_p[i]) renum[i] = (uint32_t)(d.n_u + cp++);
  d.n_p = cp;
  // owned ranges per rank (old numbering is rank-contiguous)
  d.owned_u.assign(nranks + 1, 0); d.owned_p.assign(nranks + 1, 0);
  {
    std::vector<int64_t> cnt_u(nranks, 0), cnt_p(nranks, 0);
    for (int v = 0; v < nv; ++v) if (vfirst[v] >= 0) { cnt_u[vowner[v]] += 2; cnt_p[vowner[v]] += 1; }
    for (int l = 0; l < E.nlines; ++l) if (lfirst[l] >= 0) { cnt_u[lowner[l]] += (m.elem == 0 ? 4 : 2); cnt_p[lowner[l]] += (m.elem == 0 ? 1 : 0); }
    if (dpq) for (int c = 0; c < nc; ++c) { cnt_u[m.cell_rank[c]] += 8; cnt_p[m.cell_rank[c]] += 1; }
    for (int r = 0; r < nranks; ++r) { d.owned_u[r + 1] = d.owned_u[r] + cnt_u[r]; d.owned_p[r + 1] = d.owned_p[r] + cnt_p[r]; }
  }
  // cell -> dof table in the cell-local order of FETables
  d.cell_dofs.resize((size_t)nc * nd);
  d.cell_vertices.resize((size_t)nc * nf * 2);
  for (int c = 0; c < nc; ++c) {
    uint32_t *cd = &d.cell_dofs[(size_t)c * nd];
    int i = 0;
    for (int k = 0; k < nf; ++k) {
      const int64_t b = vfirst[m.cells[(size_t)c * nf + k]];
      cd[i++] = renum[b]; cd[i++] = renum[b + 1]; cd[i++] = renum[b + 2];
    }
    for (int k = 0; k < nf; ++k) {
      const int64_t b = lfirst[E.cell_line[(size_t)c * nf + k]];
      for (int t = 0; t < dpl; ++t) cd[i++] = renum[b + t];
    }
    for (int t = 0; t < dpq; ++t) cd[i++] = renum[qfirst[c] + t];
    for (int k = 0; k < nf; ++k) {
      d.cell_vertices[((size_t)c * nf + k) * 2] = m.vx[2 * (size_t)m.cells[(size_t)c * nf + k]];
      d.cell_vertices[((size_t)c * nf + k) * 2 + 1] = m.vx[2 * (size_t)m.cells[(size_t)c * nf + k] + 1];
    }
  }
  // block sparsity from the coupling table (everything but p-p; p-p only for Mp)
  build_block_pattern(T, d.cell_dofs.data(), nc, d.n_u, false, false, d.n_u, d.n_u, d.F);
  build_block_pattern(T, d.cell_dofs.data(), nc, d.n_u, false, true, d.n_u, d.n_p, d.Bt);
  build_block_pattern(T, d.cell_dofs.data(), nc, d.n_u, true, false, d.n_p, d.n_u, d.B);
  build_block_pattern(T, d.cell_dofs.data(), nc, d.n_u, true, true, d.n_p, d.n_p, d.Mp);

  // Dirichlet lists: boundary 7 first, then {7, 6, 10}; last writer wins (std::map assignment
  // semantics of interpolate_boundary_values, NSSolverStationary.cpp:560-572).
  const double H = 0.41;
  struct BV { double shape, y; uint8_t inlet; };
  std::map<uint32_t, BV> bv;
  auto visit = [&](bool only_inlet) {
    for (const BFace &bf : m.bfaces) {
      const bool inlet = bf.bid == 7;
      if (!(inlet || (!only_inlet && (bf.bid == 6 || bf.bid == 10)))) continue;
      int a, b;
      face_vertices(m.elem, bf.face, a, b);
      const uint32_t *cd = &d.cell_dofs[(size_t)bf.cell * nd];
      const double *xa = &d.cell_vertices[((size_t)bf.cell * nf + a) * 2], *xb = &d.cell_vertices[((size_t)bf.cell * nf + b) * 2];
      auto put = [&](int ldof, double t) {
        if (T.dof_comp[ldof] == 2) return;
        const double y = xa[1] * (1.0 - t) + xb[1] * t;
        double shape = 0.0;
        if (inlet && T.dof_comp[ldof] == 0) shape = 4 * y * (H - y) / (H * H);
        bv[cd[ldof]] = BV{shape, y, (uint8_t)(inlet && T.dof_comp[ldof] == 0)};
      };
      for (int c = 0; c < 3; ++c) { put(3 * a + c, 0.0); put(3 * b + c, 1.0); }
      const int lbase = 3 * nf + dpl * bf.face;
      if (m.elem == 0) {
        const double s5 = std::sqrt(5.0), g1 = 0.5 * (1 - 1 / s5), g2 = 0.5 * (1 + 1 / s5);
        put(lbase + 0, g1); put(lbase + 1, g2); put(lbase + 2, g1); put(lbase + 3, g2);
      } else {
        put(lbase + 0, 0.5); put(lbase + 1, 0.5);
      }
    }
  };
  visit(true);
  visit(false);
  d.bc_dof.clear(); d.bc_shape.clear(); d.bc_on_inlet.clear(); d.bc_y.clear();
  for (auto &kv : bv) {
    d.bc_dof.push_back(kv.first);
    d.bc_shape.push_back(kv.second.shape);
    d.bc_on_inlet.push_back(kv.second.inlet);
    d.bc_y.push_back(kv.second.y);
  }
  d.outlet_cell.clear(); d.outlet_face.clear(); d.cylinder_cell.clear(); d.cylinder_face.clear();
  for (const BFace &bf : m.bfaces) {
    if (bf.bid == 8) { d.outlet_cell.push_back(bf.cell); d.outlet_face.push_back(bf.face); }
    if (bf.bid == 10) { d.cylinder_cell.push_back(bf.cell); d.cylinder_face.push_back(bf.face); }
  }
}

void build_local_view(const Discretisation &g, int rank, Discretisation &l) {
  if (g.is_local) throw std::invalid_argument("build_local_view needs the global discretisation");
  if (rank < 0 || rank >= g.nranks) throw std::invalid_argument("rank outside the partition");
  const FETables &T = g.fe;
  const int nd = T.ndofs, nv = g.mesh.nvpc;
  const int64_t nc = g.mesh.ncells();
  const int64_t u0 = g.owned_u[rank], u1 = g.owned_u[rank + 1], p0 = g.owned_p[rank], p1 = g.owned_p[rank + 1];
  auto owner_of = [&](uint32_t d) -> int {  // block-global dof id -> owning rank
    const bool isp = (int64_t)d >= g.n_u;
    const std::vector<int64_t> &ow = isp ? g.owned_p : g.owned_u;
    const int64_t v = isp ? (int64_t)d - g.n_u : (int64_t)d;
    return (int)(std::upper_bound(ow.begin(), ow.end(), v) - ow.begin()) - 1;
  };
  l = Discretisation();
  l.is_local = true; l.rank = rank; l.job_ranks = g.nranks; l.nranks = 1;
  l.fe = g.fe;
  l.mesh.elem = g.mesh.elem; l.mesh.nvpc = nv;
  l.n_u_owned = u1 - u0; l.n_p_owned = p1 - p0;

  // local cells (global order) and, per neighbour, the owned dofs it needs
  std::vector<int64_t> ghost_u, ghost_p;
  std::vector<std::vector<int32_t>> send_u(g.nranks), send_p(g.nranks);
  std::vector<int> owners(nd);
  for (int64_t c = 0; c < nc; ++c) {
    const uint32_t *cd = &g.cell_dofs[(size_t)c * nd];
    bool mine = false;
    unsigned long long others = 0;  // bit set of other owner ranks (<= 64 ranks)
    for (int i = 0; i < nd; ++i) {
      owners[i] = owner_of(cd[i]);
      if (owners[i] == rank) mine = true;
      else others |= 1ull << (owners[i] & 63);
    }
    if (!mine) continue;
    l.cell_global.push_back((int32_t)c);
    l.cell_owned.push_back(g.mesh.cell_rank[c] == rank);
    if (!others) continue;
    for (int i = 0; i < nd; ++i) {
      const bool isp = T.dof_comp[i] == 2;
      if (owners[i] == rank) {
        const int32_t lid = (int32_t)(isp ? (int64_t)cd[i] - g.n_u - p0 : (int64_t)cd[i] - u0);
        for (int r = 0; r < g.nranks; ++r)
          if (r != rank && (others >> (r & 63) & 1)) (isp ? send_p : send_u)[r].push_back(lid);
      } else {
        (isp ? ghost_p : ghost_u).push_back(isp ? (int64_t)cd[i] - g.n_u : (int64_t)cd[i]);
      }
    }
  }
  if (g.nranks > 64) throw std::runtime_error("more than 64 ranks are not supported by the host set-up stand-in");
  auto uniq = [](auto &v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
  uniq(ghost_u); uniq(ghost_p);
  for (auto &v : send_u) uniq(v);
  for (auto &v : send_p) uniq(v);
  l.n_u = l.n_u_owned + (int64_t)ghost_u.size();
  l.n_p = l.n_p_owned + (int64_t)ghost_p.size();
  l.l2g_u.resize(l.n_u); l.l2g_p.resize(l.n_p);
  for (int64_t i = 0; i < l.n_u_owned; ++i) l.l2g_u[i] = u0 + i;
  for (size_t i = 0; i < ghost_u.size(); ++i) l.l2g_u[l.n_u_owned + i] = ghost_u[i];
  for (int64_t i = 0; i < l.n_p_owned; ++i) l.l2g_p[i] = p0 + i;
  for (size_t i = 0; i < ghost_p.size(); ++i) l.l2g_p[l.n_p_owned + i] = ghost_p[i];
  auto local_u = [&](int64_t gid) -> int64_t {
    if (gid >= u0 && gid < u1) return gid - u0;
    return l.n_u_owned + (std::lower_bound(ghost_u.begin(), ghost_u.end(), gid) - ghost_u.begin());
  };
  auto local_p = [&](int64_t gid) -> int64_t {
    if (gid >= p0 && gid < p1) return gid - p0;
    return l.n_p_owned + (std::lower_bound(ghost_p.begin(), ghost_p.end(), gid) - ghost_p.begin());
  };
  // halo plans: ghosts are sorted by global id, i.e. grouped by owner
  auto plan = [&](Discretisation::Halo &H, const std::vector<int64_t> &ghost, const std::vector<int64_t> &owned,
                  const std::vector<std::vector<int32_t>> &send) {
    H.send_ptr.assign(1, 0); H.recv_ptr.assign(1, 0);
    size_t gpos = 0;
    for (int r = 0; r < g.nranks; ++r) {
      size_t gend = gpos;
      while (gend < ghost.size() && ghost[gend] < owned[r + 1]) ++gend;
      const bool recv = gend > gpos, snd = !send[r].empty();
      if (r != rank && (recv || snd)) {
        H.nbr.push_back(r);
        H.send_idx.insert(H.send_idx.end(), send[r].begin(), send[r].end());
        H.send_ptr.push_back((int64_t)H.send_idx.size());
        H.recv_ptr.push_back((int64_t)gend);
      } else if (recv) {
        throw std::runtime_error("ghost dof owned by this rank");
      }
      gpos = gend;
    }
  };
  plan(l.halo_u, ghost_u, g.owned_u, send_u);
  plan(l.halo_p, ghost_p, g.owned_p, send_p);

  // cell tables in local numbering
  const int64_t lc = (int64_t)l.cell_global.size();
  l.cell_dofs.resize((size_t)lc * nd);
  l.cell_vertices.resize((size_t)lc * nv * 2);
  l.mesh.material.resize(lc);
  l.mesh.cell_rank.assign(lc, 0);
  l.mesh.cells.assign((size_t)lc * nv, 0);  // vertex ids are not carried over; ncells() stays valid
  std::vector<int32_t> g2l_cell(nc, -1);
  for (int64_t k = 0; k < lc; ++k) {
    const int64_t c = l.cell_global[k];
    g2l_cell[c] = (int32_t)k;
    l.mesh.material[k] = g.mesh.material[c];
    for (int i = 0; i < nd; ++i) {
      const uint32_t d = g.cell_dofs[(size_t)c * nd + i];
      l.cell_dofs[(size_t)k * nd + i] = (T.dof_comp[i] == 2) ? (uint32_t)(l.n_u + local_p((int64_t)d - g.n_u)) : (uint32_t)local_u(d);
    }
    std::copy(&g.cell_vertices[(size_t)c * nv * 2], &g.cell_vertices[(size_t)(c + 1) * nv * 2], &l.cell_vertices[(size_t)k * nv * 2]);
  }
  l.owned_u = {0, l.n_u_owned};
  l.owned_p = {0, l.n_p_owned};
  // owned rows of the four blocks, local columns
  build_block_pattern(T, l.cell_dofs.data(), lc, l.n_u, false, false, l.n_u_owned, l.n_u, l.F);
  build_block_pattern(T, l.cell_dofs.data(), lc, l.n_u, false, true, l.n_u_owned, l.n_p, l.Bt);
  build_block_pattern(T, l.cell_dofs.data(), lc, l.n_u, true, false, l.n_p_owned, l.n_u, l.B);
  build_block_pattern(T, l.cell_dofs.data(), lc, l.n_u, true, true, l.n_p_owned, l.n_p, l.Mp);
  // boundary lists: owned constrained dofs; outlet faces of every local cell (their owned rows take the term);
  // cylinder faces of this rank's own cells (compute_lift_drag loops locally owned cells)
  for (size_t k = 0; k < g.bc_dof.size(); ++k) {
    const int64_t d = g.bc_dof[k];
    if (d < u0 || d >= u1) continue;
    l.bc_dof.push_back((uint32_t)(d - u0));
    l.bc_shape.push_back(g.bc_shape[k]);
    l.bc_on_inlet.push_back(g.bc_on_inlet[k]);
    l.bc_y.push_back(g.bc_y[k]);
  }
  for (size_t k = 0; k < g.outlet_cell.size(); ++k) {
    const int32_t lcid = g2l_cell[g.outlet_cell[k]];
    if (lcid >= 0) { l.outlet_cell.push_back(lcid); l.outlet_face.push_back(g.outlet_face[k]); }
  }
  for (size_t k = 0; k < g.cylinder_cell.size(); ++k) {
    const int c = g.cylinder_cell[k];
    if (g.mesh.cell_rank[c] == rank) { l.cylinder_cell.push_back(g2l_cell[c]); l.cylinder_face.push_back(g.cylinder_face[k]); }
  }
  for (const BFace &bf : g.mesh.bfaces)
    if (g2l_cell[bf.cell] >= 0) l.mesh.bfaces.push_back({g2l_cell[bf.cell], bf.face, bf.bid});
}

}  // namespace nsx
