// assemble.cu -- per-cell quadrature assembly of the linearised Stokes / Newton system, its
// right-hand side and the pressure mass matrix, straight into the device CSR blocks.
//
// Replaces the cell loop of NSSolverStationary::assemble_system (lab_new/src/
// NSSolverStationary.cpp:356-533; Stokes branch 383-406, Newton branch 408-452, residual 462-493,
// outlet term 503-526, scatter 528-532, Dirichlet step 540-576) and of NSSolver::assemble_system
// (lab_new/src/NSSolver.cpp:358-558; first-iteration branch 381-409, Newton branch 411-469,
// residual 479-519), plus compute_lift_drag (NSSolverStationary.cpp:835-892).
//
// Device layout.  Every FESystem shape function has one non-zero component, so the 41 x 41 (or
// 15 x 15) cell matrix of the reference splits into scalar node blocks:
//   F [(a,c),(b,c')]  velocity-velocity     Bt[(a,c),m]  velocity-pressure
//   B [m,(b,c')]      pressure-velocity     Mp[m,m']     pressure mass
// A group of TPC threads owns one cell: geometry, physical gradients and the fields at the
// quadrature points are staged in shared memory next to the reference-cell tables, then the
// threads sweep the node pairs with the quadrature sum in registers.  Cells are launched colour
// by colour (no two cells of a colour share a dof), so the scatter `val[row_start + offset] += v`
// needs no atomics and sums in a fixed order.  `offset` comes from a table of row-relative CSR
// positions that cells with the same local connectivity share (a few hundred tables on the
// structured meshes), built once with the sparsity pattern.
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "device.cuh"

namespace nsx {

// ---------------------------------------------------------------------------------------------
// host: colouring, offset tables, outlet vector
// ---------------------------------------------------------------------------------------------
namespace {

inline int find_in_row(const DevCSR &A, int64_t row, int32_t col) {
  const int32_t *b = A.h_col.data() + A.h_rowptr[row], *e = A.h_col.data() + A.h_rowptr[row + 1];
  const int32_t *it = std::lower_bound(b, e, col);
  if (it == e || *it != col) throw std::runtime_error("cell couples two dofs outside the sparsity pattern");
  return (int)(it - b);
}

// reference-cell point -> Jacobian of the (bi)linear mapping, host + device
__host__ __device__ inline void jacobian_at(int elem, const double *xv, double x, double y, double J[2][2]) {
  if (elem == 0) {
    const double d0x = -(1 - y), d0y = -(1 - x), d1x = (1 - y), d1y = -x, d2x = -y, d2y = (1 - x), d3x = y, d3y = x;
    J[0][0] = xv[0] * d0x + xv[2] * d1x + xv[4] * d2x + xv[6] * d3x;
    J[0][1] = xv[0] * d0y + xv[2] * d1y + xv[4] * d2y + xv[6] * d3y;
    J[1][0] = xv[1] * d0x + xv[3] * d1x + xv[5] * d2x + xv[7] * d3x;
    J[1][1] = xv[1] * d0y + xv[3] * d1y + xv[5] * d2y + xv[7] * d3y;
  } else {
    J[0][0] = xv[2] - xv[0]; J[0][1] = xv[4] - xv[0];
    J[1][0] = xv[3] - xv[1]; J[1][1] = xv[5] - xv[1];
  }
}

// outward unit normal and length of face f
__host__ __device__ inline void face_geometry(int elem, const double *xv, int f, double &nx, double &ny, double &len) {
  int a, b, o;
  if (elem == 0) {
    const int fv[4][3] = {{0, 2, 1}, {1, 3, 0}, {0, 1, 2}, {2, 3, 0}};
    a = fv[f][0]; b = fv[f][1]; o = fv[f][2];
  } else {
    a = f; b = (f + 1) % 3; o = (f + 2) % 3;
  }
  const double tx = xv[2 * b] - xv[2 * a], ty = xv[2 * b + 1] - xv[2 * a + 1];
  len = sqrt(tx * tx + ty * ty);
  nx = ty / len; ny = -tx / len;
  if (nx * (xv[2 * o] - xv[2 * a]) + ny * (xv[2 * o + 1] - xv[2 * a + 1]) > 0) { nx = -nx; ny = -ny; }
}

}  // namespace

void build_assembly_maps(Ctx &c) {
  const FETables &T = c.fe;
  const int nd = T.ndofs, nv = T.nvpc;
  const int64_t nc = c.ncells;
  const uint32_t *cd = c.h_cell_dofs.data();

  // --- greedy colouring over the "shares a vertex" graph; the u_x dof of a vertex names it ---
  const int64_t lu = c.n_u + c.n_ug;   // local velocity dofs (owned + ghost): the pressure ids of the cell table start here
  std::vector<int64_t> vptr(lu + 1, 0);
  for (int64_t k = 0; k < nc; ++k)
    for (int v = 0; v < nv; ++v) vptr[cd[k * nd + 3 * v] + 1]++;
  for (int64_t i = 0; i < lu; ++i) vptr[i + 1] += vptr[i];
  std::vector<int32_t> vcell(vptr[lu]);
  {
    std::vector<int64_t> fill(vptr.begin(), vptr.end() - 1);
    for (int64_t k = 0; k < nc; ++k)
      for (int v = 0; v < nv; ++v) vcell[fill[cd[k * nd + 3 * v]]++] = (int32_t)k;
  }
  std::vector<int> colour(nc, -1);
  int ncol = 0;
  for (int64_t k = 0; k < nc; ++k) {
    uint64_t used = 0;
    for (int v = 0; v < nv; ++v) {
      const uint32_t key = cd[k * nd + 3 * v];
      for (int64_t p = vptr[key]; p < vptr[key + 1]; ++p)
        if (colour[vcell[p]] >= 0) used |= 1ull << colour[vcell[p]];
    }
    int col = 0;
    while (used & (1ull << col)) ++col;
    if (col >= 63) throw std::runtime_error("cell colouring needs more than 63 colours");
    colour[k] = col;
    ncol = std::max(ncol, col + 1);
  }
  c.ncolors = ncol;
  c.color_ptr.assign(ncol + 1, 0);
  for (int64_t k = 0; k < nc; ++k) c.color_ptr[colour[k] + 1]++;
  for (int q = 0; q < ncol; ++q) c.color_ptr[q + 1] += c.color_ptr[q];
  std::vector<int32_t> cells(nc);
  {
    std::vector<int64_t> fill(c.color_ptr.begin(), c.color_ptr.end() - 1);
    for (int64_t k = 0; k < nc; ++k) cells[fill[colour[k]]++] = (int32_t)k;
  }
  c.color_cells.upload(cells, c.stream);

  // --- partitioned system: rows of Bt for the ghost velocity dofs (for S = B diag(F)^-1 Bt, spgemm.cu), from the local cells ---
  if (c.n_ug > 0 && c.Bt.nrows_ext == 0) {
    DevCSR &Bt = c.Bt;
    std::vector<std::vector<int32_t>> grow(c.n_ug);
    for (int64_t k = 0; k < nc; ++k)
      for (int i = 0; i < nd; ++i) {
        if (T.dof_comp[i] == 2) continue;
        const int64_t d = cd[k * nd + i];
        if (d < c.n_u) continue;
        for (int j = 0; j < nd; ++j)
          if (T.dof_comp[j] == 2) grow[d - c.n_u].push_back((int32_t)((int64_t)cd[k * nd + j] - lu));
      }
    Bt.h_rowptr.resize(c.n_u + c.n_ug + 1);
    for (int64_t g = 0; g < c.n_ug; ++g) {
      std::sort(grow[g].begin(), grow[g].end());
      grow[g].erase(std::unique(grow[g].begin(), grow[g].end()), grow[g].end());
      Bt.h_rowptr[c.n_u + g + 1] = Bt.h_rowptr[c.n_u + g] + (int64_t)grow[g].size();
      Bt.h_col.insert(Bt.h_col.end(), grow[g].begin(), grow[g].end());
    }
    Bt.nrows_ext = c.n_u + c.n_ug; Bt.nnz_ext = Bt.h_rowptr.back();
    std::vector<int32_t> baked(Bt.nnz_ext);
    for (int64_t k = 0; k < Bt.nnz_ext; ++k) baked[k] = (int32_t)(Bt.h_col[k] < c.n_p ? Bt.h_col[k] : Bt.h_col[k] + c.n_ug);
    Bt.rowptr.alloc_padded(Bt.h_rowptr.size(), 4, c.stream);
    NSX_CUDA(cudaMemcpyAsync(Bt.rowptr.p, Bt.h_rowptr.data(), Bt.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
    Bt.col.alloc_padded(Bt.nnz_ext, 16, c.stream);
    NSX_CUDA(cudaMemcpyAsync(Bt.col.p, baked.data(), Bt.nnz_ext * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
    Bt.val.alloc_padded(Bt.nnz_ext, 16, c.stream);
    Bt.nrb = Bt.ndesc = 0; c.nrb_u = c.nrb_p = c.ndesc_u = c.ndesc_p = 0;
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  }
  const int64_t bt_rows = c.Bt.nrows_ext ? c.Bt.nrows_ext : c.n_u;

  // --- row-relative offset tables, shared between cells with the same local connectivity ---
  const int tsz = nd * nd;
  std::vector<uint16_t> tables;
  std::vector<int32_t> cell_pat(nc);
  std::unordered_map<uint64_t, std::vector<int32_t>> seen;
  const int64_t CH = 8192;
  std::vector<uint16_t> buf((size_t)CH * tsz);
  std::vector<uint64_t> hash(CH);
  std::string err;
  for (int64_t c0 = 0; c0 < nc; c0 += CH) {
    const int64_t c1 = std::min(nc, c0 + CH);
#pragma omp parallel for schedule(static)
    for (int64_t k = c0; k < c1; ++k) {
      uint16_t *t = &buf[(size_t)(k - c0) * tsz];
      try {
        for (int i = 0; i < nd; ++i) {
          const bool ip = T.dof_comp[i] == 2;
          const int64_t ri = ip ? (int64_t)cd[k * nd + i] - lu : cd[k * nd + i];
          const bool row_owned = ri < (ip ? c.n_p : c.n_u);   // ghost rows are assembled by their owner: offset 0 into the sink
          for (int j = 0; j < nd; ++j) {
            const bool jp = T.dof_comp[j] == 2;
            const int32_t cj = (int32_t)(jp ? (int64_t)cd[k * nd + j] - lu : cd[k * nd + j]);
            const DevCSR &A = ip ? (jp ? c.Mp : c.B) : (jp ? c.Bt : c.F);
            const bool stored = row_owned || (!ip && jp && ri < bt_rows);   // ... except the ghost rows of Bt kept for the Schur product
            t[i * nd + j] = stored ? (uint16_t)find_in_row(A, ri, cj) : (uint16_t)0;
          }
        }
      } catch (const std::exception &e) {
#pragma omp critical
        err = e.what();
      }
      uint64_t h = 1469598103934665603ull;
      for (int i = 0; i < tsz; ++i) { h ^= t[i]; h *= 1099511628211ull; }
      hash[k - c0] = h;
    }
    if (!err.empty()) throw std::runtime_error(err);
    for (int64_t k = c0; k < c1; ++k) {
      const uint16_t *t = &buf[(size_t)(k - c0) * tsz];
      auto &bucket = seen[hash[k - c0]];
      int32_t id = -1;
      for (int32_t cand : bucket)
        if (std::memcmp(&tables[(size_t)cand * tsz], t, tsz * sizeof(uint16_t)) == 0) { id = cand; break; }
      if (id < 0) {
        id = (int32_t)(tables.size() / tsz);
        tables.insert(tables.end(), t, t + tsz);
        bucket.push_back(id);
      }
      cell_pat[k] = id;
    }
  }
  c.npat = (int64_t)(tables.size() / tsz);
  c.pat_off.upload(tables, c.stream);
  c.cell_pat.upload(cell_pat, c.stream);

  // --- outlet vector: -sum_faces sum_q (n . phi_i) w_f per velocity dof (NSSolverStationary.cpp:503-526) ---
  {
    std::map<uint32_t, double> acc;
    for (size_t k = 0; k < c.h_outlet_cell.size(); ++k) {
      const int64_t cell = c.h_outlet_cell[k];
      const int f = c.h_outlet_face[k];
      const double *xv = &c.h_cell_vertices[(size_t)cell * nv * 2];
      double nx, ny, len;
      face_geometry(T.elem, xv, f, nx, ny, len);
      for (int q = 0; q < T.nqf; ++q)
        for (int i = 0; i < nd; ++i) {
          if (T.dof_comp[i] == 2) continue;
          const double n_c = T.dof_comp[i] == 0 ? nx : ny;
          acc[cd[cell * nd + i]] -= n_c * T.Nvf[f][T.dof_node[i]][q] * (T.qwf[q] * len);
        }
    }
    std::vector<uint32_t> dof;
    std::vector<double> val;
    for (auto &kv : acc)
      if ((int64_t)kv.first < c.n_u) { dof.push_back(kv.first); val.push_back(kv.second); }  // owned rows only
    c.n_outlet = (int64_t)dof.size();
    c.outlet_dof.upload(dof, c.stream);
    c.outlet_unit.upload(val, c.stream);
  }
  {  // cell-table dof id -> position in a vector laid out [u owned | p owned | u ghosts | p ghosts]
    std::vector<int32_t> vmap(lu + c.n_p + c.n_pg);
    for (int64_t d = 0; d < lu; ++d) vmap[d] = (int32_t)(d < c.n_u ? d : d + c.n_p);
    for (int64_t m = 0; m < c.n_p + c.n_pg; ++m) vmap[lu + m] = (int32_t)(m < c.n_p ? c.n_u + m : c.n_u + c.n_ug + m);
    c.vmap.upload(vmap, c.stream);
  }
  // ghost velocity dofs that are Dirichlet dofs on their owner: a 0/1 mask travels through the ghost import once
  c.n_ghost_bc = 0;
  if (c.n_ug > 0 && c.comm) {
    std::vector<double> mask(c.nvec, 0.0);
    std::vector<uint32_t> bc(c.nbc);
    if (c.nbc) NSX_CUDA(cudaMemcpyAsync(bc.data(), c.bc_dof.p, c.nbc * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
    for (uint32_t d : bc) mask[d] = 1.0;
    double *tmp = c.vec[NSX_VEC_TMP1].p;
    NSX_CUDA(cudaMemcpyAsync(tmp, mask.data(), c.nvec * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    halo_exchange(c, 0, tmp);
    NSX_CUDA(cudaMemcpyAsync(mask.data(), tmp, c.nvec * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<int32_t> gbc;
    for (int64_t g = 0; g < c.n_ug; ++g) if (mask[c.n + g] != 0.0) gbc.push_back((int32_t)(c.n_u + g));
    c.n_ghost_bc = (int64_t)gbc.size();
    c.ghost_bc.upload(gbc, c.stream);
    NSX_CUDA(cudaMemsetAsync(tmp, 0, c.nvec * sizeof(double), c.stream));
  }
  c.cyl_cell.upload(c.h_cyl_cell, c.stream);
  c.cyl_face.upload(c.h_cyl_face, c.stream);
  c.face_force.alloc(2 * std::max<size_t>(1, c.h_cyl_cell.size()));
  c.d_fe.upload(&c.fe, 1, c.stream);
  c.d_owned_u.upload(c.owned_u, c.stream);
  c.bc_first.alloc(c.owned_u.size());
  NSX_CUDA(cudaStreamSynchronize(c.stream));
}

// ---------------------------------------------------------------------------------------------
// device: cell kernel
// ---------------------------------------------------------------------------------------------
namespace {

struct AsmArgs {
  int mode;
  int res_only;   // 1: residual contributions only (the matrices are left alone)
  double nu, inv_dt;
  const double *sol, *sol_old;
  double *res;
  const int64_t *F_rp, *Bt_rp, *B_rp, *Mp_rp;
  double *F_val, *Bt_val, *B_val, *Mp_val;
  const double *cell_vertices;
  const uint32_t *cell_dofs;
  const int32_t *cell_pat;
  const uint16_t *pat_off;
  const int32_t *cells;
  int ncells;
  int64_t n_u_loc, n_u_own, n_p_own;         // pressure ids of the cell table start at n_u_loc; rows >= n_*_own are ghosts
  int64_t bt_rows;                           // rows Bt stores (the owned ones, plus the ghost rows on a partitioned system)
  int64_t F_sink, Bt_sink, B_sink, Mp_sink;  // position in each value array that swallows the ghost rows' contributions
  const int32_t *vmap;
  const FETables *fe;
};

// TPC threads per cell, CPB cells per block
template <int ELEM, int NVPC, int NVN, int NPN, int NQ, int TPC, int CPB>
__global__ void __launch_bounds__(TPC *CPB) k_assemble(const AsmArgs A) {
  constexpr int ND = 2 * NVN + NPN;
  // reference-cell tables (per block)
  __shared__ double sNv[NQ][NVN], sdNv[NQ][NVN][2], sNp[NQ][NPN], sqw[NQ], sqp[NQ][2];
  __shared__ int sldv[NVN][2], sldp[NPN];
  // per cell slot
  __shared__ double sG[CPB][NQ][NVN][2], sconv[CPB][NQ][NVN], sw[CPB][NQ];
  __shared__ double suq[CPB][NQ][2], sdu[CPB][NQ][2], sgu[CPB][NQ][4], spq[CPB][NQ];
  __shared__ double sJi[CPB][NQ][4];
  __shared__ double sU[CPB][NVN][2], sUo[CPB][NVN][2], sP[CPB][NPN], sxv[CPB][NVPC * 2];
  __shared__ double se[CPB][NVN][2];
  __shared__ int64_t srb0[CPB][ND], srb1[CPB][ND];
  __shared__ uint32_t sdof[CPB][ND];

  const int tid = threadIdx.x, slot = tid / TPC, lt = tid % TPC;
  const FETables &T = *A.fe;
  for (int i = tid; i < NQ * NVN; i += TPC * CPB) {
    const int q = i / NVN, a = i % NVN;
    sNv[q][a] = T.Nv[a][q];
    sdNv[q][a][0] = T.dNv[a][q][0];
    sdNv[q][a][1] = T.dNv[a][q][1];
  }
  for (int i = tid; i < NQ * NPN; i += TPC * CPB) sNp[i / NPN][i % NPN] = T.Np[i % NPN][i / NPN];
  for (int i = tid; i < NQ; i += TPC * CPB) { sqw[i] = T.qw[i]; sqp[i][0] = T.qp[i][0]; sqp[i][1] = T.qp[i][1]; }
  for (int i = tid; i < ND; i += TPC * CPB) {
    const int comp = T.dof_comp[i], node = T.dof_node[i];
    if (comp == 2) sldp[node] = i; else sldv[node][comp] = i;
  }
  __syncthreads();

  const bool newton = (A.mode == NSX_MODE_NEWTON || A.mode == NSX_MODE_UNSTEADY_NEWTON);
  const bool unsteady = (A.mode == NSX_MODE_UNSTEADY_FIRST || A.mode == NSX_MODE_UNSTEADY_NEWTON);
  const double nu = A.nu, inv_nu = 1.0 / A.nu, inv_dt = A.inv_dt;
  const double mass = (A.mode == NSX_MODE_UNSTEADY_NEWTON) ? inv_dt : 0.0;

  for (int base = blockIdx.x * CPB; base < A.ncells; base += gridDim.x * CPB) {
    const bool active = base + slot < A.ncells;
    const int cell = active ? A.cells[base + slot] : 0;
    // ---- stage 0: connectivity, coefficients, vertices ----
    if (active) {
      for (int i = lt; i < ND; i += TPC) {
        const uint32_t d = A.cell_dofs[(int64_t)cell * ND + i];
        const uint32_t vi = (uint32_t)A.vmap[d];
        sdof[slot][i] = vi;
        const int comp = T.dof_comp[i], node = T.dof_node[i];
        const double s = A.sol[vi];
        if (comp == 2) {
          const int64_t m = (int64_t)d - A.n_u_loc;
          const bool own = m < A.n_p_own;
          sP[slot][node] = s;
          srb0[slot][i] = own ? A.B_rp[m] : A.B_sink;
          srb1[slot][i] = own ? A.Mp_rp[m] : A.Mp_sink;
        } else {
          const bool own = (int64_t)d < A.n_u_own;
          sU[slot][node][comp] = s;
          sUo[slot][node][comp] = unsteady ? A.sol_old[vi] : 0.0;
          srb0[slot][i] = own ? A.F_rp[d] : A.F_sink;
          srb1[slot][i] = (int64_t)d < A.bt_rows ? A.Bt_rp[d] : A.Bt_sink;
        }
      }
      for (int i = lt; i < NVPC * 2; i += TPC) sxv[slot][i] = A.cell_vertices[(int64_t)cell * NVPC * 2 + i];
    }
    __syncthreads();
    // ---- stage 1a: geometry per quadrature point ----
    if (active && lt < NQ) {
      double J[2][2];
      jacobian_at(ELEM, sxv[slot], sqp[lt][0], sqp[lt][1], J);
      const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      sJi[slot][lt][0] = J[1][1] / det; sJi[slot][lt][1] = -J[0][1] / det;
      sJi[slot][lt][2] = -J[1][0] / det; sJi[slot][lt][3] = J[0][0] / det;
      sw[slot][lt] = sqw[lt] * fabs(det);
    }
    __syncthreads();
    // ---- stage 1b: physical gradients ----
    if (active)
      for (int i = lt; i < NQ * NVN; i += TPC) {
        const int q = i / NVN, a = i % NVN;
        const double g0 = sdNv[q][a][0], g1 = sdNv[q][a][1];
        sG[slot][q][a][0] = sJi[slot][q][0] * g0 + sJi[slot][q][2] * g1;
        sG[slot][q][a][1] = sJi[slot][q][1] * g0 + sJi[slot][q][3] * g1;
      }
    __syncthreads();
    // ---- stage 1c: fields at the quadrature points ----
    if (active && lt < NQ) {
      const int q = lt;
      double u0 = 0, u1 = 0, o0 = 0, o1 = 0, g00 = 0, g01 = 0, g10 = 0, g11 = 0, p = 0;
#pragma unroll
      for (int a = 0; a < NVN; ++a) {
        const double n = sNv[q][a], gx = sG[slot][q][a][0], gy = sG[slot][q][a][1];
        const double c0 = sU[slot][a][0], c1 = sU[slot][a][1];
        u0 += c0 * n; u1 += c1 * n;
        o0 += sUo[slot][a][0] * n; o1 += sUo[slot][a][1] * n;
        g00 += c0 * gx; g01 += c0 * gy; g10 += c1 * gx; g11 += c1 * gy;
      }
#pragma unroll
      for (int m = 0; m < NPN; ++m) p += sP[slot][m] * sNp[q][m];
      suq[slot][q][0] = u0; suq[slot][q][1] = u1;
      sdu[slot][q][0] = unsteady ? u0 - o0 : 0.0; sdu[slot][q][1] = unsteady ? u1 - o1 : 0.0;
      sgu[slot][q][0] = g00; sgu[slot][q][1] = g01; sgu[slot][q][2] = g10; sgu[slot][q][3] = g11;  // [k][l] = d_l u_k
      spq[slot][q] = p;
    }
    __syncthreads();
    // ---- stage 1d: convection of each node function, first-iteration column term ----
    if (active) {
      for (int i = lt; i < NQ * NVN; i += TPC) {
        const int q = i / NVN, b = i % NVN;
        sconv[slot][q][b] = suq[slot][q][0] * sG[slot][q][b][0] + suq[slot][q][1] * sG[slot][q][b][1];
      }
      if (A.mode == NSX_MODE_UNSTEADY_FIRST)
        for (int i = lt; i < NVN * 2; i += TPC) {
          const int a = i / 2, cc = i % 2;
          double s = 0;
          for (int q = 0; q < NQ; ++q) s += sdu[slot][q][cc] * sNv[q][a] * inv_dt * sw[slot][q];
          se[slot][a][cc] = s;
        }
    }
    __syncthreads();
    if (active) {
      const uint16_t *off = A.pat_off + (int64_t)A.cell_pat[cell] * (ND * ND);
      // ---- F: node pairs ----
      for (int pr = lt; pr < (A.res_only ? 0 : NVN * NVN); pr += TPC) {
        const int a = pr / NVN, b = pr % NVN;
        double fb = 0, f00 = 0, f01 = 0, f10 = 0, f11 = 0;
        if (newton) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double w = sw[slot][q];
            const double naw = sNv[q][a] * w, nb = sNv[q][b];
            const double v = sG[slot][q][a][0] * sG[slot][q][b][0] + sG[slot][q][a][1] * sG[slot][q][b][1];
            const double s = naw * nb;
            fb += nu * w * v + naw * sconv[slot][q][b] + s * mass;
            f00 += s * sgu[slot][q][0]; f01 += s * sgu[slot][q][1];
            f10 += s * sgu[slot][q][2]; f11 += s * sgu[slot][q][3];
          }
          f00 += fb; f11 += fb;
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            fb += nu * sw[slot][q] * (sG[slot][q][a][0] * sG[slot][q][b][0] + sG[slot][q][a][1] * sG[slot][q][b][1]);
          f00 = fb; f11 = fb;
          if (A.mode == NSX_MODE_UNSTEADY_FIRST) {
            f00 += se[slot][a][0]; f01 += se[slot][a][0];
            f10 += se[slot][a][1]; f11 += se[slot][a][1];
          }
        }
        const int i0 = sldv[a][0], i1 = sldv[a][1], j0 = sldv[b][0], j1 = sldv[b][1];
        A.F_val[srb0[slot][i0] + off[i0 * ND + j0]] += f00;
        A.F_val[srb0[slot][i1] + off[i1 * ND + j1]] += f11;
        if (f01 != 0.0) A.F_val[srb0[slot][i0] + off[i0 * ND + j1]] += f01;
        if (f10 != 0.0) A.F_val[srb0[slot][i1] + off[i1 * ND + j0]] += f10;
      }
      // ---- Bt and B: (velocity node, component) x pressure node ----
      for (int e = lt; e < (A.res_only ? 0 : NVN * 2 * NPN); e += TPC) {
        const int m = e % NPN, ac = e / NPN, a = ac / 2, cc = ac % 2;
        double s = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) s += sG[slot][q][a][cc] * sNp[q][m] * sw[slot][q];
        const int i = sldv[a][cc], j = sldp[m];
        double bt = -s;
        if (A.mode == NSX_MODE_UNSTEADY_FIRST) bt += se[slot][a][cc];
        A.Bt_val[srb1[slot][i] + off[i * ND + j]] += bt;
        A.B_val[srb0[slot][j] + off[j * ND + i]] += newton ? s : -s;
      }
      // ---- Mp ----
      for (int e = lt; e < (A.res_only ? 0 : NPN * NPN); e += TPC) {
        const int m = e / NPN, k = e % NPN;
        double s = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) s += sNp[q][m] * sNp[q][k] * inv_nu * sw[slot][q];
        const int i = sldp[m], j = sldp[k];
        A.Mp_val[srb1[slot][i] + off[i * ND + j]] += s;
      }
      // ---- residual (Newton branches only) ----
      if (newton)
        for (int e = lt; e < ND; e += TPC) {
          double r = 0;
          if (e < NVN * 2) {
            const int a = e / 2, cc = e % 2;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const double w = sw[slot][q], n = sNv[q][a];
              const double gx = sG[slot][q][a][0], gy = sG[slot][q][a][1];
              const double gux = sgu[slot][q][2 * cc], guy = sgu[slot][q][2 * cc + 1];
              double t = -nu * (gux * gx + guy * gy);
              t -= (suq[slot][q][0] * gux + suq[slot][q][1] * guy) * n;
              t += spq[slot][q] * (cc == 0 ? gx : gy);
              t -= sdu[slot][q][cc] * n * inv_dt;
              r += t * w;
            }
            A.res[sdof[slot][sldv[a][cc]]] += r;
          } else {
            const int m = e - NVN * 2;
#pragma unroll
            for (int q = 0; q < NQ; ++q) r += (sgu[slot][q][0] + sgu[slot][q][3]) * sNp[q][m] * sw[slot][q];
            A.res[sdof[slot][sldp[m]]] += r;
          }
        }
    }
    __syncthreads();
  }
}

__global__ void k_outlet(int64_t n, const uint32_t *dof, const double *unit, double p_out, double *res) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) res[dof[i]] += p_out * unit[i];
}

// first row of each owned velocity range whose diagonal entry is non-zero
__global__ void k_bc_first_diag(int nranks, const int64_t *owned, const int64_t *rp, const int32_t *diag, const double *val,
                                unsigned long long *first) {
  const int r = blockIdx.y;
  const int64_t lo = owned[r], hi = owned[r + 1];
  for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    if ((unsigned long long)i >= first[r]) return;  // monotone: nothing earlier can come from this thread
    if (diag[i] >= 0 && val[rp[i] + diag[i]] != 0.0) { atomicMin(&first[r], (unsigned long long)i); return; }
  }
}

// MatrixTools::apply_boundary_values(bv, J, delta, r, false) -- NSSolverStationary.cpp:574-575
__global__ void k_apply_bc(int64_t nbc, const uint32_t *bc_dof, const double *bc_val, int apply_inlet, int nranks, const int64_t *owned,
                           const unsigned long long *first, const int64_t *F_rp, const int32_t *F_diag, double *F_val,
                           const int64_t *Bt_rp, double *Bt_val, double *delta, double *res) {
  const int G = 8;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  if (b >= nbc) return;
  const int64_t i = bc_dof[b];
  int r = 0;
  while (r + 1 < nranks && i >= owned[r + 1]) ++r;
  double fd = 1.0;
  const unsigned long long fi = first[r];
  if (fi < (unsigned long long)owned[r + 1]) fd = fabs(F_val[F_rp[fi] + F_diag[fi]]);
  const double v = apply_inlet ? bc_val[b] : 0.0;
  const int64_t rb = F_rp[i], re = F_rp[i + 1], dp = rb + F_diag[i];
  double dgl = F_val[dp];
  if (dgl == 0.0) dgl = fd;
  __syncwarp();
  for (int64_t k = rb + lane; k < re; k += G) F_val[k] = (k == dp) ? dgl : 0.0;
  for (int64_t k = Bt_rp[i] + lane; k < Bt_rp[i + 1]; k += G) Bt_val[k] = 0.0;
  if (lane == 0) { delta[i] = v; res[i] = v * dgl; }
}

// compute_lift_drag (NSSolverStationary.cpp:835-892): one thread per boundary-10 face
__global__ void k_lift_drag(int64_t nfaces, const int32_t *fcell, const int32_t *fface, const FETables *fe, const double *cell_vertices,
                            const uint32_t *cell_dofs, const int32_t *vmap, const double *sol, double nu, double *force) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nfaces) return;
  const FETables &T = *fe;
  const int cell = fcell[k], f = fface[k], nd = T.ndofs;
  const double *xv = cell_vertices + (int64_t)cell * T.nvpc * 2;
  const uint32_t *dofs = cell_dofs + (int64_t)cell * nd;
  double nx, ny, len;
  face_geometry(T.elem, xv, f, nx, ny, len);
  double drag = 0, lift = 0;
  for (int q = 0; q < T.nqf; ++q) {
    double x, y, J[2][2];
    face_point(T.elem, f, T.qpf[q], x, y);
    jacobian_at(T.elem, xv, x, y, J);
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double Ji[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
    double g[2][2] = {{0, 0}, {0, 0}}, p = 0;
    for (int i = 0; i < nd; ++i) {
      const double s = sol[vmap[dofs[i]]];
      const int comp = T.dof_comp[i], node = T.dof_node[i];
      if (comp == 2) p += s * T.Npf[f][node][q];
      else {
        const double g0 = T.dNvf[f][node][q][0], g1 = T.dNvf[f][node][q][1];
        g[comp][0] += s * (Ji[0][0] * g0 + Ji[1][0] * g1);
        g[comp][1] += s * (Ji[0][1] * g0 + Ji[1][1] * g1);
      }
    }
    double st[2][2];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) st[a][b] = (g[a][b] + g[b][a]) * nu;
    st[0][0] -= p; st[1][1] -= p;
    const double w = T.qwf[q] * len;
    drag += (-st[0][0] * nx - st[0][1] * ny) * w;
    lift += (-st[1][0] * nx - st[1][1] * ny) * w;
  }
  force[2 * k] = drag; force[2 * k + 1] = lift;
}

}  // namespace

void assemble_cells(Ctx &c, int mode, double nu, double dt, double p_out, bool res_only) {
  if (!res_only) { c.F.val.zero(c.stream); c.Bt.val.zero(c.stream); c.B.val.zero(c.stream); c.Mp.val.zero(c.stream); }
  c.vec[NSX_VEC_RESIDUAL].zero(c.stream);
  AsmArgs A;
  A.mode = mode; A.res_only = res_only ? 1 : 0; A.nu = nu; A.inv_dt = (dt != 0.0) ? 1.0 / dt : 0.0;
  A.sol = c.vec[NSX_VEC_SOLUTION].p; A.sol_old = c.vec[NSX_VEC_SOLUTION_OLD].p; A.res = c.vec[NSX_VEC_RESIDUAL].p;
  A.F_rp = c.F.rowptr.p; A.Bt_rp = c.Bt.rowptr.p; A.B_rp = c.B.rowptr.p; A.Mp_rp = c.Mp.rowptr.p;
  A.F_val = c.F.val.p; A.Bt_val = c.Bt.val.p; A.B_val = c.B.val.p; A.Mp_val = c.Mp.val.p;
  A.cell_vertices = c.cell_vertices.p; A.cell_dofs = c.cell_dofs.p; A.cell_pat = c.cell_pat.p; A.pat_off = c.pat_off.p;
  A.n_u_loc = c.n_u + c.n_ug; A.n_u_own = c.n_u; A.n_p_own = c.n_p; A.vmap = c.vmap.p; A.fe = c.d_fe.p;
  A.bt_rows = c.Bt.nrows_ext ? c.Bt.nrows_ext : c.n_u;
  A.F_sink = c.F.nnz + 12; A.Bt_sink = (c.Bt.nnz_ext ? c.Bt.nnz_ext : c.Bt.nnz) + 12; A.B_sink = c.B.nnz + 12; A.Mp_sink = c.Mp.nnz + 12;
  // ghost import of the state the cells read (`solution = solution_owned`, NSSolverStationary.cpp:722)
  halo_exchange(c, 0, A.sol); halo_exchange(c, 1, A.sol + c.n_u);
  if (mode >= NSX_MODE_UNSTEADY_FIRST) { halo_exchange(c, 0, A.sol_old); halo_exchange(c, 1, A.sol_old + c.n_u); }
  // the Stokes-type branches skip the residual body (NSSolverStationary.cpp:455-458): nothing for the cells to do
  const bool cells_idle = res_only && (mode == NSX_MODE_STOKES || mode == NSX_MODE_UNSTEADY_FIRST);
  for (int col = 0; col < c.ncolors && !cells_idle; ++col) {
    const int64_t lo = c.color_ptr[col], hi = c.color_ptr[col + 1];
    if (hi == lo) continue;
    A.cells = c.color_cells.p + lo; A.ncells = (int)(hi - lo);
    if (c.fe.elem == 0) {
      constexpr int CPB = 2;
      const int grid = (int)std::min<int64_t>((A.ncells + CPB - 1) / CPB, (int64_t)c.num_sms * 16);
      k_assemble<0, 4, 16, 9, 16, 128, CPB><<<grid, 128 * CPB, 0, c.stream>>>(A);
    } else {
      constexpr int CPB = 4;
      const int grid = (int)std::min<int64_t>((A.ncells + CPB - 1) / CPB, (int64_t)c.num_sms * 16);
      k_assemble<1, 3, 6, 3, 7, 32, CPB><<<grid, 32 * CPB, 0, c.stream>>>(A);
    }
    c.stat_launches++;
  }
  if (c.n_outlet) {
    k_outlet<<<(int)((c.n_outlet + 127) / 128), 128, 0, c.stream>>>(c.n_outlet, c.outlet_dof.p, c.outlet_unit.p, p_out, A.res);
    c.stat_launches++;
  }
  NSX_CUDA(cudaGetLastError());
}

namespace {
__global__ void k_clear_rows(int64_t n, const int32_t *rows, const int64_t *rp, double *val) {
  const int G = 8;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  if (b >= n) return;
  const int64_t r = rows[b];
  for (int64_t k = rp[r] + threadIdx.x % G; k < rp[r + 1]; k += G) val[k] = 0.0;
}
}  // namespace

void apply_boundary_values(Ctx &c, bool apply_inlet) {
  if (c.n_ghost_bc) {   // the ghost copies of constrained Bt rows follow their owners' (cleared) rows
    k_clear_rows<<<(int)((c.n_ghost_bc * 8 + 255) / 256), 256, 0, c.stream>>>(c.n_ghost_bc, c.ghost_bc.p, c.Bt.rowptr.p, c.Bt.val.p);
    c.stat_launches++;
  }
  if (!c.nbc) return;
  const int nr = (int)c.owned_u.size() - 1;
  NSX_CUDA(cudaMemsetAsync(c.bc_first.p, 0xff, c.bc_first.n * sizeof(unsigned long long), c.stream));
  k_bc_first_diag<<<dim3(8, nr), 256, 0, c.stream>>>(nr, c.d_owned_u.p, c.F.rowptr.p, c.F.diag.p, c.F.val.p, c.bc_first.p);
  k_apply_bc<<<(int)((c.nbc * 8 + 255) / 256), 256, 0, c.stream>>>(c.nbc, c.bc_dof.p, c.bc_val.p, apply_inlet ? 1 : 0, nr, c.d_owned_u.p, c.bc_first.p,
                                                                   c.F.rowptr.p, c.F.diag.p, c.F.val.p, c.Bt.rowptr.p, c.Bt.val.p,
                                                                   c.vec[NSX_VEC_DELTA].p, c.vec[NSX_VEC_RESIDUAL].p);
  c.stat_launches += 2;
  NSX_CUDA(cudaGetLastError());
}

void assemble(Ctx &c, int mode, bool apply_inlet, double nu, double dt, double p_out) {
  assemble_cells(c, mode, nu, dt, p_out, false);
  apply_boundary_values(c, apply_inlet);
}

namespace {
// what apply_boundary_values leaves in the vectors for homogeneous values: delta[i] = 0, r[i] = 0 * diagonal
__global__ void k_bc_vectors(int64_t nbc, const uint32_t *bc_dof, double *delta, double *res) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b < nbc) { delta[bc_dof[b]] = 0.0; res[bc_dof[b]] = 0.0; }
}
}  // namespace

// The residual vector a full assembly with homogeneous Dirichlet values would leave, bit for bit (same kernel, same colour
// order), without touching J or Mp: what the line search needs (NSSolverStationary.cpp:724-729 re-assembles everything for
// a norm; the matrices of those assemblies are never used -- the Newton loop assembles again before the next solve).
void assemble_residual(Ctx &c, int mode, double nu, double dt, double p_out) {
  assemble_cells(c, mode, nu, dt, p_out, true);
  if (c.nbc) {
    k_bc_vectors<<<(int)((c.nbc + 255) / 256), 256, 0, c.stream>>>(c.nbc, c.bc_dof.p, c.vec[NSX_VEC_DELTA].p, c.vec[NSX_VEC_RESIDUAL].p);
    c.stat_launches++;
  }
}

void lift_drag(Ctx &c, double nu, double *drag, double *lift) {
  const int64_t nf = (int64_t)c.h_cyl_cell.size();
  *drag = 0; *lift = 0;
  double *sol = c.vec[NSX_VEC_SOLUTION].p;
  halo_exchange(c, 0, sol); halo_exchange(c, 1, sol + c.n_u);
  if (nf) {
    k_lift_drag<<<(int)((nf + 63) / 64), 64, 0, c.stream>>>(nf, c.cyl_cell.p, c.cyl_face.p, c.d_fe.p, c.cell_vertices.p, c.cell_dofs.p, c.vmap.p,
                                                           sol, nu, c.face_force.p);
    c.stat_launches++;
    std::vector<double> h(2 * nf);
    NSX_CUDA(cudaMemcpyAsync(h.data(), c.face_force.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
    for (int64_t k = 0; k < nf; ++k) { *drag += h[2 * k]; *lift += h[2 * k + 1]; }  // face order, as the reference's loop
  }
  if (c.comm) {  // Utilities::MPI::sum of the two forces (NSSolverStationary.cpp:894-895)
    const double loc[2] = {*drag, *lift};
    vec_dot_dev(c, RED_SLOTS - 1, sol, sol, 0);  // makes sure the slot buffer exists (no launch for n = 0)
    double *slots = slot_ptr(c, RED_SLOTS - 4);
    NSX_CUDA(cudaMemcpyAsync(slots, loc, sizeof(loc), cudaMemcpyHostToDevice, c.stream));
    allreduce_slots(c, RED_SLOTS - 4, 2);
    double out[2];
    read_slots(c, RED_SLOTS - 4, 2, out);
    *drag = out[0]; *lift = out[1];
  }
}

}  // namespace nsx
