// decouple.cu -- component-decoupled view of the velocity block F.
//
// In the Stokes-type assembly branches (NSSolverStationary.cpp:383-406 `first or stokes`, NSSolver.cpp:381-409 `first_iter`) the
// velocity block is nu * (grad phi_i : grad phi_j): shape functions of different components have disjoint gradients, so every
// coupling between u_x and u_y dofs is an exact zero that the FESystem sparsity pattern nevertheless stores (half of F's entries).
// The reference multiplies those zeros in every SpMV and every SSOR sweep.  Here the library checks the current values on the
// device after each assembly (k_count_cross: any non-zero cross-component entry?) and, when there is none, runs the inner solves'
// F products and the Gauss-Seidel sweeps on the same-component entries only: Fd (a compacted CSR whose values are gathered from F)
// and plan variant 1 of block F.  Dropping an exact zero changes no sum, so results are the ones of the full matrix (up to the
// order of additions); ILU(0) keeps the full pattern because fill lands on those positions.  In the Newton branches the check
// finds the convective couplings and everything runs on the full F.
#include <algorithm>

#include "device.cuh"

namespace nsx {

namespace {

__global__ void k_count_cross(int64_t nnz, const uint8_t *__restrict__ cross, const double *__restrict__ val, unsigned long long *count) {
  unsigned long long local = 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    if (cross[k] && val[k] != 0.0) ++local;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

}  // namespace

const std::vector<uint8_t> &velocity_components(Ctx &c) {
  const int64_t nloc = c.n_u + c.n_ug;
  if ((int64_t)c.h_comp_u.size() == nloc) return c.h_comp_u;
  c.h_comp_u.assign(nloc, 255);
  const int nd = c.fe.ndofs;
  for (int64_t cell = 0; cell < c.ncells; ++cell)
    for (int k = 0; k < nd; ++k) {
      const int comp = c.fe.dof_comp[k];
      if (comp > 1) continue;
      const int64_t d = c.h_cell_dofs[(size_t)cell * nd + k];
      if (d < nloc) c.h_comp_u[d] = (uint8_t)comp;
    }
  for (int64_t i = 0; i < nloc; ++i)
    if (c.h_comp_u[i] > 1) throw std::logic_error("a velocity dof belongs to no cell");
  return c.h_comp_u;
}

static void build_decoupled(Ctx &c) {
  const std::vector<uint8_t> &comp = velocity_components(c);
  const DevCSR &F = c.F;
  DevCSR &D = c.Fd;
  const int64_t n = F.nrows;
  std::vector<uint8_t> cross(F.nnz);
  D.nrows = n; D.ncols = F.ncols; D.row0 = F.row0;
  D.h_rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    int64_t cnt = 0;
    for (int64_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) {
      cross[k] = comp[F.h_col[k]] != comp[i];
      cnt += !cross[k];
    }
    D.h_rowptr[i + 1] = D.h_rowptr[i] + cnt;
  }
  D.nnz = D.h_rowptr[n];
  D.h_col.resize(D.nnz);
  std::vector<int64_t> src(D.nnz);
  std::vector<int32_t> baked(D.nnz);
  const int64_t own = c.n_u, shift = c.n_p;   // device columns follow the vector layout [u owned | p owned | u ghosts | p ghosts]
  D.max_row = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t o = D.h_rowptr[i];
    for (int64_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k)
      if (!cross[k]) {
        D.h_col[o] = F.h_col[k]; src[o] = k;
        baked[o] = (int32_t)(F.h_col[k] < own ? F.h_col[k] : F.h_col[k] + shift);
        ++o;
      }
    D.max_row = std::max<int>(D.max_row, (int)(D.h_rowptr[i + 1] - D.h_rowptr[i]));
  }
  D.rowptr.alloc_padded(D.h_rowptr.size(), 4, c.stream);
  NSX_CUDA(cudaMemcpyAsync(D.rowptr.p, D.h_rowptr.data(), D.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  D.col.alloc_padded(D.nnz, 16, c.stream);
  NSX_CUDA(cudaMemcpyAsync(D.col.p, baked.data(), D.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  D.val.alloc_padded(D.nnz, 16, c.stream);
  D.nrb = D.ndesc = 0; D.pair_state = -1;   // same-component rows have no (2k, 2k+1) column pairs
  c.Fd_src.upload(src, c.stream);
  c.F_cross.upload(cross, c.stream);
  c.cross_count.alloc(2);
  NSX_CUDA(cudaStreamSynchronize(c.stream));
}

// Structure of the node view: local velocity dofs (2a, 2a + 1) are the two components of node a, and the same-component entries
// of rows 2a and 2a + 1 sit at columns (2b, 2b + 1) pair by pair.  True for deal.II's FESystem(FE^2) numbering; verified here.
static void build_node_view(Ctx &c) {
  c.node_struct = -1;
  const std::vector<uint8_t> &comp = velocity_components(c);
  const DevCSR &F = c.F;
  const int64_t n = F.nrows, nloc = c.n_u + c.n_ug;
  if ((n & 1) || (nloc & 1) || (c.n_ug && (c.n_p & 1))) return;   // ghost pairs must stay 16-byte aligned behind the owned pressure entries
  for (size_t r = 0; r < c.owned_u.size(); ++r) if (c.owned_u[r] & 1) return;
  for (int64_t d = 0; d < nloc; ++d) if (comp[d] != (d & 1)) return;
  DevCSR &K = c.Kn;
  K.nrows = n / 2; K.ncols = nloc / 2; K.row0 = 0;
  K.h_rowptr.assign(K.nrows + 1, 0);
  std::vector<int64_t> sx, sy;
  std::vector<int32_t> baked;
  K.h_col.clear();
  const int64_t own = c.n_u, shift = c.n_p;
  K.max_row = 0;
  for (int64_t a = 0; a < K.nrows; ++a) {
    const int64_t bx = F.h_rowptr[2 * a], ex = F.h_rowptr[2 * a + 1], by = ex, ey = F.h_rowptr[2 * a + 2];
    int64_t ky = by;
    for (int64_t kx = bx; kx < ex; ++kx) {
      const int32_t cx = F.h_col[kx];
      if (comp[cx] != 0) continue;
      while (ky < ey && comp[F.h_col[ky]] != 1) ++ky;
      if (ky == ey || F.h_col[ky] != cx + 1) return;
      K.h_col.push_back(cx / 2);
      baked.push_back((int32_t)(cx < own ? cx : cx + shift));
      sx.push_back(kx); sy.push_back(ky);
      ++ky;
    }
    while (ky < ey && comp[F.h_col[ky]] != 1) ++ky;
    if (ky != ey) return;   // the y row has a same-component entry without an x twin
    K.h_rowptr[a + 1] = (int64_t)K.h_col.size();
    K.max_row = std::max<int>(K.max_row, (int)(K.h_rowptr[a + 1] - K.h_rowptr[a]));
  }
  K.nnz = K.h_rowptr[K.nrows];
  K.rowptr.alloc_padded(K.h_rowptr.size(), 4, c.stream);
  NSX_CUDA(cudaMemcpyAsync(K.rowptr.p, K.h_rowptr.data(), K.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  K.col.alloc_padded(K.nnz, 16, c.stream);
  NSX_CUDA(cudaMemcpyAsync(K.col.p, baked.data(), K.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  K.val.alloc_padded(K.nnz, 16, c.stream);
  K.nrb = K.ndesc = 0; K.pair_state = -1;
  c.Kn_src_x.upload(sx, c.stream);
  c.Kn_src_y.upload(sy, c.stream);
  c.h_Kn_src = sx;
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  c.node_struct = 1;
}

namespace {
__global__ void k_count_unequal(int64_t n, const int64_t *__restrict__ sx, const int64_t *__restrict__ sy, const double *__restrict__ val, unsigned long long *count) {
  unsigned long long local = 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    if (val[sx[k]] != val[sy[k]]) ++local;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}
}  // namespace

int stokes_view(Ctx &c) {
  if (!c.decouple || !c.F.nnz) return 0;
  if (c.dec_epoch == c.matrix_epoch) return c.dec_ok ? (c.node_ok ? 2 : 1) : 0;
  if (!c.F_cross.p) build_decoupled(c);
  if (c.node_struct == 0 && c.decouple_nodes) build_node_view(c);
  const bool try_nodes = c.node_struct == 1 && c.decouple_nodes;
  NSX_CUDA(cudaMemsetAsync(c.cross_count.p, 0, 2 * sizeof(unsigned long long), c.stream));
  k_count_cross<<<c.num_sms * 8, 256, 0, c.stream>>>(c.F.nnz, c.F_cross.p, c.F.val.p, c.cross_count.p);
  if (try_nodes) k_count_unequal<<<c.num_sms * 8, 256, 0, c.stream>>>(c.Kn.nnz, c.Kn_src_x.p, c.Kn_src_y.p, c.F.val.p, c.cross_count.p + 1);
  c.stat_launches += try_nodes ? 2 : 1;
  unsigned long long local[2] = {0, 0};
  NSX_CUDA(cudaMemcpyAsync(local, c.cross_count.p, sizeof(local), cudaMemcpyDeviceToHost, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  // every rank must take the same branch: the verdicts are sums over ranks (a rank whose numbering does not pair up vetoes the node view)
  double total[2] = {(double)local[0], try_nodes ? (double)local[1] : 1.0};
  if (c.comm) {
    double *slot = slot_ptr(c, RED_SLOTS - 4);
    NSX_CUDA(cudaMemcpyAsync(slot, total, sizeof(total), cudaMemcpyHostToDevice, c.stream));
    allreduce_slots(c, RED_SLOTS - 4, 2);
    read_slots(c, RED_SLOTS - 4, 2, total);
  }
  c.dec_ok = total[0] == 0.0;
  c.node_ok = c.dec_ok && total[1] == 0.0;
  c.dec_epoch = c.matrix_epoch;
  if (c.node_ok) gather_values(c, c.Kn.nnz, c.Kn_src_x.p, c.F.val.p, c.Kn.val.p);
  if (c.dec_ok) gather_values(c, c.Fd.nnz, c.Fd_src.p, c.F.val.p, c.Fd.val.p);
  return c.dec_ok ? (c.node_ok ? 2 : 1) : 0;
}

int effective_view(Ctx &c) {
  int view = stokes_view(c);
  if (view == 2 && (c.ordering < 2 || c.stream_spmv != 3)) view = 1;   // the node view lives in the block-local sweeps and the direct SpMV
  return view;
}

}  // namespace nsx
