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

// Structure of the node view.  A velocity node a carries two dofs, dx(a) and dy(a) (found through the cell table: the two components
// of a cell-local node); deal.II numbers them next to each other for vertices but component-wise inside lines and cells, so the
// pairs are not adjacent in general.  Nodes are numbered by ascending dx.  Row dx(a) of F restricted to x columns and row dy(a)
// restricted to y columns must hit the same nodes: then K(a, b) = F(dx a, dx b) = F(dy a, dy b) is one scalar matrix (if the values
// agree, checked per assembly), kept as Kn with columns in the NODE LAYOUT of a vector: entry 2b = x component of node b, 2b + 1 = y.
// On a partitioned system only the owned entries of a vector are permuted; its ghost tail keeps the reference order, a ghost node's
// pair is addressed through a small table (negative column ids), and the ghost import packs from node-layout positions.
static void build_node_view(Ctx &c) {
  c.node_struct = -1;
  if (c.n_u & 1) return;
  const std::vector<uint8_t> &comp = velocity_components(c);
  const DevCSR &F = c.F;
  const int64_t n = F.nrows, nloc = c.n_u + c.n_ug;
  const FETables &T = c.fe;
  const int nd = T.ndofs;
  std::vector<int32_t> partner(nloc, -1);
  {
    int ldv[MAX_VN][2];
    for (int i = 0; i < nd; ++i) if (T.dof_comp[i] < 2) ldv[T.dof_node[i]][T.dof_comp[i]] = i;
    for (int64_t cell = 0; cell < c.ncells; ++cell)
      for (int a = 0; a < T.nvn; ++a) {
        const int64_t dx = c.h_cell_dofs[(size_t)cell * nd + ldv[a][0]], dy = c.h_cell_dofs[(size_t)cell * nd + ldv[a][1]];
        if (dx >= nloc || dy >= nloc || (dx < n) != (dy < n)) return;   // both components of a node have the same owner
        if ((partner[dx] >= 0 && partner[dx] != dy) || (partner[dy] >= 0 && partner[dy] != dx)) return;
        partner[dx] = (int32_t)dy; partner[dy] = (int32_t)dx;
      }
  }
  // owned nodes first (ascending x dof), then the ghost nodes (ascending x dof): ghost node g has column id nn + g on the host
  std::vector<int32_t> node_of(nloc, -1);
  c.h_node_dx.clear(); c.h_node_dy.clear();
  std::vector<int2> gpair;
  for (int64_t d = 0; d < nloc; ++d) {
    if (partner[d] < 0) return;
    if (comp[d] != 0) continue;
    if (d < n) { node_of[d] = (int32_t)c.h_node_dx.size(); c.h_node_dx.push_back((int32_t)d); c.h_node_dy.push_back(partner[d]); }
  }
  const int64_t nn = (int64_t)c.h_node_dx.size();
  if (2 * nn != n) return;
  for (int64_t d = n; d < nloc; ++d)
    if (comp[d] == 0) {   // a ghost pair sits in the ghost tail of a vector, behind the owned pressure entries
      node_of[d] = (int32_t)(nn + (int64_t)gpair.size());
      gpair.push_back(make_int2((int)(d + c.n_p), (int)(partner[d] + c.n_p)));
    }
  // preconditioner ranges (nsx_set_ranks) must hold whole nodes
  c.owned_nodes.assign(c.owned_u.size(), 0);
  for (size_t r = 0; r + 1 < c.owned_u.size(); ++r) {
    int64_t cnt = 0;
    for (int64_t d = c.owned_u[r]; d < c.owned_u[r + 1]; ++d) {
      if (partner[d] < c.owned_u[r] || partner[d] >= c.owned_u[r + 1]) return;
      cnt += comp[d] == 0;
    }
    c.owned_nodes[r + 1] = c.owned_nodes[r] + cnt;
  }
  DevCSR &K = c.Kn;
  K.nrows = nn; K.ncols = nn + (int64_t)gpair.size(); K.row0 = 0;
  K.h_rowptr.assign(nn + 1, 0);
  K.h_col.clear();
  std::vector<int64_t> sx, sy;
  std::vector<int32_t> dev_col;
  K.max_row = 0;
  for (int64_t a = 0; a < nn; ++a) {
    const int64_t rx = c.h_node_dx[a], ry = c.h_node_dy[a];
    const int32_t *yb = F.h_col.data() + F.h_rowptr[ry], *ye = F.h_col.data() + F.h_rowptr[ry + 1];
    int64_t ny = 0;
    for (const int32_t *q = yb; q < ye; ++q) ny += comp[*q] == 1;
    for (int64_t kx = F.h_rowptr[rx]; kx < F.h_rowptr[rx + 1]; ++kx) {
      const int32_t cx = F.h_col[kx];
      if (comp[cx] != 0) continue;
      const int32_t cy = partner[cx];
      const int32_t *it = std::lower_bound(yb, ye, cy);
      if (it == ye || *it != cy) return;
      const int32_t b = node_of[cx];
      K.h_col.push_back(b);
      dev_col.push_back(b < nn ? 2 * b : -(int32_t)(b - nn) - 1);   // owned: position of the pair in the node layout; ghost: -(pair id) - 1
      sx.push_back(kx); sy.push_back(F.h_rowptr[ry] + (it - yb));
      --ny;
    }
    if (ny != 0) return;   // the y row has a same-component entry without an x twin
    K.h_rowptr[a + 1] = (int64_t)K.h_col.size();
    K.max_row = std::max<int>(K.max_row, (int)(K.h_rowptr[a + 1] - K.h_rowptr[a]));
  }
  K.nnz = K.h_rowptr[nn];
  K.rowptr.alloc_padded(K.h_rowptr.size(), 4, c.stream);
  NSX_CUDA(cudaMemcpyAsync(K.rowptr.p, K.h_rowptr.data(), K.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  K.col.alloc_padded(K.nnz, 16, c.stream);
  NSX_CUDA(cudaMemcpyAsync(K.col.p, dev_col.data(), K.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  K.val.alloc_padded(K.nnz, 16, c.stream);
  K.nrb = K.ndesc = 0; K.pair_state = -1;
  c.Kn_src_x.upload(sx, c.stream);
  c.Kn_src_y.upload(sy, c.stream);
  c.h_Kn_src = sx;
  c.node_dx.upload(c.h_node_dx, c.stream);
  c.node_dy.upload(c.h_node_dy, c.stream);
  c.node_gpair.upload(gpair, c.stream);
  // ghost import of a vector in the node layout: the owned entries a neighbour needs, at their node-layout positions
  if (c.halo_u.nsend) {
    std::vector<int32_t> ref(c.halo_u.nsend), pos(c.halo_u.nsend);
    NSX_CUDA(cudaMemcpyAsync(ref.data(), c.halo_u.send_idx.p, ref.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
    for (size_t k = 0; k < ref.size(); ++k) {
      const int32_t d = ref[k];
      pos[k] = comp[d] == 0 ? 2 * node_of[d] : 2 * node_of[partner[d]] + 1;
    }
    c.halo_u.send_idx_node.upload(pos, c.stream);
  }
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

namespace {
__global__ void k_to_node_layout(int64_t nn, const int32_t *__restrict__ dx, const int32_t *__restrict__ dy, const double *__restrict__ ref, double2 *__restrict__ node) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nn; a += (int64_t)gridDim.x * blockDim.x) node[a] = make_double2(ref[dx[a]], ref[dy[a]]);
}
__global__ void k_from_node_layout(int64_t nn, const int32_t *__restrict__ dx, const int32_t *__restrict__ dy, const double2 *__restrict__ node, double *__restrict__ ref) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nn; a += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = node[a];
    ref[dx[a]] = v.x; ref[dy[a]] = v.y;
  }
}
}  // namespace

void vec_to_node_layout(Ctx &c, double *node, const double *ref) {
  const int64_t nn = c.Kn.nrows;
  k_to_node_layout<<<grid_for(nn, 256, c.num_sms * 8), 256, 0, c.stream>>>(nn, c.node_dx.p, c.node_dy.p, ref, reinterpret_cast<double2 *>(node));
  c.stat_launches++;
}
void vec_from_node_layout(Ctx &c, double *ref, const double *node) {
  const int64_t nn = c.Kn.nrows;
  k_from_node_layout<<<grid_for(nn, 256, c.num_sms * 8), 256, 0, c.stream>>>(nn, c.node_dx.p, c.node_dy.p, reinterpret_cast<const double2 *>(node), ref);
  c.stat_launches++;
}

int effective_view(Ctx &c) {
  int view = stokes_view(c);
  if (view == 2 && (c.ordering < 2 || c.stream_spmv != 3)) view = 1;   // the node view lives in the block-local sweeps and the direct SpMV
  return view;
}

}  // namespace nsx
