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
  c.cross_count.alloc(1);
  NSX_CUDA(cudaStreamSynchronize(c.stream));
}

bool decoupled_ok(Ctx &c) {
  if (!c.decouple || !c.F.nnz) return false;
  if (c.dec_epoch == c.matrix_epoch) return c.dec_ok;
  if (!c.F_cross.p) build_decoupled(c);
  NSX_CUDA(cudaMemsetAsync(c.cross_count.p, 0, sizeof(unsigned long long), c.stream));
  k_count_cross<<<c.num_sms * 8, 256, 0, c.stream>>>(c.F.nnz, c.F_cross.p, c.F.val.p, c.cross_count.p);
  c.stat_launches++;
  unsigned long long local = 0;
  NSX_CUDA(cudaMemcpyAsync(&local, c.cross_count.p, sizeof(local), cudaMemcpyDeviceToHost, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  // every rank must take the same branch: the verdict is the sum over ranks
  double total = (double)local;
  if (c.comm) {
    double *slot = slot_ptr(c, RED_SLOTS - 3);
    NSX_CUDA(cudaMemcpyAsync(slot, &total, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    allreduce_slots(c, RED_SLOTS - 3, 1);
    total = read_slot(c, RED_SLOTS - 3);
  }
  c.dec_ok = total == 0.0;
  c.dec_epoch = c.matrix_epoch;
  if (c.dec_ok) gather_values(c, c.Fd.nnz, c.Fd_src.p, c.F.val.p, c.Fd.val.p);
  return c.dec_ok;
}

}  // namespace nsx
