// spgemm.cu -- approximate Schur complement S = B diag(F)^-1 Bt of the aSIMPLE preconditioner.
//
// Replaces PreconditionaSIMPLE::initialize (NSSolverStationary.hpp:259-275, NSSolver.hpp:275-286):
// diag_element() loop -> D, D^-1; B.mmult(S, Bt, D^-1) (EpetraExt MatrixMatrix).  The sparsity of
// S depends only on the patterns of B and Bt, so the symbolic product is done once on the host;
// every call then recomputes the values on the device: one warp per row of S, the row accumulated
// in shared memory, contributions added in the (k, l) order of the row-by-row product.
#include <algorithm>

#include "device.cuh"

namespace nsx {

namespace {

// n_u_own / n_p_own: owned counts.  Device columns follow the vector layout (ghost velocity ids shifted by n_p_own, ghost pressure
// ids by the ghost velocity count): dinv is a full-layout vector, Bt's rows are indexed by the local velocity id, and only the
// owned pressure columns are kept (S is the rank-local block: what ILU(0) factors; the product S x of the CG solve is formed
// from B, diag(F)^-1 and Bt on a partitioned system, krylov.cu)
__global__ void __launch_bounds__(256) k_schur(int64_t n_p, const int64_t *__restrict__ B_rp, const int32_t *__restrict__ B_col,
                                               const double *__restrict__ B_val, const double *__restrict__ dinv,
                                               const int64_t *__restrict__ Bt_rp, const int32_t *__restrict__ Bt_col,
                                               const double *__restrict__ Bt_val, const int64_t *__restrict__ S_rp,
                                               const int32_t *__restrict__ S_col, double *__restrict__ S_val, int maxrow, int64_t n_u_own, int64_t n_p_own) {
  extern __shared__ double s_acc[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
  if (row >= n_p) return;
  double *acc = s_acc + (size_t)wib * maxrow;
  const int64_t sb = S_rp[row];
  const int sl = (int)(S_rp[row + 1] - sb);
  for (int t = lane; t < sl; t += 32) acc[t] = 0.0;
  __syncwarp();
  for (int64_t k = B_rp[row]; k < B_rp[row + 1]; ++k) {
    const int32_t mb = B_col[k];
    const double a = B_val[k] * dinv[mb];
    const int64_t m = mb < n_u_own ? mb : mb - n_p_own;
    for (int64_t l = Bt_rp[m] + lane; l < Bt_rp[m + 1]; l += 32) {
      const int32_t cj = Bt_col[l];
      if (cj >= n_p_own) continue;
      int lo = 0, hi = sl;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S_col[sb + mid] < cj) lo = mid + 1; else hi = mid;
      }
      acc[lo] += a * Bt_val[l];
    }
    __syncwarp();
  }
  for (int t = lane; t < sl; t += 32) S_val[sb + t] = acc[t];
}

}  // namespace

void schur_symbolic(Ctx &c) {
  if (c.S_symbolic) return;
  if (c.n_ug && c.Bt.nrows_ext != c.n_u + c.n_ug) throw std::logic_error("the ghost rows of Bt are built by nsx_finalize_setup");
  const DevCSR &B = c.B, &Bt = c.Bt;
  DevCSR &S = c.S;
  const int64_t n = c.n_p;
  S.nrows = S.ncols = n;
  std::vector<std::vector<int32_t>> rows(n);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
      tmp.clear();
      for (int64_t k = B.h_rowptr[i]; k < B.h_rowptr[i + 1]; ++k) {
        const int64_t m = B.h_col[k];   // local velocity id, owned or ghost: Bt holds both kinds of rows
        for (int64_t l = Bt.h_rowptr[m]; l < Bt.h_rowptr[m + 1]; ++l)
          if (Bt.h_col[l] < n) tmp.push_back(Bt.h_col[l]);   // owned pressure columns: the rank-local block
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      rows[i] = tmp;
    }
  }
  S.h_rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) S.h_rowptr[i + 1] = S.h_rowptr[i] + (int64_t)rows[i].size();
  S.nnz = S.h_rowptr[n];
  S.nrb = S.ndesc = 0;
  S.h_col.resize(S.nnz);
  std::vector<int32_t> diag(n, -1);
  S.max_row = 0;
  for (int64_t i = 0; i < n; ++i) {
    std::copy(rows[i].begin(), rows[i].end(), S.h_col.begin() + S.h_rowptr[i]);
    S.max_row = std::max<int>(S.max_row, (int)rows[i].size());
    auto it = std::lower_bound(rows[i].begin(), rows[i].end(), (int32_t)i);
    if (it != rows[i].end() && *it == i) diag[i] = (int32_t)(it - rows[i].begin());
  }
  S.rowptr.alloc_padded(S.h_rowptr.size(), 4, c.stream);
  NSX_CUDA(cudaMemcpyAsync(S.rowptr.p, S.h_rowptr.data(), S.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  S.col.alloc_padded(S.nnz, 16, c.stream);
  NSX_CUDA(cudaMemcpyAsync(S.col.p, S.h_col.data(), S.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  S.diag.upload(diag, c.stream);
  S.val.alloc_padded(S.nnz, 16, c.stream);
  c.Dvec.alloc(c.nvec);   // full vector layout: the ghost entries of diag(F)^-1 arrive through the ghost import
  c.Dinv.alloc(c.nvec);
  c.Dvec.zero(c.stream); c.Dinv.zero(c.stream);
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  tri_erase(c, NSX_BLOCK_S);
  c.S_symbolic = true;
}

void schur_complement(Ctx &c) {
  schur_symbolic(c);
  extract_diag(c, c.F, c.Dvec.p, c.Dinv.p);
  halo_exchange(c, 0, c.Dinv.p);
  const int wpb = 8;
  const size_t smem = (size_t)wpb * c.S.max_row * sizeof(double);
  if (smem > 48 * 1024) throw std::runtime_error("Schur complement row too long for the shared-memory accumulator");
  k_schur<<<(int)((c.n_p + wpb - 1) / wpb), wpb * 32, smem, c.stream>>>(c.n_p, c.B.rowptr.p, c.B.col.p, c.B.val.p, c.Dinv.p, c.Bt.rowptr.p,
                                                                       c.Bt.col.p, c.Bt.val.p, c.S.rowptr.p, c.S.col.p, c.S.val.p, c.S.max_row, c.n_u, c.n_p);
  c.stat_launches++;
  NSX_CUDA(cudaGetLastError());
}

}  // namespace nsx
