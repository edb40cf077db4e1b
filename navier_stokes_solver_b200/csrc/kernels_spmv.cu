// kernels_spmv.cu -- FP64 CSR SpMV over the Jacobian blocks (HBM-bound).
//
// Replaces TrilinosWrappers::BlockSparseMatrix::vmult / SparseMatrix::vmult / vmult_add as called
// inside every Krylov solver of the reference (NSSolverStationary.cpp:589-637,
// NSSolverStationary.hpp:140, 198, 207, 288, 292, 305; NSSolver.hpp:226, 324, 345).
//
// Layout: one CSR per block (F n_u x n_u, Bt n_u x n_p, B n_p x n_u, Mp, S), 64-bit row
// pointers, 32-bit block-local columns, FP64 values.  A sub-warp of G lanes owns a row, so the
// lanes of a warp stream 32/G consecutive rows = one contiguous span of the value / column
// arrays (coalesced, streamed with ld.global.cs so that x stays cached), x is gathered through
// the read-only path, and the row sum is a log2(G) shuffle reduction (deterministic order).
// Algorithmic bytes per product: 12 nnz + 8 (m+1) + 8 n + 8 m.
#include "device.cuh"

namespace nsx {

namespace {

template <int G>
__device__ __forceinline__ double row_dot(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                          const double *__restrict__ val, const double *__restrict__ x, int64_t row, int lane) {
  const int64_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  double s0 = 0, s1 = 0;
  int64_t k = b + lane;
  for (; k + G < e; k += 2 * G) {
    const int32_t c0 = __ldcs(col + k), c1 = __ldcs(col + k + G);
    const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + G);
    s0 += v0 * __ldg(x + c0);
    s1 += v1 * __ldg(x + c1);
  }
  if (k < e) s0 += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
  return s0 + s1;
}

template <int G>
__device__ __forceinline__ double group_sum(double s) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
  return s;
}

template <int G>
__global__ void __launch_bounds__(256) k_spmv(int64_t nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                              const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y, int add) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const bool valid = row < nrows;  // no early exit: the shuffles below use the full-warp mask
  double s = valid ? row_dot<G>(rowptr, col, val, x, row, lane) : 0.0;
  s = group_sum<G>(s);
  if (valid && lane == 0) y[row] = add ? y[row] + s : s;
}

// y_u = F x_u + Bt x_p ; y_p = B x_u in one launch (jacobian_matrix.vmult)
template <int G>
__global__ void __launch_bounds__(256) k_block_spmv(int64_t n_u, int64_t n_p,
                                                    const int64_t *__restrict__ f_rp, const int32_t *__restrict__ f_col, const double *__restrict__ f_val,
                                                    const int64_t *__restrict__ bt_rp, const int32_t *__restrict__ bt_col, const double *__restrict__ bt_val,
                                                    const int64_t *__restrict__ b_rp, const int32_t *__restrict__ b_col, const double *__restrict__ b_val,
                                                    const double *__restrict__ x, double *__restrict__ y) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
  const bool valid = row < n_u + n_p;
  double s = 0.0;
  if (!valid) {
  } else if (row < n_u) {
    s = row_dot<G>(f_rp, f_col, f_val, x, row, lane);
    s += row_dot<G>(bt_rp, bt_col, bt_val, x + n_u, row, lane);
  } else {
    s = row_dot<G>(b_rp, b_col, b_val, x, row - n_u, lane);
  }
  s = group_sum<G>(s);
  if (valid && lane == 0) y[row] = s;
}

__global__ void k_extract_diag(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ diag,
                               const double *__restrict__ val, double *d, double *dinv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = diag[i] >= 0 ? val[rowptr[i] + diag[i]] : 0.0;
  if (d) d[i] = v;
  if (dinv) dinv[i] = 1.0 / v;
}

inline int pick_group(const DevCSR &A) {
  const double avg = A.nrows ? (double)A.nnz / (double)A.nrows : 0.0;
  if (avg > 48) return 16;
  if (avg > 12) return 8;
  return 4;
}

}  // namespace

void spmv(Ctx &c, const DevCSR &A, const double *x, double *y, bool add) {
  if (!A.nrows) return;
  const int G = pick_group(A);
  const int64_t threads = A.nrows * G;
  const int grid = (int)((threads + 255) / 256);
  if (G == 16) k_spmv<16><<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, add ? 1 : 0);
  else if (G == 8) k_spmv<8><<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, add ? 1 : 0);
  else k_spmv<4><<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, add ? 1 : 0);
  c.stat_launches++; c.stat_spmv++;
}

void block_spmv(Ctx &c, const double *x, double *y) {
  const int64_t n = c.n_u + c.n_p;
  const double avg = (double)(c.F.nnz + c.Bt.nnz + c.B.nnz) / (double)n;
  if (avg > 40) {
    const int grid = (int)((n * 16 + 255) / 256);
    k_block_spmv<16><<<grid, 256, 0, c.stream>>>(c.n_u, c.n_p, c.F.rowptr.p, c.F.col.p, c.F.val.p, c.Bt.rowptr.p, c.Bt.col.p, c.Bt.val.p,
                                                  c.B.rowptr.p, c.B.col.p, c.B.val.p, x, y);
  } else {
    const int grid = (int)((n * 8 + 255) / 256);
    k_block_spmv<8><<<grid, 256, 0, c.stream>>>(c.n_u, c.n_p, c.F.rowptr.p, c.F.col.p, c.F.val.p, c.Bt.rowptr.p, c.Bt.col.p, c.Bt.val.p,
                                                 c.B.rowptr.p, c.B.col.p, c.B.val.p, x, y);
  }
  c.stat_launches++; c.stat_spmv++;
}

void extract_diag(Ctx &c, const DevCSR &A, double *d, double *dinv) {
  const int grid = (int)((A.nrows + 255) / 256);
  k_extract_diag<<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.diag.p, A.val.p, d, dinv);
  c.stat_launches++;
}

}  // namespace nsx
