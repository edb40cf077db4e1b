// amg.cu -- smoothed-aggregation AMG on the velocity block: set-up and V-cycle, all on the device.
//
// Replaces TrilinosWrappers::PreconditionAMG (Trilinos ML through deal.II's default AdditionalData:
// elliptic, one V-cycle, aggregation threshold 1e-4, Chebyshev smoother of degree 2 before and after,
// direct coarse solve, scalar constant near-null space), NSSolverStationary.hpp:184-185, 225, which the
// stationary blockTriangular preconditioner rebuilds in every solve_system (NSSolverStationary.cpp:601-604).
//
// Algorithm (the oracle restates the same one sequentially, oracle/oracle_amg.inc):
//   strength   j ~ i  iff  a_ij^2 > theta^2 |a_ii a_jj|  or  a_ji^2 > theta^2 |a_ii a_jj|
//   roots      distance-2 maximal independent set of the strength graph.  ML's greedy phase 1 (a node whose
//              whole neighbourhood is free becomes a root) builds the same kind of set in natural order; here
//              the order is a fixed pseudo-random priority, so that a Luby-type parallel sweep (a node joins
//              when it holds the largest priority among the undecided nodes within two edges) selects exactly
//              the set the sequential greedy pass selects.
//   aggregates neighbours of a root join it; the rest join the aggregate of their highest-priority aggregated
//              neighbour (ML phase 2).  Aggregates never cross an owned-range boundary (uncoupled).
//   P          (I - 4/3 / lambda_max D^-1 A) P_tent, P_tent(i, agg(i)) = 1;  R = P^T;  A_c = R A P
//   lambda_max of D^-1 A from 10 CG-Lanczos steps (ML "eigen-analysis: type" cg)
//   smoother   Chebyshev polynomial of degree 2 in D^-1 A on [lambda_max / 10, 1.1 lambda_max]
//   coarsest   <= 2000 rows (or level 10; deal.II sets "coarse: max size" = 2000 and "smoother: Chebyshev alpha" = 10 on top of ML's SA defaults): dense inverse by Gauss-Jordan with partial pivoting
// The sparse products are expand - sort - compress: every scalar product a_ik b_kj is written out with the key
// (i, j), a stable radix sort groups equal keys, and one thread per distinct key adds its run in sorted order
// (deterministic).  Level operators are plain CSR and go through the library's own SpMV kernels.
#include <cub/cub.cuh>

#include <cmath>

#include "device.cuh"

namespace nsx {

namespace {

constexpr double AMG_THRESHOLD = 1e-4, AMG_EIG_RATIO = 10.0, AMG_EIG_BOOST = 1.1;
constexpr int AMG_MAX_LEVELS = 10, AMG_COARSE_MAX = 2000, AMG_CHEBY_DEGREE = 2, AMG_DENSE_MAX = 4096;

__host__ __device__ inline uint32_t amg_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ inline unsigned long long amg_key(int64_t i) {
  return ((unsigned long long)amg_hash((uint32_t)i) << 32) | (uint32_t)(i + 1);
}

struct CsrView { int64_t n; const int64_t *rowptr; const int32_t *col; const double *val; };
inline CsrView view(const DevCSR &A) { return CsrView{A.nrows, A.rowptr.p, A.col.p, A.val.p}; }

__device__ inline double csr_entry(const CsrView &A, int64_t i, int32_t j) {
  int64_t lo = A.rowptr[i], hi = A.rowptr[i + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t cc = A.col[mid];
    if (cc < j) lo = mid + 1; else hi = mid;
  }
  return (lo < A.rowptr[i + 1] && A.col[lo] == j) ? A.val[lo] : 0.0;
}

__global__ void k_amg_diag(CsrView A, double *__restrict__ d, double *__restrict__ dinv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const double v = csr_entry(A, i, (int32_t)i);
  d[i] = v; dinv[i] = 1.0 / v;
}

// rank-local part of a square block: entries whose column lies in the owned range of the row
__global__ void k_filter_count(CsrView A, int nr, const int64_t *__restrict__ ranges, int64_t *__restrict__ cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int r = 0;
  while (r + 1 < nr && i >= ranges[r + 1]) ++r;
  const int64_t lo = ranges[r], hi = ranges[r + 1];
  int64_t m = 0;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) m += (A.col[k] >= lo && A.col[k] < hi);
  cnt[i] = m;
}
__global__ void k_filter_fill(CsrView A, int nr, const int64_t *__restrict__ ranges, const int64_t *__restrict__ out_ptr, int32_t *__restrict__ col,
                              double *__restrict__ val) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int r = 0;
  while (r + 1 < nr && i >= ranges[r + 1]) ++r;
  const int64_t lo = ranges[r], hi = ranges[r + 1];
  int64_t q = out_ptr[i];
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
    if (A.col[k] >= lo && A.col[k] < hi) { col[q] = A.col[k]; val[q] = A.val[k]; ++q; }
}

__global__ void k_amg_strong(CsrView A, const double *__restrict__ d, double t2, uint8_t *__restrict__ strong) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int32_t j = A.col[k];
    uint8_t s = 0;
    if (j != i) {
      const double bound = t2 * fabs(d[i] * d[j]), aij = A.val[k], aji = csr_entry(A, j, (int32_t)i);
      s = (aij * aij > bound || aji * aji > bound) ? 1 : 0;
    }
    strong[k] = s;
  }
}

// --- distance-2 maximal independent set (state: 0 undecided, 1 root, 2 excluded) ---
__global__ void k_mis_m1(CsrView A, const uint8_t *__restrict__ strong, const uint8_t *__restrict__ state, unsigned long long *__restrict__ m1) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  unsigned long long m = state[i] == 0 ? amg_key(i) : 0ull;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
    if (strong[k]) { const int32_t j = A.col[k]; if (state[j] == 0) m = max(m, amg_key(j)); }
  m1[i] = m;
}
__global__ void k_mis_select(CsrView A, const uint8_t *__restrict__ strong, const unsigned long long *__restrict__ m1, uint8_t *__restrict__ state) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n || state[i] != 0) return;
  unsigned long long m = m1[i];
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) if (strong[k]) m = max(m, m1[A.col[k]]);
  if (m == amg_key(i)) state[i] = 1;
}
__global__ void k_mis_t1(CsrView A, const uint8_t *__restrict__ strong, const uint8_t *__restrict__ state, uint8_t *__restrict__ t1) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  uint8_t t = state[i] == 1;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1] && !t; ++k) if (strong[k] && state[A.col[k]] == 1) t = 1;
  t1[i] = t;
}
__global__ void k_mis_exclude(CsrView A, const uint8_t *__restrict__ strong, const uint8_t *__restrict__ t1, uint8_t *__restrict__ state,
                              unsigned long long *__restrict__ undecided) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n || state[i] != 0) return;
  uint8_t t = t1[i];
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1] && !t; ++k) if (strong[k] && t1[A.col[k]]) t = 1;
  if (t) state[i] = 2; else atomicAdd(undecided, 1ull);
}
__global__ void k_root_flags(int64_t n, const uint8_t *__restrict__ state, int64_t *__restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flag[i] = state[i] == 1;
}
__global__ void k_agg_phase1(CsrView A, const uint8_t *__restrict__ strong, const uint8_t *__restrict__ state, const int64_t *__restrict__ rootid,
                             int32_t *__restrict__ agg1) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int32_t a = state[i] == 1 ? (int32_t)rootid[i] : -1;
  if (a < 0)
    for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (strong[k] && state[A.col[k]] == 1) a = (int32_t)rootid[A.col[k]];
  agg1[i] = a;
}
__global__ void k_agg_phase2(CsrView A, const uint8_t *__restrict__ strong, const int32_t *__restrict__ agg1, int32_t *__restrict__ agg,
                             unsigned long long *__restrict__ unassigned) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int32_t a = agg1[i];
  if (a < 0) {
    unsigned long long best = 0;
    for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (strong[k]) { const int32_t j = A.col[k]; if (agg1[j] >= 0 && amg_key(j) > best) { best = amg_key(j); a = agg1[j]; } }
    if (a < 0) atomicAdd(unassigned, 1ull);
  }
  agg[i] = a;
}
__global__ void k_make_ptent(int64_t n, const int32_t *__restrict__ agg, int64_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { rowptr[i] = i; col[i] = agg[i]; val[i] = 1.0; }
  if (i == n) rowptr[n] = n;
}

// --- expand / sort / compress products ---
__global__ void k_count_products(CsrView A, const int64_t *__restrict__ b_rowptr, int64_t *__restrict__ cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int64_t m = 0;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) { const int32_t j = A.col[k]; m += b_rowptr[j + 1] - b_rowptr[j]; }
  cnt[i] = m;
}
// with dinv != nullptr the left factor is I - omega D^-1 A instead of A
__global__ void k_expand_products(CsrView A, CsrView B, int64_t ncols_b, const int64_t *__restrict__ off, const double *__restrict__ dinv, double omega,
                                  unsigned long long *__restrict__ keys, double *__restrict__ vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int64_t q = off[i];
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int32_t j = A.col[k];
    double a = A.val[k];
    if (dinv) a = (j == i ? 1.0 : 0.0) - omega * dinv[i] * a;
    for (int64_t l = B.rowptr[j]; l < B.rowptr[j + 1]; ++l, ++q) {
      keys[q] = (unsigned long long)i * (unsigned long long)ncols_b + (unsigned long long)B.col[l];
      vals[q] = a * B.val[l];
    }
  }
}
__global__ void k_expand_transpose(CsrView A, unsigned long long *__restrict__ keys, double *__restrict__ vals) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    keys[k] = (unsigned long long)A.col[k] * (unsigned long long)A.n + (unsigned long long)i;
    vals[k] = A.val[k];
  }
}
__global__ void k_heads(int64_t N, const unsigned long long *__restrict__ keys, int64_t *__restrict__ head) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < N) head[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}
__global__ void k_compress(int64_t N, const unsigned long long *__restrict__ keys, const double *__restrict__ vals, const int64_t *__restrict__ head,
                           const int64_t *__restrict__ pos, unsigned long long ncols, int32_t *__restrict__ col, double *__restrict__ val,
                           int32_t *__restrict__ row) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= N || !head[k]) return;
  const unsigned long long key = keys[k];
  double s = vals[k];
  for (int64_t m = k + 1; m < N && keys[m] == key; ++m) s += vals[m];
  const int64_t q = pos[k];
  col[q] = (int32_t)(key % ncols); row[q] = (int32_t)(key / ncols); val[q] = s;
}
__global__ void k_rowptr_from_rows(int64_t nnz, const int32_t *__restrict__ row, int64_t nrows, int64_t *__restrict__ rowptr) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (nnz == 0) { if (q <= nrows) rowptr[q] = 0; return; }
  if (q >= nnz) return;
  const int64_t r = row[q], prev = q ? row[q - 1] : -1;
  for (int64_t rr = prev + 1; rr <= r; ++rr) rowptr[rr] = q;
  if (q == nnz - 1) for (int64_t rr = r + 1; rr <= nrows; ++rr) rowptr[rr] = nnz;
}

// --- smoother, coarse solve ---
__global__ void k_lanczos_start(int64_t n, double *__restrict__ r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) r[i] = 2.0 * (amg_hash((uint32_t)i ^ 0x9e3779b9U) / 4294967296.0) - 1.0;
}
__global__ void k_mul3(int64_t n, double *__restrict__ z, const double *__restrict__ d, const double *__restrict__ r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = d[i] * r[i];
}
__global__ void k_cheby_zero(int64_t n, const double *__restrict__ dinv, const double *__restrict__ x, double theta, double *__restrict__ w, double *__restrict__ y) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { const double v = dinv[i] * x[i] / theta; w[i] = v; y[i] = v; }
}
// first != 0: w = D^-1 (x - t) / theta ; else w = c1 w + c2 D^-1 (x - t) ; y += w      (t = A y)
__global__ void k_cheby_step(int64_t n, const double *__restrict__ dinv, const double *__restrict__ x, const double *__restrict__ t, int first, double c1,
                             double c2, double *__restrict__ w, double *__restrict__ y) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = first ? dinv[i] * (x[i] - t[i]) / c2 : c1 * w[i] + c2 * dinv[i] * (x[i] - t[i]);
  w[i] = v; y[i] += v;
}
__global__ void k_sub(int64_t n, double *__restrict__ r, const double *__restrict__ b, const double *__restrict__ t) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) r[i] = b[i] - t[i];
}
__global__ void k_to_dense(CsrView A, double *__restrict__ M, double *__restrict__ Inv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  for (int64_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) M[i * A.n + A.col[k]] = A.val[k];
  Inv[i * A.n + i] = 1.0;
}
// Gauss-Jordan with partial pivoting in one CTA: M -> I, Inv -> M^-1 (row-major, n <= AMG_DENSE_MAX)
__global__ void __launch_bounds__(1024) k_gauss_jordan(int n, double *__restrict__ M, double *__restrict__ Inv) {
  __shared__ double s_best[32];
  __shared__ int s_row[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < n; ++k) {
    double best = -1.0; int brow = k;
    for (int i = k + tid; i < n; i += nt) { const double v = fabs(M[(size_t)i * n + k]); if (v > best) { best = v; brow = i; } }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o); const int orow = __shfl_down_sync(0xffffffffu, brow, o);
      if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_row[tid >> 5] = brow; }
    __syncthreads();
    if (tid == 0) {
      double b = s_best[0]; int r = s_row[0];
      for (int w = 1; w < (nt >> 5); ++w) if (s_best[w] > b || (s_best[w] == b && s_row[w] < r)) { b = s_best[w]; r = s_row[w]; }
      s_piv = r;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k)
      for (int j = tid; j < 2 * n; j += nt) {
        double *X = j < n ? M : Inv; const int jj = j < n ? j : j - n;
        const double a = X[(size_t)k * n + jj]; X[(size_t)k * n + jj] = X[(size_t)p * n + jj]; X[(size_t)p * n + jj] = a;
      }
    __syncthreads();
    const double piv = M[(size_t)k * n + k];
    __syncthreads();
    for (int j = tid; j < 2 * n; j += nt) { double *X = j < n ? M : Inv; const int jj = j < n ? j : j - n; X[(size_t)k * n + jj] /= piv; }
    __syncthreads();
    // eliminate column k from every other row; the factors are read before any thread overwrites them
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int rows = min(32, n - i0);
      __shared__ double s_f[32];
      if (tid < rows) s_f[tid] = (i0 + tid == k) ? 0.0 : M[(size_t)(i0 + tid) * n + k];
      __syncthreads();
      for (int e = tid; e < rows * 2 * n; e += nt) {
        const int i = i0 + e / (2 * n), j = e % (2 * n);
        const double f = s_f[i - i0];
        if (f != 0.0) { double *X = j < n ? M : Inv; const int jj = j < n ? j : j - n; X[(size_t)i * n + jj] -= f * X[(size_t)k * n + jj]; }
      }
      __syncthreads();
    }
  }
}
__global__ void k_dense_matvec(int n, const double *__restrict__ Inv, const double *__restrict__ b, double *__restrict__ x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0;
  for (int j = lane; j < n; j += 32) s += Inv[(size_t)row * n + j] * b[j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) x[row] = s;
}

inline int g256(int64_t n) { return (int)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1); }

// largest eigenvalue of a symmetric tridiagonal by Sturm bisection (host, <= 10 x 10)
double tridiag_lambda_max(const std::vector<double> &a, const std::vector<double> &b) {
  const int m = (int)a.size();
  double lo = a[0], hi = a[0];
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r); hi = std::max(hi, a[i] + r);
  }
  for (int it = 0; it < 200; ++it) {
    const double x = 0.5 * (lo + hi);
    int below = 0; double q = 1.0;
    for (int i = 0; i < m; ++i) {
      q = a[i] - x - (i > 0 ? b[i - 1] * b[i - 1] / q : 0.0);
      if (q == 0.0) q = 1e-300;
      if (q < 0) below++;
    }
    if (below >= m) hi = x; else lo = x;
  }
  return 0.5 * (lo + hi);
}

}  // namespace

struct AmgLevel {
  const DevCSR *A = nullptr;  // level operator: level 0 points at the caller's block when it is already rank-local
  DevCSR Aown, P, R;
  DevBuf<double> d, dinv, x, b, w, t, r;
  double lmax = 0;
  int64_t n = 0;
};

struct AmgHierarchy {
  std::vector<std::unique_ptr<AmgLevel>> L;
  DevBuf<double> coarse_M, coarse_inv;
  int64_t nc = 0;
  // scratch of the set-up
  DevBuf<char> cub_tmp;
  DevBuf<unsigned long long> keys, keys2, counter;
  DevBuf<double> vals, vals2;
  DevBuf<int64_t> i64a, i64b;
  DevBuf<int32_t> rows;
  DevBuf<int64_t> ranges;
};

namespace {

void *cub_scratch(AmgHierarchy &H, size_t bytes) {
  if (H.cub_tmp.n < bytes) H.cub_tmp.alloc(bytes + bytes / 4 + 1024);
  return H.cub_tmp.p;
}

void exclusive_sum(Ctx &c, AmgHierarchy &H, const int64_t *in, int64_t *out, int64_t count) {
  size_t bytes = 0;
  NSX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, c.stream));
  void *tmp = cub_scratch(H, bytes);
  NSX_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, c.stream));
  c.stat_launches += 2;
}

int64_t read_i64(Ctx &c, const int64_t *p) {
  int64_t v = 0;
  NSX_CUDA(cudaMemcpyAsync(&v, p, sizeof v, cudaMemcpyDeviceToHost, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  return v;
}

// sorts the N (key, value) pairs in H.keys / H.vals, adds the values of equal keys and stores the result as the CSR C
// (key = row * ncols + col)
void esc_finish(Ctx &c, AmgHierarchy &H, int64_t N, int64_t nrows, int64_t ncols, DevCSR &C) {
  C.nrows = nrows; C.ncols = ncols;
  C.rowptr.alloc(nrows + 1);
  if (N == 0) {
    C.nnz = 0; C.col.alloc(1); C.val.alloc(1);
    k_rowptr_from_rows<<<g256(nrows + 1), 256, 0, c.stream>>>(0, nullptr, nrows, C.rowptr.p);
    return;
  }
  H.keys2.alloc(N); H.vals2.alloc(N);
  int bits = 1;
  while (bits < 64 && ((unsigned long long)nrows * (unsigned long long)ncols) >> bits) ++bits;
  cub::DoubleBuffer<unsigned long long> dk(H.keys.p, H.keys2.p);
  cub::DoubleBuffer<double> dv(H.vals.p, H.vals2.p);
  size_t bytes = 0;
  NSX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, N, 0, bits, c.stream));
  void *tmp = cub_scratch(H, bytes);
  NSX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, dk, dv, N, 0, bits, c.stream));
  const unsigned long long *keys = dk.Current();
  const double *vals = dv.Current();
  H.i64a.alloc(N + 1); H.i64b.alloc(N + 1);
  k_heads<<<g256(N), 256, 0, c.stream>>>(N, keys, H.i64a.p);
  NSX_CUDA(cudaMemsetAsync(H.i64a.p + N, 0, sizeof(int64_t), c.stream));
  exclusive_sum(c, H, H.i64a.p, H.i64b.p, N + 1);
  const int64_t nnz = read_i64(c, H.i64b.p + N);
  C.nnz = nnz;
  C.col.alloc(nnz); C.val.alloc(nnz);
  H.rows.alloc(nnz);
  k_compress<<<g256(N), 256, 0, c.stream>>>(N, keys, vals, H.i64a.p, H.i64b.p, (unsigned long long)ncols, C.col.p, C.val.p, H.rows.p);
  k_rowptr_from_rows<<<g256(nnz), 256, 0, c.stream>>>(nnz, H.rows.p, nrows, C.rowptr.p);
  c.stat_launches += 6;
}

// C = A B, or C = (I - omega D^-1 A) B when dinv is given
void esc_product(Ctx &c, AmgHierarchy &H, const DevCSR &A, const DevCSR &B, DevCSR &C, const double *dinv = nullptr, double omega = 0) {
  const int64_t n = A.nrows;
  H.i64a.alloc(n + 1); H.i64b.alloc(n + 1);
  k_count_products<<<g256(n), 256, 0, c.stream>>>(view(A), B.rowptr.p, H.i64a.p);
  NSX_CUDA(cudaMemsetAsync(H.i64a.p + n, 0, sizeof(int64_t), c.stream));
  exclusive_sum(c, H, H.i64a.p, H.i64b.p, n + 1);
  const int64_t N = read_i64(c, H.i64b.p + n);
  if ((double)N * 32.0 > 64e9) throw std::logic_error("AMG set-up: the expanded product does not fit the scratch budget (chunked products are not built yet)");
  H.keys.alloc(std::max<int64_t>(N, 1)); H.vals.alloc(std::max<int64_t>(N, 1));
  DevBuf<int64_t> off;  // the offsets must survive esc_finish's reuse of the scratch
  off.alloc(n + 1);
  NSX_CUDA(cudaMemcpyAsync(off.p, H.i64b.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, c.stream));
  k_expand_products<<<g256(n), 256, 0, c.stream>>>(view(A), view(B), B.ncols, off.p, dinv, omega, H.keys.p, H.vals.p);
  c.stat_launches += 2;
  esc_finish(c, H, N, n, B.ncols, C);
  NSX_CUDA(cudaStreamSynchronize(c.stream));  // `off` is freed on return
}

void esc_transpose(Ctx &c, AmgHierarchy &H, const DevCSR &A, DevCSR &T) {
  H.keys.alloc(std::max<int64_t>(A.nnz, 1)); H.vals.alloc(std::max<int64_t>(A.nnz, 1));
  k_expand_transpose<<<g256(A.nrows), 256, 0, c.stream>>>(view(A), H.keys.p, H.vals.p);
  c.stat_launches++;
  esc_finish(c, H, A.nnz, A.ncols, A.nrows, T);
}

double lambda_max(Ctx &c, AmgLevel &l) {
  const int64_t n = l.n;
  double *r = l.r.p, *z = l.w.p, *p = l.x.p, *ap = l.t.p;
  k_lanczos_start<<<g256(n), 256, 0, c.stream>>>(n, r);
  k_mul3<<<g256(n), 256, 0, c.stream>>>(n, z, l.dinv.p, r);
  vec_copy(c, p, z, n);
  c.stat_launches += 2;
  double rz = vec_dot(c, r, z, n);
  std::vector<double> al, be;
  for (int it = 0; it < 10 && it < n; ++it) {
    spmv_local(c, *l.A, p, ap);
    const double pap = vec_dot(c, p, ap, n);
    if (!(pap > 0) || !(rz > 0)) break;
    const double alpha = rz / pap;
    vec_axpy(c, r, -alpha, ap, n);
    k_mul3<<<g256(n), 256, 0, c.stream>>>(n, z, l.dinv.p, r);
    c.stat_launches++;
    const double rz_new = vec_dot(c, r, z, n);
    al.push_back(alpha);
    if (!(rz_new > 0)) break;
    const double beta = rz_new / rz;
    be.push_back(beta);
    rz = rz_new;
    vec_sadd(c, p, beta, 1.0, z, n);
  }
  const int m = (int)al.size();
  if (m == 0) throw std::logic_error("AMG set-up: no usable Lanczos step for lambda_max (matrix not positive on the start vector)");
  std::vector<double> ta(m), tb(std::max(0, m - 1));
  for (int k = 0; k < m; ++k) {
    ta[k] = 1.0 / al[k] + (k > 0 ? be[k - 1] / al[k - 1] : 0.0);
    if (k + 1 < m) tb[k] = std::sqrt(be[k]) / al[k];
  }
  return tridiag_lambda_max(ta, tb);
}

// aggregates of level l; returns their number, ids in `agg`
int64_t aggregate(Ctx &c, AmgHierarchy &H, AmgLevel &l, DevBuf<int32_t> &agg) {
  const int64_t n = l.n;
  const CsrView A = view(*l.A);
  DevBuf<uint8_t> strong, state, t1;
  DevBuf<unsigned long long> m1;
  DevBuf<int32_t> agg1;
  strong.alloc(std::max<int64_t>(l.A->nnz, 1)); state.alloc(n); t1.alloc(n); m1.alloc(n); agg1.alloc(n); agg.alloc(n);
  H.counter.alloc(1);
  k_amg_strong<<<g256(n), 256, 0, c.stream>>>(A, l.d.p, AMG_THRESHOLD * AMG_THRESHOLD, strong.p);
  NSX_CUDA(cudaMemsetAsync(state.p, 0, n, c.stream));
  c.stat_launches++;
  for (int round = 0;; ++round) {
    if (round > 1000) throw std::logic_error("AMG set-up: the independent-set sweep did not terminate");
    NSX_CUDA(cudaMemsetAsync(H.counter.p, 0, sizeof(unsigned long long), c.stream));
    k_mis_m1<<<g256(n), 256, 0, c.stream>>>(A, strong.p, state.p, m1.p);
    k_mis_select<<<g256(n), 256, 0, c.stream>>>(A, strong.p, m1.p, state.p);
    k_mis_t1<<<g256(n), 256, 0, c.stream>>>(A, strong.p, state.p, t1.p);
    k_mis_exclude<<<g256(n), 256, 0, c.stream>>>(A, strong.p, t1.p, state.p, H.counter.p);
    c.stat_launches += 4;
    if (read_i64(c, (const int64_t *)H.counter.p) == 0) break;
  }
  H.i64a.alloc(n + 1); H.i64b.alloc(n + 1);
  k_root_flags<<<g256(n), 256, 0, c.stream>>>(n, state.p, H.i64a.p);
  NSX_CUDA(cudaMemsetAsync(H.i64a.p + n, 0, sizeof(int64_t), c.stream));
  exclusive_sum(c, H, H.i64a.p, H.i64b.p, n + 1);
  const int64_t nagg = read_i64(c, H.i64b.p + n);
  NSX_CUDA(cudaMemsetAsync(H.counter.p, 0, sizeof(unsigned long long), c.stream));
  k_agg_phase1<<<g256(n), 256, 0, c.stream>>>(A, strong.p, state.p, H.i64b.p, agg1.p);
  k_agg_phase2<<<g256(n), 256, 0, c.stream>>>(A, strong.p, agg1.p, agg.p, H.counter.p);
  c.stat_launches += 3;
  if (read_i64(c, (const int64_t *)H.counter.p) != 0) throw std::logic_error("AMG aggregation left a node unassigned");
  return nagg;
}

void alloc_level_vectors(AmgLevel &l) {
  const int64_t n = std::max<int64_t>(l.n, 1);
  l.d.alloc(n); l.dinv.alloc(n); l.x.alloc(n); l.b.alloc(n); l.w.alloc(n); l.t.alloc(n); l.r.alloc(n);
}

void cheby(Ctx &c, AmgLevel &l, double *y, const double *x, bool zero_start) {
  const int64_t n = l.n;
  const double beta = AMG_EIG_BOOST * l.lmax, alpha = l.lmax / AMG_EIG_RATIO;
  const double delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha), s1 = theta * delta;
  if (zero_start) k_cheby_zero<<<g256(n), 256, 0, c.stream>>>(n, l.dinv.p, x, theta, l.w.p, y);
  else {
    spmv_local(c, *l.A, y, l.t.p);
    k_cheby_step<<<g256(n), 256, 0, c.stream>>>(n, l.dinv.p, x, l.t.p, 1, 0.0, theta, l.w.p, y);
  }
  c.stat_launches++;
  double rhok = 1.0 / s1;
  for (int k = 1; k < AMG_CHEBY_DEGREE; ++k) {
    const double rhokp1 = 1.0 / (2.0 * s1 - rhok), c1 = rhokp1 * rhok, c2 = 2.0 * rhokp1 * delta;
    rhok = rhokp1;
    spmv_local(c, *l.A, y, l.t.p);
    k_cheby_step<<<g256(n), 256, 0, c.stream>>>(n, l.dinv.p, x, l.t.p, 0, c1, c2, l.w.p, y);
    c.stat_launches++;
  }
}

void vcycle(Ctx &c, AmgHierarchy &H, size_t lev, double *x, const double *b) {
  AmgLevel &l = *H.L[lev];
  if (lev + 1 == H.L.size()) {
    k_dense_matvec<<<g256((int64_t)H.nc * 32), 256, 0, c.stream>>>((int)H.nc, H.coarse_inv.p, b, x);
    c.stat_launches++;
    return;
  }
  AmgLevel &lc = *H.L[lev + 1];
  cheby(c, l, x, b, true);
  spmv_local(c, *l.A, x, l.t.p);
  k_sub<<<g256(l.n), 256, 0, c.stream>>>(l.n, l.r.p, b, l.t.p);
  c.stat_launches++;
  spmv_local(c, l.R, l.r.p, lc.b.p);
  vcycle(c, H, lev + 1, lc.x.p, lc.b.p);
  spmv_local(c, l.P, lc.x.p, x, true);
  cheby(c, l, x, b, false);
}

}  // namespace

void amg_setup(Ctx &c, const DevCSR &F) {
  if (!c.amg) c.amg = std::shared_ptr<AmgHierarchy>(new AmgHierarchy, [](AmgHierarchy *p) { delete p; });
  AmgHierarchy &H = *c.amg;
  // everything below works on this rank's diagonal block: its dot products must not be summed over the ranks
  struct LocalScope { Ctx &c; void *comm; explicit LocalScope(Ctx &cc) : c(cc), comm(cc.comm) { c.comm = nullptr; } ~LocalScope() { c.comm = comm; } } local_scope(c);
  H.L.clear();
  H.L.emplace_back(new AmgLevel);
  {
    AmgLevel &l0 = *H.L[0];
    l0.n = F.nrows;
    const std::vector<int64_t> &owned = c.owned_u;
    if (owned.size() == 2 && c.n_ug == 0) l0.A = &F;  // one rank: the block is its own local part
    else {
      const int nr = (int)owned.size() - 1;
      H.ranges.upload(owned, c.stream);
      H.i64a.alloc(F.nrows + 1); H.i64b.alloc(F.nrows + 1);
      k_filter_count<<<g256(F.nrows), 256, 0, c.stream>>>(view(F), nr, H.ranges.p, H.i64a.p);
      NSX_CUDA(cudaMemsetAsync(H.i64a.p + F.nrows, 0, sizeof(int64_t), c.stream));
      exclusive_sum(c, H, H.i64a.p, H.i64b.p, F.nrows + 1);
      DevCSR &A = l0.Aown;
      A.nrows = A.ncols = F.nrows;
      A.nnz = read_i64(c, H.i64b.p + F.nrows);
      A.rowptr.alloc(F.nrows + 1); A.col.alloc(std::max<int64_t>(A.nnz, 1)); A.val.alloc(std::max<int64_t>(A.nnz, 1));
      NSX_CUDA(cudaMemcpyAsync(A.rowptr.p, H.i64b.p, (F.nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, c.stream));
      k_filter_fill<<<g256(F.nrows), 256, 0, c.stream>>>(view(F), nr, H.ranges.p, A.rowptr.p, A.col.p, A.val.p);
      c.stat_launches += 2;
      l0.A = &A;
    }
  }
  for (int lev = 0;; ++lev) {
    AmgLevel &l = *H.L[lev];
    const int64_t n = l.n;
    alloc_level_vectors(l);
    k_amg_diag<<<g256(n), 256, 0, c.stream>>>(view(*l.A), l.d.p, l.dinv.p);
    c.stat_launches++;
    if (n <= AMG_COARSE_MAX || lev + 1 >= AMG_MAX_LEVELS) break;
    l.lmax = lambda_max(c, l);
    DevBuf<int32_t> agg;
    const int64_t nagg = aggregate(c, H, l, agg);
    if (nagg >= n) break;
    DevCSR Pt;
    Pt.nrows = n; Pt.ncols = nagg; Pt.nnz = n;
    Pt.rowptr.alloc(n + 1); Pt.col.alloc(n); Pt.val.alloc(n);
    k_make_ptent<<<g256(n + 1), 256, 0, c.stream>>>(n, agg.p, Pt.rowptr.p, Pt.col.p, Pt.val.p);
    c.stat_launches++;
    esc_product(c, H, *l.A, Pt, l.P, l.dinv.p, 4.0 / 3.0 / l.lmax);
    esc_transpose(c, H, l.P, l.R);
    DevCSR AP;
    esc_product(c, H, *l.A, l.P, AP);
    H.L.emplace_back(new AmgLevel);
    AmgLevel &lc = *H.L.back();
    esc_product(c, H, H.L[lev]->R, AP, lc.Aown);
    lc.A = &lc.Aown;
    lc.n = nagg;
    if (c.verbose) fprintf(stderr, "[nsx amg] level %d: n %lld nnz %lld lambda_max %.4f -> %lld aggregates\n", lev, (long long)n,
                           (long long)H.L[lev]->A->nnz, H.L[lev]->lmax, (long long)nagg);
  }
  AmgLevel &lc = *H.L.back();
  H.nc = lc.n;
  if (H.nc > AMG_DENSE_MAX) throw std::logic_error("AMG coarsest level too large for the dense solve");
  H.coarse_M.alloc((size_t)H.nc * H.nc); H.coarse_inv.alloc((size_t)H.nc * H.nc);
  H.coarse_M.zero(c.stream); H.coarse_inv.zero(c.stream);
  k_to_dense<<<g256(H.nc), 256, 0, c.stream>>>(view(*lc.A), H.coarse_M.p, H.coarse_inv.p);
  k_gauss_jordan<<<1, 1024, 0, c.stream>>>((int)H.nc, H.coarse_M.p, H.coarse_inv.p);
  c.stat_launches += 2;
  // release the set-up scratch (hundreds of MB on the fine level)
  H.keys.release(); H.keys2.release(); H.vals.release(); H.vals2.release(); H.i64a.release(); H.i64b.release(); H.rows.release(); H.cub_tmp.release();
}

void amg_apply(Ctx &c, double *y, const double *x) {
  if (!c.amg || c.amg->L.empty()) throw std::logic_error("amg_apply before amg_setup");
  vcycle(c, *c.amg, 0, y, x);
}

int amg_levels(const Ctx &c) { return c.amg ? (int)c.amg->L.size() : 0; }

}  // namespace nsx
