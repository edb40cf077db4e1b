// trisolve.cu -- rank-local ILU(0) and symmetric Gauss-Seidel on one square block.
//
// Replaces TrilinosWrappers::PreconditionILU (Ifpack ILU, fill 0, overlap 0:
// NSSolverStationary.hpp:231, 325-326; NSSolver.hpp:183, 189, 244, 250, 370, 373) and
// TrilinosWrappers::PreconditionSSOR (Ifpack point relaxation, symmetric Gauss-Seidel, one sweep,
// omega 1, zero start: NSSolverStationary.hpp:160, 166).  Both are local to each rank's owned
// range in the reference (additive Schwarz with overlap 0), so couplings that cross an owned-range
// boundary are dropped when the plan is built.
//
// Both operators are a lower sweep followed by an upper sweep over the same sparse matrix:
//   SGS : w = (D + L)^-1 x ,  y = w - D^-1 U y          (== Ifpack's two Gauss-Seidel passes)
//   ILU : w = L^-1 x       ,  y = U^-1 w                 (L unit lower, U upper incl. diagonal)
// The plan stores the block permuted by the elimination order, its rows grouped by dependency level:
//   * natural order (Ifpack's; NSX_OPT_ORDERING = 0, parity runs): thousands of short levels.  A level with many rows
//     is one grid launch (a sub-warp per row); runs of consecutive small levels are chained inside ONE thread block
//     with __syncthreads() between levels, which removes the launch latency that dominates this order;
//   * multicolour order (default): a greedy colouring re-sorted by dependency level (32 levels for Q3/Q2, each a
//     contiguous row range for both sweeps) and ONE persistent launch per application, k_sweep_phased below:
//     colour phases separated by a hand-rolled grid barrier, the next phase's rows staged in shared memory with
//     cp.async while the barrier is in flight.
// The ILU(0) factorisation runs over the same level schedule (one launch per level).
#include <cooperative_groups.h>

#include <algorithm>
#include <numeric>

#include "device.cuh"

namespace cg = cooperative_groups;

namespace nsx {

namespace {

constexpr int TG = 8;          // lanes per row
constexpr int CHAIN_T = 1024;  // threads of the chained single-block kernel
constexpr int CHAIN_ROWS = 4 * (CHAIN_T / TG);  // a level with at most this many rows may be chained

struct TriArgs {
  const int64_t *rowptr;
  const int32_t *col, *diag, *perm;
  double *val;
  const double *x;  // input, original numbering, offset to the block
  double *w, *yp;   // permuted work vectors
  double *y;        // output, original numbering
};

template <bool SGS, bool FWD>
__device__ __forceinline__ void tri_row(const TriArgs &A, int32_t row, const cg::thread_block_tile<TG> &tile) {
  const int64_t b = A.rowptr[row], e = A.rowptr[row + 1], d = b + A.diag[row];
  double s = 0;
  if (FWD) {
    for (int64_t k = b + tile.thread_rank(); k < d; k += TG) s += __ldcg(A.val + k) * __ldcg(A.w + A.col[k]);
  } else {
    for (int64_t k = d + 1 + tile.thread_rank(); k < e; k += TG) s += __ldcg(A.val + k) * __ldcg(A.yp + A.col[k]);
  }
#pragma unroll
  for (int o = TG / 2; o > 0; o >>= 1) s += tile.shfl_down(s, o);
  if (tile.thread_rank() == 0) {
    const double dg = __ldcg(A.val + d);
    if (FWD) {
      const double r = A.x[A.perm[row]] - s;
      A.w[row] = SGS ? r / dg : r;
    } else {
      const double wv = __ldcg(A.w + row);
      const double r = SGS ? wv - s / dg : (wv - s) / dg;
      A.yp[row] = r;
      A.y[A.perm[row]] = r;
    }
  }
}

template <bool SGS, bool FWD>
__global__ void __launch_bounds__(256) k_tri_level(TriArgs A, const int32_t *order, int64_t r0, int64_t r1) {
  auto tile = cg::tiled_partition<TG>(cg::this_thread_block());
  const int64_t r = r0 + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / TG;
  if (r < r1) tri_row<SGS, FWD>(A, order[r], tile);
}

template <bool SGS, bool FWD>
__global__ void __launch_bounds__(CHAIN_T) k_tri_chain(TriArgs A, const int32_t *order, const int64_t *lvl, int l0, int l1) {
  auto tile = cg::tiled_partition<TG>(cg::this_thread_block());
  for (int l = l0; l < l1; ++l) {
    const int64_t r1 = lvl[l + 1];
    for (int64_t r = lvl[l] + threadIdx.x / TG; r < r1; r += CHAIN_T / TG) tri_row<SGS, FWD>(A, order[r], tile);
    __syncthreads();
  }
}

// ILU(0) of one row: for k < row in the row's pattern (ascending), l_ik = a_ik / u_kk, then
// a_ij -= l_ik u_kj for the j > k of row k that are also in row i.
__device__ __forceinline__ void ilu_row(const TriArgs &A, int32_t row, const cg::thread_block_tile<TG> &tile) {
  const int64_t b = A.rowptr[row], e = A.rowptr[row + 1], d = b + A.diag[row];
  for (int64_t p = b; p < d; ++p) {
    const int32_t k = A.col[p];
    const int64_t kd = A.rowptr[k] + A.diag[k], ke = A.rowptr[k + 1];
    const double lik = __ldcg(A.val + p) / __ldcg(A.val + kd);
    tile.sync();
    if (tile.thread_rank() == 0) A.val[p] = lik;
    for (int64_t m = kd + 1 + tile.thread_rank(); m < ke; m += TG) {
      const int32_t cj = A.col[m];
      int64_t lo = p + 1, hi = e;  // columns of this row right of k
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (A.col[mid] < cj) lo = mid + 1; else hi = mid;
      }
      if (lo < e && A.col[lo] == cj) A.val[lo] = __ldcg(A.val + lo) - lik * __ldcg(A.val + m);
    }
    tile.sync();
  }
}

__global__ void __launch_bounds__(256) k_ilu_level(TriArgs A, const int32_t *order, int64_t r0, int64_t r1) {
  auto tile = cg::tiled_partition<TG>(cg::this_thread_block());
  const int64_t r = r0 + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / TG;
  if (r < r1) ilu_row(A, order[r], tile);
}

__global__ void __launch_bounds__(CHAIN_T) k_ilu_chain(TriArgs A, const int32_t *order, const int64_t *lvl, int l0, int l1) {
  auto tile = cg::tiled_partition<TG>(cg::this_thread_block());
  for (int l = l0; l < l1; ++l) {
    const int64_t r1 = lvl[l + 1];
    for (int64_t r = lvl[l] + threadIdx.x / TG; r < r1; r += CHAIN_T / TG) ilu_row(A, order[r], tile);
    __syncthreads();
  }
}

// ---- cooperative sweep: all levels of both sweeps in one launch, grid barrier between levels ----
// With ~17 k rows per colour and ~17 k sub-warps in the grid a sub-warp owns about one row per
// level, and a level costs one barrier plus the dependent chain rowptr -> (val, col) -> w[col].
// Only the last link depends on the previous level, so the sub-warp loads the matrix entries of
// its NEXT row into registers before it enters the barrier; after the barrier only the gather of
// the work vector and the shuffle reduction remain.
constexpr int PF = 4;  // prefetched entries per lane (rows of up to PF * TG entries per triangle are fully covered)
struct RowPre {
  int32_t row;
  int32_t c[PF];
  double v[PF];
  int64_t rest_b, rest_e;
  double dg, xin;
};

template <bool FWD>
__device__ __forceinline__ void prefetch_row(const TriArgs &A, const int32_t *__restrict__ order, int64_t r, int64_t r1, int lane, RowPre &P) {
  P.row = -1;
  if (r >= r1) return;
  const int32_t row = order[r];
  P.row = row;
  const int64_t b = A.rowptr[row], e = A.rowptr[row + 1], d = b + A.diag[row];
  const int64_t lo = FWD ? b : d + 1, hi = FWD ? d : e;
#pragma unroll
  for (int t = 0; t < PF; ++t) {
    const int64_t k = lo + lane + t * TG;
    if (k < hi) { P.v[t] = __ldcg(A.val + k); P.c[t] = A.col[k]; }
    else { P.v[t] = 0.0; P.c[t] = -1; }
  }
  P.rest_b = lo + PF * TG + lane; P.rest_e = hi;
  P.dg = __ldcg(A.val + d);
  P.xin = FWD ? A.x[A.perm[row]] : 0.0;
}

template <bool SGS, bool FWD>
__device__ __forceinline__ void finish_row(const TriArgs &A, const RowPre &P, const cg::thread_block_tile<TG> &tile) {
  if (P.row < 0) return;  // uniform over the tile
  const double *src = FWD ? A.w : A.yp;
  double s = 0;
#pragma unroll
  for (int t = 0; t < PF; ++t)
    if (P.c[t] >= 0) s += P.v[t] * __ldcg(src + P.c[t]);
  for (int64_t k = P.rest_b; k < P.rest_e; k += TG) s += __ldcg(A.val + k) * __ldcg(src + A.col[k]);
#pragma unroll
  for (int o = TG / 2; o > 0; o >>= 1) s += tile.shfl_down(s, o);
  if (tile.thread_rank() == 0) {
    if (FWD) {
      const double r = P.xin - s;
      A.w[P.row] = SGS ? r / P.dg : r;
    } else {
      const double wv = __ldcg(A.w + P.row);
      const double r = SGS ? wv - s / P.dg : (wv - s) / P.dg;
      A.yp[P.row] = r;
      A.y[A.perm[P.row]] = r;
    }
  }
}

template <bool SGS>
__global__ void __launch_bounds__(256) k_sweep_coop(TriArgs A, const int32_t *__restrict__ order_f, const int64_t *__restrict__ lvl_f, int nlf,
                                                    const int32_t *__restrict__ order_b, const int64_t *__restrict__ lvl_b, int nlb) {
  cg::grid_group grid = cg::this_grid();
  auto tile = cg::tiled_partition<TG>(cg::this_thread_block());
  const int lane = tile.thread_rank();
  const int64_t sub = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / TG, nsub = (int64_t)gridDim.x * blockDim.x / TG;
  RowPre P;
  prefetch_row<true>(A, order_f, lvl_f[0] + sub, lvl_f[1], lane, P);
  for (int l = 0; l < nlf; ++l) {
    const int64_t r0 = lvl_f[l], r1 = lvl_f[l + 1];
    finish_row<SGS, true>(A, P, tile);
    for (int64_t r = r0 + sub + nsub; r < r1; r += nsub) tri_row<SGS, true>(A, order_f[r], tile);
    if (l + 1 < nlf) prefetch_row<true>(A, order_f, r1 + sub, lvl_f[l + 2], lane, P);
    else prefetch_row<false>(A, order_b, lvl_b[0] + sub, lvl_b[1], lane, P);
    grid.sync();
  }
  for (int l = 0; l < nlb; ++l) {
    const int64_t r0 = lvl_b[l], r1 = lvl_b[l + 1];
    finish_row<SGS, false>(A, P, tile);
    for (int64_t r = r0 + sub + nsub; r < r1; r += nsub) tri_row<SGS, false>(A, order_b[r], tile);
    if (l + 1 < nlb) {
      prefetch_row<false>(A, order_b, r1 + sub, lvl_b[l + 2], lane, P);
      grid.sync();
    }
  }
}

// ---- colour-phased persistent sweep ------------------------------------------------------------------------
// One CTA of FTHREADS threads per SM, a sub-warp of FT lanes per row, 2 x ncol phases (colours ascending for the lower
// sweep, descending for the upper one) separated by a hand-rolled split grid barrier (one arrival counter in L2:
// fence + red by one thread per CTA, ld.acquire spin; ~1.4 us measured on B200, tools/microbench/grid_barrier.cu).
// A phase on the critical path is: barrier -> gather of the work vector through L2 -> shuffle reduction -> store.
// Everything that does not depend on the previous colour is fetched ahead of the barrier wait: each lane stages the
// values and columns of its share of the NEXT phase's row in shared memory with cp.async (FCAP entries per lane, i.e.
// rows of up to FCAP x FT entries per triangle -- every row of the Q3/Q2 and P2/P1 patterns -- at no register cost),
// the row's diagonal and right-hand side ride in registers, and the row descriptor + permutation entry are loaded two
// phases ahead.  No DRAM round trip is left between two barriers.
constexpr int FT = 4, FCAP = 24;

__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

struct RowPreF {
  int32_t row, prm, cnt;   // cnt: staged entries of this lane
  int64_t rest_b, rest_e;  // entries beyond the staged ones (none on the supported element patterns)
  double dg, xin;
};

// split grid barrier: arrive (after the CTA's stores) ... independent loads of the next phase ... wait.  The fence of the
// arriving thread must not sit behind the prefetch loads (it would wait for their DRAM round trip), so they are issued
// between the two halves and their latency overlaps the wait.
__device__ __forceinline__ void grid_arrive(unsigned long long *counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1ull);
  }
}
__device__ __forceinline__ void grid_wait(unsigned long long *counter, unsigned long long target) {
  if (threadIdx.x == 0) {
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// sv / sc: this thread's column of the staging area, stride = blockDim.x entries
template <bool FWD, int FTHREADS>
__device__ __forceinline__ void stage_row(const TriArgs &A, int32_t row, int4 ri, int32_t prm, int lane, double *sv, int32_t *sc, RowPreF &P) {
  P.row = row; P.prm = prm; P.cnt = 0;
  if (row < 0) return;
  const int64_t b = ((int64_t)(uint32_t)ri.x) | ((int64_t)ri.y << 32), d = b + ri.z;
  const int64_t lo = FWD ? b : d + 1, hi = FWD ? d : d + 1 + ri.w;
  int cnt = 0;
  for (int64_t k = lo + lane; k < hi && cnt < FCAP; k += FT, ++cnt) {
    cp_async_8(sv + cnt * FTHREADS, A.val + k);
    cp_async_4(sc + cnt * FTHREADS, A.col + k);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  P.cnt = cnt;
  P.rest_b = lo + lane + (int64_t)FCAP * FT; P.rest_e = hi;
  P.dg = __ldcs(A.val + d);
  P.xin = FWD ? __ldg(A.x + prm) : 0.0;
}

template <bool SGS, bool FWD, int FTHREADS>
__device__ __forceinline__ void finish_row_f(const TriArgs &A, const RowPreF &P, const double *sv, const int32_t *sc, const cg::thread_block_tile<FT> &tile) {
  if (P.row < 0) return;  // uniform over the tile
  const double *src = FWD ? A.w : A.yp;
  // eight predicated gathers of the work vector in flight per trip: a row of up to 8 x FT staged entries costs one L2
  // round trip on the critical path instead of one per pair of entries
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int t0 = 0; t0 < P.cnt; t0 += 8) {
    double xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = (t0 + u < P.cnt) ? __ldcg(src + sc[(t0 + u) * FTHREADS]) : 0.0;
#pragma unroll
    for (int u = 0; u < 8; u += 4) {
      s0 += (t0 + u < P.cnt ? sv[(t0 + u) * FTHREADS] : 0.0) * xv[u];
      s1 += (t0 + u + 1 < P.cnt ? sv[(t0 + u + 1) * FTHREADS] : 0.0) * xv[u + 1];
      s2 += (t0 + u + 2 < P.cnt ? sv[(t0 + u + 2) * FTHREADS] : 0.0) * xv[u + 2];
      s3 += (t0 + u + 3 < P.cnt ? sv[(t0 + u + 3) * FTHREADS] : 0.0) * xv[u + 3];
    }
  }
  for (int64_t k = P.rest_b; k < P.rest_e; k += FT) s1 += __ldcs(A.val + k) * __ldcg(src + __ldcs(A.col + k));
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = FT / 2; o > 0; o >>= 1) s += tile.shfl_down(s, o);
  if (tile.thread_rank() == 0) {
    if (FWD) {
      const double r = P.xin - s;
      A.w[P.row] = SGS ? r / P.dg : r;
    } else {
      const double wv = __ldcg(A.w + P.row);
      const double r = SGS ? wv - s / P.dg : (wv - s) / P.dg;
      A.yp[P.row] = r;
      A.y[P.prm] = r;
    }
  }
}

template <bool SGS, int FTHREADS>
__global__ void __launch_bounds__(FTHREADS, 1) k_sweep_phased(TriArgs A, const int4 *__restrict__ rinfo, const int64_t *__restrict__ cptr, int ncol,
                                                               unsigned long long *counter, unsigned long long epoch0) {
  extern __shared__ __align__(16) unsigned char stage_raw[];
  __shared__ int64_t s_cptr[258];
  double *sv = reinterpret_cast<double *>(stage_raw) + threadIdx.x;                                   // [FCAP][FTHREADS] doubles
  int32_t *sc = reinterpret_cast<int32_t *>(stage_raw + (size_t)FCAP * FTHREADS * sizeof(double)) + threadIdx.x;  // [FCAP][FTHREADS] ints
  for (int i = threadIdx.x; i <= ncol; i += FTHREADS) s_cptr[i] = cptr[i];
  __syncthreads();
  auto tile = cg::tiled_partition<FT>(cg::this_thread_block());
  const int lane = tile.thread_rank();
  const int64_t sub = (blockIdx.x * (int64_t)FTHREADS + threadIdx.x) / FT, nsub = (int64_t)gridDim.x * FTHREADS / FT;
  const int nph = 2 * ncol;
  unsigned long long target = epoch0;
  // row of this sub-warp in phase p (-1: none)
  auto row_of = [&](int p) -> int32_t {
    if (p >= nph) return -1;
    const int col = p < ncol ? p : nph - 1 - p;
    const int64_t r = s_cptr[col] + sub;
    return r < s_cptr[col + 1] ? (int32_t)r : -1;
  };
  RowPreF P;
  // ring of row descriptors, three phases deep: slot (p % 3) holds the descriptor + permutation entry of this sub-warp's row in phase p
  constexpr int NSUB = FTHREADS / FT;
  __shared__ __align__(16) int4 s_ri[3 * NSUB];
  __shared__ int32_t s_prm[3 * NSUB];
  const int lsub = threadIdx.x / FT;
  int32_t row1 = row_of(1);
  int4 ri1 = make_int4(0, 0, 0, 0);
  int32_t prm1 = 0;
  {
    const int32_t row0 = row_of(0);
    int4 ri0 = make_int4(0, 0, 0, 0);
    int32_t prm0 = 0;
    if (row0 >= 0) { ri0 = __ldg(rinfo + row0); prm0 = __ldg(A.perm + row0); }
    if (row1 >= 0) { ri1 = __ldg(rinfo + row1); prm1 = __ldg(A.perm + row1); }
    stage_row<true, FTHREADS>(A, row0, ri0, prm0, lane, sv, sc, P);
    asm volatile("cp.async.commit_group;" ::: "memory");  // (empty) descriptor group, keeps the group count per phase uniform
    // descriptor of phase 1 into its ring slot
    if (row1 >= 0 && lane == 0) { s_ri[(1 % 3) * NSUB + lsub] = ri1; s_prm[(1 % 3) * NSUB + lsub] = prm1; }
  }
  __syncthreads();
  for (int p = 0; p < nph; ++p) {
    const bool fwd = p < ncol;
    const int col = fwd ? p : nph - 1 - p;
    // the staged row of this phase: its group is the last but one (the descriptor group of phase p + 2 follows it)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    if (fwd) finish_row_f<SGS, true, FTHREADS>(A, P, sv, sc, tile); else finish_row_f<SGS, false, FTHREADS>(A, P, sv, sc, tile);
    // colours wider than the grid: the remaining rows, staged and consumed on the spot
    for (int64_t r = s_cptr[col] + sub + nsub; r < s_cptr[col + 1]; r += nsub) {
      RowPreF Q;
      const int4 ri = __ldg(rinfo + r);
      const int32_t prm = __ldg(A.perm + r);
      if (fwd) stage_row<true, FTHREADS>(A, (int32_t)r, ri, prm, lane, sv, sc, Q); else stage_row<false, FTHREADS>(A, (int32_t)r, ri, prm, lane, sv, sc, Q);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (fwd) finish_row_f<SGS, true, FTHREADS>(A, Q, sv, sc, tile); else finish_row_f<SGS, false, FTHREADS>(A, Q, sv, sc, tile);
    }
    if (p + 1 < nph) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");   // descriptor of phase p + 1 (issued a phase ago) has landed
      grid_arrive(counter);                                    // ... and its __syncthreads publishes it to the sub-warp
      // next phase's matrix entries from the descriptor in the ring, then the descriptor of the phase after it (no register
      // ever depends on that load before the next barrier has passed)
      const int32_t rown = row_of(p + 1);
      int4 ri = make_int4(0, 0, 0, 0);
      int32_t prm = 0;
      if (rown >= 0) { ri = s_ri[((p + 1) % 3) * NSUB + lsub]; prm = s_prm[((p + 1) % 3) * NSUB + lsub]; }
      if (p + 1 < ncol) stage_row<true, FTHREADS>(A, rown, ri, prm, lane, sv, sc, P); else stage_row<false, FTHREADS>(A, rown, ri, prm, lane, sv, sc, P);
      const int32_t row2 = row_of(p + 2);
      if (row2 >= 0 && lane == 0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_ri + ((p + 2) % 3) * NSUB + lsub)), "l"(rinfo + row2) : "memory");
        cp_async_4(s_prm + ((p + 2) % 3) * NSUB + lsub, A.perm + row2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      target += gridDim.x;
      grid_wait(counter, target);
    }
  }
}

__global__ void k_gather_values(int64_t nnz, const int64_t *__restrict__ src, const double *__restrict__ a, double *__restrict__ v) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) v[k] = a[src[k]];
}

void make_schedule(const std::vector<int64_t> &lvl, std::vector<TriStep> &steps) {
  steps.clear();
  const int nl = (int)lvl.size() - 1;
  int l = 0;
  while (l < nl) {
    const int64_t rows = lvl[l + 1] - lvl[l];
    if (rows > CHAIN_ROWS) {
      steps.push_back(TriStep{0, l, l + 1});
      ++l;
    } else {
      int e = l;
      while (e < nl && lvl[e + 1] - lvl[e] <= CHAIN_ROWS) ++e;
      steps.push_back(TriStep{1, l, e});
      l = e;
    }
  }
}

// levels of the dependency DAG of the strictly lower (fwd) or upper (bwd) part
void levels_of(const TriPlan &P, bool fwd, std::vector<int64_t> &lvl, std::vector<int32_t> &order) {
  const int64_t n = P.n;
  std::vector<int32_t> level(n, 0);
  int nl = 0;
  if (fwd) {
    for (int64_t i = 0; i < n; ++i) {
      int32_t m = 0;
      for (int64_t k = P.h_rowptr[i]; k < P.h_rowptr[i] + P.h_diag[i]; ++k) m = std::max(m, level[P.h_col[k]] + 1);
      level[i] = m; nl = std::max(nl, m + 1);
    }
  } else {
    for (int64_t i = n - 1; i >= 0; --i) {
      int32_t m = 0;
      for (int64_t k = P.h_rowptr[i] + P.h_diag[i] + 1; k < P.h_rowptr[i + 1]; ++k) m = std::max(m, level[P.h_col[k]] + 1);
      level[i] = m; nl = std::max(nl, m + 1);
    }
  }
  lvl.assign(nl + 1, 0);
  for (int64_t i = 0; i < n; ++i) lvl[level[i] + 1]++;
  for (int l = 0; l < nl; ++l) lvl[l + 1] += lvl[l];
  order.resize(n);
  std::vector<int64_t> fill(lvl.begin(), lvl.end() - 1);
  for (int64_t i = 0; i < n; ++i) order[fill[level[i]]++] = (int32_t)i;
}

}  // namespace

DevCSR &block_ref(Ctx &c, int block) {
  switch (block) {
    case NSX_BLOCK_F: return c.F;
    case NSX_BLOCK_BT: return c.Bt;
    case NSX_BLOCK_B: return c.B;
    case NSX_BLOCK_MP: return c.Mp;
    case NSX_BLOCK_S: return c.S;
  }
  throw std::invalid_argument("unknown block id");
}

TriPlan &tri_plan(Ctx &c, int block, int variant, int ordering) {
  if (variant && block != NSX_BLOCK_F) throw std::invalid_argument("only block F has a component-decoupled plan");
  const int ord = ordering < 0 ? c.ordering : ordering;
  const int key = block + 16 * variant + 256 * ord;
  auto it = c.tri.find(key);
  if (it != c.tri.end()) return *it->second;
  const DevCSR &A0 = block_ref(c, block);
  if (A0.nrows > A0.ncols) throw std::invalid_argument("triangular plan needs a square block (owned rows x owned + ghost columns on a partitioned system)");
  if (A0.h_rowptr.empty()) throw std::logic_error("block pattern is not set");
  // The pattern the plan is built from: the block's own, or (variant 1) its same-component entries only.  `A` is a light
  // view (host pattern + sizes); orig[k] is the position of entry k in the block's value array.
  struct View {
    int64_t nrows, ncols;
    std::vector<int64_t> own_rp, orig;
    std::vector<int32_t> own_col;
    const std::vector<int64_t> *rp;
    const std::vector<int32_t> *col;
  } V;
  V.nrows = A0.nrows; V.ncols = A0.ncols; V.rp = &A0.h_rowptr; V.col = &A0.h_col;
  if (variant == 2) {   // velocity nodes: the scalar matrix K of F = K (x) I_2 (decouple.cu)
    if (c.node_struct != 1) throw std::logic_error("the node view of F is not available");
    V.nrows = c.Kn.nrows; V.ncols = c.Kn.ncols; V.rp = &c.Kn.h_rowptr; V.col = &c.Kn.h_col;
    V.orig = c.h_Kn_src;
  } else if (variant) {
    const std::vector<uint8_t> &comp = velocity_components(c);
    V.own_rp.assign(A0.nrows + 1, 0);
    for (int64_t i = 0; i < A0.nrows; ++i) {
      int64_t cnt = 0;
      for (int64_t k = A0.h_rowptr[i]; k < A0.h_rowptr[i + 1]; ++k) cnt += comp[A0.h_col[k]] == comp[i];
      V.own_rp[i + 1] = V.own_rp[i] + cnt;
    }
    V.own_col.resize(V.own_rp[A0.nrows]); V.orig.resize(V.own_rp[A0.nrows]);
    for (int64_t i = 0; i < A0.nrows; ++i) {
      int64_t o = V.own_rp[i];
      for (int64_t k = A0.h_rowptr[i]; k < A0.h_rowptr[i + 1]; ++k)
        if (comp[A0.h_col[k]] == comp[i]) { V.own_col[o] = A0.h_col[k]; V.orig[o] = k; ++o; }
    }
    V.rp = &V.own_rp; V.col = &V.own_col;
  }
  struct { int64_t nrows, ncols; const std::vector<int64_t> &h_rowptr; const std::vector<int32_t> &h_col; } A{V.nrows, V.ncols, *V.rp, *V.col};
  const std::vector<int64_t> &owned = variant == 2 ? c.owned_nodes : (block == NSX_BLOCK_F) ? c.owned_u : c.owned_p;
  std::unique_ptr<TriPlan> up(new TriPlan);
  TriPlan &P = *up;
  P.ordering = ord;
  P.node = variant == 2;
  const int64_t n = A.nrows;
  P.n = n;
  // group of each row: couplings between rows of different groups are dropped.  Orderings 0 / 1: the owned ranges (the
  // plan keeps exactly the rank-local diagonal block, as Ifpack does with overlap 0; ghost columns -- local ids >= n on a
  // partitioned system -- belong to no group).  Ordering 2: every owned range is cut further into spatially compact
  // blocks that one CTA sweeps out of shared memory (what the reference's preconditioners are under mpirun -n <#blocks>).
  std::vector<int32_t> range(std::max<int64_t>(n, A.ncols), -1);
  if (ord >= 2) {
    int ng = 0;
    for (size_t r = 0; r + 1 < owned.size(); ++r)
      if (owned[r + 1] > owned[r]) ng += geometric_blocks(c, block, A.h_rowptr, owned[r], owned[r + 1], ng, range, variant == 2 ? &c.h_node_dx : nullptr);
    P.nblk = ng;
  } else {
    for (size_t r = 0; r + 1 < owned.size(); ++r)
      for (int64_t i = owned[r]; i < owned[r + 1]; ++i) range[i] = (int32_t)r;
  }
  // elimination order
  std::vector<int32_t> perm(n), iperm(n);
  if (ord == 0) {
    std::iota(perm.begin(), perm.end(), 0);
  } else {
    std::vector<int32_t> colour(n, ord == 3 ? 0 : -1), stamp;   // ordering 3: one "colour", i.e. the natural order inside every block
    int ncol = ord == 3 ? 1 : 0;
    for (int64_t i = 0; i < n && ord != 3; ++i) {
      stamp.assign(ncol + 1, 0);
      for (int64_t k = A.h_rowptr[i]; k < A.h_rowptr[i + 1]; ++k) {
        const int32_t j = A.h_col[k];
        if (j != i && range[j] == range[i] && colour[j] >= 0) stamp[colour[j]] = 1;
      }
      int col = 0;
      while (col < ncol && stamp[col]) ++col;
      colour[i] = col;
      ncol = std::max(ncol, col + 1);
    }
    std::vector<int64_t> cnt(ncol + 1, 0);
    for (int64_t i = 0; i < n; ++i) cnt[colour[i] + 1]++;
    for (int q = 0; q < ncol; ++q) cnt[q + 1] += cnt[q];
    for (int64_t i = 0; i < n; ++i) perm[cnt[colour[i]]++] = (int32_t)i;
    // Greedy colouring may use more colours than the elimination order needs phases.  Re-sort the rows by their
    // dependency level under the colour order: two coupled rows keep their relative order (the later one sits at least
    // one level higher), so the triangular split -- and with it the preconditioner -- is unchanged, while the number of
    // phases drops to the depth of the dependency graph and every level is a contiguous row range for both sweeps.
    for (int64_t r = 0; r < n; ++r) iperm[perm[r]] = (int32_t)r;
    std::vector<int32_t> level(n, 0);
    int nlev = 0;
    for (int64_t r = 0; r < n; ++r) {
      const int64_t i = perm[r];
      int32_t m = 0;
      for (int64_t k = A.h_rowptr[i]; k < A.h_rowptr[i + 1]; ++k) {
        const int32_t j = A.h_col[k];
        if (j != i && range[j] == range[i] && iperm[j] < r) m = std::max(m, level[iperm[j]] + 1);
      }
      level[r] = m; nlev = std::max(nlev, m + 1);
    }
    if (ord == 1) {
      std::vector<int64_t> lp(nlev + 1, 0);
      for (int64_t r = 0; r < n; ++r) lp[level[r] + 1]++;
      for (int q = 0; q < nlev; ++q) lp[q + 1] += lp[q];
      P.cptr = lp;
      std::vector<int32_t> perm2(n);
      for (int64_t r = 0; r < n; ++r) perm2[lp[level[r]]++] = perm[r];
      perm.swap(perm2);
    } else {
      // block-local: rows sorted by (block, level), the colour order breaking ties
      std::vector<int64_t> key(n);
      for (int64_t r = 0; r < n; ++r) key[r] = ((int64_t)range[perm[r]] * nlev + level[r]) * n + r;
      std::sort(key.begin(), key.end());
      std::vector<int32_t> perm2(n);
      P.h_level.resize(n);
      P.blk_off.assign(P.nblk + 1, 0);
      for (int64_t q = 0; q < n; ++q) {
        const int64_t r = key[q] % n;
        perm2[q] = perm[r];
        P.h_level[q] = level[r];
        P.blk_off[range[perm[r]] + 1]++;
      }
      for (int b = 0; b < P.nblk; ++b) P.blk_off[b + 1] += P.blk_off[b];
      perm.swap(perm2);
    }
  }
  for (int64_t r = 0; r < n; ++r) iperm[perm[r]] = (int32_t)r;
  // permuted, filtered pattern
  P.h_rowptr.assign(n + 1, 0);
  for (int64_t r = 0; r < n; ++r) {
    const int64_t i = perm[r];
    int64_t cnt = 0;
    for (int64_t k = A.h_rowptr[i]; k < A.h_rowptr[i + 1]; ++k) cnt += range[A.h_col[k]] == range[i];
    P.h_rowptr[r + 1] = P.h_rowptr[r] + cnt;
  }
  P.nnz = P.h_rowptr[n];
  P.h_col.resize(P.nnz);
  P.h_diag.assign(n, -1);
  std::vector<int64_t> src(P.nnz);
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, int64_t>> tmp;
#pragma omp for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      const int64_t i = perm[r];
      tmp.clear();
      for (int64_t k = A.h_rowptr[i]; k < A.h_rowptr[i + 1]; ++k)
        if (range[A.h_col[k]] == range[i]) tmp.emplace_back(iperm[A.h_col[k]], k);
      std::sort(tmp.begin(), tmp.end());
      int64_t o = P.h_rowptr[r];
      for (auto &t : tmp) {
        if (t.first == r) P.h_diag[r] = (int32_t)(o - P.h_rowptr[r]);
        P.h_col[o] = t.first; src[o] = V.orig.empty() ? t.second : V.orig[t.second]; ++o;
      }
    }
  }
  for (int64_t r = 0; r < n; ++r)
    if (P.h_diag[r] < 0) throw std::runtime_error("block has a structurally missing diagonal entry");
  P.h_perm = perm;
  std::vector<int32_t> of, ob;
  levels_of(P, true, P.lvl_f, of);
  levels_of(P, false, P.lvl_b, ob);
  make_schedule(P.lvl_f, P.steps_f);
  make_schedule(P.lvl_b, P.steps_b);
  P.rowptr.upload(P.h_rowptr, c.stream);
  P.col.upload(P.h_col, c.stream);
  P.diag.upload(P.h_diag, c.stream);
  P.src.upload(src, c.stream);
  P.perm.upload(perm, c.stream);
  P.order_fwd.upload(of, c.stream);
  P.order_bwd.upload(ob, c.stream);
  P.d_lvl_f.upload(P.lvl_f, c.stream);
  P.d_lvl_b.upload(P.lvl_b, c.stream);
  if (!P.cptr.empty()) {
    std::vector<int4> ri(n);
    for (int64_t r = 0; r < n; ++r) {
      const int64_t b = P.h_rowptr[r];
      ri[r] = make_int4((int)(b & 0xffffffffll), (int)(b >> 32), P.h_diag[r], (int)(P.h_rowptr[r + 1] - b) - P.h_diag[r] - 1);
    }
    P.rinfo.upload(ri, c.stream);
    P.d_cptr.upload(P.cptr, c.stream);
    P.barrier.alloc(1);
    P.barrier.zero(c.stream);
    P.barrier_epoch = 0;
  }
  P.val.alloc(P.nnz);
  P.work.alloc(n);
  P.yp.alloc(n);
  if (P.nblk) bl_build(c, P);
  if (P.node) {   // the two vector entries of every (permuted) node row, in both layouts
    std::vector<int32_t> a(n), b(n), cc(n), dd(n);
    for (int64_t r = 0; r < n; ++r) { a[r] = c.h_node_dx[perm[r]]; b[r] = c.h_node_dy[perm[r]]; cc[r] = 2 * perm[r]; dd[r] = 2 * perm[r] + 1; }
    P.px_ref.upload(a, c.stream); P.py_ref.upload(b, c.stream); P.px_node.upload(cc, c.stream); P.py_node.upload(dd, c.stream);
  }
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  c.tri[key] = std::move(up);
  return *c.tri[key];
}

void tri_erase(Ctx &c, int block) {
  for (auto it = c.tri.begin(); it != c.tri.end();)
    if ((it->first & 15) == block) it = c.tri.erase(it); else ++it;
}

void gather_values(Ctx &c, int64_t nnz, const int64_t *src, const double *a, double *v) {
  if (!nnz) return;
  k_gather_values<<<grid_for(nnz, 256, c.num_sms * 16), 256, 0, c.stream>>>(nnz, src, a, v);
  c.stat_launches++;
}

void tri_refresh_values(Ctx &c, TriPlan &P, const DevCSR &A) {
  if (!P.nnz) return;
  k_gather_values<<<grid_for(P.nnz, 256, c.num_sms * 16), 256, 0, c.stream>>>(P.nnz, P.src.p, A.val.p, P.val.p);
  c.stat_launches++;
  P.factored = false;
  P.bl_sgs = -1;
}

static TriArgs args_of(TriPlan &P, const double *x, double *y) {
  TriArgs A;
  A.rowptr = P.rowptr.p; A.col = P.col.p; A.diag = P.diag.p; A.perm = P.perm.p; A.val = P.val.p;
  A.x = x; A.w = P.work.p; A.yp = P.yp.p; A.y = y;
  return A;
}

void ilu0_factor(Ctx &c, TriPlan &P, const DevCSR &A) {
  tri_refresh_values(c, P, A);
  TriArgs T = args_of(P, nullptr, nullptr);
  for (const TriStep &s : P.steps_f) {
    if (s.kind == 0) {
      const int64_t r0 = P.lvl_f[s.l0], r1 = P.lvl_f[s.l0 + 1];
      k_ilu_level<<<(int)(((r1 - r0) * TG + 255) / 256), 256, 0, c.stream>>>(T, P.order_fwd.p, r0, r1);
    } else {
      k_ilu_chain<<<1, CHAIN_T, 0, c.stream>>>(T, P.order_fwd.p, P.d_lvl_f.p, s.l0, s.l1);
    }
    c.stat_launches++;
  }
  NSX_CUDA(cudaGetLastError());
  P.factored = true;
  if (P.nblk) bl_refresh(c, P, false);
}

template <bool SGS>
static void sweep(Ctx &c, TriPlan &P, double *y, const double *x) {
  if (P.nblk) {
    if (SGS && P.bl_sgs != 1) bl_refresh(c, P, true);
    bl_sweep(c, P, SGS, y, x);
    return;
  }
  TriArgs T = args_of(P, x, y);
  const int nlf = (int)P.lvl_f.size() - 1, nlb = (int)P.lvl_b.size() - 1;
  if (c.coop_sweep == 1 && !P.cptr.empty() && P.cptr.size() <= 257) {
    const int ncol = (int)P.cptr.size() - 1;
    const int4 *ri = P.rinfo.p;
    const int64_t *cp = P.d_cptr.p;
    unsigned long long *counter = P.barrier.p, epoch0 = P.barrier_epoch;
    int a_ncol = ncol;
    void *args[] = {&T, &ri, &cp, &a_ncol, &counter, &epoch0};
    static const int threads = [] { const char *e = getenv("NSX_SWEEP_THREADS"); return e ? atoi(e) : 512; }();
    const void *fn = threads == 256 ? (const void *)k_sweep_phased<SGS, 256> : threads == 384 ? (const void *)k_sweep_phased<SGS, 384>
                     : threads == 768 ? (const void *)k_sweep_phased<SGS, 768> : (const void *)k_sweep_phased<SGS, 512>;
    const int nthreads = (threads == 256 || threads == 384 || threads == 768) ? threads : 512;
    const size_t smem = (size_t)FCAP * nthreads * (sizeof(double) + sizeof(int32_t));
    static std::vector<std::pair<int, const void *>> attr_done;  // (device, kernel) pairs whose dynamic shared memory limit has been raised
    if (std::find(attr_done.begin(), attr_done.end(), std::make_pair(c.device, fn)) == attr_done.end()) {
      NSX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done.emplace_back(c.device, fn);
    }
    NSX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(c.num_sms), dim3(nthreads), args, smem, c.stream));
    P.barrier_epoch += (unsigned long long)(2 * ncol - 1) * (unsigned long long)c.num_sms;
    c.stat_launches++;
    return;
  }
  if (c.coop_sweep && nlf + nlb <= 512) {
    if (!P.coop_grid) {
      int per_sm = 0;
      NSX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_coop<SGS>, 256, 0));
      const int64_t widest = std::max<int64_t>(1, (P.n * TG / std::max(1, std::min(nlf, nlb)) + 255) / 256);  // blocks an average level can fill
      P.coop_grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)per_sm * c.num_sms, widest));
    }
    const int32_t *of = P.order_fwd.p, *ob = P.order_bwd.p;
    const int64_t *lf = P.d_lvl_f.p, *lb = P.d_lvl_b.p;
    int a_nlf = nlf, a_nlb = nlb;
    void *args[] = {&T, &of, &lf, &a_nlf, &ob, &lb, &a_nlb};
    NSX_CUDA(cudaLaunchCooperativeKernel((const void *)k_sweep_coop<SGS>, dim3(P.coop_grid), dim3(256), args, 0, c.stream));
    c.stat_launches++;
    return;
  }
  for (const TriStep &s : P.steps_f) {
    if (s.kind == 0) {
      const int64_t r0 = P.lvl_f[s.l0], r1 = P.lvl_f[s.l0 + 1];
      k_tri_level<SGS, true><<<(int)(((r1 - r0) * TG + 255) / 256), 256, 0, c.stream>>>(T, P.order_fwd.p, r0, r1);
    } else {
      k_tri_chain<SGS, true><<<1, CHAIN_T, 0, c.stream>>>(T, P.order_fwd.p, P.d_lvl_f.p, s.l0, s.l1);
    }
    c.stat_launches++;
  }
  for (const TriStep &s : P.steps_b) {
    if (s.kind == 0) {
      const int64_t r0 = P.lvl_b[s.l0], r1 = P.lvl_b[s.l0 + 1];
      k_tri_level<SGS, false><<<(int)(((r1 - r0) * TG + 255) / 256), 256, 0, c.stream>>>(T, P.order_bwd.p, r0, r1);
    } else {
      k_tri_chain<SGS, false><<<1, CHAIN_T, 0, c.stream>>>(T, P.order_bwd.p, P.d_lvl_b.p, s.l0, s.l1);
    }
    c.stat_launches++;
  }
  NSX_CUDA(cudaGetLastError());
}

void ilu0_apply(Ctx &c, TriPlan &P, double *y, const double *x) {
  if (!P.factored) throw std::logic_error("ILU(0) applied before it was factored");
  sweep<false>(c, P, y, x);
}

void sgs_apply(Ctx &c, TriPlan &P, double *y, const double *x) { sweep<true>(c, P, y, x); }

}  // namespace nsx
