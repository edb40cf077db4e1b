// kernels_spmv.cu -- FP64 CSR SpMV over the Jacobian blocks (HBM-bound).
//
// Replaces TrilinosWrappers::BlockSparseMatrix::vmult / SparseMatrix::vmult / vmult_add as called
// inside every Krylov solver of the reference (NSSolverStationary.cpp:589-637,
// NSSolverStationary.hpp:140, 198, 207, 288, 292, 305; NSSolver.hpp:226, 324, 345).
//
// Kernels, in the order the dispatch prefers them (NSX_OPT_STREAM_SPMV):
//   3 (default)  k_spmv_tma<true>: persistent CTAs, a producer warp keeps a shared-memory ring of row blocks full with
//                1-D TMA bulk copies, consumer sub-warps reduce each row straight from the stage; columns of F and B
//                (velocity dofs) come in aligned pairs: one column id and one 16-byte x gather per two non-zeros
//   2            k_spmv_tma<false>: same ring, products parked in the stage, then reduced per row
//   1            k_spmv_stream / k_block_spmv_stream: plain coalesced loads through shared memory
//   0            k_spmv / k_block_spmv: a sub-warp per row (also used for the small AMG level operators)
// Layout: one CSR per block (F n_u x n_u, Bt n_u x n_p, B n_p x n_u, Mp, S), 64-bit row
// pointers, 32-bit block-local columns, FP64 values.  In kernel 0 a sub-warp of G lanes owns a row, so the
// lanes of a warp stream 32/G consecutive rows = one contiguous span of the value / column
// arrays (coalesced, streamed with ld.global.cs so that x stays cached), x is gathered through
// the read-only path, and the row sum is a log2(G) shuffle reduction (deterministic order).
// Algorithmic bytes per product: 12 nnz + 8 (m+1) + 8 n + 8 m.
#include <algorithm>

#include "device.cuh"
#include "tma.cuh"

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

// ---- streaming kernel ------------------------------------------------------------------------
// A CTA owns a run of consecutive rows holding at most SNNZ non-zeros (of one matrix, or of the two
// matrices that share those rows: F and Bt).  Phase 1 streams the run's values and columns with
// fully coalesced, independent loads (SNNZ / ST per thread in flight) and parks val * x[col] in
// shared memory; phase 2 reduces each row's segment with a sub-warp.  The HBM stream is decoupled
// from the row structure, so short rows (Bt: 13, Mp: 16 per row) stream as well as long ones.
constexpr int ST = 256;      // threads per CTA
constexpr int SNNZ = 2048;   // non-zeros per CTA
constexpr int SG = 8;        // lanes per row in the reduction phase
constexpr int SROWS = 256;   // rows per CTA at most (their row pointers are staged in shared memory)

struct StreamMat {
  const int64_t *rp;
  const int32_t *col;
  const double *val;
  const double *x;
  const int32_t *pcol = nullptr;  // paired columns (one even column id per two adjacent non-zeros), or null
  int node = 0;                   // rows are velocity nodes: y(2r, 2r+1) = sum_k val_k * x(col_k, col_k + 1)
  const int2 *gpair = nullptr;    // node rows: a column id c < 0 is the ghost node -c - 1, whose pair sits at x[gpair.x], x[gpair.y]
};

__device__ __forceinline__ void stream_rows(const int32_t *__restrict__ rb, int b, const StreamMat &A1, const StreamMat &A2, bool two,
                                            double *__restrict__ y, int add, double *prod, int *srp) {
  const int r0 = rb[b], r1 = rb[b + 1], nr = r1 - r0;
  const int64_t s1 = A1.rp[r0];
  const int n1 = (int)(A1.rp[r1] - s1);
  int64_t s2 = 0;
  int n2 = 0;
  if (two) { s2 = A2.rp[r0]; n2 = (int)(A2.rp[r1] - s2); }
  // row pointers of the block (relative), loaded alongside the streams so that phase 2 touches no global memory
  for (int i = threadIdx.x; i <= nr; i += ST) {
    srp[i] = (int)(A1.rp[r0 + i] - s1);
    if (two) srp[SROWS + 1 + i] = n1 + (int)(A2.rp[r0 + i] - s2);
  }
#pragma unroll 8
  for (int k = threadIdx.x; k < n1; k += ST) prod[k] = __ldcs(A1.val + s1 + k) * __ldg(A1.x + __ldcs(A1.col + s1 + k));
  if (two) {
#pragma unroll 2
    for (int k = threadIdx.x; k < n2; k += ST) prod[n1 + k] = __ldcs(A2.val + s2 + k) * __ldg(A2.x + __ldcs(A2.col + s2 + k));
  }
  __syncthreads();
  const int lane = threadIdx.x & (SG - 1);
  const int passes = (nr + ST / SG - 1) / (ST / SG);
  for (int ps = 0; ps < passes; ++ps) {
    const int i = ps * (ST / SG) + threadIdx.x / SG;
    double s = 0;
    if (i < nr) {
      for (int k = srp[i] + lane; k < srp[i + 1]; k += SG) s += prod[k];
      if (two)
        for (int k = srp[SROWS + 1 + i] + lane; k < srp[SROWS + 2 + i]; k += SG) s += prod[k];
    }
#pragma unroll
    for (int o = SG / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, SG);
    if (i < nr && lane == 0) y[r0 + i] = add ? y[r0 + i] + s : s;
  }
}

__global__ void __launch_bounds__(ST) k_spmv_stream(const int32_t *__restrict__ rb, StreamMat A, double *__restrict__ y, int add) {
  __shared__ double prod[SNNZ];
  __shared__ int srp[2 * (SROWS + 1)];
  stream_rows(rb, blockIdx.x, A, A, false, y, add, prod, srp);
}

// jacobian_matrix.vmult in one launch: CTAs [0, nb_u) own velocity rows (F and Bt), the rest own pressure rows (B)
__global__ void __launch_bounds__(ST) k_block_spmv_stream(const int32_t *__restrict__ rb_u, int nb_u, StreamMat F, StreamMat Bt,
                                                          const int32_t *__restrict__ rb_p, StreamMat B, double *__restrict__ y, int64_t n_u) {
  __shared__ double prod[SNNZ];
  __shared__ int srp[2 * (SROWS + 1)];
  if ((int)blockIdx.x < nb_u) stream_rows(rb_u, blockIdx.x, F, Bt, true, y, 0, prod, srp);
  else stream_rows(rb_p, blockIdx.x - nb_u, B, B, false, y + n_u, 0, prod, srp);
}

// rows grouped into runs of at most SNNZ non-zeros (summed over the one or two matrices sharing the rows)
void build_row_blocks(Ctx &c, const DevCSR &A1, const DevCSR *A2, DevBuf<int32_t> &rb, int &nb) {
  std::vector<int32_t> h{0};
  int64_t acc = 0;
  for (int64_t r = 0; r < A1.nrows; ++r) {
    int64_t len = A1.h_rowptr[r + 1] - A1.h_rowptr[r];
    if (A2) len += A2->h_rowptr[r + 1] - A2->h_rowptr[r];
    if (len > SNNZ) throw std::runtime_error("matrix row too long for the streaming SpMV");
    if (acc + len > SNNZ || r - h.back() >= SROWS) { h.push_back((int32_t)r); acc = 0; }
    acc += len;
  }
  h.push_back((int32_t)A1.nrows);
  nb = (int)h.size() - 1;
  rb.upload(h, c.stream);
}

// ---- TMA-fed persistent kernel ----------------------------------------------------------------
// The same two phases, but the value / column streams are brought into shared memory by the TMA
// engine (cp.async.bulk, 1-D bulk copies completing on an mbarrier) through a TS-stage ring that a
// dedicated producer warp keeps full.  The HBM stream no longer depends on how long the consumer
// warps stall on the x gather.  Four persistent CTAs per SM walk the row blocks grid-strided.
// Ring parameters from the sweep in profiles/ (B200, 300x100 Q3/Q2 Jacobian): 2 stages x 1792
// non-zeros, 4 CTAs per SM, 16 lanes per row (8 on paired columns): 100 us per block product.
#ifndef NSX_TS
#define NSX_TS 2
#endif
#ifndef NSX_TNNZ
#define NSX_TNNZ 1792
#endif
#ifndef NSX_DL
#define NSX_DL 16
#endif
#ifndef NSX_TMINB
#define NSX_TMINB 4
#endif
constexpr int TS = NSX_TS;            // ring stages
constexpr int TNNZ = NSX_TNNZ;        // non-zeros per row block
#ifndef NSX_TROWS
#define NSX_TROWS (NSX_TNNZ / 8)
#endif
constexpr int TROWS = NSX_TROWS;      // rows per row block at most
constexpr int TCAP = TNNZ + 32;       // elements per stage (two segments, each padded to a multiple of 8 at both ends)
constexpr int TRP = TROWS + 4;        // row pointers per matrix per stage (range padded to even ends)
constexpr int TCONS = 256;            // consumer threads (8 warps) + 1 producer warp
constexpr size_t TMA_STAGE = (size_t)TCAP * 12 + 2 * TRP * 8 + sizeof(RowBlockDesc);  // values, columns, row pointers, descriptor
constexpr size_t TMA_SMEM = TS * TMA_STAGE + 2 * TS * 8;  // ring + (full, empty) barriers

struct TmaStage {  // views into one stage of the ring
  double *val; int32_t *col; int64_t *rp1, *rp2; const RowBlockDesc *desc;
};
__device__ __forceinline__ TmaStage stage_of(unsigned char *ring, int st) {
  unsigned char *b = ring + (size_t)st * TMA_STAGE;
  TmaStage s;
  s.val = (double *)b;
  s.rp1 = (int64_t *)(b + (size_t)TCAP * 8);
  s.rp2 = s.rp1 + TRP;
  s.col = (int32_t *)(b + (size_t)TCAP * 8 + 2 * TRP * 8);
  s.desc = (const RowBlockDesc *)(b + (size_t)TCAP * 12 + 2 * TRP * 8);
  return s;
}

// DIRECT consumer: a sub-warp of DLX lanes walks each row straight out of the stage: values and columns come from shared
// memory (filled by the TMA engine, never written by the SM), x through the read-only path, the sum stays in registers.
// No product round trip through shared memory, no CTA-wide barrier.  PAIRED: the first matrix' columns come in aligned
// pairs (both velocity components of a node group) -- one column id, one 16-byte gather of x and one 16-byte read of the
// values per two non-zeros.
template <int DLX, bool PAIRED>
__device__ __forceinline__ void direct_rows(const TmaStage &S, const RowBlockDesc &d, bool two, const double *__restrict__ x1, const double *__restrict__ x2,
                                            double *__restrict__ yy, int add, int tid) {
  const double *v = S.val;
  const int32_t *cidx = S.col;
  const int64_t s1 = d.a1, s2 = d.a2 - d.c1;  // stage index = global index - s
  const int nr = d.r1 - d.r0, roff = d.r0 & 1;
  const int sl = tid & (DLX - 1);
  const unsigned hmask = DLX == 32 ? 0xffffffffu : ((1u << (DLX & 31)) - 1u) << (tid & (32 - DLX) & 31);  // sub-warps leave the row loop independently
  for (int i = tid / DLX; i < nr; i += TCONS / DLX) {
    const int b1 = (int)(S.rp1[roff + i] - s1), e1 = (int)(S.rp1[roff + i + 1] - s1);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (PAIRED) {
      const int pb = b1 >> 1, pe = e1 >> 1;
      const double2 *__restrict__ vv = reinterpret_cast<const double2 *>(v);
      const double2 z = make_double2(0.0, 0.0);
      for (int k = pb + sl; k - sl < pe; k += 4 * DLX) {
        const bool p0 = k < pe, p1 = k + DLX < pe, p2 = k + 2 * DLX < pe, p3 = k + 3 * DLX < pe;
        const int c0 = p0 ? cidx[k] : 0, c1 = p1 ? cidx[k + DLX] : 0, c2 = p2 ? cidx[k + 2 * DLX] : 0, c3 = p3 ? cidx[k + 3 * DLX] : 0;
        const double2 x0 = p0 ? __ldg(reinterpret_cast<const double2 *>(x1 + c0)) : z, xb = p1 ? __ldg(reinterpret_cast<const double2 *>(x1 + c1)) : z;
        const double2 xc = p2 ? __ldg(reinterpret_cast<const double2 *>(x1 + c2)) : z, xd = p3 ? __ldg(reinterpret_cast<const double2 *>(x1 + c3)) : z;
        const double2 v0 = p0 ? vv[k] : z, vb = p1 ? vv[k + DLX] : z, vc = p2 ? vv[k + 2 * DLX] : z, vd = p3 ? vv[k + 3 * DLX] : z;
        a0 += v0.x * x0.x; a1 += v0.y * x0.y; a2 += vb.x * xb.x; a3 += vb.y * xb.y;
        a0 += vc.x * xc.x; a1 += vc.y * xc.y; a2 += vd.x * xd.x; a3 += vd.y * xd.y;
      }
    } else {
      // four predicated gathers in flight per lane and trip; the trip count is uniform over the sub-warp
      for (int k = b1 + sl; k - sl < e1; k += 4 * DLX) {
        const bool p0 = k < e1, p1 = k + DLX < e1, p2 = k + 2 * DLX < e1, p3 = k + 3 * DLX < e1;
        const int c0 = p0 ? cidx[k] : 0, c1 = p1 ? cidx[k + DLX] : 0, c2 = p2 ? cidx[k + 2 * DLX] : 0, c3 = p3 ? cidx[k + 3 * DLX] : 0;
        const double x0 = p0 ? __ldg(x1 + c0) : 0.0, x1v = p1 ? __ldg(x1 + c1) : 0.0, x2v = p2 ? __ldg(x1 + c2) : 0.0, x3v = p3 ? __ldg(x1 + c3) : 0.0;
        const double v0 = p0 ? v[k] : 0.0, v1 = p1 ? v[k + DLX] : 0.0, v2 = p2 ? v[k + 2 * DLX] : 0.0, v3 = p3 ? v[k + 3 * DLX] : 0.0;
        a0 += v0 * x0; a1 += v1 * x1v; a2 += v2 * x2v; a3 += v3 * x3v;
      }
    }
    if (two) {
      const int b2 = (int)(S.rp2[roff + i] - s2), e2 = (int)(S.rp2[roff + i + 1] - s2);
      for (int q = b2 + sl; q - sl < e2; q += 2 * DLX) {
        const bool p0 = q < e2, p1 = q + DLX < e2;
        const int c0 = p0 ? cidx[q] : 0, c1 = p1 ? cidx[q + DLX] : 0;
        const double x0 = p0 ? __ldg(x2 + c0) : 0.0, x1v = p1 ? __ldg(x2 + c1) : 0.0;
        const double v0 = p0 ? v[q] : 0.0, v1 = p1 ? v[q + DLX] : 0.0;
        a2 += v0 * x0; a3 += v1 * x1v;
      }
    }
    double s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = DLX >> 1; o > 0; o >>= 1) s += __shfl_down_sync(hmask, s, o, DLX);
    if (sl == 0) yy[d.r0 + i] = add ? yy[d.r0 + i] + s : s;
  }
}

// NODE consumer: the matrix is the scalar K of F = K (x) I_2 over the velocity nodes, vectors are in the node layout (entry 2b / 2b + 1
// = x / y component of node b): a column id is 2b (16-byte aligned), one value multiplies the (x, y) pair, a row writes a pair.
template <bool GHOSTS>
__device__ __forceinline__ double2 node_pair(const double *__restrict__ x, const int2 *__restrict__ gp, int c) {
  if (!GHOSTS || c >= 0) return __ldg(reinterpret_cast<const double2 *>(x + c));
  const int2 g = __ldg(gp + (-c - 1));
  return make_double2(__ldg(x + g.x), __ldg(x + g.y));
}

template <int DLX, bool GHOSTS>
__device__ __forceinline__ void direct_rows_node(const TmaStage &S, const RowBlockDesc &d, const double *__restrict__ x1, const int2 *__restrict__ gp,
                                                 double *__restrict__ yy, int add, int tid) {
  const double *v = S.val;
  const int32_t *cidx = S.col;
  const int64_t s1 = d.a1;
  const int nr = d.r1 - d.r0, roff = d.r0 & 1;
  const int sl = tid & (DLX - 1);
  const unsigned hmask = DLX == 32 ? 0xffffffffu : ((1u << (DLX & 31)) - 1u) << (tid & (32 - DLX) & 31);
  const double2 z = make_double2(0.0, 0.0);
  for (int i = tid / DLX; i < nr; i += TCONS / DLX) {
    const int b1 = (int)(S.rp1[roff + i] - s1), e1 = (int)(S.rp1[roff + i + 1] - s1);
    double ax = 0, ay = 0, bx = 0, by = 0;
    for (int k = b1 + sl; k - sl < e1; k += 4 * DLX) {
      const bool p0 = k < e1, p1 = k + DLX < e1, p2 = k + 2 * DLX < e1, p3 = k + 3 * DLX < e1;
      const int c0 = p0 ? cidx[k] : 0, c1 = p1 ? cidx[k + DLX] : 0, c2 = p2 ? cidx[k + 2 * DLX] : 0, c3 = p3 ? cidx[k + 3 * DLX] : 0;
      const double2 x0 = p0 ? node_pair<GHOSTS>(x1, gp, c0) : z, xb = p1 ? node_pair<GHOSTS>(x1, gp, c1) : z;
      const double2 xc = p2 ? node_pair<GHOSTS>(x1, gp, c2) : z, xd = p3 ? node_pair<GHOSTS>(x1, gp, c3) : z;
      const double v0 = p0 ? v[k] : 0.0, vb = p1 ? v[k + DLX] : 0.0, vc = p2 ? v[k + 2 * DLX] : 0.0, vd = p3 ? v[k + 3 * DLX] : 0.0;
      ax += v0 * x0.x; ay += v0 * x0.y; bx += vb * xb.x; by += vb * xb.y;
      ax += vc * xc.x; ay += vc * xc.y; bx += vd * xd.x; by += vd * xd.y;
    }
    double sx = ax + bx, sy = ay + by;
#pragma unroll
    for (int o = DLX >> 1; o > 0; o >>= 1) { sx += __shfl_down_sync(hmask, sx, o, DLX); sy += __shfl_down_sync(hmask, sy, o, DLX); }
    if (sl == 0) {
      double2 *out = reinterpret_cast<double2 *>(yy) + (d.r0 + i);
      if (add) { const double2 old = *out; sx += old.x; sy += old.y; }
      *out = make_double2(sx, sy);
    }
  }
}

// Row blocks `first, first + stride, ...` of a list whose entries are of kind 0 (matrices M[0] and, if it has
// non-zeros there, M[1] share the rows; y offset 0) or kind 1 (matrix M[2] alone; y offset yoff1).
// NODE: 0 general rows (paired / unpaired columns), 1 node rows on one rank, 2 node rows with ghost nodes -- separate instantiations
// keep the register budget of 4 CTAs per SM for each
template <bool DIRECT, int NODE>
__device__ __forceinline__ void tma_rows(const RowBlockDesc *__restrict__ desc, int first, int stride, int nb, const StreamMat *M,
                                         double *__restrict__ y, int64_t yoff1, int add, unsigned char *ring, uint64_t *full, uint64_t *empty,
                                         RowBlockDesc *pdesc, int l2_hints) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int nit = first < nb ? (nb - first + stride - 1) / stride : 0;
  if (tid >= TCONS) {
    const uint64_t policy = l2_policy_evict_first();   // matrix streams are read once per product: leave the L2 to the vectors
    // producer warp: descriptors are fetched 32 at a time by the whole warp, lane 0 feeds the ring
    for (int it0 = 0; it0 < nit; it0 += 32) {
      __syncwarp();
      if (it0 + lane < nit) pdesc[lane] = desc[first + (it0 + lane) * stride];
      __syncwarp();
      for (int it = it0; it < min(nit, it0 + 32); ++it) {
        const int st = it % TS, use = it / TS;
        if (use > 0 && lane == 0) mbar_wait(&empty[st], (use - 1) & 1);
        __syncwarp();
        const RowBlockDesc d = pdesc[it - it0];
        const StreamMat &A1 = M[d.kind ? 2 : 0], &A2 = M[1];
        const bool two = d.kind == 0 && d.c2 > 0;
        const TmaStage S = stage_of(ring, st);
        const int ra = d.r0 & ~1, rc = ((d.r1 + 2) & ~1) - ra;  // even-aligned range of row pointers covering [r0, r1]
        // the stage's (up to) seven bulk copies are issued by seven lanes in one go
        void *dst = nullptr; const void *src = nullptr; uint32_t bytes = 0;
        switch (lane) {
          case 0: dst = (void *)S.desc; src = desc + first + it * stride; bytes = sizeof(RowBlockDesc); break;
          case 1: dst = S.rp1; src = A1.rp + ra; bytes = rc * 8; break;
          case 2: dst = S.val; src = A1.val + d.a1; bytes = d.c1 * 8; break;
          case 3: if (A1.pcol) { dst = S.col; src = A1.pcol + (d.a1 >> 1); bytes = d.c1 * 2; } else { dst = S.col; src = A1.col + d.a1; bytes = d.c1 * 4; } break;
          case 4: if (two) { dst = S.rp2; src = A2.rp + ra; bytes = rc * 8; } break;
          case 5: if (two) { dst = S.val + d.c1; src = A2.val + d.a2; bytes = d.c2 * 8; } break;
          case 6: if (two) { dst = S.col + d.c1; src = A2.col + d.a2; bytes = d.c2 * 4; } break;
          default: break;
        }
        if (lane == 0)
          mbar_expect_tx(&full[st], (uint32_t)d.c1 * (A1.pcol ? 10u : 12u) + (two ? (uint32_t)d.c2 * 12u : 0u) + (uint32_t)rc * 8u * (two ? 2u : 1u) +
                                        (uint32_t)sizeof(RowBlockDesc));
        __syncwarp();
        if (bytes) { if (l2_hints) bulk_g2s_hint(dst, src, bytes, &full[st], policy); else bulk_g2s(dst, src, bytes, &full[st]); }
      }
    }
    return;
  }
  for (int it = 0; it < nit; ++it) {
    const int st = it % TS, use = it / TS;
    const TmaStage S = stage_of(ring, st);
    mbar_wait(&full[st], use & 1);
    const RowBlockDesc d = *S.desc;
    const bool two = d.kind == 0 && d.c2 > 0;
    const double *__restrict__ x1 = M[d.kind ? 2 : 0].x, *__restrict__ x2 = M[1].x;
    const bool paired = DIRECT && M[d.kind ? 2 : 0].pcol != nullptr;
    double *v = S.val;
    const int32_t *cidx = S.col;
    double *yy = y + (d.kind ? yoff1 : 0);
    const int64_t s1 = d.a1, s2 = d.a2 - d.c1;  // stage index = global index - s
    const int nr = d.r1 - d.r0, roff = d.r0 & 1;
    if (DIRECT) {
      // lanes per row follow the block's mean row length (host-side choice): four predicated entries per lane and trip
      if (NODE) {
        const int2 *gp = M[0].gpair;
        if (d.lanes >= 8) direct_rows_node<8, NODE == 2>(S, d, x1, gp, yy, add, tid); else direct_rows_node<4, NODE == 2>(S, d, x1, gp, yy, add, tid);
      } else if (paired) direct_rows<NSX_DL / 2, true>(S, d, two, x1, x2, yy, add, tid);
      else if (d.lanes >= 16) direct_rows<16, false>(S, d, two, x1, x2, yy, add, tid);
      else direct_rows<8, false>(S, d, two, x1, x2, yy, add, tid);
    } else {
    // phase 1: products in place (every consumer thread busy, four independent gathers each)
#pragma unroll 4
    for (int k = d.o1 + tid; k < d.o1 + d.n1; k += TCONS) v[k] *= __ldg(x1 + cidx[k]);
    if (two) {
#pragma unroll 2
      for (int k = d.c1 + d.o2 + tid; k < d.c1 + d.o2 + d.n2; k += TCONS) v[k] *= __ldg(x2 + cidx[k]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // phase 2: a sub-warp of SG lanes reduces each row's segment(s); row pointers come from the stage
    constexpr int L = SG;
    const int sl = tid & (L - 1);
    const int rpp = TCONS / L, passes = (nr + rpp - 1) / rpp;
    for (int ps = 0; ps < passes; ++ps) {
      const int i = ps * rpp + tid / L;
      double s = 0;
      if (i < nr) {
        const int b1 = (int)(S.rp1[roff + i] - s1), e1 = (int)(S.rp1[roff + i + 1] - s1);
        for (int k = b1 + sl; k < e1; k += L) s += v[k];
        if (two) {
          const int b2 = (int)(S.rp2[roff + i] - s2), e2 = (int)(S.rp2[roff + i + 1] - s2);
          for (int k = b2 + sl; k < e2; k += L) s += v[k];
        }
      }
#pragma unroll
      for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, L);
      if (i < nr && sl == 0) yy[d.r0 + i] = add ? yy[d.r0 + i] + s : s;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes (products) before the TMA's next write
    }
    // release the stage (this warp's reads are done; the mbarrier orders them before the TMA's next write)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
}

// y = A x (kind-0 blocks with one matrix), or jacobian_matrix.vmult with M = {F, Bt, B}: the row blocks of the
// velocity rows (F + Bt, kind 0) come first in the list, then those of the pressure rows (B, kind 1)
template <bool DIRECT, int NODE>
__global__ void __launch_bounds__(TCONS + 32, NSX_TMINB) k_spmv_tma(const RowBlockDesc *__restrict__ desc, int nb, StreamMat M0, StreamMat M1, StreamMat M2,
                                                            double *__restrict__ y, int64_t yoff1, int add, int l2_hints) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ RowBlockDesc pdesc[32];
  __shared__ StreamMat M[3];
  uint64_t *full = (uint64_t *)(tma_smem + TS * TMA_STAGE), *empty = full + TS;
  if (threadIdx.x == 0) {
    M[0] = M0; M[1] = M1; M[2] = M2;
    for (int s = 0; s < TS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TCONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  tma_rows<DIRECT, NODE>(desc, blockIdx.x, gridDim.x, nb, M, y, yoff1, add, tma_smem, full, empty, pdesc, l2_hints);
}

void append_row_descs(std::vector<RowBlockDesc> &h, const DevCSR &A1, const DevCSR *A2, int kind) {
  int64_t acc = 0;
  int64_t r0 = 0;
  auto close = [&](int64_t r1) {
    if (r1 == r0) return;
    RowBlockDesc d{};
    const int64_t s1 = A1.h_rowptr[r0], e1 = A1.h_rowptr[r1];
    d.a1 = s1 & ~(int64_t)7; d.o1 = (int)(s1 - d.a1); d.n1 = (int)(e1 - s1);
    d.c1 = d.n1 ? (int)(((e1 + 7) & ~(int64_t)7) - d.a1) : 0;
    if (A2) {
      const int64_t s2 = A2->h_rowptr[r0], e2 = A2->h_rowptr[r1];
      d.a2 = s2 & ~(int64_t)7; d.o2 = (int)(s2 - d.a2); d.n2 = (int)(e2 - s2);
      d.c2 = d.n2 ? (int)(((e2 + 7) & ~(int64_t)7) - d.a2) : 0;
    }
    d.r0 = (int)r0; d.r1 = (int)r1; d.kind = kind;
    const double mean = (double)(d.n1 + d.n2) / (double)(r1 - r0);
    d.lanes = mean > 40 ? 16 : mean > 14 ? 8 : 4;
    h.push_back(d);
    r0 = r1;
  };
  for (int64_t r = 0; r < A1.nrows; ++r) {
    int64_t len = A1.h_rowptr[r + 1] - A1.h_rowptr[r];
    if (A2) len += A2->h_rowptr[r + 1] - A2->h_rowptr[r];
    if (len > TNNZ) throw std::runtime_error("matrix row too long for the TMA-fed SpMV");
    if (acc + len > TNNZ || r - r0 >= TROWS) { close(r); acc = 0; }
    acc += len;
  }
  close(A1.nrows);
}

// Paired columns of a block whose columns are velocity dofs: in the FESystem numbering the two components of a node group
// sit next to each other, so a row's columns come in aligned pairs (2k, 2k + 1).  Verified entry by entry on the device
// layout of the columns (ghost columns are shifted by the owned pressure count on a partitioned system); a block that
// does not pair keeps the scalar path.
void build_pairs(Ctx &c, DevCSR &A, bool cols_are_p) {
  if (A.pair_state) return;
  A.pair_state = -1;
  if (cols_are_p || A.nnz == 0 || (A.nnz & 1)) return;
  const int64_t own = c.n_u, shift = c.n_p;
  const bool baked = c.n_ug + c.n_pg > 0;
  std::vector<int32_t> pc((size_t)(A.nnz / 2));
  for (int64_t i = 0; i < A.nrows; ++i) {
    const int64_t b = A.h_rowptr[i], e = A.h_rowptr[i + 1];
    if ((b & 1) || (e & 1)) return;
    for (int64_t k = b; k < e; k += 2) {
      int64_t c0 = A.h_col[k], c1 = A.h_col[k + 1];
      if (baked) { if (c0 >= own) c0 += shift; if (c1 >= own) c1 += shift; }
      if ((c0 & 1) || c1 != c0 + 1) return;
      pc[(size_t)(k >> 1)] = (int32_t)c0;
    }
  }
  A.pcol.alloc_padded(pc.size(), 16, c.stream);
  NSX_CUDA(cudaMemcpyAsync(A.pcol.p, pc.data(), pc.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  A.pair_state = 1;
}

inline const int32_t *pairs_for(Ctx &c, const DevCSR &A_, const double *x) {
  DevCSR &A = const_cast<DevCSR &>(A_);
  if (c.stream_spmv != 3) return nullptr;
  if (!A.pair_state) build_pairs(c, A, &A_ == &c.Bt || &A_ == &c.Mp || &A_ == &c.S);
  return (A.pair_state == 1 && ((uintptr_t)x & 15) == 0) ? A.pcol.p : nullptr;
}

void tma_attr_once(const Ctx &c) {   // cudaFuncSetAttribute applies per device
  static std::vector<int> done;
  if (std::find(done.begin(), done.end(), c.device) != done.end()) return;
  NSX_CUDA(cudaFuncSetAttribute(k_spmv_tma<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
  NSX_CUDA(cudaFuncSetAttribute(k_spmv_tma<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
  NSX_CUDA(cudaFuncSetAttribute(k_spmv_tma<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
  NSX_CUDA(cudaFuncSetAttribute(k_spmv_tma<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
  done.push_back(c.device);
}

inline int pick_group(const DevCSR &A) {
  const double avg = A.nrows ? (double)A.nnz / (double)A.nrows : 0.0;
  if (avg > 48) return 16;
  if (avg > 12) return 8;
  return 4;
}

}  // namespace

void spmv(Ctx &c, const DevCSR &A_, const double *x, double *y, bool add) {
  // ghost import of the input first (no-op on one GPU): a rank that owns no row of this block still has to post its sends
  halo_exchange(c, (&A_ == &c.F || &A_ == &c.B || &A_ == &c.Fd || &A_ == &c.Kn) ? 0 : 1, x, &A_ == &c.Kn);
  if (!A_.nrows) return;
  spmv_local(c, A_, x, y, add);
}

void spmv_local(Ctx &c, const DevCSR &A_, const double *x, double *y, bool add) {
  DevCSR &A = const_cast<DevCSR &>(A_);
  if (!A.nrows) return;
  if (c.stream_spmv >= 2 && !A.h_rowptr.empty() && A.nrows < (int64_t)1 << 31) {
    tma_attr_once(c);
    if (!A.ndesc) {
      std::vector<RowBlockDesc> h;
      append_row_descs(h, A, nullptr, 0);
      A.ndesc = (int)h.size();
      A.desc.upload(h, c.stream);
    }
    const int grid = std::min(A.ndesc, NSX_TMINB * c.num_sms);
    StreamMat M{A.rowptr.p, A.col.p, A.val.p, x, (&A_ == &c.F || &A_ == &c.B) ? pairs_for(c, A_, x) : nullptr};
    if (&A_ == &c.Kn) {
      if (c.stream_spmv != 3 || (((uintptr_t)x | (uintptr_t)y) & 15)) throw std::logic_error("the node view of F needs the direct TMA SpMV and 16-byte aligned vectors");
      M.node = 1; M.gpair = c.n_ug ? c.node_gpair.p : nullptr;
    }
    if (M.node && M.gpair) k_spmv_tma<true, 2><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(A.desc.p, A.ndesc, M, M, M, y, 0, add ? 1 : 0, c.l2_hints ? 1 : 0);
    else if (M.node) k_spmv_tma<true, 1><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(A.desc.p, A.ndesc, M, M, M, y, 0, add ? 1 : 0, c.l2_hints ? 1 : 0);
    else if (c.stream_spmv == 3) k_spmv_tma<true, 0><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(A.desc.p, A.ndesc, M, M, M, y, 0, add ? 1 : 0, c.l2_hints ? 1 : 0);
    else k_spmv_tma<false, 0><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(A.desc.p, A.ndesc, M, M, M, y, 0, add ? 1 : 0, c.l2_hints ? 1 : 0);
    c.stat_launches++; c.stat_spmv++;
    return;
  }
  if (c.stream_spmv && !A.h_rowptr.empty() && A.nrows < (int64_t)1 << 31) {
    if (!A.nrb) build_row_blocks(c, A, nullptr, A.rb, A.nrb);
    k_spmv_stream<<<A.nrb, ST, 0, c.stream>>>(A.rb.p, StreamMat{A.rowptr.p, A.col.p, A.val.p, x}, y, add ? 1 : 0);
    c.stat_launches++; c.stat_spmv++;
    return;
  }
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
  halo_exchange(c, 0, x);
  halo_exchange(c, 1, x + c.n_u);
  if (c.stream_spmv >= 2 && n < (int64_t)1 << 31) {
    tma_attr_once(c);
    if (!c.ndesc_u) {
      std::vector<RowBlockDesc> h;
      append_row_descs(h, c.F, &c.Bt, 0);
      append_row_descs(h, c.B, nullptr, 1);
      c.ndesc_u = (int)h.size();
      c.desc_u.upload(h, c.stream);
    }
    const int grid = std::min(c.ndesc_u, NSX_TMINB * c.num_sms);
    const StreamMat MF{c.F.rowptr.p, c.F.col.p, c.F.val.p, x, pairs_for(c, c.F, x)}, MBt{c.Bt.rowptr.p, c.Bt.col.p, c.Bt.val.p, x + c.n_u, nullptr},
        MB{c.B.rowptr.p, c.B.col.p, c.B.val.p, x, pairs_for(c, c.B, x)};
    if (c.stream_spmv == 3) k_spmv_tma<true, 0><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(c.desc_u.p, c.ndesc_u, MF, MBt, MB, y, c.n_u, 0, c.l2_hints ? 1 : 0);
    else k_spmv_tma<false, 0><<<grid, TCONS + 32, TMA_SMEM, c.stream>>>(c.desc_u.p, c.ndesc_u, MF, MBt, MB, y, c.n_u, 0, c.l2_hints ? 1 : 0);
    c.stat_launches++; c.stat_spmv++;
    return;
  }
  if (c.stream_spmv && n < (int64_t)1 << 31) {
    if (!c.nrb_u) { build_row_blocks(c, c.F, &c.Bt, c.rb_u, c.nrb_u); build_row_blocks(c, c.B, nullptr, c.rb_p, c.nrb_p); }
    k_block_spmv_stream<<<c.nrb_u + c.nrb_p, ST, 0, c.stream>>>(c.rb_u.p, c.nrb_u, StreamMat{c.F.rowptr.p, c.F.col.p, c.F.val.p, x},
                                                                StreamMat{c.Bt.rowptr.p, c.Bt.col.p, c.Bt.val.p, x + c.n_u}, c.rb_p.p,
                                                                StreamMat{c.B.rowptr.p, c.B.col.p, c.B.val.p, x}, y, c.n_u);
    c.stat_launches++; c.stat_spmv++;
    return;
  }
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

// ---- probes: upper bounds for the SpMV on this matrix (measurement only) ----------------------
namespace {
// what 0: stream val + col only ; 1: + gather x[col] ; both reduce everything into one number per thread
template <int WHAT>
__global__ void __launch_bounds__(256) k_probe(int64_t nnz, const int32_t *__restrict__ col, const double *__restrict__ val,
                                               const double *__restrict__ x, double *sink) {
  double acc = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; k + 3 * stride < nnz; k += 4 * stride) {
    const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + stride), v2 = __ldcs(val + k + 2 * stride), v3 = __ldcs(val + k + 3 * stride);
    const int32_t c0 = __ldcs(col + k), c1 = __ldcs(col + k + stride), c2 = __ldcs(col + k + 2 * stride), c3 = __ldcs(col + k + 3 * stride);
    if (WHAT == 0) acc += v0 * c0 + v1 * c1 + v2 * c2 + v3 * c3;
    else acc += v0 * __ldg(x + c0) + v1 * __ldg(x + c1) + v2 * __ldg(x + c2) + v3 * __ldg(x + c3);
  }
  for (; k < nnz; k += stride) acc += __ldcs(val + k) * (WHAT == 0 ? (double)__ldcs(col + k) : __ldg(x + __ldcs(col + k)));
  if (acc == 1.2345e-300) *sink = acc;
}
}  // namespace

void spmv_probe(Ctx &c, const DevCSR &A, int what, const double *x, double *sink) {
  const int grid = c.num_sms * 8;
  if (what == 0) k_probe<0><<<grid, 256, 0, c.stream>>>(A.nnz, A.col.p, A.val.p, x, sink);
  else k_probe<1><<<grid, 256, 0, c.stream>>>(A.nnz, A.col.p, A.val.p, x, sink);
  c.stat_launches++;
}

void extract_diag(Ctx &c, const DevCSR &A, double *d, double *dinv) {
  const int grid = (int)((A.nrows + 255) / 256);
  k_extract_diag<<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.diag.p, A.val.p, d, dinv);
  c.stat_launches++;
}

}  // namespace nsx
