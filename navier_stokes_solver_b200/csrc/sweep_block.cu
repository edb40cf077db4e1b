// sweep_block.cu -- block-local ILU(0) / symmetric Gauss-Seidel sweeps: one CTA per block, work vector in shared memory.
//
// Replaces the same Trilinos objects as trisolve.cu (TrilinosWrappers::PreconditionILU / PreconditionSSOR,
// NSSolverStationary.hpp:160, 166, 231, 325-326; NSSolver.hpp:183, 189, 244, 250, 370, 373).  Those are additive-Schwarz
// preconditioners with overlap 0: each MPI rank factors and sweeps its own diagonal block and drops every coupling to
// another rank.  Elimination order 2 (NSX_OPT_ORDERING, the default) gives the GPU the same structure at the granularity
// it needs: the owned rows are cut into spatially compact blocks (weighted recursive coordinate bisection of the dof
// positions, two blocks per SM at the README size -- what `mpirun -n 296` with a geometric partitioner would produce) and
// ONE CTA owns a block for the whole application (two CTAs are resident per SM and hide each other's level latency):
//   * the block's slice of the work vector (<= 4096 rows) lives in shared memory, so the dependent gathers of a sweep never
//     leave the SM and a dependency level costs one named barrier among 16 warps instead of a grid-wide barrier;
//   * rows inside a block follow a multicolour order re-sorted by dependency level (32 levels for Q3/Q2); a level is one
//     or a few "passes" of up to 512 lanes, a row owning ceil(entries / Q) consecutive lanes of a warp (Q = 4 entries per
//     lane), reduced by a segmented shuffle reduction; only the end of a level costs a barrier, and only among the 16
//     consumer warps of the CTA (bar.sync);
//   * the matrix is stored per pass as FP64 values + 16-bit block-local columns, lane-major, followed by the reciprocal
//     diagonals, the lane descriptors and the row slots (about 12 bytes per non-zero including the padding), each pass
//     two contiguous spans that a producer warp brings into a shared-memory ring with the TMA engine (cp.async.bulk
//     completing on an mbarrier), up to 16 passes ahead of the consumers; ring offsets and the pass whose consumption
//     frees each region are fixed by the host when the plan is built.  The lower sweep streams the strictly lower
//     entries, the upper sweep the strictly upper ones: every stored non-zero of the block-diagonal part crosses HBM
//     exactly once per application.
//   * NODE = true (decouple.cu): the rows are velocity nodes of F = K (x) I_2; one matrix value serves the (x, y) pair of a node, the
//     work vector holds a double2 per row, and the vector entries of a pair come from two index arrays (reference or node layout).
// Elimination order 3 keeps the blocks and takes Ifpack's natural order inside (many short levels; the stronger ILU(0)).
// The reciprocal of the diagonal is stored (one rounding away from the division the CPU oracle does).
#include <algorithm>
#include <numeric>

#include "device.cuh"
#include "tma.cuh"

namespace nsx {

namespace {

constexpr int NC = 512;           // consumer threads (16 warps); one more warp feeds the ring
constexpr int NT = NC + 32;
constexpr int NSLOT = 16;         // passes in flight at most (full / empty barrier pairs)
constexpr int BLOCKS_PER_SM = 2;   // CTAs resident per SM (shared memory is split between them)
constexpr int MAX_BLOCK_ROWS = 4096;
constexpr int PASS_BYTES_MAX = 32 * 1024;
constexpr size_t SMEM_PER_CTA = (size_t)(227 * 1024 / BLOCKS_PER_SM - 1024) / 128 * 128;

// Pass format ("lane-split rows"): a row with c entries in this sweep's triangle owns k = max(1, ceil(c / Q)) consecutive
// lanes of one warp, Q entries each; rows are packed into warps first-fit.  With T = 32 x warps lanes in the pass:
//   value section : val[Q][T] (entry q of lane t at q T + t), then rdiag[rows]           (doubles)
//   index section : col[Q][T], meta[T], rowid[rows]                                        (uint16)
//   meta: bit 15 = first lane of its row, bits 14..5 = row slot in the pass, bits 4..0 = lanes of the row behind this one
// Header (two int4): {value offset, index offset (elements, from the block's bases), ring offset / 16, pass to wait for}
//                    {T | Q << 16, rows | rounds << 16 | flags << 24, value count, index count}; flags: 1 upper sweep, 2 level ends
struct PassHdr { int val_off, idx_off, ring16, wait; int tq, rrf, val_cnt, idx_cnt; };

// NODE: a row is a velocity node and stands for the vector entries perm[row] (x) and perm_y[row] (y); one matrix value serves both
template <bool SGS, bool NODE>
__global__ void __launch_bounds__(NT, BLOCKS_PER_SM) k_sweep_block(const BlkDesc *__restrict__ blks, const PassHdr *__restrict__ passes, const double *__restrict__ bl_val,
                                                        const uint16_t *__restrict__ bl_idx, const int32_t *__restrict__ perm, const int32_t *__restrict__ perm_y,
                                                        const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ scale,
                                                        double *__restrict__ v_out, const int *__restrict__ gate, int max_rows, int max_pass, int ring_bytes, int l2_hints) {
  if (gate && *gate != 0) return;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: ring, work vector (+ 8 zero slots for padding rows / entries), pass headers, barriers
  constexpr int W = NODE ? 2 : 1;   // doubles per row of the work vector
  double *xs = reinterpret_cast<double *>(smem + ring_bytes);
  PassHdr *hdr = reinterpret_cast<PassHdr *>(xs + (size_t)W * (max_rows + 8));
  uint64_t *full = reinterpret_cast<uint64_t *>(hdr + max_pass), *empty = full + NSLOT;
  const BlkDesc B = blks[blockIdx.x];
  const int tid = threadIdx.x;
  for (int p = tid; p < B.npass; p += NT) hdr[p] = passes[B.pass0 + p];
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // right-hand side of the block into shared memory (optionally scaled: v = x / *scale, also stored for the Krylov basis)
  {
    double inv = 1.0;
    if (scale) { const double a = *scale; inv = isfinite(a) ? 1.0 / a : 0.0; }
    for (int i = tid; i < B.nrows; i += NT) {
      const int32_t g = perm[B.row0 + i];
      if (NODE) {   // perm / perm_y: the vector entries of the node's two components
        const int32_t gy = perm_y[B.row0 + i];
        double2 v = make_double2(x[g], x[gy]);
        if (scale) { v.x = inv * v.x; v.y = inv * v.y; v_out[g] = v.x; v_out[gy] = v.y; }
        reinterpret_cast<double2 *>(xs)[i] = v;
      } else {
        double v = x[g];
        if (scale) { v = inv * v; v_out[g] = v; }
        xs[i] = v;
      }
    }
    if (tid < 8 * W) xs[(size_t)W * max_rows + tid] = 0.0;
  }
  __syncthreads();
  if (tid >= NC) {
    // producer warp: one lane walks the passes; a pass is issued as soon as the ring region the host assigned to it is free
    if (tid == NC) {
      const double *gval = bl_val + B.val_base;
      const uint16_t *gidx = bl_idx + B.idx_base;
      const uint64_t policy = l2_policy_evict_first();   // the matrix stream is read once per application: keep the Krylov basis in L2
      for (int p = 0; p < B.npass; ++p) {
        const PassHdr h = hdr[p];
        if (h.wait >= 0) mbar_wait(&empty[h.wait % NSLOT], (h.wait / NSLOT) & 1);
        const int st = p % NSLOT;
        unsigned char *dst = smem + (size_t)h.ring16 * 16;
        mbar_expect_tx(&full[st], (uint32_t)h.val_cnt * 8u + (uint32_t)h.idx_cnt * 2u);
        if (l2_hints) {
          bulk_g2s_hint(dst, gval + (uint32_t)h.val_off, (uint32_t)h.val_cnt * 8u, &full[st], policy);
          bulk_g2s_hint(dst + (size_t)h.val_cnt * 8, gidx + (uint32_t)h.idx_off, (uint32_t)h.idx_cnt * 2u, &full[st], policy);
        } else {
          bulk_g2s(dst, gval + (uint32_t)h.val_off, (uint32_t)h.val_cnt * 8u, &full[st]);
          bulk_g2s(dst + (size_t)h.val_cnt * 8, gidx + (uint32_t)h.idx_off, (uint32_t)h.idx_cnt * 2u, &full[st]);
        }
      }
    }
    return;
  }
  const int lane = tid & 31;
  for (int p = 0; p < B.npass; ++p) {
    const PassHdr h = hdr[p];
    const int st = p % NSLOT;
    const int T = h.tq & 0xffff, Q = h.tq >> 16;
    const int rows = h.rrf & 0xffff, rounds = (h.rrf >> 16) & 0xff, flags = h.rrf >> 24;
    mbar_wait(&full[st], (p / NSLOT) & 1);   // every warp, also the idle ones: nobody runs more than NSLOT passes ahead
    if (tid < T) {   // warp-uniform: T is a multiple of 32
      const double *sv = reinterpret_cast<const double *>(smem + (size_t)h.ring16 * 16);
      const uint16_t *sc = reinterpret_cast<const uint16_t *>(sv + h.val_cnt);
      const int m = sc[Q * T + tid];
      const int behind = m & 31;
      if (NODE) {
        const double2 *xs2 = reinterpret_cast<const double2 *>(xs);
        double ax = 0, ay = 0, bx = 0, by = 0;
        int q = 0;
        for (; q + 1 < Q; q += 2) {
          const int i0 = q * T + tid, i1 = i0 + T;
          const double v0 = sv[i0], v1 = sv[i1];
          const double2 x0 = xs2[sc[i0]], x1 = xs2[sc[i1]];
          ax = fma(v0, x0.x, ax); ay = fma(v0, x0.y, ay);
          bx = fma(v1, x1.x, bx); by = fma(v1, x1.y, by);
        }
        if (q < Q) { const int i0 = q * T + tid; const double v0 = sv[i0]; const double2 x0 = xs2[sc[i0]]; ax = fma(v0, x0.x, ax); ay = fma(v0, x0.y, ay); }
        double sx = ax + bx, sy = ay + by;
        for (int r = 0, o = 1; r < rounds; ++r, o <<= 1) {
          const double ux = __shfl_down_sync(0xffffffffu, sx, o), uy = __shfl_down_sync(0xffffffffu, sy, o);
          if (o <= behind) { sx += ux; sy += uy; }
        }
        if (m & 0x8000) {
          const int slot = (m >> 5) & 0x3ff;
          const int row = sc[Q * T + T + slot];
          const double rdiag = sv[Q * T + slot];
          double2 r0 = xs2[row];
          if (!(flags & 1)) { r0.x = (r0.x - sx) * rdiag; r0.y = (r0.y - sy) * rdiag; }
          else if (SGS) { r0.x = r0.x - sx * rdiag; r0.y = r0.y - sy * rdiag; }
          else { r0.x = (r0.x - sx) * rdiag; r0.y = (r0.y - sy) * rdiag; }
          reinterpret_cast<double2 *>(xs)[row] = r0;
        }
      } else {
      double a0 = 0, a1 = 0;
      int q = 0;
      for (; q + 1 < Q; q += 2) {
        const int i0 = q * T + tid, i1 = i0 + T;
        const double v0 = sv[i0], v1 = sv[i1];
        const int c0 = sc[i0], c1 = sc[i1];
        a0 = fma(v0, xs[c0], a0);
        a1 = fma(v1, xs[c1], a1);
      }
      if (q < Q) { const int i0 = q * T + tid; a0 = fma(sv[i0], xs[sc[i0]], a0); }
      double s = a0 + a1;
      for (int r = 0, o = 1; r < rounds; ++r, o <<= 1) {   // segmented reduction: the first lane of a row collects its lanes
        const double up = __shfl_down_sync(0xffffffffu, s, o);
        if (o <= behind) s += up;
      }
      if (m & 0x8000) {
        const int slot = (m >> 5) & 0x3ff;
        const int row = sc[Q * T + T + slot];
        const double rdiag = sv[Q * T + slot];
        const double r0 = xs[row];
        if (!(flags & 1)) xs[row] = (r0 - s) * rdiag;            // w = (D + L)^-1 x   |  w = L^-1 x (rdiag = 1)
        else xs[row] = SGS ? r0 - s * rdiag : (r0 - s) * rdiag;  // y = w - D^-1 U y   |  y = U^-1 w
      }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);   // this warp has finished reading the pass
    if (flags & 2) asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");   // the level's results are visible to the next level
  }
  for (int i = tid; i < B.nrows; i += NC) {
    if (NODE) { const double2 v = reinterpret_cast<const double2 *>(xs)[i]; y[perm[B.row0 + i]] = v.x; y[perm_y[B.row0 + i]] = v.y; }
    else y[perm[B.row0 + i]] = xs[i];
  }
}

// kind (low 2 bits of the map): 0 value, 1 zero, 2 reciprocal diagonal of a lower-sweep pass, 3 of an upper-sweep pass
__global__ void k_bl_refresh(int64_t n, const int64_t *__restrict__ map, const double *__restrict__ val, double *__restrict__ out, int sgs) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = map[k];
    const int kind = (int)(m & 3);
    double v = 0.0;
    if (kind == 0) v = val[m >> 2];
    else if (kind == 2) v = sgs ? 1.0 / val[m >> 2] : 1.0;
    else if (kind == 3) v = 1.0 / val[m >> 2];
    out[k] = v;
  }
}

struct Pt { double x, y, w; int32_t id; };

// weighted recursive coordinate bisection: pts[lo, hi) into `parts` groups of (nearly) equal weight
void rcb(std::vector<Pt> &pts, int64_t lo, int64_t hi, int parts, int first, std::vector<int32_t> &grp) {
  if (parts <= 1 || hi - lo <= 1) {
    for (int64_t i = lo; i < hi; ++i) grp[pts[i].id] = first;
    return;
  }
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300, wsum = 0;
  for (int64_t i = lo; i < hi; ++i) {
    x0 = std::min(x0, pts[i].x); x1 = std::max(x1, pts[i].x);
    y0 = std::min(y0, pts[i].y); y1 = std::max(y1, pts[i].y);
    wsum += pts[i].w;
  }
  const bool alongx = (x1 - x0) >= (y1 - y0);
  std::sort(pts.begin() + lo, pts.begin() + hi, [alongx](const Pt &a, const Pt &b) {
    const double ka = alongx ? a.x : a.y, kb = alongx ? b.x : b.y;
    if (ka != kb) return ka < kb;
    const double sa = alongx ? a.y : a.x, sb = alongx ? b.y : b.x;
    if (sa != sb) return sa < sb;
    return a.id < b.id;
  });
  const int left = parts / 2;
  const double target = wsum * left / parts;
  double acc = 0;
  int64_t cut = lo;
  while (cut < hi - 1 && acc + pts[cut].w * 0.5 < target) acc += pts[cut++].w;
  cut = std::max(lo + 1, std::min(hi - 1, cut));
  rcb(pts, lo, cut, left, first, grp);
  rcb(pts, cut, hi, parts - left, first + left, grp);
}

}  // namespace

int geometric_blocks(Ctx &c, int block, const std::vector<int64_t> &rowptr, int64_t lo, int64_t hi, int first_group, std::vector<int32_t> &grp,
                     const std::vector<int32_t> *row_dof) {
  const int64_t n = hi - lo;
  std::vector<int32_t> dof_row;   // node rows: dof -> row (the x dof of a node names it)
  if (row_dof) {
    dof_row.assign(c.n_u + c.n_ug, -1);
    for (size_t r = 0; r < row_dof->size(); ++r) dof_row[(*row_dof)[r]] = (int32_t)r;
  }
  const bool pressure = block != NSX_BLOCK_F;
  const int64_t off = pressure ? c.n_u + c.n_ug : 0;           // position of the block's dofs in the cell table's numbering
  const int64_t nloc = pressure ? c.n_p + c.n_pg : c.n_u + c.n_ug;
  // position of a dof = mean centroid of the cells that hold it (both components of a node coincide)
  std::vector<double> sx(n, 0.0), sy(n, 0.0);
  std::vector<int32_t> cnt(n, 0);
  const int nd = c.fe.ndofs, nv = c.fe.nvpc;
  for (int64_t cell = 0; cell < c.ncells; ++cell) {
    double cx = 0, cy = 0;
    for (int v = 0; v < nv; ++v) { cx += c.h_cell_vertices[((size_t)cell * nv + v) * 2]; cy += c.h_cell_vertices[((size_t)cell * nv + v) * 2 + 1]; }
    cx /= nv; cy /= nv;
    for (int k = 0; k < nd; ++k) {
      int64_t d = (int64_t)c.h_cell_dofs[(size_t)cell * nd + k] - off;
      if (d < 0 || d >= nloc) continue;
      if (row_dof) { d = dof_row[d]; if (d < 0) continue; }   // node rows take the position of their x dof
      if (d < lo || d >= hi) continue;
      sx[d - lo] += cx; sy[d - lo] += cy; cnt[d - lo]++;
    }
  }
  std::vector<Pt> pts(n);
  for (int64_t i = 0; i < n; ++i) {
    const double m = cnt[i] ? 1.0 / cnt[i] : 0.0;
    pts[i] = Pt{sx[i] * m, sy[i] * m, (double)(rowptr[lo + i + 1] - rowptr[lo + i]), (int32_t)(lo + i)};
  }
  // one block per SM while that gives blocks of 512 .. 4096 rows; fewer blocks below, whole waves of blocks above (the
  // weights are non-zero counts, so a block of short rows may hold more rows than the average: MAX_BLOCK_ROWS leaves room)
  int64_t parts;
  if (c.block_rows > 0) parts = (n + c.block_rows - 1) / c.block_rows;
  else {
    // BLOCKS_PER_SM co-resident CTAs per SM hide each other's level latency; blocks of 512 .. 2048 rows, whole waves above.  On N
    // ranks every rank cuts its own range: one block per SM there, so that the total number of blocks (and with it the strength of
    // the block-Jacobi preconditioner) does not grow faster than the machine
    const int per_sm = c.nranks > 1 ? 1 : BLOCKS_PER_SM;
    const int64_t slots = (int64_t)std::max(1, c.num_sms) * per_sm, per = n / slots;
    // (no block below 512 rows: the unsteady aSIMPLE's single ILU(0) application per iteration on the reference's mesh -- 52 k
    // velocity nodes -- converges under the reference's iteration cap with 102 blocks of 512 nodes and misses it with 203 of 256)
    if (per <= 512) parts = std::min<int64_t>(slots, (n + 256) / 512);
    else parts = slots * ((per + 2047) / 2048);
  }
  parts = std::max<int64_t>(1, parts);
  rcb(pts, 0, n, (int)parts, first_group, grp);
  return (int)parts;
}

static size_t bl_fixed_smem(int max_rows, int max_pass, bool node) {
  return (size_t)(max_rows + 8) * 8 * (node ? 2 : 1) + (size_t)max_pass * sizeof(PassHdr) + 2 * NSLOT * 8;
}

void bl_build(Ctx &c, TriPlan &P) {
  const int nb = P.nblk;
  std::vector<BlkDesc> blk(nb);
  std::vector<std::vector<PassHdr>> pass(nb);
  std::vector<std::vector<int64_t>> vmap(nb);
  std::vector<std::vector<uint16_t>> idx(nb);
  int max_rows = 0;
  for (int b = 0; b < nb; ++b) max_rows = std::max<int>(max_rows, (int)(P.blk_off[b + 1] - P.blk_off[b]));
  max_rows = (max_rows + 1) & ~1;   // keeps the areas behind the work vector 16-byte aligned
  if (max_rows > MAX_BLOCK_ROWS) throw std::runtime_error("block-local sweep: a block exceeds the shared-memory work vector; lower NSX_OPT_BLOCK_ROWS");
  const int dummy = max_rows;   // slot of padding rows / entries (kept at zero by the kernel)
  std::string err;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nb; ++b) {
    const int64_t r0 = P.blk_off[b], r1 = P.blk_off[b + 1];
    std::vector<PassHdr> &ps = pass[b];
    std::vector<int64_t> &vm = vmap[b];
    std::vector<uint16_t> &ix = idx[b];
    std::vector<int64_t> rows;
    struct Seg { int64_t row; int lane0, k; };
    std::vector<Seg> segs;
    std::vector<int> room;   // free lanes per warp of the pass under construction
    std::vector<uint16_t> bank_used;
    std::vector<int> left;
    for (int dir = 0; dir < 2; ++dir) {
      // levels ascending for the lower sweep, descending for the upper one; the rows of a level are contiguous
      int64_t a = dir == 0 ? r0 : r1;
      while (dir == 0 ? a < r1 : a > r0) {
        int64_t e = a;
        if (dir == 0) { while (e < r1 && P.h_level[e] == P.h_level[a]) ++e; }
        else { while (e > r0 && P.h_level[e - 1] == P.h_level[a - 1]) --e; }
        const int64_t lo = std::min(a, e), hi = std::max(a, e);
        auto count = [&](int64_t r) -> int { return dir == 0 ? P.h_diag[r] : (int)(P.h_rowptr[r + 1] - P.h_rowptr[r]) - P.h_diag[r] - 1; };
        rows.resize(hi - lo);
        std::iota(rows.begin(), rows.end(), lo);
        std::stable_sort(rows.begin(), rows.end(), [&](int64_t u, int64_t v) { return count(u) > count(v); });
        const int maxc = rows.empty() ? 0 : count(rows[0]);
        // entries per lane: 4, or up to NSX_OPT_SWEEP_Q when the level's longest row has that many (fewer lanes per row, fewer
        // shuffle rounds, more padding in the stream)
        int Q = 4;
        while (Q < c.sweep_q && Q < maxc) Q *= 2;
        while (Q * 32 < maxc) Q *= 2;   // a row's lanes stay inside one warp
        if (maxc == 0) Q = 1;
        // warps of a pass: what the ring takes (values + columns + lane / row tables of a pass <= PASS_BYTES_MAX)
        const int max_warps = std::max(1, std::min(NC / 32, (PASS_BYTES_MAX - 64) / (32 * (Q * 10 + 2 + 10))));
        size_t at = 0;
        while (at < rows.size()) {
          // first-fit packing of the rows' lane groups into the warps of one pass
          segs.clear(); room.clear();
          int maxk = 1;
          size_t next = at;
          for (; next < rows.size(); ++next) {
            const int k = std::max(1, (count(rows[next]) + Q - 1) / Q);
            int w = 0;
            while (w < (int)room.size() && room[w] < k) ++w;
            if (w == (int)room.size()) {
              if ((int)room.size() == max_warps) break;
              room.push_back(32);
            }
            segs.push_back(Seg{rows[next], w * 32 + (32 - room[w]), k});
            room[w] -= k;
            maxk = std::max(maxk, k);
          }
          const int T = 32 * (int)room.size(), nr = (int)segs.size(), nrp = (nr + 7) / 8 * 8;
          int rounds = 0;
          while ((1 << rounds) < maxk) ++rounds;
          PassHdr h;
          h.val_off = (int)vm.size(); h.idx_off = (int)ix.size(); h.ring16 = 0; h.wait = -1;
          h.tq = T | (Q << 16);
          h.rrf = nrp | (rounds << 16) | ((dir | (next == rows.size() ? 2 : 0)) << 24);
          h.val_cnt = Q * T + nrp; h.idx_cnt = Q * T + T + nrp;
          if ((size_t)h.val_cnt * 8 + (size_t)h.idx_cnt * 2 > (size_t)PASS_BYTES_MAX) {
#pragma omp critical
            err = "block-local sweep: a pass exceeds the ring";
          }
          ps.push_back(h);
          const size_t v0 = vm.size(), i0 = ix.size();
          vm.resize(v0 + h.val_cnt, 1);            // kind 1 = zero
          ix.resize(i0 + h.idx_cnt, (uint16_t)dummy);
          for (int t = 0; t < T; ++t) ix[i0 + (size_t)Q * T + t] = 0;   // meta of unused lanes: not a head, nothing behind
          // Entries of a row may sit in any of its lanes' slots: place them so that the lanes of a half-warp gather from
          // distinct 8-byte banks of the work vector where possible (a 64-bit shared-memory load is served per half-warp,
          // 16 banks of 8 bytes)
          bank_used.assign((size_t)(T / 16) * Q, 0);
          for (int sgi = 0; sgi < nr; ++sgi) {
            const Seg &sg = segs[sgi];
            const int64_t r = sg.row;
            const int cnt = count(r);
            const int64_t kb = P.h_rowptr[r] + (dir == 0 ? 0 : P.h_diag[r] + 1);
            left.resize(cnt);
            std::iota(left.begin(), left.end(), 0);
            for (int l = 0; l < sg.k; ++l) {
              const int t = sg.lane0 + l;
              for (int q = 0; q < Q && !left.empty(); ++q) {
                uint16_t &used = bank_used[(size_t)(t / 16) * Q + q];
                size_t pick = 0;
                for (size_t u = 0; u < left.size(); ++u)
                  if (!(used >> ((P.h_col[kb + left[u]] - r0) & 15) & 1)) { pick = u; break; }
                const int e2 = left[pick];
                left.erase(left.begin() + pick);
                used |= (uint16_t)(1u << ((P.h_col[kb + e2] - r0) & 15));
                vm[v0 + (size_t)q * T + t] = (kb + e2) << 2;
                ix[i0 + (size_t)q * T + t] = (uint16_t)(P.h_col[kb + e2] - r0);
              }
              ix[i0 + (size_t)Q * T + t] = (uint16_t)((l == 0 ? 0x8000 : 0) | (sgi << 5) | (sg.k - 1 - l));
            }
            vm[v0 + (size_t)Q * T + sgi] = ((P.h_rowptr[r] + P.h_diag[r]) << 2) | (dir == 0 ? 2 : 3);
            ix[i0 + (size_t)Q * T + T + sgi] = (uint16_t)(r - r0);
          }
          at = next;
        }
        a = e;
      }
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  int64_t nval = 0, nidx = 0, npass = 0;
  int max_pass = 0;
  for (int b = 0; b < nb; ++b) {
    blk[b].val_base = nval; blk[b].idx_base = nidx;
    blk[b].pass0 = (int)npass; blk[b].npass = (int)pass[b].size();
    blk[b].row0 = (int)P.blk_off[b]; blk[b].nrows = (int)(P.blk_off[b + 1] - P.blk_off[b]);
    if (vmap[b].size() >= ((size_t)1 << 31) || idx[b].size() >= ((size_t)1 << 31)) throw std::runtime_error("block-local sweep: block stream exceeds 2^31 entries");
    nval += (int64_t)vmap[b].size(); nidx += (int64_t)idx[b].size(); npass += blk[b].npass;
    max_pass = std::max(max_pass, blk[b].npass);
  }
  // ring schedule: every pass gets a fixed region of the shared-memory ring and the number of the pass whose consumption
  // frees it (passes are consumed in order), so the producer warp needs no bookkeeping
  const size_t fixed = bl_fixed_smem(max_rows, max_pass, P.node);
  if (fixed + 2 * (size_t)PASS_BYTES_MAX > SMEM_PER_CTA) throw std::runtime_error("block-local sweep: shared memory budget exceeded");
  const int ring_bytes = (int)((SMEM_PER_CTA - fixed) / 128 * 128);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nb; ++b) {
    std::vector<PassHdr> &ps = pass[b];
    std::vector<int> beg(ps.size()), end(ps.size());
    int head = 0, oldest = 0;   // oldest: first pass still resident
    for (int p = 0; p < (int)ps.size(); ++p) {
      const int bytes = (ps[p].val_cnt * 8 + ps[p].idx_cnt * 2 + 127) / 128 * 128;
      if (head + bytes > ring_bytes) head = 0;
      int wait = p - NSLOT;   // the barrier pair of this slot must have been released
      // evict every resident pass that overlaps [head, head + bytes)
      while (oldest < p) {
        bool overlap = false;
        for (int q = oldest; q < p; ++q)
          if (beg[q] < head + bytes && head < end[q]) { overlap = true; break; }
        if (!overlap) break;
        wait = std::max(wait, oldest);
        ++oldest;
      }
      beg[p] = head; end[p] = head + bytes;
      ps[p].ring16 = head / 16;
      ps[p].wait = wait;
      if (wait >= 0) oldest = std::max(oldest, wait + 1);
      head += bytes;
    }
  }
  std::vector<int64_t> all_map(nval);
  std::vector<uint16_t> all_idx(nidx);
  std::vector<PassHdr> all_pass(npass);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nb; ++b) {
    std::copy(vmap[b].begin(), vmap[b].end(), all_map.begin() + blk[b].val_base);
    std::copy(idx[b].begin(), idx[b].end(), all_idx.begin() + blk[b].idx_base);
    std::copy(pass[b].begin(), pass[b].end(), all_pass.begin() + blk[b].pass0);
  }
  P.bl_nval = nval;
  P.bl_max_rows = max_rows; P.bl_max_pass = max_pass; P.bl_ring = ring_bytes;
  P.bl_map.upload(all_map, c.stream);
  P.bl_idx.alloc_padded(nidx, 64, c.stream);
  NSX_CUDA(cudaMemcpyAsync(P.bl_idx.p, all_idx.data(), nidx * sizeof(uint16_t), cudaMemcpyHostToDevice, c.stream));
  P.bl_val.alloc_padded(nval, 16, c.stream);
  P.bl_pass.alloc(npass * sizeof(PassHdr));
  NSX_CUDA(cudaMemcpyAsync(P.bl_pass.p, all_pass.data(), npass * sizeof(PassHdr), cudaMemcpyHostToDevice, c.stream));
  P.bl_blk.upload(blk, c.stream);
  P.bl_sgs = -1;
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  if (c.verbose) {
    fprintf(stderr, "[nsx] block-local sweep plan: %d blocks, rows/block max %d, %lld passes (max %d per block), %.1f MB of values + %.1f MB of indices for %lld non-zeros (%.2f B per non-zero), ring %d KB\n",
            nb, max_rows, (long long)npass, max_pass, nval * 8e-6, nidx * 2e-6, (long long)P.nnz, (nval * 8.0 + nidx * 2.0) / (double)std::max<int64_t>(1, P.nnz), ring_bytes / 1024);
  }
}

void bl_refresh(Ctx &c, TriPlan &P, bool sgs) {
  if (!P.bl_nval) { P.bl_sgs = sgs; return; }
  k_bl_refresh<<<grid_for(P.bl_nval, 256, c.num_sms * 16), 256, 0, c.stream>>>(P.bl_nval, P.bl_map.p, P.val.p, P.bl_val.p, sgs ? 1 : 0);
  c.stat_launches++;
  P.bl_sgs = sgs ? 1 : 0;
}

void bl_sweep(Ctx &c, TriPlan &P, bool sgs, double *y, const double *x, const double *scale, double *v_out, const int *gate, bool node_layout) {
  if (P.bl_sgs != (sgs ? 1 : 0)) throw std::logic_error("block-local sweep: the stream does not hold the values of this preconditioner");
  if (!P.nblk || !P.n) return;
  const size_t smem = (size_t)P.bl_ring + bl_fixed_smem(P.bl_max_rows, P.bl_max_pass, P.node);
  if (node_layout && !P.node) throw std::logic_error("only a node plan works on vectors in the node layout");
  const int32_t *px = P.node ? (node_layout ? P.px_node.p : P.px_ref.p) : P.perm.p, *py = P.node ? (node_layout ? P.py_node.p : P.py_ref.p) : nullptr;
  typedef void (*Kernel)(const BlkDesc *, const PassHdr *, const double *, const uint16_t *, const int32_t *, const int32_t *, const double *, double *, const double *, double *,
                         const int *, int, int, int, int);
  const int which = (sgs ? 1 : 0) + (P.node ? 2 : 0);
  const Kernel kernels[4] = {k_sweep_block<false, false>, k_sweep_block<true, false>, k_sweep_block<false, true>, k_sweep_block<true, true>};
  static std::map<std::pair<int, int>, size_t> attr;   // (device, kernel) -> limit already granted
  size_t &lim = attr[{c.device, which}];
  if (lim < smem) {
    NSX_CUDA(cudaFuncSetAttribute(kernels[which], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lim = smem;
  }
  const PassHdr *ph = reinterpret_cast<const PassHdr *>(P.bl_pass.p);
  kernels[which]<<<P.nblk, NT, smem, c.stream>>>(P.bl_blk.p, ph, P.bl_val.p, P.bl_idx.p, px, py, x, y, scale, v_out, gate, P.bl_max_rows, P.bl_max_pass, P.bl_ring, c.l2_hints ? 1 : 0);
  c.stat_launches++;
}

}  // namespace nsx
