// sweep_block.cu -- block-local ILU(0) / symmetric Gauss-Seidel sweeps: one CTA per block, work vector in shared memory.
//
// Replaces the same Trilinos objects as trisolve.cu (TrilinosWrappers::PreconditionILU / PreconditionSSOR,
// NSSolverStationary.hpp:160, 166, 231, 325-326; NSSolver.hpp:183, 189, 244, 250, 370, 373).  Those are additive-Schwarz
// preconditioners with overlap 0: each MPI rank factors and sweeps its own diagonal block and drops every coupling to
// another rank.  Elimination order 2 (NSX_OPT_ORDERING, the default) gives the GPU the same structure at the granularity
// it needs: the owned rows are cut into spatially compact blocks (weighted recursive coordinate bisection of the dof
// positions, one block per SM at the README size -- what `mpirun -n 148` with a geometric partitioner would produce) and
// ONE CTA owns a block for the whole application:
//   * the block's slice of the work vector (<= 8192 rows = 64 KB) lives in shared memory, so the dependent gathers of a
//     sweep never leave the SM and a dependency level costs one __syncthreads instead of a grid-wide barrier;
//   * rows inside a block follow a multicolour order re-sorted by dependency level (32 levels for Q3/Q2); a level is cut
//     into "passes" of up to NT / lanes rows, 4 .. 32 lanes per row depending on the row length;
//   * the matrix is stored per pass as an ELL slice [entry quad][row slot][lane] of FP64 values + 16-bit block-local
//     columns, followed by the reciprocal diagonals and the row slots -- 10 bytes per non-zero, each pass one
//     contiguous span that the TMA engine (cp.async.bulk completing on an mbarrier) brings into a 4-stage ring while
//     earlier passes compute.  The lower sweep streams the strictly lower entries, the upper sweep the strictly upper
//     ones: every stored non-zero of the block-diagonal part crosses HBM exactly once per application.
// The reciprocal of the diagonal is stored (one rounding away from the division the CPU oracle does).
#include <algorithm>
#include <numeric>

#include "device.cuh"
#include "tma.cuh"

namespace nsx {

namespace {

constexpr int NT = 512;           // threads per CTA
constexpr int NSTAGE = 4;         // ring stages
constexpr int PASS_ENTRIES = 4096;  // matrix entries per pass at most (values 32 KB + columns 8 KB)
constexpr int MAX_BLOCK_ROWS = 8192;

// pass header: x = offset of the value section (doubles, from the block's val_base), y = offset of the index section
// (uint16 units, from idx_base), z = rows (padded to 8) | quads << 16, w = log2(lanes) | backward << 8
__device__ __forceinline__ int hdr_rows(const int4 &h) { return h.z & 0xffff; }
__device__ __forceinline__ int hdr_quads(const int4 &h) { return (h.z >> 16) & 0xffff; }
__device__ __forceinline__ int hdr_llog(const int4 &h) { return h.w & 0xff; }
__device__ __forceinline__ int hdr_bwd(const int4 &h) { return (h.w >> 8) & 1; }

template <bool SGS>
__global__ void __launch_bounds__(NT, 1) k_sweep_block(const BlkDesc *__restrict__ blks, const int4 *__restrict__ passes, const double *__restrict__ bl_val,
                                                        const uint16_t *__restrict__ bl_idx, const int32_t *__restrict__ perm, const double *__restrict__ x,
                                                        double *__restrict__ y, const double *__restrict__ scale, double *__restrict__ v_out,
                                                        const int *__restrict__ gate, int max_rows, int max_pass, int stage_val, int stage_idx) {
  if (gate && *gate != 0) return;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: stages (values then indices, each 128-byte aligned), work vector, pass headers, barriers
  const size_t stage_bytes = (size_t)stage_val * 8 + (size_t)stage_idx * 2;
  double *xs = reinterpret_cast<double *>(smem + NSTAGE * stage_bytes);
  int4 *hdr = reinterpret_cast<int4 *>(xs + max_rows + 8);
  uint64_t *full = reinterpret_cast<uint64_t *>(hdr + max_pass);
  const BlkDesc B = blks[blockIdx.x];
  const int tid = threadIdx.x;
  const double *gval = bl_val + B.val_base;
  const uint16_t *gidx = bl_idx + B.idx_base;
  auto issue = [&](int p) {  // thread 0: bulk copies of pass p into its stage
    const int4 h = hdr[p];
    const int st = p % NSTAGE;
    const uint32_t cnt = (uint32_t)hdr_quads(h) * hdr_rows(h) * (1u << hdr_llog(h)) + hdr_rows(h);
    unsigned char *dst = smem + st * stage_bytes;
    mbar_expect_tx(&full[st], cnt * 10u);
    bulk_g2s(dst, gval + (uint32_t)h.x, cnt * 8u, &full[st]);
    bulk_g2s(dst + (size_t)stage_val * 8, gidx + (uint32_t)h.y, cnt * 2u, &full[st]);
  };
  for (int p = tid; p < B.npass; p += NT) hdr[p] = passes[B.pass0 + p];
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int p = 0; p < NSTAGE && p < B.npass; ++p) issue(p);
  // right-hand side of the block into shared memory (optionally scaled: v = x / *scale, also stored for the Krylov basis)
  {
    const double inv = scale ? 1.0 / *scale : 1.0;
    for (int i = tid; i < B.nrows; i += NT) {
      const int32_t g = perm[B.row0 + i];
      double v = x[g];
      if (scale) { v *= inv; v_out[g] = v; }
      xs[i] = v;
    }
    if (tid < 8) xs[max_rows + tid] = 0.0;  // slot of the padding rows / padding entries
  }
  __syncthreads();
  for (int p = 0; p < B.npass; ++p) {
    const int4 h = hdr[p];
    const int st = p % NSTAGE;
    const int rows = hdr_rows(h), quads = hdr_quads(h), llog = hdr_llog(h);
    const double *sv = reinterpret_cast<const double *>(smem + st * stage_bytes);
    const uint16_t *sc = reinterpret_cast<const uint16_t *>(smem + st * stage_bytes + (size_t)stage_val * 8);
    mbar_wait(&full[st], (p / NSTAGE) & 1);
    const int slot = tid >> llog, lane = tid & ((1 << llog) - 1);
    const int stride = rows << llog;   // entries per quad
    const bool active = slot < rows;
    double a0 = 0, a1 = 0;
    if (active) {
      int q = 0;
      for (; q + 1 < quads; q += 2) {
        const int i0 = q * stride + tid, i1 = i0 + stride;
        const double v0 = sv[i0], v1 = sv[i1];
        const int c0 = sc[i0], c1 = sc[i1];
        a0 = fma(v0, xs[c0], a0);
        a1 = fma(v1, xs[c1], a1);
      }
      if (q < quads) { const int i0 = q * stride + tid; a0 = fma(sv[i0], xs[sc[i0]], a0); }
    }
    double s = a0 + a1;
    for (int o = (1 << llog) >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (active && lane == 0) {
      const int tail = quads * stride;
      const int row = sc[tail + slot];
      const double dinv = sv[tail + slot];
      const double r = xs[row];
      if (!hdr_bwd(h)) xs[row] = (r - s) * dinv;            // w = (D + L)^-1 x   |  w = L^-1 x (dinv = 1)
      else xs[row] = SGS ? r - s * dinv : (r - s) * dinv;   // y = w - D^-1 U y   |  y = U^-1 w
    }
    __syncthreads();   // the level's results are visible, the stage is free
    if (tid == 0 && p + NSTAGE < B.npass) issue(p + NSTAGE);
  }
  for (int i = tid; i < B.nrows; i += NT) y[perm[B.row0 + i]] = xs[i];
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

int geometric_blocks(Ctx &c, int block, const DevCSR &A, int64_t lo, int64_t hi, int first_group, std::vector<int32_t> &grp) {
  const int64_t n = hi - lo;
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
      const int64_t d = (int64_t)c.h_cell_dofs[(size_t)cell * nd + k] - off;
      if (d < 0 || d >= nloc || d < lo || d >= hi) continue;
      sx[d - lo] += cx; sy[d - lo] += cy; cnt[d - lo]++;
    }
  }
  std::vector<Pt> pts(n);
  for (int64_t i = 0; i < n; ++i) {
    const double m = cnt[i] ? 1.0 / cnt[i] : 0.0;
    pts[i] = Pt{sx[i] * m, sy[i] * m, (double)(A.h_rowptr[lo + i + 1] - A.h_rowptr[lo + i]), (int32_t)(lo + i)};
  }
  int64_t target = c.block_rows > 0 ? c.block_rows : std::max<int64_t>(512, std::min<int64_t>(4096, n / std::max(1, c.num_sms)));
  target = std::min<int64_t>(target, MAX_BLOCK_ROWS / 2);   // weights are non-zero counts: leave room for blocks of short rows
  int64_t parts = std::max<int64_t>(1, (n + target - 1) / target);
  if (parts > c.num_sms && c.block_rows <= 0) parts = (parts + c.num_sms - 1) / c.num_sms * c.num_sms;   // whole waves
  rcb(pts, 0, n, (int)parts, first_group, grp);
  return (int)parts;
}

void bl_build(Ctx &c, TriPlan &P) {
  const int nb = P.nblk;
  std::vector<BlkDesc> blk(nb);
  std::vector<std::vector<int4>> pass(nb);
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
    std::vector<int4> &ps = pass[b];
    std::vector<int64_t> &vm = vmap[b];
    std::vector<uint16_t> &ix = idx[b];
    std::vector<int64_t> rows;
    for (int dir = 0; dir < 2; ++dir) {
      // levels ascending for the lower sweep, descending for the upper one; rows of a level are contiguous
      int64_t a = dir == 0 ? r0 : r1;
      while (dir == 0 ? a < r1 : a > r0) {
        int64_t e;
        if (dir == 0) { e = a; while (e < r1 && P.h_level[e] == P.h_level[a]) ++e; }
        else { e = a; while (e > r0 && P.h_level[e - 1] == P.h_level[a - 1]) --e; }
        const int64_t lo = std::min(a, e), hi = std::max(a, e);
        auto count = [&](int64_t r) -> int { return dir == 0 ? P.h_diag[r] : (int)(P.h_rowptr[r + 1] - P.h_rowptr[r]) - P.h_diag[r] - 1; };
        rows.resize(hi - lo);
        std::iota(rows.begin(), rows.end(), lo);
        std::stable_sort(rows.begin(), rows.end(), [&](int64_t u, int64_t v) { return count(u) > count(v); });
        size_t at = 0;
        while (at < rows.size()) {
          const int W = count(rows[at]);
          const int llog = W <= 24 ? 2 : W <= 48 ? 3 : W <= 96 ? 4 : 5;
          const int lanes = 1 << llog, quads = (W + lanes - 1) / lanes;
          int cap = NT / lanes;
          if (quads) cap = std::min(cap, PASS_ENTRIES / (quads * lanes) / 8 * 8);
          if (cap < 8) {
#pragma omp critical
            err = "block-local sweep: a matrix row is too long for one pass";
            cap = 8;
          }
          const int nr = (int)std::min<size_t>(cap, rows.size() - at);
          const int nrp = (nr + 7) / 8 * 8;
          int4 h;
          h.x = (int)vm.size(); h.y = (int)ix.size();
          h.z = nrp | (quads << 16); h.w = llog | (dir << 8);
          ps.push_back(h);
          for (int q = 0; q < quads; ++q)
            for (int s = 0; s < nrp; ++s)
              for (int l = 0; l < lanes; ++l) {
                const int e2 = q * lanes + l;
                if (s < nr && e2 < count(rows[at + s])) {
                  const int64_t r = rows[at + s];
                  const int64_t k = P.h_rowptr[r] + (dir == 0 ? e2 : P.h_diag[r] + 1 + e2);
                  vm.push_back(k << 2);
                  ix.push_back((uint16_t)(P.h_col[k] - r0));
                } else { vm.push_back(1); ix.push_back((uint16_t)dummy); }
              }
          for (int s = 0; s < nrp; ++s) {
            if (s < nr) {
              const int64_t r = rows[at + s];
              vm.push_back(((P.h_rowptr[r] + P.h_diag[r]) << 2) | (dir == 0 ? 2 : 3));
              ix.push_back((uint16_t)(r - r0));
            } else { vm.push_back(1); ix.push_back((uint16_t)dummy); }
          }
          at += nr;
        }
        a = e;
      }
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  int64_t nval = 0, npass = 0;
  int max_pass = 0, stage = 0;
  for (int b = 0; b < nb; ++b) {
    blk[b].val_base = nval; blk[b].idx_base = nval;   // both streams hold one element per slot
    blk[b].pass0 = (int)npass; blk[b].npass = (int)pass[b].size();
    blk[b].row0 = (int)P.blk_off[b]; blk[b].nrows = (int)(P.blk_off[b + 1] - P.blk_off[b]);
    if (vmap[b].size() >= ((size_t)1 << 31)) throw std::runtime_error("block-local sweep: block stream exceeds 2^31 entries");
    nval += (int64_t)vmap[b].size(); npass += blk[b].npass;
    max_pass = std::max(max_pass, blk[b].npass);
    for (const int4 &h : pass[b]) stage = std::max(stage, ((h.z >> 16) & 0xffff) * (h.z & 0xffff) * (1 << (h.w & 0xff)) + (h.z & 0xffff));
  }
  std::vector<int64_t> all_map(nval);
  std::vector<uint16_t> all_idx(nval);
  std::vector<int4> all_pass(npass);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nb; ++b) {
    std::copy(vmap[b].begin(), vmap[b].end(), all_map.begin() + blk[b].val_base);
    std::copy(idx[b].begin(), idx[b].end(), all_idx.begin() + blk[b].idx_base);
    std::copy(pass[b].begin(), pass[b].end(), all_pass.begin() + blk[b].pass0);
  }
  P.bl_nval = nval;
  P.bl_max_rows = max_rows; P.bl_max_pass = (max_pass + 3) / 4 * 4;
  P.bl_stage_val = (stage + 15) / 16 * 16;   // 128-byte multiples for both sections
  P.bl_stage_idx = (stage + 63) / 64 * 64;
  P.bl_map.upload(all_map, c.stream);
  P.bl_idx.alloc_padded(nval, 64, c.stream);
  NSX_CUDA(cudaMemcpyAsync(P.bl_idx.p, all_idx.data(), nval * sizeof(uint16_t), cudaMemcpyHostToDevice, c.stream));
  P.bl_val.alloc_padded(nval, 16, c.stream);
  P.bl_pass.upload(all_pass, c.stream);
  P.bl_blk.upload(blk, c.stream);
  P.bl_sgs = -1;
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  if (c.verbose) {
    fprintf(stderr, "[nsx] block-local sweep plan: %d blocks, rows/block max %d, %lld passes (max %d per block), %lld stream slots for %lld non-zeros (%.1f%% padding), stage %d entries\n",
            nb, max_rows, (long long)npass, max_pass, (long long)nval, (long long)P.nnz, 100.0 * (double)(nval - P.nnz) / (double)std::max<int64_t>(1, P.nnz), stage);
  }
}

void bl_refresh(Ctx &c, TriPlan &P, bool sgs) {
  if (!P.bl_nval) { P.bl_sgs = sgs; return; }
  k_bl_refresh<<<grid_for(P.bl_nval, 256, c.num_sms * 16), 256, 0, c.stream>>>(P.bl_nval, P.bl_map.p, P.val.p, P.bl_val.p, sgs ? 1 : 0);
  c.stat_launches++;
  P.bl_sgs = sgs ? 1 : 0;
}

void bl_sweep(Ctx &c, TriPlan &P, bool sgs, double *y, const double *x, const double *scale, double *v_out, const int *gate) {
  if (P.bl_sgs != (sgs ? 1 : 0)) throw std::logic_error("block-local sweep: the stream does not hold the values of this preconditioner");
  if (!P.nblk || !P.n) return;
  const size_t smem = (size_t)NSTAGE * ((size_t)P.bl_stage_val * 8 + (size_t)P.bl_stage_idx * 2) + (size_t)(P.bl_max_rows + 8) * 8 +
                      (size_t)P.bl_max_pass * 16 + NSTAGE * 8;
  if (smem > 227 * 1024) throw std::runtime_error("block-local sweep: shared memory budget exceeded");
  static std::map<std::pair<int, int>, size_t> attr;   // (device, kernel) -> limit already granted
  size_t &lim = attr[{c.device, sgs ? 1 : 0}];
  if (lim < smem) {
    if (sgs) NSX_CUDA(cudaFuncSetAttribute(k_sweep_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else NSX_CUDA(cudaFuncSetAttribute(k_sweep_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lim = smem;
  }
  if (sgs)
    k_sweep_block<true><<<P.nblk, NT, smem, c.stream>>>(P.bl_blk.p, P.bl_pass.p, P.bl_val.p, P.bl_idx.p, P.perm.p, x, y, scale, v_out, gate,
                                                         P.bl_max_rows, P.bl_max_pass, P.bl_stage_val, P.bl_stage_idx);
  else
    k_sweep_block<false><<<P.nblk, NT, smem, c.stream>>>(P.bl_blk.p, P.bl_pass.p, P.bl_val.p, P.bl_idx.p, P.perm.p, x, y, scale, v_out, gate,
                                                          P.bl_max_rows, P.bl_max_pass, P.bl_stage_val, P.bl_stage_idx);
  c.stat_launches++;
}

}  // namespace nsx
