// device.cuh -- context and shared declarations of the sm_100a library behind include/nsx.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsx.h"
#include "fe.hpp"

namespace nsx {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct NoConvergence : std::runtime_error {
  int last_step; double last_residual;
  NoConvergence(int s, double r, const char *who = "")
      : std::runtime_error(std::string("SolverControl::NoConvergence") + (who[0] ? std::string(" in ") + who : std::string()) + " at step " + std::to_string(s) +
                           ", residual " + std::to_string(r)),
        last_step(s), last_residual(r) {}
};

#define NSX_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      throw nsx::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count) NSX_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  // `pad` extra elements behind the `count` the buffer reports (bulk copies round their ranges up to 16 bytes)
  // (zeroed on the caller's stream: a legacy-stream memset would race with later copies on a non-blocking stream)
  void alloc_padded(size_t count, size_t pad, cudaStream_t s) {
    release();
    NSX_CUDA(cudaMalloc(&p, (count + pad) * sizeof(T)));
    NSX_CUDA(cudaMemsetAsync(p, 0, (count + pad) * sizeof(T), s));
    n = count;
  }
  void upload(const T *h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) NSX_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T> &h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void zero(cudaStream_t s) { if (n) NSX_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// one row block of the TMA-fed SpMV: 16-byte aligned element ranges of the value / column arrays of the
// one or two matrices that share rows [r0, r1): aligned start a, aligned count c, front pad o, real count n
struct alignas(16) RowBlockDesc { long long a1, a2; int c1, o1, n1, c2, o2, n2, r0, r1, kind, lanes, pad[2]; };  // 64 bytes: bulk-copied into the ring

// Device CSR block (pattern fixed once, values rewritten by every assembly).
struct DevCSR {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  int64_t row0 = 0;  // global (block-local) index of the first stored row: ranks store their owned rows only
  // Bt on a partitioned system also keeps the rows of the GHOST velocity dofs behind the owned ones (assembled redundantly from
  // the local cells, complete for owned pressure columns): what EpetraExt imports for B diag(F)^-1 Bt in the reference
  int64_t nrows_ext = 0, nnz_ext = 0;   // rows / entries including that extension (0: none)
  DevBuf<int64_t> rowptr;
  DevBuf<int32_t> col;
  DevBuf<double> val;
  DevBuf<int32_t> diag;  // offset of the diagonal inside each row (square blocks), -1 if absent
  DevBuf<int32_t> pcol;       // paired columns (kernels_spmv.cu build_pairs), when the block's columns pair up
  int pair_state = 0;         // 0 not examined, 1 paired, -1 does not pair
  DevBuf<RowBlockDesc> desc;  // row blocks of the TMA-fed SpMV
  int ndesc = 0;
  DevBuf<int32_t> rb;   // row blocks of the streaming SpMV (runs of rows with a bounded non-zero count)
  int nrb = 0;
  std::vector<int64_t> h_rowptr;  // host copies of the pattern (symbolic work is done on the host)
  std::vector<int32_t> h_col;
  int max_row = 0;
};

// Level-scheduled triangular machinery shared by ILU(0) and SGS on one square block.  The plan
// holds the rank-local part of the block (couplings that cross an owned-range boundary are
// dropped: the reference's Ifpack preconditioners run with overlap 0), permuted by the
// elimination order (natural as Ifpack, or multicolour), rows grouped by dependency level.
struct TriStep { int kind, l0, l1; };   // kind 0: one wide level as a grid launch; 1: levels [l0,l1) chained in one CTA
// one block of the block-local sweep: its slice of the value / index streams, its passes, its rows
struct alignas(16) BlkDesc { long long val_base, idx_base; int pass0, npass, row0, nrows; };
struct TriPlan {
  int64_t n = 0, nnz = 0;
  int ordering = 0;
  DevBuf<int64_t> rowptr;               // permuted, filtered pattern, columns sorted, permuted column ids
  DevBuf<int32_t> col;
  DevBuf<int32_t> diag;                 // offset of the diagonal inside each row
  DevBuf<int64_t> src;                  // position in the block's value array of each entry
  DevBuf<int32_t> perm;                 // new -> old row
  DevBuf<int32_t> order_fwd, order_bwd; // rows sorted by forward / backward dependency level
  DevBuf<int64_t> d_lvl_f, d_lvl_b;     // level pointers into order_*
  DevBuf<double> val;                   // filtered values (SGS) or LU factors (ILU)
  DevBuf<double> work, yp;              // intermediate vectors (permuted numbering)
  bool factored = false;
  long long built_at = -1;              // solve (Ctx::solve_seq) whose matrix values the numeric data hold; -1: none yet
  int coop_grid = 0;                    // grid of the cooperative sweep (0: not sized yet)
  // colour-phased persistent sweep (multicolour order): rows of a colour are contiguous in the permuted numbering
  std::vector<int64_t> cptr;            // colour pointers (empty for the natural order)
  DevBuf<int64_t> d_cptr;
  DevBuf<int4> rinfo;                   // per row: start of the row in the plan (2 x int32 = int64), lower count, upper count
  DevBuf<unsigned long long> barrier;   // arrival counter of the in-kernel grid barrier (monotonic across launches)
  unsigned long long barrier_epoch = 0;
  std::vector<int64_t> h_rowptr, lvl_f, lvl_b;
  std::vector<int32_t> h_col, h_diag, h_perm;
  std::vector<TriStep> steps_f, steps_b;
  // block-local sweeps (ordering 2, sweep_block.cu): the rows are cut into spatially compact blocks, one CTA each; block b
  // owns the permuted rows [blk_off[b], blk_off[b+1]), sorted by their dependency level inside the block (h_level)
  int nblk = 0;
  std::vector<int64_t> blk_off;
  std::vector<int32_t> h_level;
  DevBuf<double> bl_val;                // value sections of all passes (ELL slices + reciprocal diagonals)
  DevBuf<uint16_t> bl_idx;              // index sections (block-local columns + row slots)
  DevBuf<int64_t> bl_map;               // per entry of bl_val: (index into val << 2) | kind
  DevBuf<unsigned char> bl_pass;        // pass headers (sweep_block.cu PassHdr)
  DevBuf<BlkDesc> bl_blk;               // block descriptors
  int64_t bl_nval = 0;
  int bl_max_rows = 0, bl_max_pass = 0, bl_ring = 0;   // sizes of the shared-memory areas (rows, passes, ring bytes)
  int bl_sgs = -1;                      // what the stream currently holds: 1 SGS values, 0 ILU factors, -1 nothing
  bool node = false;                    // rows are velocity nodes: every row stands for the (u_x, u_y) pair of a node
  DevBuf<int32_t> px_ref, py_ref, px_node, py_node;   // vector entries of the pair of permuted row r: reference layout / node layout
};

struct AmgHierarchy;  // amg.cu

// Ghost import plan of one block on this rank (the role of Epetra_Import in the reference's ghosted vectors)
struct HaloPlan {
  std::vector<int> nbr;                 // neighbour ranks
  std::vector<int64_t> send_ptr, recv_ptr;
  DevBuf<int32_t> send_idx;             // owned local ids to pack, grouped by neighbour
  DevBuf<int32_t> send_idx_node;        // the same entries at their positions in the node layout of a vector (decouple.cu)
  DevBuf<double> send_buf;
  int64_t nsend = 0;
};

struct Ctx {
  int rank = 0, nranks = 1, device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int verbose = 0;
  int ordering = 2;      // ILU / SGS elimination order: 0 natural (Ifpack), 1 multicolour over the owned range, 2 (default) multicolour inside CTA-local blocks
  bool l2_hints = true;   // NSX_OPT_L2_HINTS: TMA matrix streams carry an evict-first L2 policy
  // NSX_OPT_PRECOND_LAG: numeric preconditioner data (ILU factors, sweep values, AMG hierarchy, Schur complement) built in solve s
  // serve the solves s+1 .. s+lag as well (0, default: rebuilt in every solve, as the reference's initialize() calls do)
  int precond_lag = 0;
  int sweep_q = 4;        // NSX_OPT_SWEEP_Q: entries per lane of the block-local sweeps' passes (4, 8 or 16)
  long long solve_seq = 0, schur_built_at = -1, amg_built_at = -1, stat_precond_builds = 0;
  bool ordering_auto = true;   // NSX_OPT_ORDERING never set: preconditioners that are a single ILU(0) application per iteration use ordering 3
  int block_rows = 0;    // ordering 2: target rows per block (0: n / #SMs clamped to [512, 4096])
  bool host_inner = false;  // inner FGMRES recurrences on the host (round-1 behaviour) instead of the device
  // device-driven inner FGMRES (krylov.cu): recurrence state in device memory, its verdicts mirrored in a mapped host record
  DevBuf<unsigned char> fg_dev;
  void *fg_rec = nullptr;        // pinned, mapped
  long long fg_seq = 0;
  int ortho = 2;        // 0 modified Gram-Schmidt chain (as deal.II), 1 batched classical Gram-Schmidt twice, 2 (default) as 1 for the outer solver and a conditional second pass for the inner FGMRES solves
  int stream_spmv = 3;  // 3: TMA-fed persistent SpMV, rows reduced straight from the stage, paired columns (default); 2: same ring, products staged in shared memory; 1: streaming with plain loads; 0: sub-warp per row
  DevBuf<RowBlockDesc> desc_u, desc_p;
  int ndesc_u = 0, ndesc_p = 0;
  DevBuf<int32_t> rb_u, rb_p;  // row blocks of the Jacobian block SpMV: velocity rows (F + Bt), pressure rows (B)
  int nrb_u = 0, nrb_p = 0;
  int coop_sweep = 1;   // multicolour sweeps as one cooperative launch with grid barriers between colours (2: the older level-phased kernel)
  int num_sms = 148;

  // discretisation
  FETables fe;
  bool have_disc = false, finalized = false;
  // n_u / n_p / n count the OWNED dofs of this rank (everything, on one GPU); n_ug / n_pg its ghosts.  Every
  // vector is allocated with nvec = n + n_ug + n_pg entries, laid out [u owned | p owned | u ghosts | p ghosts]:
  // BLAS-1 kernels run over the first n entries, SpMV column indices are baked to that layout at set-up.
  int64_t ncells = 0, n_u = 0, n_p = 0, n = 0, n_ug = 0, n_pg = 0, nvec = 0;
  HaloPlan halo_u, halo_p;
  void *comm = nullptr;           // ncclComm_t (comm.cu); null on one GPU
  int64_t stat_halo = 0, stat_allreduce = 0;
  DevBuf<int32_t> vmap;           // local block dof id (cell table) -> index in a vector
  std::vector<uint8_t> h_cell_owned;
  std::vector<double> h_cell_vertices;
  std::vector<uint32_t> h_cell_dofs;
  DevBuf<double> cell_vertices;
  DevBuf<uint32_t> cell_dofs;
  DevCSR F, Bt, B, Mp, S;
  // Component-decoupled view of F (decouple.cu): in the Stokes-type branches the couplings between the two velocity
  // components are exact zeros (the vector Laplacian is two scalar ones).  When the current values say so (checked on the
  // device after every assembly), the SGS sweeps and the F products of the inner solves run on the same-component entries
  // only -- half the bytes, bit-identical sums -- through Fd and the plan variant 1 of block F.
  DevCSR Fd;
  DevBuf<int64_t> Fd_src;          // Fd entry -> index into F.val
  DevBuf<uint8_t> F_cross;         // per F entry: 1 = couples different components
  DevBuf<unsigned long long> cross_count;
  std::vector<uint8_t> h_comp_u;   // component of every local velocity dof (owned + ghost)
  bool decouple = true;            // NSX_OPT_DECOUPLE (0 off, 1 same-component view only, 2 = default: node view where it holds)
  bool decouple_nodes = true;
  long long matrix_epoch = 1, dec_epoch = 0;   // F.val changed / decision taken at
  bool dec_ok = false;
  // Node view: when moreover F(dx a, dx b) == F(dy a, dy b) bit for bit (F = K (x) I_2, the vector Laplacian of the Stokes-type
  // branches) one scalar matrix K over the velocity NODES serves both components: Kn holds K with columns in the node layout of a
  // vector (entry 2b / 2b + 1 = x / y component of node b), plan variant 2 of block F sweeps node pairs, the inner FGMRES on F runs
  // in the node layout (its SpMV multiplies (x, y) pairs with one matrix value) and permutes only its right-hand side and result.
  DevCSR Kn;
  DevBuf<int64_t> Kn_src_x, Kn_src_y;   // Kn entry -> index of its (x,x) / (y,y) twin in F.val
  std::vector<int64_t> h_Kn_src;
  std::vector<int32_t> h_node_dx, h_node_dy;   // the two dofs of every velocity node (nodes numbered by ascending x dof)
  DevBuf<int32_t> node_dx, node_dy;
  DevBuf<int2> node_gpair;                     // ghost nodes: positions of their (x, y) entries in a vector's ghost tail
  std::vector<int64_t> owned_nodes;            // owned_u in nodes
  DevBuf<double> node_b, node_x;               // right-hand side / iterate of an inner solve in the node layout
  int node_struct = 0;             // 0 not examined, 1 the numbering pairs up, -1 it does not
  bool node_ok = false;
  std::vector<int64_t> owned_u{0, 0}, owned_p{0, 0};
  // Dirichlet
  DevBuf<uint32_t> bc_dof;
  DevBuf<double> bc_val;
  int64_t nbc = 0;
  // faces
  std::vector<int32_t> h_outlet_cell, h_outlet_face, h_cyl_cell, h_cyl_face;
  DevBuf<double> outlet_unit;     // -sum_faces sum_q (n . phi_i) w_f per outlet velocity dof (times p_out at use)
  DevBuf<uint32_t> outlet_dof;
  int64_t n_outlet = 0;
  DevBuf<unsigned long long> bc_first;  // per owned range: first row with a non-zero diagonal
  DevBuf<int32_t> cyl_cell, cyl_face;
  // assembly: cells grouped by colour (no two cells of one colour share a dof => the scatter
  // needs no atomics and the summation order is fixed), and per cell the id of its table of
  // row-relative CSR offsets (identical tables are shared between cells)
  int ncolors = 0;
  std::vector<int64_t> color_ptr;
  DevBuf<int32_t> color_cells;
  DevBuf<int32_t> cell_pat;
  DevBuf<uint16_t> pat_off;       // npat * ndofs * ndofs
  int64_t npat = 0;
  DevBuf<FETables> d_fe;
  DevBuf<double> face_force;      // per cylinder face: (drag, lift)
  DevBuf<int64_t> d_owned_u;
  // vectors (block layout [u | p])
  DevBuf<double> vec[7];
  // reductions
  DevBuf<double> red_partial;     // per-block partial sums
  DevBuf<double> red_result;      // device scalars
  DevBuf<unsigned int> red_counter;
  double *h_scalars = nullptr;    // pinned
  // preconditioner plans, keyed by block
  std::map<int, std::unique_ptr<TriPlan>> tri;
  std::shared_ptr<AmgHierarchy> amg;  // shared_ptr: deleter bound where the type is complete (amg.cu)
  // aSIMPLE
  DevBuf<double> Dvec, Dinv, delta_p, tmp_u, tmp_p, tmp_s;
  DevBuf<int32_t> ghost_bc;       // ghost velocity dofs that are Dirichlet dofs on their owner (their Bt rows are cleared too)
  int64_t n_ghost_bc = 0;
  bool S_symbolic = false;
  // Krylov workspaces (outer block vectors, inner velocity / pressure vectors)
  std::vector<DevBuf<double>> work_outer, work_inner_u, work_inner_p;
  // statistics
  int64_t stat_inner_F = 0, stat_inner_S = 0, stat_applies = 0, stat_launches = 0, stat_spmv = 0;
  int last_step = 0;
  double last_residual = 0;
  // parameters of the assembly launch timed by nsx_time_kernel
  int time_mode = NSX_MODE_NEWTON;
  double time_nu = 1.0 / 90.0, time_dt = 0.01;
  // L2 flush buffer for timing
  DevBuf<char> flush;
};

// ---- kernels_vec.cu -------------------------------------------------------------------------
constexpr int RED_SLOTS = 192;  // device scalar slots: 80 per solver nesting depth + scratch
struct VecList { const double *v[32]; };
void vec_copy(Ctx &c, double *y, const double *x, int64_t n);
void vec_set(Ctx &c, double *y, double a, int64_t n);
void vec_scale(Ctx &c, double *y, double a, int64_t n);
void vec_axpy(Ctx &c, double *y, double a, const double *x, int64_t n);                 // y += a x
void vec_sadd(Ctx &c, double *y, double s, double a, const double *x, int64_t n);      // y = s y + a x
void vec_equ(Ctx &c, double *y, double a, const double *x, int64_t n);                 // y = a x
void vec_mul(Ctx &c, double *y, const double *d, int64_t n);                           // y *= d (element-wise)
// y += (sign * *dev_coef) x, coefficient read on the device (no host round trip)
void vec_axpy_dev(Ctx &c, double *y, double sign, const double *dev_coef, const double *x, int64_t n);
// slot <- a . b (deterministic two-stage reduction; result stays on the device)
void vec_dot_dev(Ctx &c, int slot, const double *a, const double *b, int64_t n);
// w += (sign * *dev_coef) x ; slot <- w . v      (deal.II add_and_dot; v may alias w)
void vec_add_and_dot_dev(Ctx &c, int slot, double *w, double sign, const double *dev_coef, const double *x, const double *v, int64_t n);
double *slot_ptr(Ctx &c, int slot);
// batched classical Gram-Schmidt pass: slot0+m <- v_m . w (m < k) ; then w -= sum_m slot[m] v_m, slot_norm <- w . w
void vec_multi_dot_dev(Ctx &c, int slot0, const VecList &V, int k, const double *w, int64_t n);
void vec_multi_axpy_norm_dev(Ctx &c, int slot_norm, const VecList &V, int k, int slot_coef, double *w, int64_t n, bool reduce = true);   // reduce: sum the norm over ranks
// ---- device-driven inner FGMRES (SolverFGMRES on one block; deal.II's recurrences restated on the device) ----
// Recurrence state of one solve in device memory.  Instead of deal.II's Householder least-squares solve of the whole
// Hessenberg matrix in every iteration, the columns are reduced by Givens rotations as they arrive: the residual of the
// (j+1) x j problem deal.II checks at step j is |g[j]|, and the same y comes out of the back substitution.
struct FgDev {
  double R[30][32];     // rotated Hessenberg columns (upper triangular)
  double g[32], cs[32], sn[32], y[32], hsave[32];
  double a, tol, res;   // a: norm of the vector the next basis vector is scaled from
  int it, max_it, ny, gate;   // gate: 0 go on, 1 second Gram-Schmidt pass wanted, 2 converged, 3 failed
};
struct FgRec { long long seq; int gate, it, ny, pad; double res; };   // mirror of the verdicts in mapped host memory
FgDev *fg_state(Ctx &c);   // allocates the state + record on first use
// res = sqrt(*beta2); verdict of SolverControl::check(it0, res); starts a restart cycle
void fg_begin(Ctx &c, const double *beta2, double tol, int max_it, int it0);
// column j of the cycle: mode 0 = coefficients slots[0..j] + square norm slots[j+1] (modified Gram-Schmidt chain);
// 1 = first classical pass (slots[0..j], slots[64]), asks for a second pass after heavy cancellation; 2 = that second
// pass (adds slots[32..], norm slots[65]); 3 = two passes done up front; 4 = as 1 with the square norm from Pythagoras:
// slots[j+1] = |w|^2 before the update came with the coefficients in one reduction (partitioned runs)
void fg_step(Ctx &c, const double *slots, int j, int mode);
bool vec_multi_axpy_norm_fg(Ctx &c, int slot_norm, const VecList &V, int k, int slot_coef, double *w, int64_t n, const double *slots, int j, int mode);
FgRec fg_wait(Ctx &c);     // spins on the mapped record until the last fg_begin / fg_step has landed
// v = x / *a (zero if *a is not finite), skipped when *gate != 0
void vec_scale_to_dev(Ctx &c, double *v, const double *x, const double *a, const int *gate, int64_t n);
// x += sum_{m < *count} coef[m] V_m  (count and coefficients read on the device)
void vec_multi_add_dev(Ctx &c, double *x, const VecList &V, const double *coef, const int *count, int64_t n);
// blocking reads of device scalars
double read_slot(Ctx &c, int slot);
void read_slots(Ctx &c, int first, int count, double *out);
double vec_dot(Ctx &c, const double *a, const double *b, int64_t n);
void flush_l2_cache(Ctx &c, int round);
constexpr int FP64_PEAK_ITERS = 8192;
double fp64_peak_launch(Ctx &c, double *sink);   // launches the FP64 FMA peak kernel; returns the flops of the launch
double vec_norm(Ctx &c, const double *a, int64_t n);
double vec_add_and_dot(Ctx &c, double *w, double a, const double *x, const double *v, int64_t n);

// ---- kernels_spmv.cu ------------------------------------------------------------------------
void spmv(Ctx &c, const DevCSR &A, const double *x, double *y, bool add = false);
void spmv_local(Ctx &c, const DevCSR &A, const double *x, double *y, bool add = false);  // no ghost import (rank-local operators)
void block_spmv(Ctx &c, const double *x, double *y);  // y_u = F x_u + Bt x_p ; y_p = B x_u
void extract_diag(Ctx &c, const DevCSR &A, double *d, double *dinv);
void spmv_probe(Ctx &c, const DevCSR &A, int what, const double *x, double *sink);  // measurement only

// ---- assemble.cu ----------------------------------------------------------------------------
void build_assembly_maps(Ctx &c);
void assemble_cells(Ctx &c, int mode, double nu, double dt, double p_out, bool res_only = false);
void assemble_residual(Ctx &c, int mode, double nu, double dt, double p_out);
void apply_boundary_values(Ctx &c, bool apply_inlet);
void assemble(Ctx &c, int mode, bool apply_inlet, double nu, double dt, double p_out);
void lift_drag(Ctx &c, double nu, double *drag, double *lift);

// ---- trisolve.cu ----------------------------------------------------------------------------
TriPlan &tri_plan(Ctx &c, int block, int variant = 0, int ordering = -1);   // ordering < 0: the context's (NSX_OPT_ORDERING)
void tri_erase(Ctx &c, int block);   // drops every cached plan of a block   // block F only: variant 1 same-component couplings, variant 2 velocity nodes (Ctx::Kn)
void gather_values(Ctx &c, int64_t nnz, const int64_t *src, const double *a, double *v);   // v[k] = a[src[k]]
void tri_refresh_values(Ctx &c, TriPlan &P, const DevCSR &A);   // permuted copy of the values
void ilu0_factor(Ctx &c, TriPlan &P, const DevCSR &A);
void ilu0_apply(Ctx &c, TriPlan &P, double *y, const double *x);
void sgs_apply(Ctx &c, TriPlan &P, double *y, const double *x);
DevCSR &block_ref(Ctx &c, int block);

// ---- sweep_block.cu -------------------------------------------------------------------------
// cuts the rows [lo, hi) of a square block into spatially compact groups (weighted recursive coordinate bisection of the
// dof positions implied by the cell table, weights = row lengths of `rowptr`); grp[i] = first_group + k
int geometric_blocks(Ctx &c, int block, const std::vector<int64_t> &rowptr, int64_t lo, int64_t hi, int first_group, std::vector<int32_t> &grp,
                     const std::vector<int32_t> *row_dof = nullptr);   // row_dof: rows are velocity nodes, row i sits where dof (*row_dof)[i] does
void bl_build(Ctx &c, TriPlan &P);                       // streams + descriptors of the block-local sweep
void bl_refresh(Ctx &c, TriPlan &P, bool sgs);           // stream values from P.val (after tri_refresh_values / ilu0_factor)
// y = M^-1 x.  Fused variants for the inner FGMRES: x is scaled by 1 / *scale on the way in and the scaled vector is
// also stored to v_out; the launch does nothing when *gate != 0 (speculative launch behind a device-side decision).
// A node plan reads / writes vectors in the reference layout (entries dx(a), dy(a)) or, node_layout = true, in the node layout.
void bl_sweep(Ctx &c, TriPlan &P, bool sgs, double *y, const double *x, const double *scale = nullptr, double *v_out = nullptr, const int *gate = nullptr,
              bool node_layout = false);

// ---- decouple.cu ----------------------------------------------------------------------------
const std::vector<uint8_t> &velocity_components(Ctx &c);
// The cheapest exact view of F on its current values (decision cached per assembly; refreshes the view's values):
// 0 the full matrix, 1 same-component entries only (cross-component entries are all exact zeros), 2 one scalar matrix over the
// velocity nodes (moreover the two diagonal component blocks are bit-identical)
int stokes_view(Ctx &c);
int effective_view(Ctx &c);
void vec_to_node_layout(Ctx &c, double *node, const double *ref);     // node[2a], node[2a+1] = ref[dx(a)], ref[dy(a)]
void vec_from_node_layout(Ctx &c, double *ref, const double *node);   // stokes_view, limited to what the selected kernels support (node view: orderings 2 / 3 and the direct SpMV)

// ---- spgemm.cu ------------------------------------------------------------------------------
void schur_symbolic(Ctx &c);
void schur_complement(Ctx &c);  // S = B diag(F)^-1 Bt, fills Dvec / Dinv

// ---- amg.cu ---------------------------------------------------------------------------------
void amg_setup(Ctx &c, const DevCSR &A);
void amg_apply(Ctx &c, double *y, const double *x);

// ---- comm.cu --------------------------------------------------------------------------------
// blk 0: velocity ghosts of the vector starting at `base` (block or velocity-only vector);
// blk 1: pressure ghosts of the pressure part starting at `base`.  No-ops on one GPU.
void halo_exchange(Ctx &c, int blk, const double *base, bool node_layout = false);   // node_layout: the owned velocity entries sit in the node layout
void allreduce_slots(Ctx &c, int slot, int count);   // in-place sum over ranks of device scalar slots
void comm_destroy(Ctx &c);

// ---- krylov.cu ------------------------------------------------------------------------------
int solve_system(Ctx &c, int flavour, int solver, int prec, double tol, int max_it, double alpha, double *final_res);
void precond_apply_once(Ctx &c, int flavour, int prec, double alpha, const double *src, double *dst);

inline int grid_for(int64_t n, int block, int max_blocks) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace nsx

struct nsx_ctx : nsx::Ctx {};
