// capi.cu -- the C ABI of include/nsx.h over the device library.
#include <algorithm>
#include <cstring>

#include "device.cuh"

using namespace nsx;

namespace {

template <class Fn>
int guarded(nsx_ctx *ctx, Fn &&fn) {
  if (!ctx) return NSX_E_BADARG;
  try {
    NSX_CUDA(cudaSetDevice(ctx->device));
    fn();
    return NSX_OK;
  } catch (const NoConvergence &e) {
    ctx->err = e.what();
    ctx->last_step = e.last_step; ctx->last_residual = e.last_residual;
    return NSX_E_NOCONV;
  } catch (const std::invalid_argument &e) { ctx->err = e.what(); return NSX_E_BADARG; }
  catch (const CudaError &e) { ctx->err = e.what(); return NSX_E_CUDA; }
  catch (const std::logic_error &e) { ctx->err = e.what(); return NSX_E_STATE; }
  catch (const std::exception &e) { ctx->err = e.what(); return NSX_E_CUDA; }
}

// `cols_are_p`: the block's columns index pressure dofs.  Host copies keep the local column ids; the device
// copy is baked to the vector layout [u owned | p owned | u ghosts | p ghosts] (relative to the start of the
// velocity or of the pressure part), so that SpMV kernels index x directly.
void set_block(Ctx &c, DevCSR &A, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int32_t *col, bool cols_are_p) {
  A.nrows = nrows; A.ncols = ncols;
  A.h_rowptr.assign(rowptr, rowptr + nrows + 1);
  A.nnz = rowptr[nrows];
  A.h_col.assign(col, col + A.nnz);
  for (int64_t k = 0; k < A.nnz; ++k)
    if (col[k] < 0 || col[k] >= ncols) throw std::invalid_argument("pattern column out of range");
  A.rowptr.alloc_padded(A.h_rowptr.size(), 4, c.stream);
  NSX_CUDA(cudaMemcpyAsync(A.rowptr.p, A.h_rowptr.data(), A.h_rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  A.col.alloc_padded(A.nnz, 16, c.stream);
  std::vector<int32_t> baked;
  const int32_t *dev_col = A.h_col.data();
  const int64_t own = cols_are_p ? c.n_p : c.n_u, shift = cols_are_p ? c.n_ug : c.n_p;
  if (c.n_ug + c.n_pg > 0) {
    baked.resize(A.nnz);
    for (int64_t k = 0; k < A.nnz; ++k) baked[k] = (int32_t)(col[k] < own ? col[k] : col[k] + shift);
    dev_col = baked.data();
  }
  NSX_CUDA(cudaMemcpyAsync(A.col.p, dev_col, A.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  A.val.alloc_padded(A.nnz, 16, c.stream);  // val[nnz + 12] is the sink of the ghost rows' contributions (assemble.cu)
  A.nrb = A.ndesc = 0;
  A.nrows_ext = A.nnz_ext = 0;
  A.pair_state = 0; A.pcol.release();
  c.nrb_u = c.nrb_p = c.ndesc_u = c.ndesc_p = 0;
  A.max_row = 0;
  for (int64_t i = 0; i < nrows; ++i) {
    A.max_row = std::max<int>(A.max_row, (int)(rowptr[i + 1] - rowptr[i]));
    if (!std::is_sorted(col + rowptr[i], col + rowptr[i + 1])) throw std::invalid_argument("pattern rows must have ascending columns");
  }
  if (&A == &c.F || &A == &c.Mp) {  // square blocks: owned rows x (owned + ghost) columns, diagonal among the owned
    std::vector<int32_t> diag(nrows, -1);
    for (int64_t i = 0; i < nrows; ++i) {
      const int32_t *b = col + rowptr[i], *e = col + rowptr[i + 1];
      const int32_t *it = std::lower_bound(b, e, (int32_t)i);
      if (it != e && *it == i) diag[i] = (int32_t)(it - b);
    }
    A.diag.upload(diag, c.stream);
  }
  NSX_CUDA(cudaStreamSynchronize(c.stream));
}

double *vec_of(Ctx &c, int which) {
  if (which < 0 || which > NSX_VEC_TMP1) throw std::invalid_argument("unknown vector id");
  if (!c.vec[which].p) throw std::logic_error("vectors are allocated by nsx_set_discretisation");
  return c.vec[which].p;
}

// the plan behind the query id NSX_BLOCK_F_DECOUPLED: the view of F the sweeps use on the current values
TriPlan &stokes_plan(Ctx &c) {
  const int view = effective_view(c);
  if (view == 0) throw std::logic_error("F couples the velocity components on its current values: there is no decoupled view");
  return tri_plan(c, NSX_BLOCK_F, view);
}

void need_final(Ctx &c) {
  if (!c.finalized) throw std::logic_error("nsx_finalize_setup has not been called");
}

}  // namespace

extern "C" {

int nsx_create(int rank, int nranks, int device_id, void *stream, nsx_ctx **out) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks) return NSX_E_BADARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device_id < 0 || device_id >= ndev) return NSX_E_CUDA;  // no CPU fallback
  nsx_ctx *c = new nsx_ctx;
  c->rank = rank; c->nranks = nranks; c->device = device_id;
  try {
    NSX_CUDA(cudaSetDevice(device_id));
    if (stream) c->stream = (cudaStream_t)stream;
    else { NSX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    NSX_CUDA(cudaGetDeviceProperties(&prop, device_id));
    c->num_sms = prop.multiProcessorCount;
  } catch (const std::exception &) { delete c; return NSX_E_CUDA; }
  *out = c;
  return NSX_OK;
}

int nsx_destroy(nsx_ctx *ctx) {
  if (!ctx) return NSX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(*ctx);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->fg_rec) cudaFreeHost(ctx->fg_rec);
  cudaStream_t s = ctx->own_stream ? ctx->stream : nullptr;
  delete ctx;
  if (s) cudaStreamDestroy(s);
  return NSX_OK;
}

const char *nsx_last_error(const nsx_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nsx_set_option(nsx_ctx *ctx, int option, int64_t value) {
  return guarded(ctx, [&] {
    switch (option) {
      case NSX_OPT_ORDERING:
        if (value < 0 || value > 3) throw std::invalid_argument("ordering must be 0 (natural), 1 (multicolour), 2 (multicolour inside CTA-local blocks) or 3 (natural inside CTA-local blocks)");
        ctx->ordering = (int)value; ctx->ordering_auto = false; break;
      case NSX_OPT_BLOCK_ROWS:
        if (value < 0 || value > 4096) throw std::invalid_argument("block rows must lie in [0, 4096] (0: automatic)");
        ctx->block_rows = (int)value; ctx->tri.clear(); ctx->schur_built_at = ctx->amg_built_at = -1; break;
      case NSX_OPT_VERBOSE: ctx->verbose = (int)value; break;
      case NSX_OPT_ORTHO:
        if (value < 0 || value > 2) throw std::invalid_argument("orthogonalisation must be 0, 1 or 2");
        ctx->ortho = (int)value; break;
      case NSX_OPT_COOP_SWEEP:
        if (value < 0 || value > 2) throw std::invalid_argument("sweep kernel must be 0 (a launch per level), 1 (colour-phased persistent) or 2 (level-phased cooperative)");
        ctx->coop_sweep = (int)value; break;
      case NSX_OPT_STREAM_SPMV:
        if (value < 0 || value > 3) throw std::invalid_argument("SpMV kernel must be 0, 1, 2 or 3");
        ctx->stream_spmv = (int)value; break;
      case NSX_OPT_HOST_INNER: ctx->host_inner = value != 0; break;
      case NSX_OPT_L2_HINTS: ctx->l2_hints = value != 0; break;
      case NSX_OPT_SWEEP_Q:
        if (value != 4 && value != 8 && value != 16) throw std::invalid_argument("entries per lane must be 4, 8 or 16");
        ctx->sweep_q = (int)value; ctx->tri.clear(); ctx->schur_built_at = ctx->amg_built_at = -1; break;
      case NSX_OPT_PRECOND_LAG:
        if (value < 0 || value > 1000000) throw std::invalid_argument("preconditioner lag must be a solve count >= 0");
        ctx->precond_lag = (int)value; break;
      case NSX_OPT_DECOUPLE:
        if (value < 0 || value > 2) throw std::invalid_argument("decouple must be 0 (off), 1 (same-component view) or 2 (node view where it holds)");
        ctx->decouple = value != 0; ctx->decouple_nodes = value == 2; ctx->dec_epoch = 0; break;
      default: throw std::invalid_argument("unknown option");
    }
  });
}

int64_t nsx_get_stat(const nsx_ctx *ctx, int stat) {
  if (!ctx) return -1;
  auto levels = [&](int block) -> int64_t {   // of the block's most recently built plan
    int64_t l = 0;
    for (const auto &kv : ctx->tri) if ((kv.first & 15) == block) l = (int64_t)kv.second->lvl_f.size() - 1;
    return l;
  };
  switch (stat) {
    case NSX_STAT_INNER_F: return ctx->stat_inner_F;
    case NSX_STAT_INNER_S: return ctx->stat_inner_S;
    case NSX_STAT_PRECOND_APPLIES: return ctx->stat_applies;
    case NSX_STAT_KERNEL_LAUNCHES: return ctx->stat_launches;
    case NSX_STAT_LEVELS_F: return levels(NSX_BLOCK_F);
    case NSX_STAT_LEVELS_MP: return levels(NSX_BLOCK_MP);
    case NSX_STAT_LEVELS_S: return levels(NSX_BLOCK_S);
    case NSX_STAT_SPMV_CALLS: return ctx->stat_spmv;
    case NSX_STAT_ASSEMBLY_COLOURS: return ctx->ncolors;
    case NSX_STAT_ASSEMBLY_TABLES: return ctx->npat;
    case NSX_STAT_LAST_STEP: return ctx->last_step;
    case NSX_STAT_HALO_EXCHANGES: return ctx->stat_halo;
    case NSX_STAT_ALLREDUCES: return ctx->stat_allreduce;
    case NSX_STAT_PRECOND_BUILDS: return ctx->stat_precond_builds;
    case NSX_STAT_SWEEP_BYTES_F: {   // stored bytes one application of the F sweeps streams: the plan of the view found by the last check
      int view = (ctx->dec_epoch == ctx->matrix_epoch && ctx->dec_ok) ? (ctx->node_ok ? 2 : 1) : 0;
      if (view == 2 && (ctx->ordering < 2 || ctx->stream_spmv != 3)) view = 1;
      auto it = ctx->tri.find(NSX_BLOCK_F + 16 * view + 256 * ctx->ordering);
      if (it == ctx->tri.end()) return 0;
      const TriPlan &P = *it->second;
      return P.nblk ? (int64_t)(P.bl_nval * 8 + (int64_t)P.bl_idx.n * 2) : (int64_t)P.nnz * 12;
    }
    case NSX_STAT_SPMV_BYTES_F: {    // stored matrix bytes one F product of the inner solves reads (values + columns)
      int view = (ctx->dec_epoch == ctx->matrix_epoch && ctx->dec_ok) ? (ctx->node_ok ? 2 : 1) : 0;
      if (view == 2 && (ctx->ordering < 2 || ctx->stream_spmv != 3)) view = 1;
      if (view == 2) return ctx->Kn.nnz * 12;
      if (view == 1) return ctx->Fd.nnz * 12;
      return ctx->F.nnz * (ctx->F.pair_state == 1 ? 10 : 12);
    }
    case NSX_STAT_F_DECOUPLED: return (ctx->dec_epoch == ctx->matrix_epoch && ctx->dec_ok) ? (ctx->node_ok ? 2 : 1) : 0;
  }
  return -1;
}

int nsx_set_discretisation(nsx_ctx *ctx, int elem, int64_t n_cells, const double *cell_vertices, const uint32_t *cell_dofs, int64_t n_u,
                           int64_t n_p) {
  return guarded(ctx, [&] {
    if ((elem != 0 && elem != 1) || n_cells <= 0 || !cell_vertices || !cell_dofs || n_u <= 0 || n_p <= 0)
      throw std::invalid_argument("bad discretisation");
    Ctx &c = *ctx;
    build_fe_tables(elem, c.fe);
    c.ncells = n_cells; c.n_u = n_u; c.n_p = n_p; c.n = n_u + n_p;
    c.n_ug = c.n_pg = 0; c.nvec = c.n;
    c.halo_u = HaloPlan(); c.halo_p = HaloPlan();
    c.h_cell_owned.clear();
    c.h_cell_vertices.assign(cell_vertices, cell_vertices + (size_t)n_cells * c.fe.nvpc * 2);
    c.h_cell_dofs.assign(cell_dofs, cell_dofs + (size_t)n_cells * c.fe.ndofs);
    for (uint32_t d : c.h_cell_dofs)
      if ((int64_t)d >= n_u + n_p) throw std::invalid_argument("cell dof index out of range");
    c.cell_vertices.upload(c.h_cell_vertices, c.stream);
    c.cell_dofs.upload(c.h_cell_dofs, c.stream);
    for (auto &v : c.vec) { v.alloc(c.nvec); v.zero(c.stream); }
    c.owned_u = {0, n_u}; c.owned_p = {0, n_p};
    c.have_disc = true; c.finalized = false; c.S_symbolic = false;
    c.tri.clear(); c.schur_built_at = c.amg_built_at = -1;
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_set_pattern(nsx_ctx *ctx, int block, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int32_t *col) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    if (!rowptr || !col || block < NSX_BLOCK_F || block > NSX_BLOCK_MP) throw std::invalid_argument("bad pattern block");
    const bool cols_p = !(block == NSX_BLOCK_F || block == NSX_BLOCK_B);
    const int64_t er = (block == NSX_BLOCK_F || block == NSX_BLOCK_BT) ? c.n_u : c.n_p;
    const int64_t ec = cols_p ? c.n_p + c.n_pg : c.n_u + c.n_ug;
    if (nrows != er || ncols != ec) throw std::invalid_argument("pattern shape does not match the block (owned rows x owned + ghost columns)");
    set_block(c, block_ref(c, block), nrows, ncols, rowptr, col, cols_p);
    c.finalized = false;
    tri_erase(c, block);
    if (block == NSX_BLOCK_F) {
      c.F_cross.release(); c.Fd = DevCSR(); c.Kn = DevCSR(); c.node_struct = 0; c.h_comp_u.clear(); c.dec_epoch = 0;
    }
    c.matrix_epoch++;
    if (block == NSX_BLOCK_B || block == NSX_BLOCK_BT) { c.S_symbolic = false; c.schur_built_at = -1; tri_erase(c, NSX_BLOCK_S); }
  });
}

int nsx_set_faces(nsx_ctx *ctx, int kind, int64_t n, const int32_t *cell, const int32_t *face) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    if (n < 0 || (n && (!cell || !face))) throw std::invalid_argument("bad face list");
    for (int64_t k = 0; k < n; ++k)
      if (cell[k] < 0 || cell[k] >= c.ncells || face[k] < 0 || face[k] >= c.fe.nfaces) throw std::invalid_argument("face out of range");
    if (kind == 8) { c.h_outlet_cell.assign(cell, cell + n); c.h_outlet_face.assign(face, face + n); }
    else if (kind == 10) { c.h_cyl_cell.assign(cell, cell + n); c.h_cyl_face.assign(face, face + n); }
    else throw std::invalid_argument("face kind must be 8 (outlet) or 10 (cylinder)");
    c.finalized = false;
  });
}

int nsx_set_dirichlet(nsx_ctx *ctx, int64_t n, const uint32_t *dof, const double *inlet_value) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    if (n < 0 || (n && (!dof || !inlet_value))) throw std::invalid_argument("bad Dirichlet list");
    for (int64_t k = 0; k < n; ++k)
      if ((int64_t)dof[k] >= c.n_u) throw std::invalid_argument("only velocity dofs can be constrained");
    c.nbc = n;
    c.bc_dof.upload(dof, n, c.stream);
    c.bc_val.upload(inlet_value, n, c.stream);
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_set_ranks(nsx_ctx *ctx, int nranks, const int64_t *owned_u, const int64_t *owned_p) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    if (nranks < 1 || !owned_u || !owned_p || owned_u[0] != 0 || owned_p[0] != 0 || owned_u[nranks] != c.n_u || owned_p[nranks] != c.n_p)
      throw std::invalid_argument("owned ranges must tile [0, n_u) and [0, n_p)");
    for (int r = 0; r < nranks; ++r)
      if (owned_u[r + 1] < owned_u[r] || owned_p[r + 1] < owned_p[r]) throw std::invalid_argument("owned ranges must be ascending");
    c.owned_u.assign(owned_u, owned_u + nranks + 1);
    c.owned_p.assign(owned_p, owned_p + nranks + 1);
    c.tri.clear(); c.schur_built_at = c.amg_built_at = -1;
    c.finalized = false;
  });
}

int nsx_set_partition(nsx_ctx *ctx, int64_t n_u_owned, int64_t n_p_owned) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    const int64_t lu = c.n_u + c.n_ug, lp = c.n_p + c.n_pg;  // local (owned + ghost) counts given to nsx_set_discretisation
    if (n_u_owned <= 0 || n_u_owned > lu || n_p_owned <= 0 || n_p_owned > lp) throw std::invalid_argument("owned counts must lie in (0, local count]");
    for (int b = NSX_BLOCK_F; b <= NSX_BLOCK_MP; ++b)
      if (!block_ref(c, b).h_rowptr.empty()) throw std::logic_error("nsx_set_partition must precede nsx_set_pattern");
    c.n_u = n_u_owned; c.n_p = n_p_owned; c.n = c.n_u + c.n_p;
    c.n_ug = lu - n_u_owned; c.n_pg = lp - n_p_owned;
    c.nvec = lu + lp;
    c.owned_u = {0, c.n_u}; c.owned_p = {0, c.n_p};
    c.tri.clear(); c.schur_built_at = c.amg_built_at = -1;
    c.finalized = false;
  });
}

int nsx_set_halo(nsx_ctx *ctx, int block, int n_neighbours, const int32_t *neighbour, const int64_t *send_ptr, const int32_t *send_idx,
                 const int64_t *recv_ptr) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    if ((block != 0 && block != 1) || n_neighbours < 0 || (n_neighbours && (!neighbour || !send_ptr || !recv_ptr))) throw std::invalid_argument("bad halo plan");
    HaloPlan &H = block ? c.halo_p : c.halo_u;
    H = HaloPlan();
    const int64_t own = block ? c.n_p : c.n_u, ghosts = block ? c.n_pg : c.n_ug;
    if (n_neighbours == 0) { if (ghosts) throw std::invalid_argument("ghost dofs without a neighbour to fill them"); return; }
    if (send_ptr[0] != 0 || recv_ptr[0] != 0 || recv_ptr[n_neighbours] != ghosts) throw std::invalid_argument("halo plan does not cover the ghost dofs");
    for (int i = 0; i < n_neighbours; ++i) {
      if (neighbour[i] < 0 || neighbour[i] >= c.nranks || neighbour[i] == c.rank) throw std::invalid_argument("bad neighbour rank");
      if (send_ptr[i + 1] < send_ptr[i] || recv_ptr[i + 1] < recv_ptr[i]) throw std::invalid_argument("halo ranges must be ascending");
    }
    H.nsend = send_ptr[n_neighbours];
    if (H.nsend && !send_idx) throw std::invalid_argument("bad halo plan");
    for (int64_t k = 0; k < H.nsend; ++k)
      if (send_idx[k] < 0 || send_idx[k] >= own) throw std::invalid_argument("only owned dofs can be sent");
    H.nbr.assign(neighbour, neighbour + n_neighbours);
    H.send_ptr.assign(send_ptr, send_ptr + n_neighbours + 1);
    H.recv_ptr.assign(recv_ptr, recv_ptr + n_neighbours + 1);
    H.send_idx.upload(send_idx, (size_t)H.nsend, c.stream);
    H.send_buf.alloc((size_t)std::max<int64_t>(1, H.nsend));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_halo_exchange(nsx_ctx *ctx, int which) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    double *v = vec_of(c, which);
    halo_exchange(c, 0, v);
    halo_exchange(c, 1, v + c.n_u);
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_finalize_setup(nsx_ctx *ctx) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (!c.have_disc) throw std::logic_error("nsx_set_discretisation must come first");
    for (int b = NSX_BLOCK_F; b <= NSX_BLOCK_MP; ++b)
      if (block_ref(c, b).h_rowptr.empty()) throw std::logic_error("all four block patterns must be set before nsx_finalize_setup");
    if ((c.n_ug && c.halo_u.nbr.empty()) || (c.n_pg && c.halo_p.nbr.empty())) throw std::logic_error("ghost dofs need nsx_set_halo");
    build_assembly_maps(c);
    c.finalized = true;
  });
}

int nsx_vec_upload(nsx_ctx *ctx, int which, const double *host) {
  return guarded(ctx, [&] {
    if (!host) throw std::invalid_argument("null host pointer");
    NSX_CUDA(cudaMemcpyAsync(vec_of(*ctx, which), host, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NSX_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int nsx_vec_download(nsx_ctx *ctx, int which, double *host) {
  return guarded(ctx, [&] {
    if (!host) throw std::invalid_argument("null host pointer");
    NSX_CUDA(cudaMemcpyAsync(host, vec_of(*ctx, which), ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NSX_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int nsx_vec_download_ghosts(nsx_ctx *ctx, int which, double *host_u_ghosts, double *host_p_ghosts) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    const double *v = vec_of(c, which);
    if (c.n_ug && host_u_ghosts) NSX_CUDA(cudaMemcpyAsync(host_u_ghosts, v + c.n, c.n_ug * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (c.n_pg && host_p_ghosts) NSX_CUDA(cudaMemcpyAsync(host_p_ghosts, v + c.n + c.n_ug, c.n_pg * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_vec_set(nsx_ctx *ctx, int which, double value) {
  return guarded(ctx, [&] { vec_set(*ctx, vec_of(*ctx, which), value, ctx->n); });
}

int nsx_vec_copy(nsx_ctx *ctx, int dst, int src) {
  return guarded(ctx, [&] { vec_copy(*ctx, vec_of(*ctx, dst), vec_of(*ctx, src), ctx->n); });
}

int nsx_assemble(nsx_ctx *ctx, int mode, int apply_inlet, double nu, double dt, double p_out, double *residual_l2) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    need_final(c);
    if (mode < NSX_MODE_STOKES || mode > NSX_MODE_UNSTEADY_NEWTON) throw std::invalid_argument("unknown assembly mode");
    if (!(nu > 0.0)) throw std::invalid_argument("nu must be positive");
    if (mode >= NSX_MODE_UNSTEADY_FIRST && !(dt > 0.0)) throw std::invalid_argument("dt must be positive");
    assemble(c, mode, apply_inlet != 0, nu, dt, p_out);
    c.matrix_epoch++;
    if (residual_l2) *residual_l2 = vec_norm(c, c.vec[NSX_VEC_RESIDUAL].p, c.n);
  });
}

int nsx_assemble_residual(nsx_ctx *ctx, int mode, double nu, double dt, double p_out, double *residual_l2) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    need_final(c);
    if (mode < NSX_MODE_STOKES || mode > NSX_MODE_UNSTEADY_NEWTON) throw std::invalid_argument("unknown assembly mode");
    if (!(nu > 0.0)) throw std::invalid_argument("nu must be positive");
    if (mode >= NSX_MODE_UNSTEADY_FIRST && !(dt > 0.0)) throw std::invalid_argument("dt must be positive");
    assemble_residual(c, mode, nu, dt, p_out);
    if (residual_l2) *residual_l2 = vec_norm(c, c.vec[NSX_VEC_RESIDUAL].p, c.n);
  });
}

int nsx_assemble_cells(nsx_ctx *ctx, int mode, double nu, double dt, double p_out) {
  return guarded(ctx, [&] {
    need_final(*ctx);
    assemble_cells(*ctx, mode, nu, dt, p_out);
    ctx->matrix_epoch++;
    NSX_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int nsx_solve(nsx_ctx *ctx, int flavour, int solver, int prec, double tol, int max_it, double alpha, int *iterations, double *final_residual) {
  int rc = guarded(ctx, [&] {
    Ctx &c = *ctx;
    need_final(c);
    if (flavour != NSX_STATIONARY && flavour != NSX_UNSTEADY) throw std::invalid_argument("unknown flavour");
    double fr = 0;
    const int it = solve_system(c, flavour, solver, prec, tol, max_it, alpha, &fr);
    c.last_step = it; c.last_residual = fr;
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
  if (ctx && (rc == NSX_OK || rc == NSX_E_NOCONV)) {
    if (iterations) *iterations = ctx->last_step;
    if (final_residual) *final_residual = ctx->last_residual;
  }
  return rc;
}

int nsx_save_eval_point(nsx_ctx *ctx) {
  return guarded(ctx, [&] { vec_copy(*ctx, vec_of(*ctx, NSX_VEC_EVAL), vec_of(*ctx, NSX_VEC_SOLUTION), ctx->n); });
}

int nsx_update(nsx_ctx *ctx, double alpha) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    // solution_owned = evaluation_point; solution_owned.add(alpha, delta_owned); solution = solution_owned
    vec_copy(c, vec_of(c, NSX_VEC_SOLUTION), vec_of(c, NSX_VEC_EVAL), c.n);
    vec_axpy(c, vec_of(c, NSX_VEC_SOLUTION), alpha, vec_of(c, NSX_VEC_DELTA), c.n);
  });
}

int nsx_copy_old(nsx_ctx *ctx) {
  return guarded(ctx, [&] { vec_copy(*ctx, vec_of(*ctx, NSX_VEC_SOLUTION_OLD), vec_of(*ctx, NSX_VEC_SOLUTION), ctx->n); });
}

int nsx_lift_drag(nsx_ctx *ctx, double nu, double *drag_force, double *lift_force) {
  return guarded(ctx, [&] {
    need_final(*ctx);
    if (!drag_force || !lift_force) throw std::invalid_argument("null output pointer");
    lift_drag(*ctx, nu, drag_force, lift_force);
  });
}

// ---- test / measurement hooks ----------------------------------------------------------------

int nsx_get_block_nnz(nsx_ctx *ctx, int block, int64_t *nnz) {
  return guarded(ctx, [&] { *nnz = block_ref(*ctx, block).nnz; });
}

int nsx_get_block_pattern(nsx_ctx *ctx, int block, int64_t *rowptr, int32_t *col) {
  return guarded(ctx, [&] {
    const DevCSR &A = block_ref(*ctx, block);
    std::copy(A.h_rowptr.begin(), A.h_rowptr.end(), rowptr);
    std::copy(A.h_col.begin(), A.h_col.end(), col);
  });
}

int nsx_get_block_values(nsx_ctx *ctx, int block, double *values) {
  return guarded(ctx, [&] {
    const DevCSR &A = block_ref(*ctx, block);
    NSX_CUDA(cudaMemcpyAsync(values, A.val.p, A.nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NSX_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int nsx_set_block_values(nsx_ctx *ctx, int block, const double *values) {
  return guarded(ctx, [&] {
    DevCSR &A = block_ref(*ctx, block);
    NSX_CUDA(cudaMemcpyAsync(A.val.p, values, A.nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NSX_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->matrix_epoch++;
  });
}

int nsx_spmv(nsx_ctx *ctx, int block, int vec_x, int vec_y) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (vec_x == vec_y) throw std::invalid_argument("x and y must differ");
    const double *x = vec_of(c, vec_x);
    double *y = vec_of(c, vec_y);
    if (block == NSX_BLOCK_J) block_spmv(c, x, y);
    else {
      const DevCSR &A = block_ref(c, block);
      if (!A.nrows) throw std::logic_error("block is not set");
      spmv(c, A, x, y);
    }
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_inner_apply(nsx_ctx *ctx, int block, int kind, int vec_x, int vec_y) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    if (vec_x == vec_y) throw std::invalid_argument("x and y must differ");
    const double *x = vec_of(c, vec_x);
    double *y = vec_of(c, vec_y);
    const DevCSR &A = block_ref(c, block);
    const int view = block == NSX_BLOCK_F ? effective_view(c) : 0;   // the view the solvers would use on the current values
    if (kind == 0) { TriPlan &P = tri_plan(c, block, view); tri_refresh_values(c, P, A); sgs_apply(c, P, y, x); }
    else if (kind == 1) { TriPlan &P = tri_plan(c, block, view == 2 ? 2 : 0); ilu0_factor(c, P, A); ilu0_apply(c, P, y, x); }
    else if (kind == 2) { amg_setup(c, A); amg_apply(c, y, x); }
    else throw std::invalid_argument("inner preconditioner kind must be 0 SGS, 1 ILU(0), 2 AMG");
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_ilu0_factor(nsx_ctx *ctx, int block, double *lu_values, int32_t *perm) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    const DevCSR &A = block_ref(c, block);
    TriPlan &P = tri_plan(c, block);
    ilu0_factor(c, P, A);
    // factors back in the layout of the block's value array (entries dropped by the rank filter read 0)
    std::vector<double> pv(P.nnz);
    std::vector<int64_t> src(P.nnz);
    NSX_CUDA(cudaMemcpyAsync(pv.data(), P.val.p, P.nnz * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaMemcpyAsync(src.data(), P.src.p, P.nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
    std::fill(lu_values, lu_values + A.nnz, 0.0);
    for (int64_t k = 0; k < P.nnz; ++k) lu_values[src[k]] = pv[k];
    if (perm) std::copy(P.h_perm.begin(), P.h_perm.end(), perm);
  });
}

int nsx_schur(nsx_ctx *ctx) {
  return guarded(ctx, [&] { schur_complement(*ctx); NSX_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

int nsx_precond_apply(nsx_ctx *ctx, int flavour, int prec, double alpha, int vec_src, int vec_dst) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    need_final(c);
    if (vec_src == vec_dst) throw std::invalid_argument("src and dst must differ");
    precond_apply_once(c, flavour, prec, alpha, vec_of(c, vec_src), vec_of(c, vec_dst));
    NSX_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int nsx_get_ordering(nsx_ctx *ctx, int block, int32_t *perm) {
  return guarded(ctx, [&] {
    TriPlan &P = block == NSX_BLOCK_F_DECOUPLED ? stokes_plan(*ctx) : tri_plan(*ctx, block);
    if (P.node) for (size_t r = 0; r < P.h_perm.size(); ++r) { perm[2 * r] = ctx->h_node_dx[P.h_perm[r]]; perm[2 * r + 1] = ctx->h_node_dy[P.h_perm[r]]; }   // a node = its two dofs, x first
    else std::copy(P.h_perm.begin(), P.h_perm.end(), perm);
  });
}

int nsx_get_sweep_blocks(nsx_ctx *ctx, int block, int32_t *n_blocks, int64_t *offsets) {
  return guarded(ctx, [&] {
    TriPlan &P = block == NSX_BLOCK_F_DECOUPLED ? stokes_plan(*ctx) : tri_plan(*ctx, block);
    if (!n_blocks) throw std::invalid_argument("null output pointer");
    *n_blocks = P.nblk;
    if (offsets) for (size_t b = 0; b < P.blk_off.size(); ++b) offsets[b] = P.blk_off[b] * (P.node ? 2 : 1);   // a node = two dofs
  });
}

int nsx_check_decoupled(nsx_ctx *ctx, int *yes) {
  return guarded(ctx, [&] {
    if (!yes) throw std::invalid_argument("null output pointer");
    need_final(*ctx);
    *yes = effective_view(*ctx);
  });
}

int nsx_synchronize(nsx_ctx *ctx) {
  return guarded(ctx, [&] { NSX_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

int nsx_time_kernel(nsx_ctx *ctx, int what, int reps, int flush_l2, double *ms_per_launch) {
  return guarded(ctx, [&] {
    Ctx &c = *ctx;
    need_final(c);
    if (reps < 1 || !ms_per_launch) throw std::invalid_argument("bad timing request");
    if (flush_l2 && !c.flush.p) { c.flush.alloc((size_t)512 << 20); c.flush.zero(c.stream); }
    cudaEvent_t e0, e1;
    NSX_CUDA(cudaEventCreate(&e0));
    NSX_CUDA(cudaEventCreate(&e1));
    double *x = c.vec[NSX_VEC_TMP0].p, *y = c.vec[NSX_VEC_TMP1].p;
    TriPlan *P = nullptr;
    const int view = (what == 1 || (what >= 5 && what <= 7)) ? effective_view(c) : 0;   // what the inner solves would run on the current values
    if (what == 5) { P = &tri_plan(c, NSX_BLOCK_F, view); tri_refresh_values(c, *P, c.F); }
    if (what == 6 || what == 7) { P = &tri_plan(c, NSX_BLOCK_F, view == 2 ? 2 : 0); ilu0_factor(c, *P, c.F); }
    double total = 0;
    const bool back_to_back = flush_l2 == 2;   // one event pair around all launches (for kernels whose input exceeds L2)
    if (back_to_back) NSX_CUDA(cudaEventRecord(e0, c.stream));
    for (int r = 0; r < reps; ++r) {
      if (flush_l2 == 1) flush_l2_cache(c, r);
      if (!back_to_back) NSX_CUDA(cudaEventRecord(e0, c.stream));
      switch (what) {
        case 0: block_spmv(c, x, y); break;
        case 1: spmv(c, view == 2 ? c.Kn : view == 1 ? c.Fd : c.F, x, y); break;   // (Kn: the timing vectors stand for node-layout ones)
        case 2: assemble_cells(c, c.time_mode, c.time_nu, c.time_dt, 1.0); break;
        case 3: vec_dot_dev(c, RED_SLOTS - 2, x, y, c.n); break;
        case 4: vec_axpy(c, y, 1e-9, x, c.n); break;
        case 5: sgs_apply(c, *P, y, x); break;
        case 6: ilu0_apply(c, *P, y, x); break;
        case 7: ilu0_factor(c, *P, c.F); break;
        case 8: spmv_probe(c, c.F, 0, x, y); break;   // stream F's values + columns only
        case 9: spmv_probe(c, c.F, 1, x, y); break;   // ... plus the x gather
        case 20: fp64_peak_launch(c, y); break;        // FP64 FMA peak: 148 x 8 x 256 threads x 8192 x 8 FMAs = 39.7 GFlop per launch
        default: throw std::invalid_argument("unknown kernel id");
      }
      if (back_to_back) continue;
      NSX_CUDA(cudaEventRecord(e1, c.stream));
      NSX_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      NSX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      total += ms;
    }
    if (back_to_back) {
      NSX_CUDA(cudaEventRecord(e1, c.stream));
      NSX_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      NSX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      total = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_launch = total / reps;
  });
}

int nsx_set_time_params(nsx_ctx *ctx, int mode, double nu, double dt) {
  return guarded(ctx, [&] { ctx->time_mode = mode; ctx->time_nu = nu; ctx->time_dt = dt; });
}

}  // extern "C"
