// comm.cu -- the two exchanges of the row-partitioned path: ghost import and scalar reduction.
//
// The reference runs one MPI rank per subdomain; every rank owns a contiguous row range of each block
// (NSSolverStationary.cpp:226-242).  Two kinds of traffic cross ranks inside the hot path:
//   * ghost import before a matrix-vector product or an assembly (`solution = solution_owned`,
//     NSSolverStationary.cpp:722; Epetra's Import inside every vmult),
//   * the sums behind dot products and norms of the Krylov solvers (MPI_Allreduce inside Trilinos).
// Here: one process per GPU, NCCL over NVLink.  The ghost import packs the owned entries a neighbour needs
// (one gather kernel) and posts one grouped ncclSend / ncclRecv per neighbour straight into the ghost tail of
// the destination vector; reductions are an in-place ncclAllReduce on the device scalar slots the reduction
// kernels already write, so the host still reads one number per synchronisation point.
// NCCL is loaded at run time (dlopen): a single-GPU user needs no NCCL at all, and inside a torch process the
// library torch already loaded is the one that gets used.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "device.cuh"

namespace nsx {

namespace {

// the handful of NCCL entry points we use, with the public NCCL 2.x C signatures
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

struct CommError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

Nccl &nccl() {
  static Nccl N;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      N.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (N.lib) break;
    }
    if (!N.lib) { err = std::string("cannot load NCCL: ") + dlerror(); return; }
    auto sym = [&](const char *s) {
      void *p = dlsym(N.lib, s);
      if (!p) err = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    N.GetUniqueId = (decltype(N.GetUniqueId))sym("ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))sym("ncclCommInitRank");
    N.CommDestroy = (decltype(N.CommDestroy))sym("ncclCommDestroy");
    N.AllReduce = (decltype(N.AllReduce))sym("ncclAllReduce");
    N.Send = (decltype(N.Send))sym("ncclSend");
    N.Recv = (decltype(N.Recv))sym("ncclRecv");
    N.GroupStart = (decltype(N.GroupStart))sym("ncclGroupStart");
    N.GroupEnd = (decltype(N.GroupEnd))sym("ncclGroupEnd");
    N.GetErrorString = (decltype(N.GetErrorString))sym("ncclGetErrorString");
  });
  if (!err.empty()) throw CommError(err);
  return N;
}

#define NSX_NCCL(call)                                                                                   \
  do {                                                                                                   \
    ncclResult_t r_ = (call);                                                                            \
    if (r_ != 0) throw CommError(std::string(#call) + ": " + nccl().GetErrorString(r_));                 \
  } while (0)

__global__ void k_pack(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = x[idx[i]];
}

}  // namespace

void halo_exchange(Ctx &c, int blk, const double *base_, bool node_layout) {
  HaloPlan &H = blk ? c.halo_p : c.halo_u;
  if (!c.comm || H.nbr.empty()) return;
  Nccl &N = nccl();
  double *base = const_cast<double *>(base_);  // the ghost tail of an input vector is scratch by construction
  double *ghost = blk ? base + c.n_p + c.n_ug : base + c.n;
  if (H.nsend) {
    if (node_layout && (blk != 0 || !H.send_idx_node.p)) throw std::logic_error("no ghost import plan for the node layout");
    k_pack<<<(int)((H.nsend + 255) / 256), 256, 0, c.stream>>>(H.nsend, node_layout ? H.send_idx_node.p : H.send_idx.p, base, H.send_buf.p);
    c.stat_launches++;
  }
  ncclComm_t comm = (ncclComm_t)c.comm;
  NSX_NCCL(N.GroupStart());
  for (size_t i = 0; i < H.nbr.size(); ++i) {
    const int64_t ns = H.send_ptr[i + 1] - H.send_ptr[i], nr = H.recv_ptr[i + 1] - H.recv_ptr[i];
    if (ns) NSX_NCCL(N.Send(H.send_buf.p + H.send_ptr[i], (size_t)ns, kNcclFloat64, H.nbr[i], comm, c.stream));
    if (nr) NSX_NCCL(N.Recv(ghost + H.recv_ptr[i], (size_t)nr, kNcclFloat64, H.nbr[i], comm, c.stream));
  }
  NSX_NCCL(N.GroupEnd());
  c.stat_halo++;
}

void allreduce_slots(Ctx &c, int slot, int count) {
  if (!c.comm) return;
  Nccl &N = nccl();
  double *p = slot_ptr(c, slot);
  NSX_NCCL(N.AllReduce(p, p, (size_t)count, kNcclFloat64, kNcclSum, (ncclComm_t)c.comm, c.stream));
  c.stat_allreduce++;
}

void comm_destroy(Ctx &c) {
  if (!c.comm) return;
  try { nccl().CommDestroy((ncclComm_t)c.comm); } catch (...) {}
  c.comm = nullptr;
}

}  // namespace nsx

using namespace nsx;

extern "C" {

int nsx_comm_unique_id(void *id128) {
  if (!id128) return NSX_E_BADARG;
  try {
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != 0) return NSX_E_COMM;
    std::memcpy(id128, &id, sizeof(id));
    return NSX_OK;
  } catch (const std::exception &) { return NSX_E_COMM; }
}

int nsx_comm_init(nsx_ctx *ctx, const void *id128) {
  if (!ctx || !id128) return NSX_E_BADARG;
  try {
    NSX_CUDA(cudaSetDevice(ctx->device));
    if (ctx->nranks == 1) return NSX_OK;
    comm_destroy(*ctx);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NSX_NCCL(nccl().CommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->comm = comm;
    return NSX_OK;
  } catch (const CudaError &e) { ctx->err = e.what(); return NSX_E_CUDA; }
  catch (const std::exception &e) { ctx->err = e.what(); return NSX_E_COMM; }
}

}  // extern "C"
