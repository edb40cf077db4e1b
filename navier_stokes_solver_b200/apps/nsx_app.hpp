// nsx_app.hpp -- what the two executables share: rank discovery, the host set-up stand-in (include/nsx_host.h
// in place of deal.II's setup()), the hand-over of its arrays to the device library (include/nsx.h), rank-0
// printing, and a minimal VTU writer.
//
// In the reference this is NSSolverStationary::setup() / NSSolver::setup() (lab_new/src/NSSolverStationary.cpp:
// 3-315, NSSolver.cpp:3-311) on top of deal.II; the hand-over below is the adapter INTEGRATION.md describes.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <ctime>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <memory>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsx.h"
#include "../../include/nsx_host.h"

namespace app {

// the exception classes the reference lets escape (deal.II's SolverControl::NoConvergence; std::invalid_argument)
struct NoConvergence : std::runtime_error { using std::runtime_error::runtime_error; };

// one process per GPU: rank / size from the launcher's environment (torchrun, mpirun, srun), 1 rank otherwise
struct Ranks {
  int rank = 0, size = 1, local_rank = 0;
  Ranks() {
    auto geti = [](std::initializer_list<const char *> names, int dflt) {
      for (const char *n : names) if (const char *v = std::getenv(n)) return std::atoi(v);
      return dflt;
    };
    rank = geti({"NSX_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
    size = geti({"NSX_NRANKS", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
    local_rank = geti({"NSX_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}, rank);
  }
};

// ConditionalOStream of the reference: rank 0 prints (NSSolverStationary.hpp:387)
struct Pcout {
  bool on;
  explicit Pcout(bool active) : on(active) {}
  template <class T> Pcout &operator<<(const T &v) { if (on) std::cout << v; return *this; }
  Pcout &operator<<(std::ostream &(*m)(std::ostream &)) { if (on) std::cout << m; return *this; }
  Pcout &operator<<(std::ios_base &(*m)(std::ios_base &)) { if (on) std::cout << m; return *this; }
};

// NSX_PRINT_DIGITS=n: after a coefficient line of the reference (6 significant digits), the same value with n digits
template <class Out> inline void print_more_digits(Out &pcout, const char *label, double value) {
  if (const char *e = std::getenv("NSX_PRINT_DIGITS")) {
    char buf[96];
    std::snprintf(buf, sizeof buf, "  [nsx] %s = %.*e", label, std::max(1, std::atoi(e)) - 1, value);
    pcout << buf << std::endl;
  }
}

inline void check(nsx_ctx *ctx, int rc, const char *what) {
  if (rc == NSX_OK) return;
  const std::string msg = std::string(what) + ": " + (ctx ? nsx_last_error(ctx) : "no context");
  if (rc == NSX_E_NOCONV) throw NoConvergence(msg);
  if (rc == NSX_E_BADARG) throw std::invalid_argument(ctx ? nsx_last_error(ctx) : what);
  throw std::runtime_error(msg + " (nsx status " + std::to_string(rc) + ")");
}

// NCCL id hand-over between the processes of one job: rank 0 writes the 128 bytes to a file, the others wait for it.
// The name carries a job-unique token when the launcher provides one (torchrun's run id, the SLURM job + step, Open MPI's
// job id, or NSX_JOB_ID), else the rendezvous port and the parent pid.  Rank 0 removes a leftover of the same name first,
// creates the file exclusively (no symlink following) and removes it again once nsx_comm_init has returned (the NCCL
// bootstrap is collective: every rank has read the id by then); the other ranks ignore files older than their own start.
// The directory (NSX_ID_DIR, default /tmp) must be visible to every rank: a launch that spans nodes has to set it.
inline std::string comm_id_path(const Ranks &r) {
  auto env = [](const char *n) -> std::string { const char *v = std::getenv(n); return v ? v : ""; };
  const std::string dir = env("NSX_ID_DIR");
  long nodes = 1;
  if (!env("SLURM_NNODES").empty()) nodes = std::atol(env("SLURM_NNODES").c_str());
  else if (!env("OMPI_MCA_orte_num_nodes").empty()) nodes = std::atol(env("OMPI_MCA_orte_num_nodes").c_str());
  else if (!env("LOCAL_WORLD_SIZE").empty() && std::atol(env("LOCAL_WORLD_SIZE").c_str()) > 0)
    nodes = (r.size + std::atol(env("LOCAL_WORLD_SIZE").c_str()) - 1) / std::atol(env("LOCAL_WORLD_SIZE").c_str());
  if (nodes > 1 && dir.empty())
    throw std::runtime_error("this launch spans " + std::to_string(nodes) + " nodes: set NSX_ID_DIR to a directory every node can see (the NCCL id is handed over through a file)");
  std::string token = env("NSX_JOB_ID");
  if (token.empty()) token = env("TORCHELASTIC_RUN_ID");
  if (token.empty() && !env("SLURM_JOB_ID").empty()) token = "slurm" + env("SLURM_JOB_ID") + "." + env("SLURM_STEP_ID");
  if (token.empty()) token = env("OMPI_MCA_ess_base_jobid");
  if (token.empty() || token == "none") token = "p" + env("MASTER_PORT") + "_" + std::to_string((long)getppid());
  for (char &ch : token) if (!(std::isalnum((unsigned char)ch) || ch == '.' || ch == '_' || ch == '-')) ch = '_';
  return (dir.empty() ? std::string("/tmp") : dir) + "/nsx_comm_id_" + std::to_string((long)getuid()) + "_" + token;
}

inline void exchange_comm_id(const Ranks &r, unsigned char id[128], const std::string &path) {
  const time_t started = time(nullptr);
  if (r.rank == 0) {
    if (nsx_comm_unique_id(id) != NSX_OK) throw std::runtime_error("nsx_comm_unique_id failed (NCCL not loadable?)");
    const std::string tmp = path + ".tmp";
    unlink(path.c_str());
    unlink(tmp.c_str());
    const int fd = open(tmp.c_str(), O_CREAT | O_EXCL | O_WRONLY | O_NOFOLLOW, 0600);
    if (fd < 0) throw std::runtime_error("cannot create the NCCL id file " + tmp + ": " + std::strerror(errno));
    const ssize_t w = write(fd, id, 128);
    close(fd);
    if (w != 128 || std::rename(tmp.c_str(), path.c_str()) != 0) { unlink(tmp.c_str()); throw std::runtime_error("cannot write the NCCL id file " + path); }
  } else {
    for (int tries = 0;; ++tries) {
      struct stat st;
      if (lstat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == getuid() && st.st_size == 128 && st.st_mtime >= started - 120) {
        std::ifstream f(path, std::ios::binary);
        if (f && f.read((char *)id, 128)) break;
      }
      if (tries > 6000) throw std::runtime_error("timed out waiting for the NCCL id file " + path);
      usleep(10000);
    }
  }
}

// Owns the discretisation (global + this rank's view) and the device context.
struct Problem {
  Ranks ranks;
  nsx_disc *global = nullptr, *view = nullptr;  // view == global on one rank
  nsx_ctx *ctx = nullptr;
  int64_t n_u_owned = 0, n_p_owned = 0;

  mutable std::future<void> vtu_writer;   // the output file being written behind the solver (write_vtu)

  ~Problem() {
    if (vtu_writer.valid()) vtu_writer.wait();
    if (ctx) nsx_destroy(ctx);
    if (view && view != global) nsx_disc_free(view);
    if (global) nsx_disc_free(global);
  }
  int64_t info(int what) const { return nsx_disc_info(view, what); }
  int64_t ginfo(int what) const { return nsx_disc_info(global, what); }
  template <class T> const T *arr(int what, int64_t *count = nullptr) const {
    int64_t c = 0;
    const void *p = nsx_disc_array(view, what, &c);
    if (count) *count = c;
    return static_cast<const T *>(p);
  }

  void make_mesh(bool from_file, const std::string &mesh_file, int nx, int ny) {
    global = from_file ? nsx_disc_from_gmsh(mesh_file.c_str(), ranks.size) : nsx_disc_generate(nx, ny, 0, ranks.size);
    if (!global) throw std::runtime_error(nsx_host_last_error());
    view = ranks.size > 1 ? nsx_disc_local(global, ranks.rank) : global;
    if (!view) throw std::runtime_error(nsx_host_last_error());
  }

  // hands the outputs of setup() to the device library (the adapter of INTEGRATION.md section 2)
  void to_device(double inlet_amplitude) {
    int ndev_rc = nsx_create(ranks.rank, ranks.size, ranks.local_rank, nullptr, &ctx);
    if (ndev_rc != NSX_OK) throw std::runtime_error("nsx_create failed: no usable CUDA device (this build has no CPU path)");
    // library options from the environment; unset = the library's defaults
    const struct { const char *env; int opt; } opts[] = {{"NSX_ORDERING", NSX_OPT_ORDERING}, {"NSX_ORTHO", NSX_OPT_ORTHO}, {"NSX_BLOCK_ROWS", NSX_OPT_BLOCK_ROWS},
                                                        {"NSX_DECOUPLE", NSX_OPT_DECOUPLE}, {"NSX_HOST_INNER", NSX_OPT_HOST_INNER}, {"NSX_VERBOSE", NSX_OPT_VERBOSE},
                                                        {"NSX_L2_HINTS", NSX_OPT_L2_HINTS}, {"NSX_PRECOND_LAG", NSX_OPT_PRECOND_LAG}, {"NSX_SWEEP_Q", NSX_OPT_SWEEP_Q}};
    for (const auto &o : opts)
      if (const char *v = std::getenv(o.env)) check(ctx, nsx_set_option(ctx, o.opt, std::atoll(v)), o.env);
    const bool local = ranks.size > 1;
    const int64_t n_u = info(NSX_DI_N_U), n_p = info(NSX_DI_N_P);
    n_u_owned = local ? info(NSX_DI_N_U_OWNED) : n_u;
    n_p_owned = local ? info(NSX_DI_N_P_OWNED) : n_p;
    check(ctx, nsx_set_discretisation(ctx, (int)info(NSX_DI_ELEM), info(NSX_DI_NCELLS), arr<double>(NSX_DA_CELL_VERTICES),
                                      arr<uint32_t>(NSX_DA_CELL_DOFS), n_u, n_p), "nsx_set_discretisation");
    if (local) {
      check(ctx, nsx_set_partition(ctx, n_u_owned, n_p_owned), "nsx_set_partition");
      const int base[2] = {NSX_DA_HALO_U_NBR, NSX_DA_HALO_P_NBR};
      for (int blk = 0; blk < 2; ++blk) {
        int64_t nn = 0;
        const int32_t *nbr = arr<int32_t>(base[blk], &nn);
        check(ctx, nsx_set_halo(ctx, blk, (int)nn, nbr, arr<int64_t>(base[blk] + 1), arr<int32_t>(base[blk] + 2), arr<int64_t>(base[blk] + 3)), "nsx_set_halo");
      }
    }
    const int blocks[4] = {NSX_BLOCK_F, NSX_BLOCK_BT, NSX_BLOCK_B, NSX_BLOCK_MP};
    const int rp[4] = {NSX_DA_F_ROWPTR, NSX_DA_BT_ROWPTR, NSX_DA_B_ROWPTR, NSX_DA_MP_ROWPTR};
    const int64_t rows[4] = {n_u_owned, n_u_owned, n_p_owned, n_p_owned}, cols[4] = {n_u, n_p, n_u, n_p};
    for (int b = 0; b < 4; ++b)
      check(ctx, nsx_set_pattern(ctx, blocks[b], rows[b], cols[b], arr<int64_t>(rp[b]), arr<int32_t>(rp[b] + 1)), "nsx_set_pattern");
    int64_t cnt = 0;
    const int32_t *oc = arr<int32_t>(NSX_DA_OUTLET_CELL, &cnt);
    check(ctx, nsx_set_faces(ctx, 8, cnt, oc, arr<int32_t>(NSX_DA_OUTLET_FACE)), "nsx_set_faces(8)");
    const int32_t *cc = arr<int32_t>(NSX_DA_CYL_CELL, &cnt);
    check(ctx, nsx_set_faces(ctx, 10, cnt, cc, arr<int32_t>(NSX_DA_CYL_FACE)), "nsx_set_faces(10)");
    int64_t nbc = 0;
    const uint32_t *bc = arr<uint32_t>(NSX_DA_BC_DOF, &nbc);
    std::vector<double> inlet((size_t)nbc);
    nsx_disc_inlet_values(view, inlet_amplitude, inlet.data());
    check(ctx, nsx_set_dirichlet(ctx, nbc, bc, inlet.data()), "nsx_set_dirichlet");
    if (local) {
      unsigned char id[128];
      const std::string id_path = comm_id_path(ranks);
      exchange_comm_id(ranks, id, id_path);
      const int rc_comm = nsx_comm_init(ctx, id);
      if (ranks.rank == 0) unlink(id_path.c_str());   // every rank has joined (or the bootstrap failed): the file has served
      check(ctx, rc_comm, "nsx_comm_init");
    } else {
      check(ctx, nsx_set_ranks(ctx, 1, arr<int64_t>(NSX_DA_OWNED_U), arr<int64_t>(NSX_DA_OWNED_P)), "nsx_set_ranks");
    }
    check(ctx, nsx_finalize_setup(ctx), "nsx_finalize_setup");
  }

  // Minimal stand-in for DataOut::write_vtu_with_pvtu_record (NSSolverStationary.cpp:765-800; deal.II keeps the real
  // one): this rank's owned cells as linear cells with the vertex values of velocity and pressure.
  // Asynchronous: the calling thread only takes a snapshot of the solution (device -> host, plus the ghost values on a
  // partitioned run); formatting and writing happen on a worker thread while the solver goes on -- one file in flight, the
  // next call (or the destructor) waits for it and rethrows what it threw.  NSX_SYNC_OUTPUT=1 writes on the calling thread.
  struct VtuJob {
    std::vector<double> sol, gu, gp;
    std::vector<int64_t> cells;
    const double *cv; const uint32_t *cd;
    int64_t nvpc, dpc, n_u, n_u_owned, n_p_owned;
    int rank;
    std::string name;
    void run() const {
      auto u_at = [&](uint32_t d) { return d < (uint64_t)n_u_owned ? sol[d] : gu[d - n_u_owned]; };
      auto p_at = [&](uint32_t d) { const int64_t q = (int64_t)d - n_u; return q < n_p_owned ? sol[n_u_owned + q] : gp[q - n_p_owned]; };
      std::ofstream f(name);
      if (!f) throw std::runtime_error("cannot open " + name);
      const int64_t nc = (int64_t)cells.size(), np = nc * nvpc;
      f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n"
        << "<Piece NumberOfPoints=\"" << np << "\" NumberOfCells=\"" << nc << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
      f.precision(12);
      for (int64_t c : cells) for (int v = 0; v < nvpc; ++v) f << cv[(c * nvpc + v) * 2] << " " << cv[(c * nvpc + v) * 2 + 1] << " 0\n";
      f << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n";
      // deal.II numbers quad vertices lexicographically; VTK_QUAD wants them counter-clockwise
      const int quad_order[4] = {0, 1, 3, 2}, tri_order[3] = {0, 1, 2};
      for (int64_t k = 0; k < nc; ++k) { for (int v = 0; v < nvpc; ++v) f << k * nvpc + (nvpc == 4 ? quad_order[v] : tri_order[v]) << " "; f << "\n"; }
      f << "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n";
      for (int64_t k = 1; k <= nc; ++k) f << k * nvpc << "\n";
      f << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
      for (int64_t k = 0; k < nc; ++k) f << (nvpc == 4 ? 9 : 5) << "\n";
      f << "</DataArray>\n</Cells>\n<PointData Vectors=\"velocity\" Scalars=\"pressure\">\n<DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n";
      // FESystem cell-local order: vertex v carries [u_x, u_y, p] at 3v, 3v+1, 3v+2
      for (int64_t c : cells) for (int v = 0; v < nvpc; ++v) f << u_at(cd[c * dpc + 3 * v]) << " " << u_at(cd[c * dpc + 3 * v + 1]) << " 0\n";
      f << "</DataArray>\n<DataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\">\n";
      for (int64_t c : cells) for (int v = 0; v < nvpc; ++v) f << p_at(cd[c * dpc + 3 * v + 2]) << "\n";
      f << "</DataArray>\n</PointData>\n<CellData Scalars=\"partitioning\">\n<DataArray type=\"Float64\" Name=\"partitioning\" format=\"ascii\">\n";
      for (int64_t k = 0; k < nc; ++k) f << rank << "\n";
      f << "</DataArray>\n</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
      if (!f) throw std::runtime_error("error writing " + name);
    }
  };

  void write_vtu(const std::string &stem, unsigned index) const {
    auto job = std::make_shared<VtuJob>();
    const int64_t ncells = info(NSX_DI_NCELLS);
    job->nvpc = info(NSX_DI_NVPC); job->dpc = info(NSX_DI_DOFS_PER_CELL); job->n_u = info(NSX_DI_N_U);
    job->n_u_owned = n_u_owned; job->n_p_owned = n_p_owned; job->rank = ranks.rank;
    job->sol.resize((size_t)(n_u_owned + n_p_owned));
    check(ctx, nsx_vec_download(ctx, NSX_VEC_SOLUTION, job->sol.data()), "nsx_vec_download");
    if (ranks.size > 1) {
      job->gu.resize((size_t)(job->n_u - n_u_owned)); job->gp.resize((size_t)(info(NSX_DI_N_P) - n_p_owned));
      check(ctx, nsx_halo_exchange(ctx, NSX_VEC_SOLUTION), "nsx_halo_exchange");
      check(ctx, nsx_vec_download_ghosts(ctx, NSX_VEC_SOLUTION, job->gu.data(), job->gp.data()), "nsx_vec_download_ghosts");
    }
    job->cv = arr<double>(NSX_DA_CELL_VERTICES);   // the discretisation outlives the writer (~Problem waits for it)
    job->cd = arr<uint32_t>(NSX_DA_CELL_DOFS);
    const uint8_t *owned = ranks.size > 1 ? arr<uint8_t>(NSX_DA_CELL_OWNED) : nullptr;
    for (int64_t c = 0; c < ncells; ++c) if (!owned || owned[c]) job->cells.push_back(c);
    char name[256];
    snprintf(name, sizeof name, "%s_%u.%d.vtu", stem.c_str(), index, ranks.rank);
    job->name = name;
    if (vtu_writer.valid()) vtu_writer.get();   // one file in flight
    const char *sync = std::getenv("NSX_SYNC_OUTPUT");
    if (sync && std::atoi(sync)) job->run();
    else vtu_writer = std::async(std::launch::async, [job] { job->run(); });
  }
};

}  // namespace app
