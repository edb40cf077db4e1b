// nsx_dealii_adapter.cc -- see nsx_dealii_adapter.h.  Compiled only where deal.II (with Trilinos and MPI) is installed.
#include "nsx_dealii_adapter.h"

#include <deal.II/base/mpi.h>
#include <deal.II/fe/component_mask.h>
#include <deal.II/lac/trilinos_sparse_matrix.h>

#include <algorithm>
#include <stdexcept>
#include <string>

namespace nsx_dealii {

namespace {

void ck(nsx_ctx *ctx, int rc, const char *what) {
  if (rc != NSX_OK) throw std::runtime_error(std::string(what) + ": " + (ctx ? nsx_last_error(ctx) : "no context"));
}

// Local numbering of one block on this rank: the owned dofs first (ascending global id), then the locally relevant, not owned
// ones grouped by owner rank and ascending inside a group -- the order nsx_set_halo expects for the ghost ranges.
struct LocalBlock {
  std::vector<dealii::types::global_dof_index> l2g;   // local -> block-global id
  std::map<dealii::types::global_dof_index, int32_t> g2l;
  int64_t n_owned = 0;
  std::vector<int32_t> nbr;                            // neighbour ranks that own ghosts of ours
  std::vector<int64_t> recv_ptr{0};
};

LocalBlock number_block(const dealii::IndexSet &owned, const dealii::IndexSet &relevant, const std::vector<dealii::IndexSet> &owned_per_rank, int my_rank) {
  LocalBlock L;
  for (auto i : owned) L.l2g.push_back(i);
  L.n_owned = (int64_t)L.l2g.size();
  dealii::IndexSet ghosts = relevant;
  ghosts.subtract_set(owned);
  for (int r = 0; r < (int)owned_per_rank.size(); ++r) {
    if (r == my_rank) continue;
    dealii::IndexSet from_r = ghosts & owned_per_rank[r];
    if (from_r.n_elements() == 0) continue;
    L.nbr.push_back(r);
    for (auto i : from_r) L.l2g.push_back(i);
    L.recv_ptr.push_back((int64_t)L.l2g.size() - L.n_owned);
  }
  for (size_t l = 0; l < L.l2g.size(); ++l) L.g2l[L.l2g[l]] = (int32_t)l;
  return L;
}

// owned rows of a Trilinos block in local numbering, columns ascending
void push_pattern(nsx_ctx *ctx, int block, const dealii::TrilinosWrappers::SparseMatrix &A, const LocalBlock &rows, const LocalBlock &cols) {
  std::vector<int64_t> rowptr{0};
  std::vector<int32_t> col;
  for (int64_t lr = 0; lr < rows.n_owned; ++lr) {
    const auto gr = rows.l2g[lr];
    const size_t begin = col.size();
    for (auto it = A.begin(gr); it != A.end(gr); ++it) col.push_back(cols.g2l.at(it->column()));
    std::sort(col.begin() + begin, col.end());
    rowptr.push_back((int64_t)col.size());
  }
  ck(ctx, nsx_set_pattern(ctx, block, rows.n_owned, (int64_t)cols.l2g.size(), rowptr.data(), col.data()), "nsx_set_pattern");
}

}  // namespace

template <int dim>
nsx_ctx *hand_over(const SetupView<dim> &s, const dealii::Function<dim> &inlet, int device_id) {
  using namespace dealii;
  const int rank = Utilities::MPI::this_mpi_process(s.comm), nranks = Utilities::MPI::n_mpi_processes(s.comm);
  nsx_ctx *ctx = nullptr;
  if (nsx_create(rank, nranks, device_id, nullptr, &ctx) != NSX_OK) throw std::runtime_error("nsx_create failed: no usable CUDA device");

  // owned index sets of every rank (block-local ids): who owns which ghost
  std::vector<IndexSet> owned_u = Utilities::MPI::all_gather(s.comm, s.block_owned_dofs[0]);
  std::vector<IndexSet> owned_p = Utilities::MPI::all_gather(s.comm, s.block_owned_dofs[1]);
  const LocalBlock U = number_block(s.block_owned_dofs[0], s.block_relevant_dofs[0], owned_u, rank);
  const LocalBlock P = number_block(s.block_owned_dofs[1], s.block_relevant_dofs[1], owned_p, rank);
  const types::global_dof_index n_u_global = s.block_owned_dofs[0].size();

  // cells: locally owned ones plus the ghost cells (they touch owned dofs; their contributions to owned rows are assembled
  // redundantly on this rank, which replaces compress(VectorOperation::add), NSSolverStationary.cpp:535-537)
  const unsigned dpc = s.fe.n_dofs_per_cell();
  std::vector<uint32_t> cell_dofs;
  std::vector<double> cell_vertices;
  std::vector<int32_t> outlet_cell, outlet_face, cyl_cell, cyl_face;
  std::vector<types::global_dof_index> idx(dpc);
  int32_t local_cell = 0;
  for (const auto &cell : s.dof_handler.active_cell_iterators()) {
    if (!(cell->is_locally_owned() || cell->is_ghost())) continue;
    cell->get_dof_indices(idx);
    for (unsigned i = 0; i < dpc; ++i) {
      const bool pressure = idx[i] >= n_u_global;
      const int32_t l = pressure ? (int32_t)U.l2g.size() + P.g2l.at(idx[i] - n_u_global) : U.g2l.at(idx[i]);
      cell_dofs.push_back((uint32_t)l);
    }
    for (unsigned v = 0; v < cell->n_vertices(); ++v)
      for (unsigned k = 0; k < dim; ++k) cell_vertices.push_back(cell->vertex(v)[k]);
    if (cell->is_locally_owned() && cell->at_boundary())
      for (unsigned f = 0; f < cell->n_faces(); ++f)
        if (cell->face(f)->at_boundary()) {
          if (cell->face(f)->boundary_id() == 8) { outlet_cell.push_back(local_cell); outlet_face.push_back((int32_t)f); }
          if (cell->face(f)->boundary_id() == 10) { cyl_cell.push_back(local_cell); cyl_face.push_back((int32_t)f); }
        }
    ++local_cell;
  }
  ck(ctx, nsx_set_discretisation(ctx, s.simplex ? 1 : 0, local_cell, cell_vertices.data(), cell_dofs.data(), (int64_t)U.l2g.size(), (int64_t)P.l2g.size()),
     "nsx_set_discretisation");

  if (nranks > 1) {
    ck(ctx, nsx_set_partition(ctx, U.n_owned, P.n_owned), "nsx_set_partition");
    // send lists: what each neighbour's ghost range asks of us, in ITS order = ascending global id inside our owned set
    const LocalBlock *blocks[2] = {&U, &P};
    const std::vector<IndexSet> *all_owned[2] = {&owned_u, &owned_p};
    for (int b = 0; b < 2; ++b) {
      const LocalBlock &L = *blocks[b];
      const IndexSet &mine = s.block_owned_dofs[b];
      // every rank's relevant set, to know which of my dofs it imports
      std::vector<IndexSet> relevant = Utilities::MPI::all_gather(s.comm, s.block_relevant_dofs[b]);
      std::vector<int32_t> nbr, send_idx;
      std::vector<int64_t> send_ptr{0}, recv_ptr{0};
      for (int r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        const IndexSet wanted = relevant[r] & mine;                                   // what r imports from me
        const auto pos = std::find(L.nbr.begin(), L.nbr.end(), r);                    // what I import from r
        const int64_t n_recv = pos == L.nbr.end() ? 0 : L.recv_ptr[pos - L.nbr.begin() + 1] - L.recv_ptr[pos - L.nbr.begin()];
        if (wanted.n_elements() == 0 && n_recv == 0) continue;
        nbr.push_back(r);
        for (auto g : wanted) send_idx.push_back(L.g2l.at(g));
        send_ptr.push_back((int64_t)send_idx.size());
        recv_ptr.push_back(recv_ptr.back() + n_recv);
      }
      (void)all_owned;
      ck(ctx, nsx_set_halo(ctx, b, (int)nbr.size(), nbr.data(), send_ptr.data(), send_idx.data(), recv_ptr.data()), "nsx_set_halo");
    }
  }

  push_pattern(ctx, NSX_BLOCK_F, s.jacobian_matrix.block(0, 0), U, U);
  push_pattern(ctx, NSX_BLOCK_BT, s.jacobian_matrix.block(0, 1), U, P);
  push_pattern(ctx, NSX_BLOCK_B, s.jacobian_matrix.block(1, 0), P, U);
  push_pattern(ctx, NSX_BLOCK_MP, s.pressure_mass.block(1, 1), P, P);
  ck(ctx, nsx_set_faces(ctx, 8, (int64_t)outlet_cell.size(), outlet_cell.data(), outlet_face.data()), "nsx_set_faces(8)");
  ck(ctx, nsx_set_faces(ctx, 10, (int64_t)cyl_cell.size(), cyl_cell.data(), cyl_face.data()), "nsx_set_faces(10)");

  // Dirichlet dofs: boundaries 7 (inlet), 6 and 10 (no slip), velocity components only (NSSolverStationary.cpp:541-572).  The
  // library applies `inlet_value` in the one non-homogeneous assembly and zero afterwards.
  {
    std::map<types::global_dof_index, double> inlet_values, zero_values;
    const ComponentMask velocity_mask = []() { std::vector<bool> m(dim + 1, true); m[dim] = false; return ComponentMask(m); }();
    Functions::ZeroFunction<dim> zero(dim + 1);
    std::map<types::boundary_id, const Function<dim> *> in{{7, &inlet}}, walls{{6, &zero}, {10, &zero}};
    VectorTools::interpolate_boundary_values(s.dof_handler, in, inlet_values, velocity_mask);
    VectorTools::interpolate_boundary_values(s.dof_handler, walls, zero_values, velocity_mask);
    for (const auto &kv : zero_values) inlet_values.emplace(kv.first, 0.0);          // first writer wins, as the reference's map
    std::vector<uint32_t> dof;
    std::vector<double> val;
    for (const auto &kv : inlet_values) {                                             // std::map: ascending global id
      const auto it = U.g2l.find(kv.first);
      if (it == U.g2l.end() || it->second >= U.n_owned) continue;                     // constrained rows are owned rows
      dof.push_back((uint32_t)it->second);
      val.push_back(kv.second);
    }
    // local ids ascend with the global ids inside the owned range
    ck(ctx, nsx_set_dirichlet(ctx, (int64_t)dof.size(), dof.data(), val.data()), "nsx_set_dirichlet");
  }

  if (nranks > 1) {
    unsigned char id[128] = {0};
    if (rank == 0) ck(ctx, nsx_comm_unique_id(id), "nsx_comm_unique_id");
    MPI_Bcast(id, 128, MPI_BYTE, 0, s.comm);
    ck(ctx, nsx_comm_init(ctx, id), "nsx_comm_init");
  } else {
    const int64_t ou[2] = {0, U.n_owned}, op[2] = {0, P.n_owned};
    ck(ctx, nsx_set_ranks(ctx, 1, ou, op), "nsx_set_ranks");
  }
  ck(ctx, nsx_finalize_setup(ctx), "nsx_finalize_setup");
  return ctx;
}

void download_solution(nsx_ctx *ctx, dealii::TrilinosWrappers::MPI::BlockVector &solution_owned) {
  const auto n_u = solution_owned.block(0).locally_owned_size(), n_p = solution_owned.block(1).locally_owned_size();
  std::vector<double> host(n_u + n_p);
  ck(ctx, nsx_vec_download(ctx, NSX_VEC_SOLUTION, host.data()), "nsx_vec_download");
  size_t k = 0;
  for (unsigned b = 0; b < 2; ++b)
    for (auto i : solution_owned.block(b).locally_owned_elements()) solution_owned.block(b)[i] = host[k++];
  solution_owned.compress(dealii::VectorOperation::insert);
}

template nsx_ctx *hand_over<2>(const SetupView<2> &, const dealii::Function<2> &, int);

}  // namespace nsx_dealii
