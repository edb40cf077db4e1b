// Host-side stand-in for the set-up work that deal.II does in the reference
// (lab_new/src/NSSolverStationary.cpp:3-315 == NSSolver.cpp:3-311): mesh generation / gmsh
// reading, partitioning, DoF numbering with component-wise renumbering, block sparsity and the
// Dirichlet / boundary-face lists.  Its outputs are exactly the arrays the device C ABI
// (include/nsx.h) takes, so a deal.II adapter can replace this file without touching the device
// code.  None of this is on the timed hot path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fe.hpp"

namespace nsx {

struct BFace {
  int cell, face, bid;
};

struct Mesh {
  int elem = 0;  // 0 quads (Q3/Q2), 1 triangles (P2/P1)
  int nvpc = 4;
  std::vector<double> vx;        // 2 * n_vertices
  std::vector<int> cells;        // nvpc * n_cells
  std::vector<int> material;     // per cell
  std::vector<BFace> bfaces;     // boundary faces, cell-major, face-minor order
  std::vector<int> cell_rank;    // subdomain id per cell
  int ncells() const { return (int)(cells.size() / nvpc); }
  int nverts() const { return (int)(vx.size() / 2); }
};

struct CSRPattern {
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

struct Discretisation {
  Mesh mesh;
  FETables fe;
  int nranks = 1;
  int64_t n_u = 0, n_p = 0;
  std::vector<uint32_t> cell_dofs;       // ncells * ndofs, block-global numbering (p offset by n_u)
  std::vector<double> cell_vertices;     // ncells * nvpc * 2
  std::vector<int64_t> owned_u, owned_p; // nranks+1 offsets of the owned ranges in each block
  CSRPattern F, Bt, B, Mp;
  // Dirichlet data (velocity dofs only), ascending dof order
  std::vector<uint32_t> bc_dof;
  std::vector<double> bc_shape;          // inlet profile 4*y*(H-y)/H^2 with unit amplitude (0 off the inlet)
  std::vector<uint8_t> bc_on_inlet;      // 1 where the last writer was the inlet function
  std::vector<double> bc_y;              // support-point y coordinate (for the literal inlet formula)
  // boundary faces by id
  std::vector<int> outlet_cell, outlet_face;      // boundary id 8
  std::vector<int> cylinder_cell, cylinder_face;  // boundary id 10

  // ---- one rank's share (build_local_view); unused in the global discretisation ----
  // Local numbering per block: owned dofs first (global order), then ghosts ascending by global id, which
  // groups them by owner because owned ranges are contiguous.  In a local view n_u / n_p count owned + ghost
  // dofs, cell_dofs / patterns / boundary lists use local ids, and the patterns hold the owned rows only.
  bool is_local = false;
  int rank = 0, job_ranks = 1;
  int64_t n_u_owned = 0, n_p_owned = 0;
  std::vector<int64_t> l2g_u, l2g_p;     // local -> global id inside the block
  std::vector<int32_t> cell_global;      // local cell -> global cell
  std::vector<uint8_t> cell_owned;       // 1 where the cell's subdomain is this rank
  struct Halo {                          // ghost import plan of one block (what Epetra_Import holds in the reference)
    std::vector<int32_t> nbr;            // neighbour ranks, ascending
    std::vector<int64_t> send_ptr;       // per neighbour: range in send_idx
    std::vector<int32_t> send_idx;       // owned local ids whose values the neighbour needs, ascending
    std::vector<int64_t> recv_ptr;       // per neighbour: range of ghost slots (0 = first ghost) it fills
  } halo_u, halo_p;
};

// Rectangular channel 2.2 x 0.41 with nx x ny cells, cells whose centre lies inside the circle
// (0.2, 0.205), r = 0.05 removed (reference: NSSolverStationary.cpp:8-95).  With
// triangles = true every kept quad is split along its (v0,v3) diagonal -- a test-only helper that
// gives small simplex meshes with the same boundary ids.
void generate_mesh(int nx, int ny, bool triangles, Mesh &m);

// Gmsh 2.2 ASCII reader: triangles (type 2) become cells, lines (type 1) give boundary ids
// (reference: GridIn::read_msh at NSSolverStationary.cpp:155-160).
void read_gmsh2(const std::string &path, Mesh &m);

// Deterministic stand-in for GridTools::partition_triangulation (METIS is not available):
// equal-count strips by cell-centre x coordinate.
void partition_strips(Mesh &m, int nranks);

// FE + DoF numbering + sparsity + boundary lists (NSSolverStationary.cpp:114-314).
void build_discretisation(Discretisation &d);

// The share of rank `rank` of a partitioned discretisation: the cells that touch one of its dofs (its own cells
// plus a one-cell ghost layer, so that every owned matrix row can be assembled without exchanging matrix
// values), local dof numbering, owned-row sparsity, ghost import plans and the local boundary lists
// (reference: locally_owned / locally_relevant IndexSets, NSSolverStationary.cpp:226-242, and the
// mpi_communicator-aware sparsity at :276-305).
void build_local_view(const Discretisation &g, int rank, Discretisation &l);

}  // namespace nsx
