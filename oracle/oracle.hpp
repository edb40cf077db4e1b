// oracle.hpp -- data structures of the CPU restatement (TEST INFRASTRUCTURE ONLY, see oracle.cpp).
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace orc {

typedef std::vector<double> Vec;

struct CSR {
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

// reference-cell tables the oracle computes for itself (independent of the product's fe.hpp)
struct RefCell {
  int elem, nvpc, ndofs, nq, nqf, nfaces, nvn, npn;
  int comp[41], node[41];
  double qp[16][2], qw[16], qfp[4], qfw[4];
};

struct Problem {
  RefCell rc;
  int ncells = 0;
  int64_t n_u = 0, n_p = 0;
  std::vector<double> cell_vertices;
  std::vector<uint32_t> cell_dofs;
  CSR F, Bt, B, Mp, S;
  Vec solution, solution_old, delta, residual, eval_point;
  std::vector<uint32_t> bc_dof;
  std::vector<double> bc_val;  // inlet values (used when apply_inlet)
  std::vector<int> outlet_cell, outlet_face, cyl_cell, cyl_face;
  int nranks = 1;
  std::vector<int64_t> owned_u, owned_p;  // rank-local preconditioner blocks (Ifpack overlap 0)
  // optional sub-blocks + elimination sequences of the ILU / SGS sweeps (index 0 velocity, 1 pressure): see LocalBlocks
  std::vector<int32_t> ord[2], blk[2];
  std::vector<int64_t> ord_off[2];
  // bookkeeping of the last solve
  long inner_F_iters = 0, inner_S_iters = 0, precond_applies = 0;
  double lift_force = 0, drag_force = 0;
};

typedef std::function<void(Vec &dst, const Vec &src)> Op;

}  // namespace orc
