// C ABI over hostsetup.cpp (see include/nsx_host.h).
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include "../../include/nsx_host.h"
#include "hostsetup.hpp"

struct nsx_disc {
  nsx::Discretisation d;
  std::vector<int32_t> bfaces_flat;
};

static thread_local std::string g_host_err;

extern "C" {

const char *nsx_host_last_error(void) { return g_host_err.c_str(); }

static nsx_disc *finish(nsx_disc *h, int nranks) {
  nsx::partition_strips(h->d.mesh, nranks);
  nsx::build_discretisation(h->d);
  for (auto &bf : h->d.mesh.bfaces) {
    h->bfaces_flat.push_back(bf.cell);
    h->bfaces_flat.push_back(bf.face);
    h->bfaces_flat.push_back(bf.bid);
  }
  return h;
}

nsx_disc *nsx_disc_generate(int nx, int ny, int triangles, int nranks) {
  nsx_disc *h = nullptr;
  try {
    h = new nsx_disc;
    nsx::generate_mesh(nx, ny, triangles != 0, h->d.mesh);
    return finish(h, nranks);
  } catch (const std::exception &e) {
    g_host_err = e.what();
    delete h;
    return nullptr;
  }
}

nsx_disc *nsx_disc_from_gmsh(const char *path, int nranks) {
  nsx_disc *h = nullptr;
  try {
    h = new nsx_disc;
    nsx::read_gmsh2(path, h->d.mesh);
    return finish(h, nranks);
  } catch (const std::exception &e) {
    g_host_err = e.what();
    delete h;
    return nullptr;
  }
}

nsx_disc *nsx_disc_local(const nsx_disc *global, int rank) {
  nsx_disc *h = nullptr;
  try {
    if (!global) throw std::invalid_argument("null discretisation");
    h = new nsx_disc;
    nsx::build_local_view(global->d, rank, h->d);
    for (auto &bf : h->d.mesh.bfaces) {
      h->bfaces_flat.push_back(bf.cell);
      h->bfaces_flat.push_back(bf.face);
      h->bfaces_flat.push_back(bf.bid);
    }
    return h;
  } catch (const std::exception &e) {
    g_host_err = e.what();
    delete h;
    return nullptr;
  }
}

void nsx_disc_free(nsx_disc *d) { delete d; }

int64_t nsx_disc_info(const nsx_disc *h, int what) {
  const nsx::Discretisation &d = h->d;
  switch (what) {
    case NSX_DI_ELEM: return d.mesh.elem;
    case NSX_DI_NCELLS: return d.mesh.ncells();
    case NSX_DI_NVERTS: return d.mesh.nverts();
    case NSX_DI_N_U: return d.n_u;
    case NSX_DI_N_P: return d.n_p;
    case NSX_DI_DOFS_PER_CELL: return d.fe.ndofs;
    case NSX_DI_NQ: return d.fe.nq;
    case NSX_DI_NQF: return d.fe.nqf;
    case NSX_DI_NRANKS: return d.nranks;
    case NSX_DI_NBC: return (int64_t)d.bc_dof.size();
    case NSX_DI_NVPC: return d.mesh.nvpc;
    case NSX_DI_IS_LOCAL: return d.is_local ? 1 : 0;
    case NSX_DI_RANK: return d.rank;
    case NSX_DI_JOB_RANKS: return d.is_local ? d.job_ranks : d.nranks;
    case NSX_DI_N_U_OWNED: return d.is_local ? d.n_u_owned : d.n_u;
    case NSX_DI_N_P_OWNED: return d.is_local ? d.n_p_owned : d.n_p;
  }
  return -1;
}

#define RET(vec) do { *count = (int64_t)(vec).size(); return (vec).data(); } while (0)

const void *nsx_disc_array(const nsx_disc *h, int what, int64_t *count) {
  const nsx::Discretisation &d = h->d;
  int64_t dummy;
  if (!count) count = &dummy;
  switch (what) {
    case NSX_DA_CELL_DOFS: RET(d.cell_dofs);
    case NSX_DA_CELL_VERTICES: RET(d.cell_vertices);
    case NSX_DA_CELL_RANK: RET(d.mesh.cell_rank);
    case NSX_DA_OWNED_U: RET(d.owned_u);
    case NSX_DA_OWNED_P: RET(d.owned_p);
    case NSX_DA_F_ROWPTR: RET(d.F.rowptr);
    case NSX_DA_F_COL: RET(d.F.col);
    case NSX_DA_BT_ROWPTR: RET(d.Bt.rowptr);
    case NSX_DA_BT_COL: RET(d.Bt.col);
    case NSX_DA_B_ROWPTR: RET(d.B.rowptr);
    case NSX_DA_B_COL: RET(d.B.col);
    case NSX_DA_MP_ROWPTR: RET(d.Mp.rowptr);
    case NSX_DA_MP_COL: RET(d.Mp.col);
    case NSX_DA_BC_DOF: RET(d.bc_dof);
    case NSX_DA_BC_SHAPE: RET(d.bc_shape);
    case NSX_DA_BC_ON_INLET: RET(d.bc_on_inlet);
    case NSX_DA_BC_Y: RET(d.bc_y);
    case NSX_DA_OUTLET_CELL: RET(d.outlet_cell);
    case NSX_DA_OUTLET_FACE: RET(d.outlet_face);
    case NSX_DA_CYL_CELL: RET(d.cylinder_cell);
    case NSX_DA_CYL_FACE: RET(d.cylinder_face);
    case NSX_DA_BFACES: RET(h->bfaces_flat);
    case NSX_DA_MATERIAL: RET(d.mesh.material);
    case NSX_DA_L2G_U: RET(d.l2g_u);
    case NSX_DA_L2G_P: RET(d.l2g_p);
    case NSX_DA_CELL_GLOBAL: RET(d.cell_global);
    case NSX_DA_CELL_OWNED: RET(d.cell_owned);
    case NSX_DA_HALO_U_NBR: RET(d.halo_u.nbr);
    case NSX_DA_HALO_U_SEND_PTR: RET(d.halo_u.send_ptr);
    case NSX_DA_HALO_U_SEND_IDX: RET(d.halo_u.send_idx);
    case NSX_DA_HALO_U_RECV_PTR: RET(d.halo_u.recv_ptr);
    case NSX_DA_HALO_P_NBR: RET(d.halo_p.nbr);
    case NSX_DA_HALO_P_SEND_PTR: RET(d.halo_p.send_ptr);
    case NSX_DA_HALO_P_SEND_IDX: RET(d.halo_p.send_idx);
    case NSX_DA_HALO_P_RECV_PTR: RET(d.halo_p.recv_ptr);
    case NSX_DA_FE_TABLES: *count = (int64_t)sizeof(nsx::FETables); return &d.fe;
  }
  *count = 0;
  return nullptr;
}

void nsx_disc_inlet_values(const nsx_disc *h, double u, double *values) {
  const nsx::Discretisation &d = h->d;
  const double H = 0.41;
  for (size_t i = 0; i < d.bc_dof.size(); ++i) {
    const double y = d.bc_y[i];
    values[i] = d.bc_on_inlet[i] ? 4 * u * y * (H - y) / (H * H) : 0.0;
  }
}

}  // extern "C"
