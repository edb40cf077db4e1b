/* nsx_host.h -- C ABI of the host-side set-up stand-in.
 *
 * In the reference all of this is done by deal.II inside NSSolverStationary::setup() /
 * NSSolver::setup() (lab_new/src/NSSolverStationary.cpp:3-315, NSSolver.cpp:3-311): mesh
 * generation or GridIn::read_msh, partitioning, distribute_dofs + component_wise, IndexSets,
 * block sparsity.  Where deal.II is available an adapter fills the device ABI (nsx.h) from the
 * real DoFHandler; here these entry points produce the same arrays from scratch.
 * Nothing in this header runs on the GPU.
 */
#ifndef NSX_HOST_H
#define NSX_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsx_disc nsx_disc;

/* Generated channel mesh (reference: NSSolverStationary.cpp:8-112).  triangles != 0 splits every
 * quad into two P2/P1 triangles (test helper).  nranks > 1 partitions the cells into strips. */
nsx_disc *nsx_disc_generate(int nx, int ny, int triangles, int nranks);
/* Gmsh 2.2 mesh (reference: NSSolverStationary.cpp:146-176), P2/P1. */
nsx_disc *nsx_disc_from_gmsh(const char *path, int nranks);
/* One rank's share of a partitioned discretisation (what the locally_owned / locally_relevant index sets,
 * the owned rows of the sparsity pattern and the Epetra import plans are in the reference,
 * NSSolverStationary.cpp:226-305): the cells touching a dof of `rank` (own cells + one ghost layer), local dof
 * numbering per block (owned first, then ghosts ascending by global id), owned-row patterns with local
 * columns, ghost import plans, local boundary lists.  The result answers the same queries as a global
 * nsx_disc; NSX_DI_N_U / N_P then count owned + ghost dofs. */
nsx_disc *nsx_disc_local(const nsx_disc *global, int rank);
void nsx_disc_free(nsx_disc *d);
/* last error text of a failed nsx_disc_* call on this thread */
const char *nsx_host_last_error(void);

enum nsx_disc_info_t {
  NSX_DI_ELEM = 0,      /* 0 Q3/Q2, 1 P2/P1 */
  NSX_DI_NCELLS = 1,
  NSX_DI_NVERTS = 2,
  NSX_DI_N_U = 3,
  NSX_DI_N_P = 4,
  NSX_DI_DOFS_PER_CELL = 5,
  NSX_DI_NQ = 6,
  NSX_DI_NQF = 7,
  NSX_DI_NRANKS = 8,
  NSX_DI_NBC = 9,
  NSX_DI_NVPC = 10,
  NSX_DI_IS_LOCAL = 11,   /* 1 for the result of nsx_disc_local */
  NSX_DI_RANK = 12,
  NSX_DI_JOB_RANKS = 13,  /* ranks of the partition a local view was cut from */
  NSX_DI_N_U_OWNED = 14,
  NSX_DI_N_P_OWNED = 15
};
int64_t nsx_disc_info(const nsx_disc *d, int what);

enum nsx_disc_array_t {
  NSX_DA_CELL_DOFS = 0,      /* uint32 [ncells * dofs_per_cell] */
  NSX_DA_CELL_VERTICES = 1,  /* double [ncells * nvpc * 2] */
  NSX_DA_CELL_RANK = 2,      /* int32  [ncells] */
  NSX_DA_OWNED_U = 3,        /* int64  [nranks + 1] */
  NSX_DA_OWNED_P = 4,        /* int64  [nranks + 1] */
  NSX_DA_F_ROWPTR = 10, NSX_DA_F_COL = 11,    /* int64 / int32 */
  NSX_DA_BT_ROWPTR = 12, NSX_DA_BT_COL = 13,
  NSX_DA_B_ROWPTR = 14, NSX_DA_B_COL = 15,
  NSX_DA_MP_ROWPTR = 16, NSX_DA_MP_COL = 17,
  NSX_DA_BC_DOF = 20,        /* uint32 [nbc] ascending */
  NSX_DA_BC_SHAPE = 21,      /* double [nbc] unit-amplitude inlet profile */
  NSX_DA_BC_ON_INLET = 22,   /* uint8  [nbc] */
  NSX_DA_BC_Y = 23,          /* double [nbc] support point y */
  NSX_DA_OUTLET_CELL = 30, NSX_DA_OUTLET_FACE = 31,      /* int32 */
  NSX_DA_CYL_CELL = 32, NSX_DA_CYL_FACE = 33,            /* int32 */
  NSX_DA_BFACES = 34,        /* int32 [n * 3] (cell, face, boundary id) */
  NSX_DA_MATERIAL = 35,      /* int32 [ncells] */
  NSX_DA_FE_TABLES = 40,     /* the raw nsx::FETables struct (bytes) */
  /* local views only */
  NSX_DA_L2G_U = 50, NSX_DA_L2G_P = 51,      /* int64: local -> global id inside the block */
  NSX_DA_CELL_GLOBAL = 52,                   /* int32: local cell -> global cell */
  NSX_DA_CELL_OWNED = 53,                    /* uint8: the cell belongs to this rank's subdomain */
  NSX_DA_HALO_U_NBR = 60, NSX_DA_HALO_U_SEND_PTR = 61, NSX_DA_HALO_U_SEND_IDX = 62, NSX_DA_HALO_U_RECV_PTR = 63,  /* int32 / int64 / int32 / int64 */
  NSX_DA_HALO_P_NBR = 64, NSX_DA_HALO_P_SEND_PTR = 65, NSX_DA_HALO_P_SEND_IDX = 66, NSX_DA_HALO_P_RECV_PTR = 67
};
/* Returns a pointer into the discretisation (valid until nsx_disc_free) and the element count. */
const void *nsx_disc_array(const nsx_disc *d, int what, int64_t *count);

/* Dirichlet values of the inlet profile 4*u*y*(H-y)/(H*H), H = 0.41, for amplitude u, in the
 * order of NSX_DA_BC_DOF (reference: InletVelocity::value, NSSolverStationary.hpp:81-89). */
void nsx_disc_inlet_values(const nsx_disc *d, double amplitude, double *values);

#ifdef __cplusplus
}
#endif
#endif
