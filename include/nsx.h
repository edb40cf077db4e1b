/* nsx.h -- thin C ABI of the B200 (sm_100a) assembly + linear-algebra hot path.
 *
 * The reference (HliasGit/navier_stokes_solver) has no FFI: its hot path is reached through the
 * C++ members NSSolverStationary::assemble_system / solve_system (lab_new/src/
 * NSSolverStationary.cpp:317-577, 579-647) and NSSolver::assemble_system / solve_system
 * (lab_new/src/NSSolver.cpp:313-599, 601-672), plus the vector updates of the Newton / time
 * loops (NSSolverStationary.cpp:698, 715-729; NSSolver.cpp:707, 724-738, 820) and
 * compute_lift_drag (NSSolverStationary.cpp:802-897).  Each entry point below names the member
 * (file:line) whose device work it replaces.  Plain C types only: the caller owns host buffers,
 * the library owns device memory.  One context per GPU; calls on one context are not
 * thread-safe (as the reference objects).  Every call returns an nsx_status.
 */
#ifndef NSX_H
#define NSX_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsx_ctx nsx_ctx;

enum nsx_status {
  NSX_OK = 0,
  NSX_E_NOCONV = 1, /* <-> deal.II SolverControl::NoConvergence (never caught by the reference) */
  NSX_E_BADARG = 2, /* <-> std::invalid_argument (NSSolverStationary.cpp:641-643) */
  NSX_E_CUDA = 3,
  NSX_E_COMM = 4,
  NSX_E_STATE = 5   /* call order violated (e.g. assemble before the pattern is set) */
};

enum nsx_block { NSX_BLOCK_F = 0, NSX_BLOCK_BT = 1, NSX_BLOCK_B = 2, NSX_BLOCK_MP = 3, NSX_BLOCK_S = 4, NSX_BLOCK_J = 5,
                 NSX_BLOCK_F_DECOUPLED = 6 /* query id: the same-component view of F (nsx_get_ordering / nsx_get_sweep_blocks) */ };
enum nsx_vec { NSX_VEC_SOLUTION = 0, NSX_VEC_SOLUTION_OLD = 1, NSX_VEC_DELTA = 2, NSX_VEC_RESIDUAL = 3, NSX_VEC_EVAL = 4,
               NSX_VEC_TMP0 = 5, NSX_VEC_TMP1 = 6 };
/* assembly branches: STOKES = `global_first_iter || computing_stokes` (NSSolverStationary.cpp:383-406),
 * NEWTON (408-452), UNSTEADY_FIRST = `first_iter` (NSSolver.cpp:381-409), UNSTEADY_NEWTON (411-469) */
enum nsx_mode { NSX_MODE_STOKES = 0, NSX_MODE_NEWTON = 1, NSX_MODE_UNSTEADY_FIRST = 2, NSX_MODE_UNSTEADY_NEWTON = 3 };
enum nsx_flavour { NSX_STATIONARY = 0, NSX_UNSTEADY = 1 }; /* which header's preconditioner internals */
enum nsx_option {
  NSX_OPT_ORDERING = 0,   /* elimination order of ILU(0)/SGS: 0 natural (as Ifpack), 1 multicolour over the whole owned range,
                             2 (default) the owned range cut into spatially compact blocks (one CTA each, the structure of the
                             reference's overlap-0 Ifpack preconditioners under mpirun -n <#blocks>), multicolour inside a block;
                             3 the same blocks with Ifpack's natural (ascending index) order inside each: many more dependency
                             levels per block, the stronger ILU(0) (what a single ILU application per iteration needs) */
  NSX_OPT_VERBOSE = 1,
  NSX_OPT_ORTHO = 2,      /* Gram-Schmidt of GMRES/FGMRES: 0 modified chain (as deal.II), 1 batched classical, two passes (default),
                             2 as 1 for the outer solver, one pass + conditional second pass for the inner FGMRES solves */
  NSX_OPT_COOP_SWEEP = 3, /* ILU/SGS sweeps: 1 colour-phased persistent kernel with its own grid barrier (default, multicolour order), 2 level-phased cooperative launch, 0 one launch per level */
  NSX_OPT_STREAM_SPMV = 4, /* SpMV kernel: 3 TMA-fed persistent, rows reduced from the stage, paired velocity columns (default); 2 TMA-fed, products staged; 1 streaming with plain loads; 0 sub-warp per row */
  NSX_OPT_BLOCK_ROWS = 5, /* ordering 2: target rows per block (0 = automatic: rows / #SMs clamped to [512, 4096]) */
  NSX_OPT_DECOUPLE = 7,   /* Exact, cheaper views of F, chosen from its current VALUES (checked on the device after each assembly):
                             1: while every coupling between the two velocity components is an exact zero (the Stokes-type branches)
                             the inner solves' F products and the Gauss-Seidel sweeps run on the same-component entries only;
                             2 (default): when moreover F(u_x a, u_x b) == F(u_y a, u_y b) bit for bit, i.e. F = K (x) I_2, one scalar
                             matrix over the velocity nodes serves both components in the SpMV and in the SGS / ILU(0) sweeps
                             (orderings 2 and 3); 0: always the full pattern */
  NSX_OPT_L2_HINTS = 8,   /* 1 (default): the TMA copies of matrix values / columns carry an evict-first L2 policy, so that the streams
                             (read once per launch) do not push the Krylov basis out of the 126 MB L2; 0: no hint */
  NSX_OPT_PRECOND_LAG = 9, /* n > 0: the numeric preconditioner data (ILU(0) factors, Gauss-Seidel values, AMG hierarchy, Schur
                             complement) built in one solve also serve the next n solves -- Newton iterations / line-search states
                             change the matrices little, and FGMRES tolerates a lagged preconditioner; 0 (default): every solve
                             rebuilds them, as the initialize() calls of the reference do (NSSolverStationary.cpp:583,601,621) */
  NSX_OPT_SWEEP_Q = 10,   /* block-local sweeps: entries of a row handled by one lane (4 default, 8, 16): fewer lanes and shuffle rounds per
                             row against more padding in the matrix stream */
  NSX_OPT_HOST_INNER = 6  /* 1: the inner FGMRES solves run their recurrences on the host (one stream synchronisation per
                             iteration, round-1 behaviour); 0 (default): device-side Givens / convergence decision, the host
                             polls a mapped record and launches the next sweep speculatively */
};
enum nsx_stat {
  NSX_STAT_INNER_F = 0, NSX_STAT_INNER_S = 1, NSX_STAT_PRECOND_APPLIES = 2, NSX_STAT_KERNEL_LAUNCHES = 3,
  NSX_STAT_LEVELS_F = 4, NSX_STAT_LEVELS_MP = 5, NSX_STAT_LEVELS_S = 6, NSX_STAT_SPMV_CALLS = 7,
  NSX_STAT_ASSEMBLY_COLOURS = 8, NSX_STAT_ASSEMBLY_TABLES = 9, NSX_STAT_LAST_STEP = 10,
  NSX_STAT_HALO_EXCHANGES = 11, NSX_STAT_ALLREDUCES = 12,
  NSX_STAT_PRECOND_BUILDS = 16, /* numeric preconditioner builds since nsx_create (see NSX_OPT_PRECOND_LAG) */
  NSX_STAT_SWEEP_BYTES_F = 14, NSX_STAT_SPMV_BYTES_F = 15, /* stored bytes per launch of the F sweeps / the inner solves' F product */
  NSX_STAT_F_DECOUPLED = 13 /* view found by the last check of NSX_OPT_DECOUPLE: 0 full, 1 same-component, 2 nodes */
};

/* ctor of the solver objects (NSSolverStationary.hpp:339-351): one context per rank / GPU.
 * stream: a cudaStream_t to run on (NULL = a stream the library creates). */
int nsx_create(int rank, int nranks, int device_id, void *stream, nsx_ctx **out);
int nsx_destroy(nsx_ctx *ctx);
const char *nsx_last_error(const nsx_ctx *ctx);
int nsx_set_option(nsx_ctx *ctx, int option, int64_t value);
int64_t nsx_get_stat(const nsx_ctx *ctx, int stat);

/* Outputs of setup() that the hot path consumes (NSSolverStationary.cpp:114-314).
 * elem: 0 Q3/Q2 quads, 1 P2/P1 triangles.  cell_dofs: block-global ids in deal.II's cell-local
 * FESystem order (cell->get_dof_indices, NSSolverStationary.cpp:528).  cell_vertices: per cell
 * nvpc x (x, y). */
int nsx_set_discretisation(nsx_ctx *ctx, int elem, int64_t n_cells, const double *cell_vertices,
                           const uint32_t *cell_dofs, int64_t n_u, int64_t n_p);
/* Sparsity of one block (NSSolverStationary.cpp:264-305): F, BT, B of the Jacobian and MP. */
int nsx_set_pattern(nsx_ctx *ctx, int block, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int32_t *col);
/* Boundary faces: kind 8 = outlet Neumann term (NSSolverStationary.cpp:503-526), kind 10 = cylinder
 * (lift/drag, NSSolverStationary.cpp:840-843). */
int nsx_set_faces(nsx_ctx *ctx, int kind, int64_t n, const int32_t *cell, const int32_t *face);
/* Dirichlet velocity dofs (ascending) and the inlet values of the one non-homogeneous application
 * (interpolate_boundary_values, NSSolverStationary.cpp:541-572). */
int nsx_set_dirichlet(nsx_ctx *ctx, int64_t n, const uint32_t *dof, const double *inlet_value);
/* Owned ranges of the velocity / pressure blocks per rank (block_owned_dofs, NSSolverStationary.cpp:237-240):
 * the inner ILU / SGS / AMG are local to each range (Ifpack overlap 0). */
int nsx_set_ranks(nsx_ctx *ctx, int nranks, const int64_t *owned_u, const int64_t *owned_p);
/* ---- one rank of a row-partitioned run (mpirun -n N in the reference; one process per GPU here) ----
 * The reference partitions the cells, and every rank owns one contiguous range of each block
 * (locally_owned_dofs / block_owned_dofs, NSSolverStationary.cpp:226-242).  A rank hands this library its own
 * share in LOCAL numbering: per block the owned dofs first, then the ghosts (locally relevant, not owned)
 * grouped by owner.  nsx_set_discretisation then takes the local cells (own cells plus the ghost layer whose
 * contributions reach owned rows), cell_dofs in local ids (pressure ids offset by the local velocity count)
 * and n_u / n_p = owned + ghost counts; nsx_set_partition says how many of them are owned; nsx_set_pattern
 * takes the OWNED rows of each block with local column ids; nsx_set_dirichlet / nsx_set_faces take local ids;
 * block vectors passed to nsx_vec_upload / download hold the owned entries [velocity | pressure] only.
 * Call order: create, set_discretisation, set_partition, set_halo x2, set_pattern x4, faces, dirichlet,
 * comm_init, finalize_setup.  assemble / solve / lift_drag / halo_exchange are collective. */
int nsx_set_partition(nsx_ctx *ctx, int64_t n_u_owned, int64_t n_p_owned);
/* Ghost import plan of one block (0 velocity, 1 pressure) -- what Epetra_Import holds for the reference's
 * ghosted vectors (NSSolverStationary.cpp:309-314): to neighbour i go the owned entries
 * send_idx[send_ptr[i] .. send_ptr[i+1]); from it come the ghosts recv_ptr[i] .. recv_ptr[i+1) (0 = first ghost),
 * in the same order on both sides. */
int nsx_set_halo(nsx_ctx *ctx, int block, int n_neighbours, const int32_t *neighbour, const int64_t *send_ptr, const int32_t *send_idx,
                 const int64_t *recv_ptr);
/* NCCL bootstrap: rank 0 creates the 128-byte id, the launcher broadcasts it (MPI_Bcast, torch.distributed, a file),
 * every rank then joins.  No-op for nranks == 1. */
int nsx_comm_unique_id(void *id128);
int nsx_comm_init(nsx_ctx *ctx, const void *id128);

/* Builds the device-side gather maps; must follow the setters and precede assemble/solve. */
int nsx_finalize_setup(nsx_ctx *ctx);

/* Block vectors [velocity | pressure] of length n_u + n_p (NSSolverStationary.hpp:451-463). */
int nsx_vec_upload(nsx_ctx *ctx, int which, const double *host);
int nsx_vec_download(nsx_ctx *ctx, int which, double *host);
/* ghost import of one block vector, and a look at the imported values (tests) */
int nsx_halo_exchange(nsx_ctx *ctx, int which);
int nsx_vec_download_ghosts(nsx_ctx *ctx, int which, double *host_u_ghosts, double *host_p_ghosts);
/* device-side `v = value` and `dst = src` (Trilinos `vector = 0.0`, `a = b`; NSSolverStationary.cpp:338-340, 715) */
int nsx_vec_set(nsx_ctx *ctx, int which, double value);
int nsx_vec_copy(nsx_ctx *ctx, int dst, int src);

/* assemble_system (NSSolverStationary.cpp:317-577, NSSolver.cpp:313-599): J, Mp, r from the current
 * solution (and solution_old), then the Dirichlet rows; returns ||r||_2 (the l2_norm() the Newton
 * loop takes next, NSSolverStationary.cpp:698). */
int nsx_assemble(nsx_ctx *ctx, int mode, int apply_inlet, double nu, double dt, double p_out, double *residual_l2);
/* The residual of assemble_system alone, for the line search (NSSolverStationary.cpp:718-735, NSSolver.cpp:727-740): the reference
 * re-assembles J, Mp and r for every trial step length and only takes ||r||; its Newton loop assembles again before the next
 * solve, so the matrices of the trial assemblies are never read.  This call leaves in `residual` (and in the Dirichlet entries
 * of `delta`) exactly the bits nsx_assemble(mode, apply_inlet = 0, ...) would, and does not touch the matrices, which therefore
 * still belong to the previous state until the next nsx_assemble. */
int nsx_assemble_residual(nsx_ctx *ctx, int mode, double nu, double dt, double p_out, double *residual_l2);
/* solve_system (NSSolverStationary.cpp:579-647, NSSolver.cpp:601-672): solver 0 GMRES / 1 FGMRES /
 * 2 BiCGStab; prec 0 blockDiagonal / 1 blockTriangular / 2 aSIMPLE; delta is the warm start. */
int nsx_solve(nsx_ctx *ctx, int flavour, int solver, int prec, double tol, int max_it, double alpha,
              int *iterations, double *final_residual);
/* Newton / time-loop vector updates:
 *   save_eval_point: evaluation_point = solution                    (NSSolverStationary.cpp:715)
 *   update:          solution = evaluation_point + alpha * delta    (NSSolverStationary.cpp:720-722)
 *   copy_old:        solution_old = solution                        (NSSolver.cpp:820) */
int nsx_save_eval_point(nsx_ctx *ctx);
int nsx_update(nsx_ctx *ctx, double alpha);
int nsx_copy_old(nsx_ctx *ctx);
/* compute_lift_drag (NSSolverStationary.cpp:802-897): forces on boundary id 10. */
int nsx_lift_drag(nsx_ctx *ctx, double nu, double *drag_force, double *lift_force);

/* ---- test / measurement hooks (no reference counterpart) ---- */
/* the cell loop + compress of assemble_system without the Dirichlet step (structural identities) */
int nsx_assemble_cells(nsx_ctx *ctx, int mode, double nu, double dt, double p_out);
int nsx_get_block_nnz(nsx_ctx *ctx, int block, int64_t *nnz);
int nsx_get_block_pattern(nsx_ctx *ctx, int block, int64_t *rowptr, int32_t *col);
int nsx_get_block_values(nsx_ctx *ctx, int block, double *values);
int nsx_set_block_values(nsx_ctx *ctx, int block, const double *values);
/* y = A x on device vectors; block NSX_BLOCK_J = the whole Jacobian on block vectors */
int nsx_spmv(nsx_ctx *ctx, int block, int vec_x, int vec_y);
/* y = M^-1 x with one inner preconditioner on one block: kind 0 SGS, 1 ILU(0), 2 AMG */
int nsx_inner_apply(nsx_ctx *ctx, int block, int kind, int vec_x, int vec_y);
int nsx_ilu0_factor(nsx_ctx *ctx, int block, double *lu_values, int32_t *perm);
int nsx_schur(nsx_ctx *ctx);
/* dst = P^-1 src for the block preconditioner built from the current matrices */
int nsx_precond_apply(nsx_ctx *ctx, int flavour, int prec, double alpha, int vec_src, int vec_dst);
/* times `reps` launches of one kernel with CUDA events on the context's stream; ms per launch.
 * what: 0 Jacobian block SpMV, 1 F SpMV, 2 assembly (matrix + rhs kernels), 3 dot, 4 axpy, 5 SGS(F) apply,
 * 6 ILU(F) apply, 7 ILU(F) factorisation, 20 an FP64 FMA peak kernel (num_SMs x 8 x 256 threads x 8192 x 8 FMAs).
 * nsx_set_time_params picks the assembly branch that `what` = 2 times. */
int nsx_set_time_params(nsx_ctx *ctx, int mode, double nu, double dt);
/* flush_l2: 1 = L2 flushed before every launch, one event pair per launch; 0 = no flush; 2 = `reps` launches back to back inside one
 * event pair (for kernels whose input is larger than L2: no launch / event overhead in the figure). */
int nsx_time_kernel(nsx_ctx *ctx, int what, int reps, int flush_l2, double *ms_per_launch);
int nsx_synchronize(nsx_ctx *ctx);
/* elimination order used for a block's ILU/SGS (new -> old), for oracle parity */
int nsx_get_ordering(nsx_ctx *ctx, int block, int32_t *perm);
/* ordering 2: the blocks of the block-local sweeps -- block b eliminates perm[offsets[b] .. offsets[b+1]) in that order and
 * drops its couplings to other blocks; n_blocks = 0 for orderings 0 / 1.  offsets (n_blocks + 1 entries) may be NULL. */
int nsx_get_sweep_blocks(nsx_ctx *ctx, int block, int32_t *n_blocks, int64_t *offsets);
/* runs the check of NSX_OPT_DECOUPLE on the current values of F: *yes = the view the next solve will use (0 full matrix, 1 same-component
 * entries, 2 velocity nodes); NSX_BLOCK_F_DECOUPLED queries that view's sweep blocks / elimination order in dof numbering */
int nsx_check_decoupled(nsx_ctx *ctx, int *yes);

#ifdef __cplusplus
}
#endif
#endif
