"""ctypes binding of the product library libnsx.so: the host set-up stand-in (include/nsx_host.h)
and the CUDA hot path behind the C ABI of include/nsx.h.  Nothing here touches the CPU oracle."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBNSX = os.environ.get("NSX_LIB") or os.path.join(ROOT, "navier_stokes_solver_b200", "libnsx.so")
LIBNSX_HOST = os.environ.get("NSX_HOST_LIB") or os.path.join(ROOT, "navier_stokes_solver_b200", "libnsx_host.so")

c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)

# enum mirrors (include/nsx_host.h, include/nsx.h)
DI = dict(ELEM=0, NCELLS=1, NVERTS=2, N_U=3, N_P=4, DOFS_PER_CELL=5, NQ=6, NQF=7, NRANKS=8, NBC=9, NVPC=10,
          IS_LOCAL=11, RANK=12, JOB_RANKS=13, N_U_OWNED=14, N_P_OWNED=15)
DA = dict(CELL_DOFS=(0, np.uint32), CELL_VERTICES=(1, np.float64), CELL_RANK=(2, np.int32),
          OWNED_U=(3, np.int64), OWNED_P=(4, np.int64),
          F_ROWPTR=(10, np.int64), F_COL=(11, np.int32), BT_ROWPTR=(12, np.int64), BT_COL=(13, np.int32),
          B_ROWPTR=(14, np.int64), B_COL=(15, np.int32), MP_ROWPTR=(16, np.int64), MP_COL=(17, np.int32),
          BC_DOF=(20, np.uint32), BC_SHAPE=(21, np.float64), BC_ON_INLET=(22, np.uint8), BC_Y=(23, np.float64),
          OUTLET_CELL=(30, np.int32), OUTLET_FACE=(31, np.int32), CYL_CELL=(32, np.int32), CYL_FACE=(33, np.int32),
          BFACES=(34, np.int32), MATERIAL=(35, np.int32), FE_TABLES=(40, np.uint8),
          L2G_U=(50, np.int64), L2G_P=(51, np.int64), CELL_GLOBAL=(52, np.int32), CELL_OWNED=(53, np.uint8),
          HALO_U_NBR=(60, np.int32), HALO_U_SEND_PTR=(61, np.int64), HALO_U_SEND_IDX=(62, np.int32), HALO_U_RECV_PTR=(63, np.int64),
          HALO_P_NBR=(64, np.int32), HALO_P_SEND_PTR=(65, np.int64), HALO_P_SEND_IDX=(66, np.int32), HALO_P_RECV_PTR=(67, np.int64))
BLOCK_F, BLOCK_BT, BLOCK_B, BLOCK_MP, BLOCK_S, BLOCK_J, BLOCK_F_DECOUPLED = 0, 1, 2, 3, 4, 5, 6
MODE_STOKES, MODE_NEWTON, MODE_UNSTEADY_FIRST, MODE_UNSTEADY_NEWTON = 0, 1, 2, 3
VEC_SOLUTION, VEC_SOLUTION_OLD, VEC_DELTA, VEC_RESIDUAL, VEC_EVAL, VEC_TMP0, VEC_TMP1 = 0, 1, 2, 3, 4, 5, 6
STATIONARY, UNSTEADY = 0, 1
NSX_OK, NSX_E_NOCONV, NSX_E_BADARG, NSX_E_CUDA, NSX_E_COMM, NSX_E_STATE = 0, 1, 2, 3, 4, 5
OPT_ORDERING, OPT_VERBOSE, OPT_ORTHO, OPT_COOP_SWEEP, OPT_STREAM_SPMV, OPT_BLOCK_ROWS, OPT_HOST_INNER, OPT_DECOUPLE, OPT_L2_HINTS, OPT_PRECOND_LAG, OPT_SWEEP_Q = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
STAT = dict(INNER_F=0, INNER_S=1, PRECOND_APPLIES=2, KERNEL_LAUNCHES=3, LEVELS_F=4, LEVELS_MP=5, LEVELS_S=6,
            SPMV_CALLS=7, ASSEMBLY_COLOURS=8, ASSEMBLY_TABLES=9, LAST_STEP=10, HALO_EXCHANGES=11, ALLREDUCES=12, F_DECOUPLED=13, SWEEP_BYTES_F=14, SPMV_BYTES_F=15, PRECOND_BUILDS=16)
# every entry point include/nsx.h declares (tests check that the library exports each one)
NSX_SYMBOLS = ["nsx_create", "nsx_destroy", "nsx_last_error", "nsx_set_option", "nsx_get_stat", "nsx_set_discretisation",
               "nsx_set_pattern", "nsx_set_faces", "nsx_set_dirichlet", "nsx_set_ranks", "nsx_finalize_setup",
               "nsx_vec_upload", "nsx_vec_download", "nsx_vec_set", "nsx_vec_copy", "nsx_assemble", "nsx_assemble_residual", "nsx_solve", "nsx_save_eval_point", "nsx_update",
               "nsx_copy_old", "nsx_lift_drag", "nsx_assemble_cells", "nsx_get_block_nnz", "nsx_get_block_pattern",
               "nsx_get_block_values", "nsx_set_block_values", "nsx_spmv", "nsx_inner_apply", "nsx_ilu0_factor",
               "nsx_schur", "nsx_precond_apply", "nsx_set_time_params", "nsx_time_kernel", "nsx_synchronize",
               "nsx_get_ordering", "nsx_set_partition", "nsx_set_halo", "nsx_comm_unique_id", "nsx_comm_init",
               "nsx_halo_exchange", "nsx_vec_download_ghosts", "nsx_get_sweep_blocks", "nsx_check_decoupled"]
NSX_HOST_SYMBOLS = ["nsx_disc_generate", "nsx_disc_from_gmsh", "nsx_disc_local", "nsx_disc_free", "nsx_host_last_error", "nsx_disc_info",
                    "nsx_disc_array", "nsx_disc_inlet_values"]

_nsx = None
_nsx_host = None


def nsx_host():
    """The host set-up stand-in (include/nsx_host.h): plain C++, no CUDA, never loads the device library."""
    global _nsx_host
    if _nsx_host is None:
        if not os.path.exists(LIBNSX_HOST):
            raise RuntimeError(f"{LIBNSX_HOST} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIBNSX_HOST)
        L.nsx_disc_generate.restype = C.c_void_p
        L.nsx_disc_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.nsx_disc_from_gmsh.restype = C.c_void_p
        L.nsx_disc_from_gmsh.argtypes = [C.c_char_p, C.c_int]
        L.nsx_disc_free.argtypes = [C.c_void_p]
        L.nsx_disc_local.restype = C.c_void_p
        L.nsx_disc_local.argtypes = [C.c_void_p, C.c_int]
        L.nsx_disc_info.restype = C.c_int64
        L.nsx_disc_info.argtypes = [C.c_void_p, C.c_int]
        L.nsx_disc_array.restype = C.c_void_p
        L.nsx_disc_array.argtypes = [C.c_void_p, C.c_int, c_i64p]
        L.nsx_disc_inlet_values.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.nsx_host_last_error.restype = C.c_char_p
        _nsx_host = L
    return _nsx_host


def nsx():
    """The product library (include/nsx.h).  Loading works without a GPU; compute calls need one."""
    global _nsx
    if _nsx is None:
        if not os.path.exists(LIBNSX):
            raise RuntimeError(f"{LIBNSX} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIBNSX)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.nsx_create.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
        L.nsx_destroy.argtypes = [vp]
        L.nsx_last_error.restype = C.c_char_p
        L.nsx_last_error.argtypes = [vp]
        L.nsx_set_option.argtypes = [vp, i32, i64]
        L.nsx_get_stat.restype = i64
        L.nsx_get_stat.argtypes = [vp, i32]
        L.nsx_set_discretisation.argtypes = [vp, i32, i64, vp, vp, i64, i64]
        L.nsx_set_pattern.argtypes = [vp, i32, i64, i64, vp, vp]
        L.nsx_set_faces.argtypes = [vp, i32, i64, vp, vp]
        L.nsx_set_dirichlet.argtypes = [vp, i64, vp, vp]
        L.nsx_set_ranks.argtypes = [vp, i32, vp, vp]
        L.nsx_finalize_setup.argtypes = [vp]
        L.nsx_vec_upload.argtypes = [vp, i32, vp]
        L.nsx_vec_download.argtypes = [vp, i32, vp]
        L.nsx_vec_set.argtypes = [vp, i32, dbl]
        L.nsx_vec_copy.argtypes = [vp, i32, i32]
        L.nsx_assemble.argtypes = [vp, i32, i32, dbl, dbl, dbl, c_dp]
        L.nsx_assemble_cells.argtypes = [vp, i32, dbl, dbl, dbl]
        L.nsx_assemble_residual.argtypes = [vp, i32, dbl, dbl, dbl, c_dp]
        L.nsx_solve.argtypes = [vp, i32, i32, i32, dbl, i32, dbl, C.POINTER(i32), c_dp]
        L.nsx_save_eval_point.argtypes = [vp]
        L.nsx_update.argtypes = [vp, dbl]
        L.nsx_copy_old.argtypes = [vp]
        L.nsx_lift_drag.argtypes = [vp, dbl, c_dp, c_dp]
        L.nsx_get_block_nnz.argtypes = [vp, i32, c_i64p]
        L.nsx_get_block_pattern.argtypes = [vp, i32, vp, vp]
        L.nsx_get_block_values.argtypes = [vp, i32, vp]
        L.nsx_set_block_values.argtypes = [vp, i32, vp]
        L.nsx_spmv.argtypes = [vp, i32, i32, i32]
        L.nsx_inner_apply.argtypes = [vp, i32, i32, i32, i32]
        L.nsx_ilu0_factor.argtypes = [vp, i32, vp, vp]
        L.nsx_schur.argtypes = [vp]
        L.nsx_precond_apply.argtypes = [vp, i32, i32, dbl, i32, i32]
        L.nsx_set_time_params.argtypes = [vp, i32, dbl, dbl]
        L.nsx_time_kernel.argtypes = [vp, i32, i32, i32, c_dp]
        L.nsx_synchronize.argtypes = [vp]
        L.nsx_get_ordering.argtypes = [vp, i32, vp]
        L.nsx_get_sweep_blocks.argtypes = [vp, i32, C.POINTER(i32), vp]
        L.nsx_check_decoupled.argtypes = [vp, C.POINTER(i32)]
        L.nsx_set_partition.argtypes = [vp, i64, i64]
        L.nsx_set_halo.argtypes = [vp, i32, i32, vp, vp, vp, vp]
        L.nsx_comm_unique_id.argtypes = [vp]
        L.nsx_comm_init.argtypes = [vp, vp]
        L.nsx_halo_exchange.argtypes = [vp, i32]
        L.nsx_vec_download_ghosts.argtypes = [vp, i32, vp, vp]
        _nsx = L
    return _nsx


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Disc:
    """Host-side discretisation (stand-in for deal.II's set-up) through include/nsx_host.h."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(nsx_host().nsx_host_last_error().decode())
        self.h = handle
        L = nsx_host()
        for k, v in DI.items():
            setattr(self, k.lower(), int(L.nsx_disc_info(self.h, v)))
        self.n = self.n_u + self.n_p
        self.n_owned = self.n_u_owned + self.n_p_owned

    def local(self, rank):
        """This rank's share (own cells + ghost layer, local numbering, owned-row patterns, halo plans)."""
        return Disc(nsx_host().nsx_disc_local(self.h, rank))

    def owned_global_ids(self):
        """Global block-vector positions [velocity | pressure] of a local view's owned entries (n_u_global needed
        for the pressure offset is passed by the caller through gather/scatter helpers below)."""
        return self.array("L2G_U")[: self.n_u_owned], self.array("L2G_P")[: self.n_p_owned]

    def scatter_owned(self, global_vec, n_u_global):
        """Owned entries [u | p] of this rank taken from a global block vector."""
        gu, gp = self.owned_global_ids()
        return np.concatenate([global_vec[gu], global_vec[n_u_global + gp]])

    def gather_owned(self, local_vec, global_vec, n_u_global):
        """Writes this rank's owned entries into a global block vector."""
        gu, gp = self.owned_global_ids()
        global_vec[gu] = local_vec[: self.n_u_owned]
        global_vec[n_u_global + gp] = local_vec[self.n_u_owned:]

    @classmethod
    def generate(cls, nx, ny, triangles=False, nranks=1):
        return cls(nsx_host().nsx_disc_generate(nx, ny, int(triangles), nranks))

    @classmethod
    def from_gmsh(cls, path, nranks=1):
        return cls(nsx_host().nsx_disc_from_gmsh(path.encode(), nranks))

    def array(self, name):
        code, dt = DA[name]
        cnt = C.c_int64()
        p = nsx_host().nsx_disc_array(self.h, code, C.byref(cnt))
        if cnt.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (cnt.value * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt)

    def inlet_values(self, amplitude):
        v = np.zeros(self.nbc)
        nsx_host().nsx_disc_inlet_values(self.h, amplitude, ptr(v))
        return v

    def pattern(self, name):
        return self.array(name + "_ROWPTR"), self.array(name + "_COL")

    def __del__(self):
        try:
            nsx_host().nsx_disc_free(self.h)
        except Exception:
            pass


def synthetic_state(disc, seed=1234, noise=1e-2):
    """The synthetic Newton state of SURVEY.md section 8(d): parabolic u_x, small noise, linear p.
    Evaluated per dof from the support points implied by the cell table (vertex/GLL positions are
    not needed to the last digit for a synthetic state: the noise dominates)."""
    rng = np.random.default_rng(seed)
    n_u, n_p = disc.n_u, disc.n_p
    cd = disc.array("CELL_DOFS").reshape(disc.ncells, disc.dofs_per_cell)
    cv = disc.array("CELL_VERTICES").reshape(disc.ncells, disc.nvpc, 2)
    # approximate support point: cell centre (enough for a smooth synthetic field)
    cen = cv.mean(axis=1)
    x = np.zeros(disc.n)
    y = np.zeros(disc.n)
    for i in range(disc.dofs_per_cell):
        x[cd[:, i]] = cen[:, 0]
        y[cd[:, i]] = cen[:, 1]
    sol = np.zeros(disc.n)
    ux = 4 * 0.1 * y[:n_u] * (0.41 - y[:n_u]) / 0.41 ** 2
    even = (np.arange(n_u) % 2) == 0  # not exact component split; irrelevant for a synthetic state
    sol[:n_u] = np.where(even, ux, 0.0) + rng.uniform(-noise, noise, n_u)
    sol[n_u:] = 1 + (2.2 - x[n_u:]) * 0.05 + rng.uniform(-noise, noise, n_p)
    return sol


class NsxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nsx error {code}: {msg}")
        self.code = code


class Device:
    """One GPU context of the hot path (include/nsx.h) filled from a Disc.  Every method is a thin call
    through the C ABI; there is no CPU fallback (construction fails without a CUDA device)."""

    def __init__(self, disc, device_id=0, inlet_amplitude=0.1, ordering=None, stream=None, ortho=None, comm_id=None, block_rows=None):
        """disc: a global Disc (one GPU) or a local view (Disc.local(rank)) of a partitioned run; for the
        latter comm_id is the 128-byte NCCL id shared by all ranks (Device.new_comm_id() on rank 0)."""
        L = nsx()
        self.disc = disc
        local = bool(disc.is_local)
        # sizes of the block vectors this context exchanges with the host: the owned entries
        self.n_u, self.n_p = (disc.n_u_owned, disc.n_p_owned) if local else (disc.n_u, disc.n_p)
        self.n = self.n_u + self.n_p
        h = C.c_void_p()
        rc = L.nsx_create(disc.rank if local else 0, disc.job_ranks if local else 1, device_id, stream, C.byref(h))
        if rc:
            raise NsxError(rc, "nsx_create failed (no CUDA device? the product has no CPU path)")
        self.h = h
        if ordering is not None:
            self._ck(L.nsx_set_option(self.h, OPT_ORDERING, ordering))
        if ortho is not None:
            self._ck(L.nsx_set_option(self.h, OPT_ORTHO, ortho))
        if block_rows is not None:
            self._ck(L.nsx_set_option(self.h, OPT_BLOCK_ROWS, block_rows))
        cd = np.ascontiguousarray(disc.array("CELL_DOFS"))
        cv = np.ascontiguousarray(disc.array("CELL_VERTICES"))
        self._ck(L.nsx_set_discretisation(self.h, disc.elem, disc.ncells, ptr(cv), ptr(cd), disc.n_u, disc.n_p))
        if local:
            self._ck(L.nsx_set_partition(self.h, disc.n_u_owned, disc.n_p_owned))
            for blk, nm in ((0, "U"), (1, "P")):
                nbr, sp, si, rp = (np.ascontiguousarray(disc.array(f"HALO_{nm}_{k}")) for k in ("NBR", "SEND_PTR", "SEND_IDX", "RECV_PTR"))
                self._ck(L.nsx_set_halo(self.h, blk, len(nbr), ptr(nbr), ptr(sp), ptr(si), ptr(rp)))
        shapes = {BLOCK_F: (self.n_u, disc.n_u), BLOCK_BT: (self.n_u, disc.n_p),
                  BLOCK_B: (self.n_p, disc.n_u), BLOCK_MP: (self.n_p, disc.n_p)}
        self.shapes = shapes
        for blk, name in ((BLOCK_F, "F"), (BLOCK_BT, "BT"), (BLOCK_B, "B"), (BLOCK_MP, "MP")):
            rp, col = disc.pattern(name)
            self._ck(L.nsx_set_pattern(self.h, blk, shapes[blk][0], shapes[blk][1], ptr(rp), ptr(col)))
        oc, of = disc.array("OUTLET_CELL"), disc.array("OUTLET_FACE")
        self._ck(L.nsx_set_faces(self.h, 8, len(oc), ptr(oc), ptr(of)))
        cc, cf = disc.array("CYL_CELL"), disc.array("CYL_FACE")
        self._ck(L.nsx_set_faces(self.h, 10, len(cc), ptr(cc), ptr(cf)))
        bc_dof = disc.array("BC_DOF")
        bc_val = disc.inlet_values(inlet_amplitude)
        self._ck(L.nsx_set_dirichlet(self.h, len(bc_dof), ptr(bc_dof), ptr(bc_val)))
        if local:
            if disc.job_ranks > 1:
                if comm_id is None:
                    raise ValueError("a partitioned run needs the shared NCCL id (comm_id)")
                buf = (C.c_char * 128).from_buffer_copy(bytes(comm_id))
                self._ck(L.nsx_comm_init(self.h, buf))
        else:
            ou, op = disc.array("OWNED_U"), disc.array("OWNED_P")
            self._ck(L.nsx_set_ranks(self.h, disc.nranks, ptr(ou), ptr(op)))
        self._ck(L.nsx_finalize_setup(self.h))

    @staticmethod
    def new_comm_id():
        buf = (C.c_char * 128)()
        rc = nsx().nsx_comm_unique_id(buf)
        if rc:
            raise NsxError(rc, "ncclGetUniqueId failed (is NCCL loadable?)")
        return bytes(buf)

    def halo_exchange(self, which):
        self._ck(nsx().nsx_halo_exchange(self.h, which))

    def download_ghosts(self, which):
        gu = np.zeros(self.disc.n_u - self.n_u)
        gp = np.zeros(self.disc.n_p - self.n_p)
        self._ck(nsx().nsx_vec_download_ghosts(self.h, which, ptr(gu), ptr(gp)))
        return gu, gp

    def _ck(self, rc):
        if rc:
            raise NsxError(rc, nsx().nsx_last_error(self.h).decode())

    def last_error(self):
        return nsx().nsx_last_error(self.h).decode()

    def set_option(self, opt, value):
        self._ck(nsx().nsx_set_option(self.h, opt, value))

    def stat(self, name):
        return int(nsx().nsx_get_stat(self.h, STAT[name]))

    def upload(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        self._ck(nsx().nsx_vec_upload(self.h, which, ptr(x)))

    def download(self, which):
        x = np.empty(self.n)
        self._ck(nsx().nsx_vec_download(self.h, which, ptr(x)))
        return x

    def upload_ptr(self, which, address):
        self._ck(nsx().nsx_vec_upload(self.h, which, C.c_void_p(address)))

    def download_ptr(self, which, address):
        self._ck(nsx().nsx_vec_download(self.h, which, C.c_void_p(address)))

    def vec_set(self, which, value):
        self._ck(nsx().nsx_vec_set(self.h, which, value))

    def vec_copy(self, dst, src):
        self._ck(nsx().nsx_vec_copy(self.h, dst, src))

    def assemble(self, mode, apply_inlet, nu, dt=0.01, p_out=1.0):
        r = C.c_double()
        self._ck(nsx().nsx_assemble(self.h, mode, int(apply_inlet), nu, dt, p_out, C.byref(r)))
        return r.value

    def assemble_cells(self, mode, nu, dt=0.01, p_out=1.0):
        self._ck(nsx().nsx_assemble_cells(self.h, mode, nu, dt, p_out))

    def assemble_residual(self, mode, nu, dt=0.01, p_out=1.0):
        """||r|| of a full assembly with homogeneous Dirichlet values, matrices untouched (line-search trials)"""
        r = C.c_double()
        self._ck(nsx().nsx_assemble_residual(self.h, mode, nu, dt, p_out, C.byref(r)))
        return r.value

    def solve(self, flavour, solver, prec, tol, max_it=20000, alpha=0.5):
        it, fr = C.c_int(), C.c_double()
        rc = nsx().nsx_solve(self.h, flavour, solver, prec, tol, max_it, alpha, C.byref(it), C.byref(fr))
        if rc not in (NSX_OK, NSX_E_NOCONV, NSX_E_BADARG):
            self._ck(rc)
        return rc, it.value, fr.value

    def save_eval_point(self):
        self._ck(nsx().nsx_save_eval_point(self.h))

    def update(self, alpha):
        self._ck(nsx().nsx_update(self.h, alpha))

    def copy_old(self):
        self._ck(nsx().nsx_copy_old(self.h))

    def lift_drag(self, nu):
        d, l = C.c_double(), C.c_double()
        self._ck(nsx().nsx_lift_drag(self.h, nu, C.byref(d), C.byref(l)))
        return d.value, l.value

    def nnz(self, block):
        n = C.c_int64()
        self._ck(nsx().nsx_get_block_nnz(self.h, block, C.byref(n)))
        return n.value

    def values(self, block):
        v = np.empty(self.nnz(block))
        self._ck(nsx().nsx_get_block_values(self.h, block, ptr(v)))
        return v

    def set_values(self, block, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nnz(block)
        self._ck(nsx().nsx_set_block_values(self.h, block, ptr(v)))

    def pattern(self, block):
        nrows = self.n_u if block in (BLOCK_F, BLOCK_BT) else self.n_p
        rp = np.empty(nrows + 1, dtype=np.int64)
        col = np.empty(self.nnz(block), dtype=np.int32)
        self._ck(nsx().nsx_get_block_pattern(self.h, block, ptr(rp), ptr(col)))
        return rp, col

    def csr(self, block):
        import scipy.sparse as sp
        rp, col = self.pattern(block)
        ncols = self.disc.n_u if block in (BLOCK_F, BLOCK_B) else self.disc.n_p   # owned + ghost columns on a partitioned system
        if block == BLOCK_S:
            ncols = self.n_p
        return sp.csr_matrix((self.values(block), col, rp), shape=(len(rp) - 1, ncols))

    def _block_vec(self, block, x, cols=True):
        """pad a block-sized array into a full [u | p] vector at the offset the C ABI expects"""
        full = np.zeros(self.n)
        full[: x.size] = x
        return full

    def spmv(self, block, x):
        """y = A x through the device kernel; x sized to the block's columns (the whole vector for BLOCK_J)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.upload(VEC_TMP0, self._block_vec(block, x))
        self._ck(nsx().nsx_spmv(self.h, block, VEC_TMP0, VEC_TMP1))
        nrows = self.n if block == BLOCK_J else (self.n_u if block in (BLOCK_F, BLOCK_BT) else self.n_p)
        return self.download(VEC_TMP1)[:nrows]

    def inner_apply(self, block, kind, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.upload(VEC_TMP0, self._block_vec(block, x))
        self._ck(nsx().nsx_inner_apply(self.h, block, kind, VEC_TMP0, VEC_TMP1))
        return self.download(VEC_TMP1)[: x.size]

    def ilu0_factor(self, block):
        lu = np.empty(self.nnz(block))
        nrows = self.n_u if block == BLOCK_F else self.n_p
        perm = np.empty(nrows, dtype=np.int32)
        self._ck(nsx().nsx_ilu0_factor(self.h, block, ptr(lu), ptr(perm)))
        return lu, perm

    def ordering(self, block):
        nrows = self.n_u if block in (BLOCK_F, BLOCK_F_DECOUPLED) else self.n_p
        perm = np.empty(nrows, dtype=np.int32)
        self._ck(nsx().nsx_get_ordering(self.h, block, ptr(perm)))
        return perm

    def view(self):
        """The exact view of F the next solve uses on the current values (NSX_OPT_DECOUPLE): 0 the full matrix, 1 same-component
        entries only (cross-component couplings are exact zeros), 2 one scalar matrix over the velocity nodes (F = K (x) I_2)."""
        yes = C.c_int32()
        self._ck(nsx().nsx_check_decoupled(self.h, C.byref(yes)))
        return int(yes.value)

    def decoupled(self):
        return self.view() > 0

    def sweep_plan_id(self, block, ilu=False):
        """the plan id a Gauss-Seidel (or, ilu=True, an ILU(0)) sweep on `block` uses right now: ILU(0) only follows the node view"""
        if block != BLOCK_F:
            return block
        v = self.view()
        return BLOCK_F_DECOUPLED if (v == 2 or (v == 1 and not ilu)) else block

    def sgs_block_id(self, block):
        return self.sweep_plan_id(block, ilu=False)

    def sweep_blocks(self, block):
        """(offsets, perm) of the block-local sweeps (ordering 2): block b eliminates perm[offsets[b]:offsets[b+1]] in that order;
        None for the other orderings."""
        nb = C.c_int32()
        self._ck(nsx().nsx_get_sweep_blocks(self.h, block, C.byref(nb), None))
        if nb.value == 0:
            return None
        off = np.empty(nb.value + 1, dtype=np.int64)
        self._ck(nsx().nsx_get_sweep_blocks(self.h, block, C.byref(nb), ptr(off)))
        return off, self.ordering(block)

    def schur(self):
        self._ck(nsx().nsx_schur(self.h))
        return self.csr(BLOCK_S)

    def precond_apply(self, flavour, prec, src, dst0=None, alpha=0.5):
        self.upload(VEC_TMP0, src)
        self.upload(VEC_TMP1, np.zeros(self.n) if dst0 is None else dst0)
        self._ck(nsx().nsx_precond_apply(self.h, flavour, prec, alpha, VEC_TMP0, VEC_TMP1))
        return self.download(VEC_TMP1)

    def time_kernel(self, what, reps=20, flush_l2=True):
        ms = C.c_double()
        self._ck(nsx().nsx_time_kernel(self.h, what, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def set_time_params(self, mode, nu, dt=0.01):
        self._ck(nsx().nsx_set_time_params(self.h, mode, nu, dt))

    def synchronize(self):
        self._ck(nsx().nsx_synchronize(self.h))

    def close(self):
        if getattr(self, "h", None):
            nsx().nsx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
