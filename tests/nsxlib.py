"""ctypes helpers shared by the tests: the product library (libnsx.so: host set-up + CUDA path
behind the C ABI of include/nsx.h, include/nsx_host.h) and the CPU oracle (oracle/liboracle.so,
test infrastructure only)."""
import ctypes as C
import gzip
import os
import shutil
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBNSX = os.path.join(ROOT, "navier_stokes_solver_b200", "libnsx.so")
LIBORC = os.path.join(ROOT, "oracle", "liboracle.so")

c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)

# enum mirrors (include/nsx_host.h, include/nsx.h)
DI = dict(ELEM=0, NCELLS=1, NVERTS=2, N_U=3, N_P=4, DOFS_PER_CELL=5, NQ=6, NQF=7, NRANKS=8, NBC=9, NVPC=10)
DA = dict(CELL_DOFS=(0, np.uint32), CELL_VERTICES=(1, np.float64), CELL_RANK=(2, np.int32),
          OWNED_U=(3, np.int64), OWNED_P=(4, np.int64),
          F_ROWPTR=(10, np.int64), F_COL=(11, np.int32), BT_ROWPTR=(12, np.int64), BT_COL=(13, np.int32),
          B_ROWPTR=(14, np.int64), B_COL=(15, np.int32), MP_ROWPTR=(16, np.int64), MP_COL=(17, np.int32),
          BC_DOF=(20, np.uint32), BC_SHAPE=(21, np.float64), BC_ON_INLET=(22, np.uint8), BC_Y=(23, np.float64),
          OUTLET_CELL=(30, np.int32), OUTLET_FACE=(31, np.int32), CYL_CELL=(32, np.int32), CYL_FACE=(33, np.int32),
          BFACES=(34, np.int32), MATERIAL=(35, np.int32), FE_TABLES=(40, np.uint8))
BLOCK_F, BLOCK_BT, BLOCK_B, BLOCK_MP, BLOCK_S, BLOCK_J = 0, 1, 2, 3, 4, 5
MODE_STOKES, MODE_NEWTON, MODE_UNSTEADY_FIRST, MODE_UNSTEADY_NEWTON = 0, 1, 2, 3
VEC_SOLUTION, VEC_SOLUTION_OLD, VEC_DELTA, VEC_RESIDUAL, VEC_EVAL = 0, 1, 2, 3, 4

_nsx = None
_orc = None


def nsx():
    """The product library.  Loading works without a GPU; compute calls need one."""
    global _nsx
    if _nsx is None:
        if not os.path.exists(LIBNSX):
            raise RuntimeError(f"{LIBNSX} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIBNSX)
        L.nsx_disc_generate.restype = C.c_void_p
        L.nsx_disc_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.nsx_disc_from_gmsh.restype = C.c_void_p
        L.nsx_disc_from_gmsh.argtypes = [C.c_char_p, C.c_int]
        L.nsx_disc_free.argtypes = [C.c_void_p]
        L.nsx_disc_info.restype = C.c_int64
        L.nsx_disc_info.argtypes = [C.c_void_p, C.c_int]
        L.nsx_disc_array.restype = C.c_void_p
        L.nsx_disc_array.argtypes = [C.c_void_p, C.c_int, c_i64p]
        L.nsx_disc_inlet_values.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.nsx_host_last_error.restype = C.c_char_p
        _nsx = L
    return _nsx


def orc():
    global _orc
    if _orc is None:
        if not os.path.exists(LIBORC):
            raise RuntimeError(f"{LIBORC} is missing: run `make -C oracle`")
        L = C.CDLL(LIBORC)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_pattern.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_set_faces.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_ranks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_vec.restype = C.c_void_p
        L.orc_vec.argtypes = [C.c_void_p, C.c_int]
        L.orc_block_values.restype = C.c_void_p
        L.orc_block_values.argtypes = [C.c_void_p, C.c_int, c_i64p]
        L.orc_block_pattern.restype = C.c_int64
        L.orc_block_pattern.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orc_assemble.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_dp]
        L.orc_assemble_cells.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                C.POINTER(C.c_int), c_dp, c_i64p]
        L.orc_spmv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_inner_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ilu0_factor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_schur.argtypes = [C.c_void_p]
        L.orc_precond_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_lift_drag.argtypes = [C.c_void_p, C.c_double, c_dp, c_dp]
        L.orc_newton_stationary.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int,
                                            C.c_void_p, C.c_int64, c_i64p, c_dp, c_dp]
        L.orc_run_unsteady.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                       C.c_double, C.c_int, C.c_void_p, C.c_int64, c_i64p, c_dp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_threads.argtypes = [C.c_int]
        _orc = L
    return _orc


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Disc:
    """Host-side discretisation (stand-in for deal.II's set-up) through include/nsx_host.h."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(nsx().nsx_host_last_error().decode())
        self.h = handle
        L = nsx()
        for k, v in DI.items():
            setattr(self, k.lower(), int(L.nsx_disc_info(self.h, v)))
        self.n = self.n_u + self.n_p

    @classmethod
    def generate(cls, nx, ny, triangles=False, nranks=1):
        return cls(nsx().nsx_disc_generate(nx, ny, int(triangles), nranks))

    @classmethod
    def from_gmsh(cls, path, nranks=1):
        return cls(nsx().nsx_disc_from_gmsh(path.encode(), nranks))

    def array(self, name):
        code, dt = DA[name]
        cnt = C.c_int64()
        p = nsx().nsx_disc_array(self.h, code, C.byref(cnt))
        if cnt.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (cnt.value * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt)

    def inlet_values(self, amplitude):
        v = np.zeros(self.nbc)
        nsx().nsx_disc_inlet_values(self.h, amplitude, ptr(v))
        return v

    def pattern(self, name):
        return self.array(name + "_ROWPTR"), self.array(name + "_COL")

    def __del__(self):
        try:
            nsx().nsx_disc_free(self.h)
        except Exception:
            pass


def golden_mesh_path():
    """tests/golden/new_mesh.msh.gz (the reference's lab_new/mesh/new_mesh.msh input) unpacked to a temp file."""
    src = os.path.join(ROOT, "tests", "golden", "new_mesh.msh.gz")
    dst = os.path.join(tempfile.gettempdir(), "nsx_new_mesh.msh")
    if not os.path.exists(dst):
        with gzip.open(src, "rb") as f, open(dst + ".tmp", "wb") as g:
            shutil.copyfileobj(f, g)
        os.replace(dst + ".tmp", dst)
    return dst


class Oracle:
    """CPU restatement of the reference path on a Disc."""

    def __init__(self, disc, inlet_amplitude=0.1):
        L = orc()
        self.disc = disc
        self.cd = np.ascontiguousarray(disc.array("CELL_DOFS"))
        self.cv = np.ascontiguousarray(disc.array("CELL_VERTICES"))
        self.h = L.orc_create(disc.elem, disc.ncells, ptr(self.cv), ptr(self.cd), disc.n_u, disc.n_p)
        self.n_u, self.n_p, self.n = disc.n_u, disc.n_p, disc.n
        shapes = {BLOCK_F: (disc.n_u, disc.n_u), BLOCK_BT: (disc.n_u, disc.n_p),
                  BLOCK_B: (disc.n_p, disc.n_u), BLOCK_MP: (disc.n_p, disc.n_p)}
        self.patterns = {}
        for blk, name in ((BLOCK_F, "F"), (BLOCK_BT, "BT"), (BLOCK_B, "B"), (BLOCK_MP, "MP")):
            rp, col = disc.pattern(name)
            self.patterns[blk] = (rp, col, shapes[blk])
            L.orc_set_pattern(self.h, blk, shapes[blk][0], shapes[blk][1], ptr(rp), ptr(col))
        oc, of = disc.array("OUTLET_CELL"), disc.array("OUTLET_FACE")
        L.orc_set_faces(self.h, 8, len(oc), ptr(oc), ptr(of))
        cc, cf = disc.array("CYL_CELL"), disc.array("CYL_FACE")
        L.orc_set_faces(self.h, 10, len(cc), ptr(cc), ptr(cf))
        self.bc_dof = disc.array("BC_DOF")
        self.bc_val = disc.inlet_values(inlet_amplitude)
        L.orc_set_dirichlet(self.h, len(self.bc_dof), ptr(self.bc_dof), ptr(self.bc_val))
        ou, op = disc.array("OWNED_U"), disc.array("OWNED_P")
        L.orc_set_ranks(self.h, disc.nranks, ptr(ou), ptr(op))

    def vec(self, which):
        p = orc().orc_vec(self.h, which)
        return np.frombuffer((C.c_double * self.n).from_address(p), dtype=np.float64)

    def values(self, block):
        nnz = C.c_int64()
        p = orc().orc_block_values(self.h, block, C.byref(nnz))
        if nnz.value == 0:
            return np.zeros(0)
        return np.frombuffer((C.c_double * nnz.value).from_address(p), dtype=np.float64)

    def block_pattern(self, block):
        rp, col = C.c_void_p(), C.c_void_p()
        nrows = orc().orc_block_pattern(self.h, block, C.byref(rp), C.byref(col))
        rowptr = np.frombuffer((C.c_int64 * (nrows + 1)).from_address(rp.value), dtype=np.int64)
        cols = np.frombuffer((C.c_int32 * int(rowptr[-1])).from_address(col.value), dtype=np.int32)
        return rowptr, cols

    def csr(self, block):
        import scipy.sparse as sp
        if block == BLOCK_S:
            rp, col = self.block_pattern(block)
            return sp.csr_matrix((self.values(block).copy(), col.copy(), rp.copy()), shape=(self.n_p, self.n_p))
        rp, col, shape = self.patterns[block]
        return sp.csr_matrix((self.values(block).copy(), col, rp), shape=shape)

    def jacobian(self):
        import scipy.sparse as sp
        return sp.bmat([[self.csr(BLOCK_F), self.csr(BLOCK_BT)], [self.csr(BLOCK_B), None]], format="csr")

    def assemble(self, mode, apply_inlet, nu, dt=0.01, p_out=1.0):
        r = C.c_double()
        rc = orc().orc_assemble(self.h, mode, int(apply_inlet), nu, dt, p_out, C.byref(r))
        if rc:
            raise RuntimeError(orc().orc_last_error().decode())
        return r.value

    def assemble_cells(self, mode, nu, dt=0.01, p_out=1.0):
        rc = orc().orc_assemble_cells(self.h, mode, nu, dt, p_out)
        if rc:
            raise RuntimeError(orc().orc_last_error().decode())

    def solve(self, flavour, solver, prec, tol, max_it=20000, alpha=0.5):
        it, fr = C.c_int(), C.c_double()
        inner = np.zeros(3, dtype=np.int64)
        rc = orc().orc_solve(self.h, flavour, solver, prec, tol, max_it, alpha, C.byref(it), C.byref(fr),
                             inner.ctypes.data_as(c_i64p))
        return rc, it.value, fr.value, inner

    def spmv(self, block, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        nrows = self.n if block == BLOCK_J else (self.n_p if block in (BLOCK_B, BLOCK_MP, BLOCK_S) else self.n_u)
        y = np.zeros(nrows)
        orc().orc_spmv(self.h, block, ptr(x), ptr(y))
        return y

    def inner_apply(self, block, kind, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        rc = orc().orc_inner_apply(self.h, block, kind, ptr(x), ptr(y))
        if rc:
            raise RuntimeError(orc().orc_last_error().decode())
        return y

    def ilu0_factor(self, block):
        lu = np.zeros_like(self.values(block))
        orc().orc_ilu0_factor(self.h, block, ptr(lu))
        return lu

    def schur(self):
        orc().orc_schur(self.h)
        return self.csr(BLOCK_S)

    def precond_apply(self, flavour, prec, src, dst0=None, alpha=0.5):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.zeros(self.n) if dst0 is None else np.array(dst0, dtype=np.float64)
        rc = orc().orc_precond_apply(self.h, flavour, prec, alpha, ptr(src), ptr(dst))
        if rc:
            raise RuntimeError(orc().orc_last_error().decode())
        return dst

    def lift_drag(self, nu):
        d, l = C.c_double(), C.c_double()
        orc().orc_lift_drag(self.h, nu, C.byref(d), C.byref(l))
        return d.value, l.value

    def newton_stationary(self, Re, solver, prec, tol, max_newton_total=0, cap=100000):
        log = np.zeros((cap, 8))
        nlog, nu, u = C.c_int64(), C.c_double(), C.c_double()
        rc = orc().orc_newton_stationary(self.h, Re, solver, prec, tol, max_newton_total, ptr(log), cap,
                                         C.byref(nlog), C.byref(nu), C.byref(u))
        return rc, log[: min(nlog.value, cap)], nu.value, u.value

    def run_unsteady(self, Re, T, dt, solver, prec, tol, n_steps_max=0, cap=100000):
        log = np.zeros((cap, 8))
        nlog, nu = C.c_int64(), C.c_double()
        rc = orc().orc_run_unsteady(self.h, Re, T, dt, solver, prec, tol, n_steps_max, ptr(log), cap,
                                    C.byref(nlog), C.byref(nu))
        return rc, log[: min(nlog.value, cap)], nu.value

    def __del__(self):
        try:
            orc().orc_destroy(self.h)
        except Exception:
            pass


def synthetic_state(disc, seed=1234):
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
    sol[:n_u] = np.where(even, ux, 0.0) + rng.uniform(-1e-2, 1e-2, n_u)
    sol[n_u:] = 1 + (2.2 - x[n_u:]) * 0.05 + rng.uniform(-1e-2, 1e-2, n_p)
    return sol
