"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBORC = os.path.join(ROOT, "oracle", "liboracle.so")
c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)
BLOCK_F, BLOCK_BT, BLOCK_B, BLOCK_MP, BLOCK_S, BLOCK_J = 0, 1, 2, 3, 4, 5
_orc = None


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def orc():
    global _orc
    if _orc is None:
        if not os.path.exists(LIBORC):
            raise RuntimeError(f"{LIBORC} is missing: run `make -C oracle`")
        # OpenMP threads that sleep instead of spinning at barriers: a loaded host (other jobs, more threads than free cores) then
        # costs the oracle a little instead of orders of magnitude (read by libgomp when it is first loaded)
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        L = C.CDLL(LIBORC)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_pattern.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_set_faces.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_ranks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
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

    def set_blocks(self, which, ord_off=None, order=None):
        """Sub-blocks + elimination sequences of the ILU / SGS sweeps (which: 0 velocity rows, 1 pressure rows)."""
        if ord_off is None:
            rc = orc().orc_set_blocks(self.h, which, 0, None, None)
        else:
            ord_off = np.ascontiguousarray(ord_off, dtype=np.int64)
            order = np.ascontiguousarray(order, dtype=np.int32)
            rc = orc().orc_set_blocks(self.h, which, len(ord_off) - 1, ptr(ord_off), ptr(order))
        if rc:
            raise ValueError("orc_set_blocks: the sequences must cover every row exactly once")

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


