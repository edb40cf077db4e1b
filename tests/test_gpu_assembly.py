"""GPU assembly (nsx_assemble through the C ABI) against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): assembled matrix entries match to 1e-12 relative.  Entries are
compared against the largest magnitude of their block: structurally present entries that cancel to
~0 (e.g. the u_x-u_y couplings) have no meaningful entry-wise relative error."""
import numpy as np
import pytest

import nsxlib as N

pytestmark = pytest.mark.gpu

RTOL = 1e-12
MODES = [N.MODE_STOKES, N.MODE_NEWTON, N.MODE_UNSTEADY_FIRST, N.MODE_UNSTEADY_NEWTON]


def discs():
    return {
        "quad": N.Disc.generate(30, 12),
        "tri": N.Disc.generate(22, 9, triangles=True),
    }


@pytest.fixture(scope="module")
def setups():
    out = {}
    for k, d in discs().items():
        out[k] = (d, N.Oracle(d), N.Device(d))
    return out


def state(d, seed):
    rng = np.random.default_rng(seed)
    sol = N.synthetic_state(d, seed)
    old = sol + rng.uniform(-1e-2, 1e-2, d.n)
    return sol, old


def block_close(a, b, what):
    scale = np.abs(b).max()
    err = np.abs(a - b).max()
    assert err <= RTOL * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("elem", ["quad", "tri"])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("apply_inlet", [False, True])
def test_assemble_matches_oracle(setups, elem, mode, apply_inlet):
    d, orc, dev = setups[elem]
    sol, old = state(d, 7 + mode)
    nu, dt = 1.0 / 90.0, 0.01
    orc.vec(0)[:] = sol
    orc.vec(1)[:] = old
    orc.vec(2)[:] = 0.0
    dev.upload(N.VEC_SOLUTION, sol)
    dev.upload(N.VEC_SOLUTION_OLD, old)
    dev.upload(N.VEC_DELTA, np.zeros(d.n))
    r_o = orc.assemble(mode, apply_inlet, nu, dt)
    r_d = dev.assemble(mode, apply_inlet, nu, dt)
    for blk, name in ((N.BLOCK_F, "F"), (N.BLOCK_BT, "Bt"), (N.BLOCK_B, "B"), (N.BLOCK_MP, "Mp")):
        block_close(dev.values(blk), orc.values(blk), f"{elem} mode {mode} {name}")
    block_close(dev.download(N.VEC_RESIDUAL), orc.vec(3), "residual")
    # Dirichlet values land in delta (the warm start of the Krylov solve)
    np.testing.assert_array_equal(dev.download(N.VEC_DELTA), orc.vec(2))
    assert abs(r_d - r_o) <= 1e-12 * max(r_o, 1e-300)


@pytest.mark.parametrize("elem", ["quad", "tri"])
def test_cells_only_and_structure(setups, elem):
    """Before the Dirichlet step the Stokes-branch Jacobian is symmetric and B = Bt^T."""
    d, orc, dev = setups[elem]
    sol, _ = state(d, 3)
    dev.upload(N.VEC_SOLUTION, sol)
    dev.assemble_cells(N.MODE_STOKES, 0.1)
    F, Bt, B, Mp = (dev.csr(b) for b in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B, N.BLOCK_MP))
    assert abs(F - F.T).max() <= 1e-13 * abs(F).max()
    assert abs(B - Bt.T).max() <= 1e-13 * abs(B).max()
    # Mp row sums = integral of psi_i / nu ; the total is the fluid area / nu
    cv = d.array("CELL_VERTICES").reshape(d.ncells, d.nvpc, 2)
    if d.elem == 0:
        area = np.sum((cv[:, 1, 0] - cv[:, 0, 0]) * (cv[:, 2, 1] - cv[:, 0, 1]))
    else:
        a, b, c = cv[:, 0], cv[:, 1], cv[:, 2]
        area = 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])).sum()
    assert abs(Mp.sum() * 0.1 - area) <= 1e-12 * area
    # constant pressure is in the kernel of Bt's interior rows: Bt 1 = -int div(phi_i) = boundary flux only
    ones = np.ones(d.n_p)
    flux = Bt @ ones
    bc = d.array("BC_DOF")
    out = d.array("CELL_DOFS").reshape(d.ncells, -1)[d.array("OUTLET_CELL")].ravel()
    interior = np.ones(d.n_u, bool)
    interior[bc] = False
    interior[out[out < d.n_u]] = False
    assert np.abs(flux[interior]).max() <= 1e-13


def test_lift_drag(setups):
    for elem in ("quad", "tri"):
        d, orc, dev = setups[elem]
        sol, _ = state(d, 11)
        orc.vec(0)[:] = sol
        dev.upload(N.VEC_SOLUTION, sol)
        do, lo = orc.lift_drag(0.05)
        dd, ld = dev.lift_drag(0.05)
        assert len(d.array("CYL_CELL")) > 0
        assert abs(dd - do) <= 1e-12 * abs(do) and abs(ld - lo) <= 1e-12 * abs(lo)


def test_new_mesh_newton(tmp_path):
    """The reference's -M input (lab_new/mesh/new_mesh.msh, P2/P1) at full size."""
    d = N.Disc.from_gmsh(N.golden_mesh_path())
    orc, dev = N.Oracle(d), N.Device(d)
    sol = N.synthetic_state(d, 5)
    orc.vec(0)[:] = sol
    dev.upload(N.VEC_SOLUTION, sol)
    r_o = orc.assemble(N.MODE_NEWTON, False, 0.05)
    r_d = dev.assemble(N.MODE_NEWTON, False, 0.05)
    for blk in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B, N.BLOCK_MP):
        block_close(dev.values(blk), orc.values(blk), f"new_mesh block {blk}")
    assert abs(r_d - r_o) <= 1e-12 * r_o
    assert dev.stat("ASSEMBLY_COLOURS") < 40


@pytest.mark.parametrize("elem", ["quad", "tri"])
@pytest.mark.parametrize("mode", [N.MODE_STOKES, N.MODE_NEWTON, N.MODE_UNSTEADY_FIRST, N.MODE_UNSTEADY_NEWTON])
def test_residual_only_assembly_is_the_full_assembly_s_residual(elem, mode):
    """nsx_assemble_residual (the line search's assembly): the residual vector, its norm and the Dirichlet entries of delta are
    bit for bit what the full assembly with homogeneous values leaves; the matrices are not touched."""
    d = N.Disc.generate(14, 6) if elem == "quad" else N.Disc.generate(12, 6, triangles=True)
    dev = N.Device(d)
    rng = np.random.default_rng(5)
    sol, old = N.synthetic_state(d, 3), N.synthetic_state(d, 4)
    dev.upload(N.VEC_SOLUTION, sol)
    dev.upload(N.VEC_SOLUTION_OLD, old)
    dev.upload(N.VEC_DELTA, rng.uniform(-1, 1, d.n))
    r_full = dev.assemble(mode, False, 1 / 30.0, 0.01)
    res_full, delta_full = dev.download(N.VEC_RESIDUAL), dev.download(N.VEC_DELTA)
    mats = {b: dev.values(b) for b in (N.BLOCK_F, N.BLOCK_BT, N.BLOCK_B, N.BLOCK_MP)}
    # another state: the residual-only call must follow the state, the matrices must stay
    dev.upload(N.VEC_SOLUTION, sol * 1.5)
    r_other = dev.assemble_residual(mode, 1 / 30.0, 0.01)
    for b, v in mats.items():
        np.testing.assert_array_equal(dev.values(b), v)
    if mode in (N.MODE_NEWTON, N.MODE_UNSTEADY_NEWTON):
        assert r_other != r_full
    dev.upload(N.VEC_SOLUTION, sol)
    dev.upload(N.VEC_DELTA, rng.uniform(-1, 1, d.n))
    r_only = dev.assemble_residual(mode, 1 / 30.0, 0.01)
    assert r_only == r_full
    np.testing.assert_array_equal(dev.download(N.VEC_RESIDUAL), res_full)
    bc = d.array("BC_DOF")
    assert (dev.download(N.VEC_DELTA)[bc] == 0).all() and (delta_full[bc] == 0).all()
