"""smoke(): one small invocation of the hot path on cuda:0 (assemble + solve + update + lift/drag
through the C ABI), checked against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    sys.path.insert(0, ROOT)
    from navier_stokes_solver_b200 import binding as B
    from oracle.pyoracle import Oracle   # checker only

    d = B.Disc.generate(24, 10)
    dev = B.Device(d, device_id=0, ordering=0)
    orc = Oracle(d)
    sol = B.synthetic_state(d, 1, noise=1e-4)
    orc.vec(0)[:] = sol
    dev.upload(B.VEC_SOLUTION, sol)
    nu = 0.1
    r_o = orc.assemble(B.MODE_NEWTON, True, nu)
    r_d = dev.assemble(B.MODE_NEWTON, True, nu)
    for blk in (B.BLOCK_F, B.BLOCK_BT, B.BLOCK_B, B.BLOCK_MP):
        a, b = dev.values(blk), orc.values(blk)
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), f"assembly mismatch in block {blk}"
    assert abs(r_d - r_o) <= 1e-12 * r_o
    rc_o, it_o, _, _ = orc.solve(B.STATIONARY, 1, 2, 1e-12, 2000)
    rc_d, it_d, _ = dev.solve(B.STATIONARY, 1, 2, 1e-12, 2000)
    assert rc_o == 0 and rc_d == 0, (rc_o, rc_d)
    x_o, x_d = orc.vec(2), dev.download(B.VEC_DELTA)
    assert np.linalg.norm(x_d - x_o) <= 1e-8 * np.linalg.norm(x_o)
    # the same system once more in the library's default configuration (elimination order 2: CTA-local blocks swept out of shared
    # memory by the TMA-fed kernel; inner FGMRES recurrences on the device; batched Gram-Schmidt), against the oracle given the
    # same blocks and sequences: Ifpack's overlap-0 semantics at one rank per block
    dev.set_option(B.OPT_BLOCK_ROWS, 512)
    dev.set_option(B.OPT_ORDERING, 2)
    dev.upload(B.VEC_DELTA, np.zeros(d.n))
    orc.vec(2)[:] = 0
    dev.assemble(B.MODE_NEWTON, True, nu)
    for which, block in ((0, B.BLOCK_F), (1, B.BLOCK_MP)):   # aSIMPLE sweeps are ILU(0): always the full pattern
        off, perm = dev.sweep_blocks(block)
        orc.set_blocks(which, off, perm)
    rc_b, it_b, _, _ = orc.solve(B.STATIONARY, 1, 2, 1e-12, 4000)
    rc_m, it_m, _ = dev.solve(B.STATIONARY, 1, 2, 1e-12, 4000)
    assert rc_m == 0 and rc_b == 0, (rc_m, rc_b)
    x_m, x_b = dev.download(B.VEC_DELTA), orc.vec(2).copy()
    assert np.linalg.norm(x_m - x_b) <= 1e-8 * np.linalg.norm(x_b)
    assert np.linalg.norm(x_m - x_o) <= 1e-8 * np.linalg.norm(x_o)
    orc.set_blocks(0); orc.set_blocks(1)
    dev.set_option(B.OPT_ORDERING, 0)
    dev.save_eval_point()
    dev.update(1.0)
    orc.vec(0)[:] = sol + x_o
    dd, ld = dev.lift_drag(nu)
    do, lo = orc.lift_drag(nu)
    assert abs(dd - do) <= 1e-6 * abs(do) and abs(ld - lo) <= 1e-6 * max(abs(lo), abs(do))
    print(f"smoke ok: {d.ncells} cells, {d.n} dofs, FGMRES+aSIMPLE iterations gpu {it_d} / oracle {it_o} (block-local order: gpu {it_m} / oracle {it_b}), "
          f"kernel launches {dev.stat('KERNEL_LAUNCHES')}")
