#!/usr/bin/env python
"""Small solves through every kernel family of the library, meant to run under compute-sanitizer (no torch import: ctypes only):
  compute-sanitizer --tool memcheck  python tools/sanitize_small.py --cap 10
  compute-sanitizer --tool racecheck python tools/sanitize_small.py --short
Covers: assembly (4 modes) + residual-only, Dirichlet step, block SpMV, node / same-component / full views of F, block-local SGS and
ILU(0) sweeps (orderings 2 and 3), natural and multicolour orders, device-driven inner FGMRES, CG, Schur product, AMG, lift / drag.
(On the pool this round was built on compute-sanitizer is closed; without it the script is a 10-second run through all families.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from navier_stokes_solver_b200 import binding as B  # noqa: E402

short = "--short" in sys.argv
cap = 3 if short else 4000
if "--cap" in sys.argv:
    cap = int(sys.argv[sys.argv.index("--cap") + 1])


def run(tri, ordering, cases):
    d = B.Disc.generate(10, 5, triangles=True) if tri else B.Disc.generate(12, 6)
    dev = B.Device(d, ordering=ordering, block_rows=256 if ordering in (2, 3) else None)
    sol = B.synthetic_state(d, 7, noise=1e-4)
    for flavour, solver, prec, mode in cases:
        dev.upload(B.VEC_SOLUTION, sol); dev.upload(B.VEC_SOLUTION_OLD, sol); dev.upload(B.VEC_DELTA, np.zeros(d.n))
        r = dev.assemble(mode, True, 0.1, 0.01)
        rr = dev.assemble_residual(mode, 0.1, 0.01)
        rc, it, fr = dev.solve(flavour, solver, prec, 1e-9 * r, cap)
        print(f"tri={int(tri)} ordering={ordering} flavour={flavour} solver={solver} prec={prec} mode={mode}: view {dev.view()} "
              f"|r|={r:.3e} res-only {rr:.3e} rc={rc} its={it} final {fr:.2e}", flush=True)
        if cap >= 4000 and rc != 0:
            raise SystemExit("solve did not converge")
    dev.lift_drag(0.1)


S, U = B.STATIONARY, B.UNSTEADY
run(False, None, [(S, 1, 0, B.MODE_STOKES), (S, 1, 0, B.MODE_NEWTON), (S, 1, 2, B.MODE_NEWTON), (U, 1, 2, B.MODE_UNSTEADY_FIRST),
                  (U, 1, 2, B.MODE_UNSTEADY_NEWTON), (U, 1, 0, B.MODE_UNSTEADY_NEWTON)])
if not short:
    run(True, None, [(S, 1, 0, B.MODE_STOKES), (S, 0, 1, B.MODE_NEWTON), (S, 2, 1, B.MODE_STOKES), (U, 1, 1, B.MODE_UNSTEADY_NEWTON)])
    run(False, 0, [(S, 1, 0, B.MODE_STOKES), (U, 1, 2, B.MODE_UNSTEADY_NEWTON)])
    run(True, 1, [(S, 1, 0, B.MODE_NEWTON)])
print("sanitize_small: done")
