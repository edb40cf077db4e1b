#!/usr/bin/env python
"""Config 3 diagnosis on the GPU: first solve of the first time step of `NSSolver -M x -r 100 -T 8,0.01 -s 1 -p 2` (first_iter branch,
nu = 1) on the reference's mesh, outer FGMRES + unsteady aSIMPLE, continued in chunks of the reference's own cap (100 000 iterations,
NSSolver.cpp:604) with the increment kept as the warm start, so that the residual after every chunk shows whether the iteration
converges at all.  usage: config3_probe.py [gmsh|tri:NX,NY|NX,NY] [ordering] [chunks] [tol] [prec]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nsxlib as N  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "gmsh"
ordering = int(sys.argv[2]) if len(sys.argv) > 2 else 0
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 10
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
prec = int(sys.argv[5]) if len(sys.argv) > 5 else 2
if mesh == "gmsh":
    d = N.Disc.from_gmsh(N.golden_mesh_path())
elif mesh.startswith("tri:"):
    d = N.Disc.generate(*[int(v) for v in mesh[4:].split(",")], triangles=True)
else:
    d = N.Disc.generate(*[int(v) for v in mesh.split(",")])
dev = N.Device(d, inlet_amplitude=0.3, ordering=ordering)
dev.upload(N.VEC_SOLUTION, np.zeros(d.n)); dev.upload(N.VEC_SOLUTION_OLD, np.zeros(d.n)); dev.upload(N.VEC_DELTA, np.zeros(d.n))
for label, mode, nu in (("first_iter nu=1", N.MODE_UNSTEADY_FIRST, 1.0),):
    r0 = dev.assemble(mode, True, nu, 0.01)
    hist = []
    total = 0
    t0 = time.perf_counter()
    for c in range(chunks):
        rc, it, fr = dev.solve(N.UNSTEADY, 1, prec, tol, 100000)
        total += it
        hist.append((total, fr))
        print(json.dumps({"mesh": mesh, "dofs": d.n, "ordering": ordering, "prec": prec, "case": label, "chunk": c, "rc": rc, "iterations_total": total,
                          "residual": fr, "rhs_norm": r0, "elapsed_s": time.perf_counter() - t0}), flush=True)
        if rc == 0:
            break
