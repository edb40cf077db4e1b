#!/usr/bin/env python
"""Runs the CPU oracle's FULL first-step solve of a bench configuration (README: -m 300,100 -s 1 -p 0 -t 1e-10, Stokes branch,
nu = 1/10) and records its iteration counts in profiles/oracle_full_solves.json.  bench.py extrapolates its bounded CPU samples
to these counts.  Takes hours at 300x100 (5.1 h on 7 threads of the 8-core build container); not something for a bench lease.
usage: oracle_full_solve.py NX,NY BLOCKS [THREADS]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from navier_stokes_solver_b200 import binding as B  # noqa: E402
from oracle.pyoracle import Oracle, orc  # noqa: E402

mesh, blocks = sys.argv[1], int(sys.argv[2])
threads = int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count()
nx, ny = (int(v) for v in mesh.split(","))
orc().orc_set_threads(threads)
o = Oracle(B.Disc.generate(nx, ny, nranks=blocks))
t0 = time.perf_counter()
r0 = o.assemble(0, True, 0.1)
t_asm = time.perf_counter() - t0
t0 = time.perf_counter()
rc, it, fr, inner = o.solve(0, 1, 0, 1e-10, 20000)
t_solve = time.perf_counter() - t0
rec = {"mesh": mesh, "solver": 1, "prec": 0, "tol": 1e-10, "blocks": blocks, "threads": threads, "rc": rc, "outer": it,
       "inner_F": int(inner[0]), "inner_Mp": int(inner[1]), "final_residual": fr, "t_assemble_s": t_asm, "t_solve_s": t_solve,
       "host": "build container (8 cores)", "residual0": r0}
print(json.dumps(rec))
p = os.path.join(ROOT, "profiles", "oracle_full_solves.json")
recs = json.load(open(p)) if os.path.exists(p) else []
recs = [r for r in recs if not (r["mesh"] == mesh and r["blocks"] == blocks)] + [rec]
json.dump(recs, open(p, "w"), indent=1)
