"""Phase timing of one Newton step on the GPU (diagnostic; prints one line per phase)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navier_stokes_solver_b200 import binding as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", default="300,100")
ap.add_argument("--ordering", type=int, default=1)
ap.add_argument("--solver", type=int, default=1)
ap.add_argument("--prec", type=int, default=0)
ap.add_argument("--flavour", type=int, default=0)
ap.add_argument("--cap", type=int, default=5)
ap.add_argument("--tol", type=float, default=1e-10)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--full", type=int, default=0)
ap.add_argument("--kernels-only", action="store_true")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--kernels", default="5,6,7,0,1,2,3,4")
ap.add_argument("--spmv", type=int, default=3)
args = ap.parse_args()
nx, ny = map(int, args.mesh.split(","))


def T(label, fn):
    t = time.perf_counter()
    r = fn()
    dev.synchronize() if "dev" in globals() else None
    print(f"{label:40s} {time.perf_counter() - t:9.4f} s", flush=True)
    return r


t = time.perf_counter()
d = B.Disc.generate(nx, ny)
print(f"{'host set-up (mesh, dofs, sparsity)':40s} {time.perf_counter() - t:9.4f} s   cells {d.ncells} dofs {d.n}", flush=True)
t = time.perf_counter()
dev = B.Device(d, ordering=args.ordering)
print(f"{'device set-up (upload, colours, tables)':40s} {time.perf_counter() - t:9.4f} s   colours {dev.stat('ASSEMBLY_COLOURS')} tables {dev.stat('ASSEMBLY_TABLES')}", flush=True)
nu = 0.1
dev.set_option(B.OPT_STREAM_SPMV, args.spmv)
if args.mode == 1:
    dev.upload(B.VEC_SOLUTION, B.synthetic_state(d, 1234, noise=1e-4))
T("assemble (first)", lambda: dev.assemble(args.mode, True, nu))
T("assemble (second)", lambda: dev.assemble(args.mode, True, nu))
for k in [int(x) for x in args.kernels.split(',')]:
    T(f"time_kernel {k} first call (plan build)", lambda: dev.time_kernel(k, 1, False))
    ms = dev.time_kernel(k, args.reps, True)
    print(f"   kernel {k}: {ms:.4f} ms   levels_F {dev.stat('LEVELS_F')}", flush=True)
for cap in ([] if args.kernels_only else [2, args.cap] + ([args.full] if args.full else [])):
    dev.vec_set(B.VEC_SOLUTION, 0.0) if args.mode != 1 else None
    dev.vec_set(B.VEC_DELTA, 0.0)
    dev.assemble(args.mode, True, nu)
    l0 = dev.stat("KERNEL_LAUNCHES")
    rc, it, fr = T(f"solve capped at {cap} outer its", lambda: dev.solve(args.flavour, args.solver, args.prec, args.tol, cap))
    print(f"   rc {rc} it {it} res {fr:.3e} inner_F {dev.stat('INNER_F')} inner_S {dev.stat('INNER_S')} applies {dev.stat('PRECOND_APPLIES')} "
          f"launches {dev.stat('KERNEL_LAUNCHES') - l0}", flush=True)
