#!/usr/bin/env python
"""Development probe (GPU): kernel timings of the sweep / SpMV at a mesh size and one README step, for a set of option values.
usage: quick_step.py NX,NY [--ordering 2] [--host-inner 0] [--ortho 2] [--steps 1] [--block-rows 0] [--no-step]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from navier_stokes_solver_b200 import binding as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("mesh")
ap.add_argument("--ordering", type=int, default=2)
ap.add_argument("--host-inner", type=int, default=0)
ap.add_argument("--ortho", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--block-rows", type=int, default=0)
ap.add_argument("--no-step", action="store_true")
ap.add_argument("--no-timing", action="store_true")
ap.add_argument("--max-outer", type=int, default=20000)
ap.add_argument("--prec", type=int, default=0)
ap.add_argument("--verbose", type=int, default=1)
ap.add_argument("--l2-hints", type=int, default=1)
ap.add_argument("--sweep-q", type=int, default=4)
a = ap.parse_args()
nx, ny = (int(v) for v in a.mesh.split(","))
t0 = time.perf_counter()
d = B.Disc.generate(nx, ny)
t1 = time.perf_counter()
dev = B.Device(d, ordering=a.ordering, ortho=a.ortho, block_rows=a.block_rows or None)
dev.set_option(B.OPT_VERBOSE, a.verbose)
dev.set_option(B.OPT_HOST_INNER, a.host_inner)
dev.set_option(B.OPT_L2_HINTS, a.l2_hints)
dev.set_option(B.OPT_SWEEP_Q, a.sweep_q)
out = {"mesh": a.mesh, "cells": d.ncells, "dofs": d.n, "ordering": a.ordering, "host_inner": a.host_inner, "ortho": a.ortho,
       "disc_s": t1 - t0, "device_setup_s": time.perf_counter() - t1}
nu = 0.1
dev.upload(B.VEC_SOLUTION, np.zeros(d.n))
dev.upload(B.VEC_DELTA, np.zeros(d.n))
r0 = dev.assemble(B.MODE_STOKES, True, nu)
dev.upload(B.VEC_TMP0, np.random.default_rng(42).uniform(-1, 1, d.n))
t2 = time.perf_counter()
k = {}
for name, w in (() if a.no_timing else (("block_spmv", 0), ("spmv_F", 1), ("sgs_F", 5), ("ilu_apply_F", 6), ("dot", 3))):
    dev.time_kernel(w, 3, 0)
    k[name + "_b2b_ms"] = dev.time_kernel(w, 30, 2)
    k[name + "_flushed_ms"] = dev.time_kernel(w, 10, 1)
out["plan_and_timing_s"] = time.perf_counter() - t2
out["kernels"] = k
out["levels_F"] = dev.stat("LEVELS_F")
nnzF = dev.nnz(B.BLOCK_F)
if not a.no_timing:
    out["sgs_GBps_algorithmic"] = (12 * nnzF + 8 * (d.n_u + 1) + 32 * d.n_u) / (k["sgs_F_b2b_ms"] * 1e-3) / 1e9
print(json.dumps(out), flush=True)
if not a.no_step:
    for s in range(a.steps):
        dev.upload(B.VEC_SOLUTION, np.zeros(d.n))
        dev.upload(B.VEC_DELTA, np.zeros(d.n))
        l0 = dev.stat("KERNEL_LAUNCHES")
        t = time.perf_counter()
        r0 = dev.assemble(B.MODE_STOKES, True, nu)
        rc, it, fr = dev.solve(B.STATIONARY, 1, a.prec, 1e-10, a.max_outer)
        dev.save_eval_point(); dev.update(1.0)
        r1 = dev.assemble(B.MODE_STOKES, False, nu)
        dev.synchronize()
        t = time.perf_counter() - t
        print(json.dumps({"step": s, "s": t, "rc": rc, "outer": it, "inner_F": dev.stat("INNER_F"), "inner_S": dev.stat("INNER_S"), "res": fr,
                          "r0": r0, "r1": r1, "launches": dev.stat("KERNEL_LAUNCHES") - l0,
                          "us_per_inner": 1e6 * t / max(1, dev.stat("INNER_F"))}), flush=True)
