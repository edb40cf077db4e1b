#!/usr/bin/env python
"""bench.py -- seconds per Newton step of the stationary solver on the README configuration
(-m 300,100 -r 100 -s 1 -t 1e-10 -p 0: Q3/Q2, FGMRES + blockDiagonal), and the block SpMV's HBM
bandwidth against the measured peak.

A step = the work of one iteration of NSSolverStationary::solve_newton's inner loop
(lab_new/src/NSSolverStationary.cpp:683-741) at the start of the run: assemble_system (Stokes
branch, inlet imposed) -> solve_system (outer Krylov + block preconditioner with its inner solves,
preconditioners rebuilt as the reference does) -> evaluation_point / solution update -> the
line-search re-assembly and its residual norm.  Every step restarts from the same state
(solution = 0, delta = 0) so that all steps do identical work.

  value : step time with all inputs resident in HBM (device-side reset of the state)
  e2e   : the same step through the C ABI with HOST buffers: the state is uploaded from pinned host
          memory and the new solution is read back inside the timed region
  roofline : the kernel with the largest share of the step (the ILU / SGS sweep), algorithmic bytes / CUDA-event time
  spmv_roofline : the Jacobian block SpMV the metric names
  cpu_baseline : the CPU oracle (port of the reference path) on the host cores, bounded sample, extrapolated (says so)

--budget-s (default 780): a safety net for the driver's wall-clock limit.  The first warm-up step is timed; when W + K steps
plus the kernel timings and the CPU sample would not fit the budget, the step COUNT is lowered (never the work inside a
step), the JSON line carries the counts actually run and config.budget_guard says what was asked for.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "s per Newton step"
UNIT = "s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def spmv_bytes(nnz_j, n):
    # 12 B per non-zero (value + column), row pointers, x read once, y written once
    return 12 * nnz_j + 4 * (n + 2) + 16 * n


def parse_mesh(s):
    a, b = s.split(",")
    return int(a), int(b)


class CpuSampler:
    """Bounded samples of the step on the CPU oracle: both assemblies in full, the solve capped at a number of outer iterations
    (raised once so that the capped solve lasts about `target_s` seconds, when target_s > 0).
    The reference runs one MPI rank per core (mpirun -n N) and its Ifpack / ML inner preconditioners are rank-local, so the
    oracle gets the mesh partitioned into one rank-local block per host thread: its SGS / ILU sweeps then run in parallel
    over the blocks exactly as the reference's ranks would."""

    def __init__(self, nx, ny, solver, prec, tol, nu, threads=None):
        from navier_stokes_solver_b200 import binding as B
        from oracle.pyoracle import Oracle, orc
        if not threads:
            # one rank per ~20 000 unknowns at least, as nobody runs a 3 k-DoF mesh on 16 MPI ranks (and OpenMP regions over a few
            # thousand entries cost more to start than to run)
            probe = B.Disc.generate(nx, ny)
            threads = int(max(1, min(int(orc().orc_get_threads()), probe.n // 20000)))
        orc().orc_set_threads(threads)
        self.threads = int(orc().orc_get_threads())
        self.o = Oracle(B.Disc.generate(nx, ny, nranks=self.threads))
        self.solver, self.prec, self.tol, self.nu = solver, prec, tol, nu
        self.cap = None

    def _capped(self, cap):
        o = self.o
        o.vec(0)[:] = 0
        o.vec(2)[:] = 0
        t0 = time.perf_counter()
        o.assemble(0, True, self.nu)
        t_asm = time.perf_counter() - t0
        t0 = time.perf_counter()
        rc, it, fr, inner = o.solve(0, self.solver, self.prec, self.tol, cap)
        return {"t_asm": t_asm, "t_solve": time.perf_counter() - t0, "outer_done": max(it, 1), "converged": rc == 0,
                "threads": self.threads, "inner_F": int(inner[0]), "inner_S": int(inner[1])}

    def sample(self, outer_cap, target_s=0.0):
        if self.cap is None:   # calibrate once: how many outer iterations last about target_s
            r = self._capped(outer_cap)
            self.cap = outer_cap
            if target_s > 0 and not r["converged"] and r["t_solve"] < 0.5 * target_s:
                self.cap = int(min(400, max(outer_cap + 1, outer_cap * target_s / max(r["t_solve"], 1e-3))))
                r = self._capped(self.cap)
            return r
        return self._capped(self.cap)


def outer_total_fixture(mesh, solver, prec, blocks=None):
    """Iteration counts of the FULL first-step solve by the CPU oracle itself (profiles/oracle_full_solves.json, written by
    tools/oracle_full_solve.py in the build container: the solve takes hours of 8 cores, far beyond a bench lease).  Prefers the
    record taken with `blocks` rank-local blocks (the preconditioner depends on the partition)."""
    p = os.path.join(ROOT, "profiles", "oracle_full_solves.json")
    best = None
    if os.path.exists(p):
        with open(p) as f:
            for rec in json.load(f):
                if rec["mesh"] == mesh and rec["solver"] == solver and rec["prec"] == prec:
                    if best is None or (blocks is not None and rec.get("blocks") == blocks):
                        best = rec
    return best


def extrapolation(smp, fixture, cpu_outer_total, fallback_outer, fallback_text):
    """(scale of the capped solve's time, text).  The cost of the solve is its inner iterations, so the sample is scaled by the
    ratio of inner F iterations when the oracle's full solve is on record; by outer iterations otherwise."""
    if smp["converged"]:
        return 1.0, smp["outer_done"], None
    if cpu_outer_total:
        return max(1.0, cpu_outer_total / smp["outer_done"]), cpu_outer_total, f"linearly in outer iterations to {cpu_outer_total} (--cpu-outer-total)"
    if fixture:
        src = f"{fixture.get('source', 'the oracle`s own full solve')}, {fixture.get('blocks', '?')} blocks, profiles/oracle_full_solves.json"
        if fixture.get("inner_F") and smp.get("inner_F"):
            return (max(1.0, fixture["inner_F"] / smp["inner_F"]), fixture["outer"],
                    f"by inner F iterations, {smp['inner_F']} done of {fixture['inner_F']} ({fixture['outer']} outer iterations) in {src}")
        return max(1.0, fixture["outer"] / smp["outer_done"]), fixture["outer"], f"linearly in outer iterations to {fixture['outer']}, the count of {src}"
    if fallback_outer:
        return max(1.0, fallback_outer / smp["outer_done"]), fallback_outer, f"linearly in outer iterations to {fallback_outer}, the count of {fallback_text}"
    return 1.0, smp["outer_done"], "no full count known: NOT extrapolated"


def workload_name(args, nx, ny):
    return (f"StationaryNSSolver -m {nx},{ny} -r 100 -s {args.solver} -t {args.tol:g} -p {args.prec}: first Newton step "
            "(assemble + solve + update + line-search assembly)")


def run_reference(args, emit):
    """--impl reference: the reference's CPU implementation of the path.  deal.II / Trilinos / MPI are not installable in this
    image, so the timed code is the oracle port of the reference path (oracle/, OpenMP over all host cores, one rank-local
    preconditioner block per thread as under mpirun).  A whole step of the 300x100 configuration takes the CPU the better
    part of an hour, so each bench step is a BOUNDED SAMPLE: both assemblies in full plus the solve capped at a number of outer
    iterations that lasts about --cpu-sample-s seconds; `value` is that sample extrapolated to the full solve by the ratio of
    inner F iterations (where the time goes), the full counts being those of the oracle's own complete solve of this step
    (profiles/oracle_full_solves.json: 685 outer / 172 072 inner iterations with 16 rank-local blocks, 5.1 h in the build
    container), or linearly in outer iterations to --cpu-outer-total.  The line says `extrapolated`, and carries the measured
    seconds beside the extrapolated ones; meshes small enough are simply solved in full (no extrapolation)."""
    from navier_stokes_solver_b200 import binding as B
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny = parse_mesh(args.mesh)
    d = B.Disc.generate(nx, ny)
    nu = 1.0 / 10.0
    small = d.n < 40000   # solved in full
    times, meas = [], []
    smp = None
    sampler = CpuSampler(nx, ny, args.solver, args.prec, args.tol, nu)
    fixture = outer_total_fixture(args.mesh, args.solver, args.prec, sampler.threads)
    scale, total, how = 1.0, 0, None
    for s in range(args.warmup + args.steps):
        smp = sampler.sample(20000 if small else args.cpu_outer_cap, target_s=0.0 if small else 0.5 * args.cpu_sample_s)
        scale, total, how = extrapolation(smp, fixture, args.cpu_outer_total, 0, "")
        if s >= args.warmup:
            times.append(2 * smp["t_asm"] + smp["t_solve"] * scale)
            meas.append(2 * smp["t_asm"] + smp["t_solve"])
    val = float(np.mean(times))
    extrapolated = not smp["converged"]
    cores = smp["threads"]
    sample = (f"oracle port, mesh cut into {cores} rank-local blocks on {cores} host threads (as mpirun -n {cores}): 2 full assemblies "
              f"({smp['t_asm']:.2f} s each) + solve " +
              (f"run to convergence ({smp['outer_done']} outer iterations, {smp['t_solve']:.2f} s)" if not extrapolated else
               f"capped at {smp['outer_done']} outer iterations ({smp['t_solve']:.2f} s measured), extrapolated {how}"))
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "extrapolated": extrapolated,
            "measured_sample_s": float(np.mean(meas)),
            "config": {"workload": workload_name(args, nx, ny), "cells": d.ncells, "dofs": d.n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "extrapolated": extrapolated,
                             "measured_s": float(np.mean(meas)), "outer_done": smp["outer_done"], "outer_total": total or smp["outer_done"], "scale": scale},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_unsteady(args, emit):
    """--workload unsteady: seconds per time step of NSSolver on the reference's gmsh mesh (tests/golden/new_mesh.msh.gz,
    P2/P1, 117 273 DoFs), Re = 100, dt = 0.01, FGMRES + aSIMPLE (BASELINE config 3).  A step is what NSSolver::solve does per time
    step (lab_new/src/NSSolver.cpp:814-835): solution_old = solution, solve_newton with its Reynolds continuation 1, 11, ..., 91
    (each stage: assemble / solve / line search until the residual stalls), lift and drag.  One GPU."""
    import ctypes
    import gzip
    import shutil
    import tempfile
    import torch
    from navier_stokes_solver_b200 import binding as B
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("the unsteady workload runs on one GPU (aSIMPLE on a partitioned system is not built yet)")
    if args.unsteady_mesh == "gmsh":
        mesh = os.path.join(tempfile.gettempdir(), "nsx_bench_new_mesh.msh")
        if not os.path.exists(mesh):
            with gzip.open(os.path.join(ROOT, "tests", "golden", "new_mesh.msh.gz"), "rb") as f, open(mesh, "wb") as g:
                shutil.copyfileobj(f, g)
        d = B.Disc.from_gmsh(mesh)
        mesh_flag = "-M"
    elif args.unsteady_mesh.startswith("tri:"):
        d = B.Disc.generate(*parse_mesh(args.unsteady_mesh[4:]), triangles=True)   # P2/P1 on the generated channel (every quad split in two)
        mesh_flag = f"-M [generated {args.unsteady_mesh[4:]} grid split into triangles]"
    else:
        d = B.Disc.generate(*parse_mesh(args.unsteady_mesh))
        mesh_flag = f"-m {args.unsteady_mesh}"
    stream = torch.cuda.Stream()
    dev = B.Device(d, inlet_amplitude=0.3, ordering=args.ordering, ortho=args.ortho, stream=ctypes.c_void_p(stream.cuda_stream))
    Re, dt, tol = 100.0, 0.01, args.tol if args.tol != 1e-10 else 1e-6   # the CLI default tolerance of the unsteady binary
    state = {"apply_first": True, "n": 0}
    log = {}

    def time_step():
        dev.copy_old()
        first_iter = True
        outer, solves, assemblies = 0, 0, 0
        nu = 1.0
        cur = 1.0
        while cur <= Re:
            nu = 1.0 / cur
            n_iter, res, prev = 0, 1e-9 + 1, 0.0
            while n_iter < 10 and res > 1e-9:
                if first_iter:
                    first_iter = False
                    fi = n_iter == 0
                else:
                    fi = False
                res = dev.assemble(B.MODE_UNSTEADY_FIRST if fi else B.MODE_UNSTEADY_NEWTON, fi and state["apply_first"], nu, dt)
                assemblies += 1
                prev = res + 1 if n_iter == 0 else prev
                if res <= 1e-9:
                    break
                rc, it, fr = dev.solve(B.UNSTEADY, 1, 2, tol, 100000)
                if rc != 0:
                    raise SystemExit(f"solve failed rc={rc} in time step {state['n'] + 1}, Re stage {cur:g}, Newton iteration {n_iter}, after {solves} solves "
                                     f"({outer} outer iterations) of this step: {dev.last_error()}")
                outer += it
                solves += 1
                if it == 0:
                    break
                dev.save_eval_point()
                alpha = 1.0
                while alpha > 1e-12:
                    dev.update(alpha)
                    res = dev.assemble(B.MODE_UNSTEADY_NEWTON, False, nu, dt)
                    assemblies += 1
                    if res <= prev:
                        break
                    alpha *= 0.1
                prev = res
                n_iter += 1
            cur += 10.0
        state["apply_first"] = False
        state["n"] += 1
        drag, lift = dev.lift_drag(nu)
        log.update(outer=outer, solves=solves, assemblies=assemblies, drag=drag, lift=lift)

    for _ in range(args.warmup):
        time_step()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.synchronize(); torch.cuda.synchronize()
    l0 = dev.stat("KERNEL_LAUNCHES")
    e0.record(stream)
    for _ in range(args.steps):
        time_step()
    e1.record(stream)
    dev.synchronize(); torch.cuda.synchronize()
    val = e0.elapsed_time(e1) / 1e3 / args.steps
    launches = dev.stat("KERNEL_LAUNCHES") - l0
    clocks = sampler.stop()
    U_avg = 2 * 0.3 / 3
    emit({"metric": "s per time step", "value": val, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3,
          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "reference mesh (lab_new/mesh/new_mesh.msh) or generated channel mesh, zero initial state",
          "config": {"workload": f"NSSolver {mesh_flag} -r 100 -T 8,0.01 -s 1 -t {tol:g} -p 2: time steps {args.warmup + 1}..{args.warmup + args.steps} (10 Reynolds stages per step)",
                     "cells": d.ncells, "dofs": d.n, "outer_iterations_last_step": log["outer"], "solves_last_step": log["solves"],
                     "assemblies_last_step": log["assemblies"], "drag_coefficient": 2 * log["drag"] / (U_avg ** 2 * 0.1), "lift_coefficient": 2 * log["lift"] / (U_avg ** 2 * 0.1),
                     "l2_policy": "whole-step timing; the 117 k-DoF matrices (43 MB) fit L2, as they do in the reference configuration"},
          "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 * (log["assemblies"] + 2 * log["solves"]) + 16},
          "gpu_launches": launches, "clocks": clocks})
    dev.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nsx", choices=["nsx", "reference"])
    ap.add_argument("--workload", default="stationary", choices=["stationary", "unsteady"],
                    help="stationary: README configuration (default, the headline); unsteady: BASELINE config 3, seconds per time step")
    ap.add_argument("--mesh", default="300,100")
    ap.add_argument("--solver", type=int, default=1)
    ap.add_argument("--prec", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--ordering", type=int, default=None, help="ILU/SGS elimination order: 0 natural (as Ifpack), 1 multicolour over the owned range, 2 multicolour inside CTA-local blocks, 3 natural inside CTA-local blocks; default: the library's choice (2; 3 for the unsteady aSIMPLE)")
    ap.add_argument("--budget-s", type=float, default=780.0, help="wall-clock budget of the whole run; the step count shrinks to fit (0: off)")
    ap.add_argument("--cpu-sample-s", type=float, default=15.0, help="length of the capped CPU solve of the cpu_baseline / reference arm")
    ap.add_argument("--unsteady-mesh", default="gmsh", help="unsteady workload: 'gmsh' = the reference's new_mesh.msh (P2/P1), X,Y = generated Q3/Q2 mesh, tri:X,Y = generated P2/P1 mesh")
    ap.add_argument("--ortho", type=int, default=None, help="Gram-Schmidt variant (NSX_OPT_ORTHO); default: the library's")
    ap.add_argument("--cpu-outer-cap", type=int, default=2)
    ap.add_argument("--cpu-outer-total", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel-reps", type=int, default=50)
    args = ap.parse_args()
    t_process = time.perf_counter()
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter)
    # is routed to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)
    if args.workload == "unsteady":
        return run_unsteady(args, emit)

    import ctypes
    import torch
    from navier_stokes_solver_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks (python -m torch.distributed.run --nproc-per-node {args.gpus} ...), found WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, ny = parse_mesh(args.mesh)
    nu = 1.0 / 10.0   # first Reynolds stage of the continuation (NSSolverStationary.cpp:662-665)
    t0 = time.perf_counter()
    # the reference's scheme: the cells are partitioned, every rank owns a contiguous row range of each block
    g = B.Disc.generate(nx, ny, nranks=world)
    stream = torch.cuda.Stream(device=local_rank)   # the library runs on this stream, so that torch events bracket its work
    if world > 1:
        d = g.local(rank)
        ids = [B.Device.new_comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dev = B.Device(d, device_id=local_rank, ordering=args.ordering, ortho=args.ortho, comm_id=ids[0], stream=ctypes.c_void_p(stream.cuda_stream))
    else:
        d = g
        dev = B.Device(d, device_id=local_rank, ordering=args.ordering, ortho=args.ortho, stream=ctypes.c_void_p(stream.cuda_stream))
    setup_s = time.perf_counter() - t0

    n = dev.n   # owned entries of this rank
    pinned_in = torch.zeros(2 * n, dtype=torch.float64).pin_memory()
    pinned_out = torch.zeros(n, dtype=torch.float64).pin_memory()
    stats = {}

    def step(events=None):
        """One Newton iteration through the C ABI with HOST buffers: state in, new solution out."""
        if events:
            events[0].record(stream)
        dev.upload_ptr(B.VEC_SOLUTION, pinned_in.data_ptr())
        dev.upload_ptr(B.VEC_DELTA, pinned_in.data_ptr() + 8 * n)
        if events:
            events[1].record(stream)
        r0 = dev.assemble(B.MODE_STOKES, True, nu)
        rc, it, fr = dev.solve(B.STATIONARY, args.solver, args.prec, args.tol, 20000)
        if rc != 0:
            raise SystemExit(f"solve failed rc={rc} it={it} res={fr}: {dev.last_error()}")
        dev.save_eval_point()
        dev.update(1.0)
        r1 = dev.assemble(B.MODE_STOKES, False, nu)
        if events:
            events[2].record(stream)
        dev.download_ptr(B.VEC_SOLUTION, pinned_out.data_ptr())
        if events:
            events[3].record(stream)
        stats.update(outer=it, inner_F=dev.stat("INNER_F"), inner_S=dev.stat("INNER_S"), applies=dev.stat("PRECOND_APPLIES"),
                     r0=r0, r1=r1, final_res=fr)

    def barrier():
        if world > 1:
            dist.barrier()
        dev.synchronize()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return float(tt.item())

    # budget guard: time the first warm-up step, then fit W + K steps, the kernel timings and the CPU sample into --budget-s by
    # lowering the step COUNT (never the work inside a step); the line reports the counts actually run
    warmup, steps = args.warmup, args.steps
    guard = None
    if warmup + steps > 0:
        barrier()
        t1 = time.perf_counter()
        step()
        barrier()
        t1 = max_over_ranks(time.perf_counter() - t1)
        done = 1
        if args.budget_s > 0:
            reserve = 40.0 + (0.0 if (args.no_cpu or world > 1) else 2.5 * args.cpu_sample_s + 10.0)
            left = args.budget_s - max_over_ranks(time.perf_counter() - t_process) - reserve
            afford = int(left / max(t1, 1e-3))   # further steps that fit
            want = max(0, warmup - 1) + steps
            if afford < want:
                w2 = min(warmup, 3)
                k2 = max(1, afford - max(0, w2 - 1))
                if afford < max(0, w2 - 1) + 1:
                    w2, k2 = max(1, min(w2, afford)), 1
                guard = {"budget_s": args.budget_s, "requested_steps": steps, "requested_warmup": warmup, "first_step_s": t1,
                         "note": "step count lowered to fit the wall-clock budget; every step still does the full work"}
                warmup, steps = w2, k2
        for _ in range(max(0, warmup - done)):
            step()
        if warmup == 0:
            warmup = 1   # the timed first step was a warm-up in effect
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # timed region: K steps between barriers, CUDA events on the library's stream; the host<->device copies are
    # bracketed by their own events so that the device-resident time (value) is the same steps minus the copies
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = dev.stat("KERNEL_LAUNCHES")
    spmv0 = dev.stat("SPMV_CALLS")
    t_wall = time.perf_counter()
    e_begin.record(stream)
    for k in range(steps):
        step(evs[k])
    e_end.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = dev.stat("KERNEL_LAUNCHES") - l0
    spmv_calls = dev.stat("SPMV_CALLS") - spmv0
    total_ms = e_begin.elapsed_time(e_end)
    copy_ms = sum(e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e in evs)
    e2e_s = max_over_ranks(total_ms) / 1e3 / steps
    value = max_over_ranks(total_ms - copy_ms) / 1e3 / steps
    wall_s = max_over_ranks(t_wall) / steps
    launches_all = int(sum_over_ranks(float(launches)))
    h2d = int(sum_over_ranks(16.0 * n))
    d2h = int(sum_over_ranks(8.0 * n + 16))

    # kernel timings (CUDA events on the library's stream, L2 flushed between launches); rank 0's share at N > 1
    reps = args.kernel_reps
    rng = np.random.default_rng(42)
    dev.upload(B.VEC_TMP0, rng.uniform(-1, 1, n))
    dev.set_time_params(B.MODE_NEWTON, 1.0 / 90.0)
    gstate = B.synthetic_state(g, 1234)
    dev.upload(B.VEC_SOLUTION, d.scatter_owned(gstate, g.n_u) if world > 1 else gstate)
    # the matrix kernels first, on the matrices of the last step (the views of F follow the VALUES: DESIGN.md section 3); the assembly
    # timing, which overwrites them with a Newton-branch matrix, last.  Same order on every rank (the SpMVs import ghosts).
    names = [("block_spmv", 0), ("spmv_F", 1), ("sgs_F", 5), ("ilu_apply_F", 6), ("ilu_factor_F", 7), ("dot", 3), ("axpy", 4), ("assembly_newton", 2)]
    for _, w in names:
        dev.time_kernel(w, 5, True)
    k_ms = {name: dev.time_kernel(w, reps, True) for name, w in names}
    # kernels whose input (330-510 MB) exceeds the 126 MB L2: also `reps` launches back to back inside ONE event pair, which
    # leaves the ~5 us of per-launch event / launch overhead out of a ~100 us figure.  The roofline figures use this one.
    big = [x for x in names if x[0] in ("block_spmv", "spmv_F", "sgs_F")]
    k_b2b = {name: dev.time_kernel(w, reps, 2) for name, w in big}
    k_flushed = dict(k_ms)
    if g.n > 300000:   # only when the matrices really exceed L2
        k_ms.update(k_b2b)
    # FP64 FMA peak of this GPU, measured in the same run (the denominator of the assembly's FP64 utilisation)
    dev.time_kernel(20, 2, False)
    fp64_ms = dev.time_kernel(20, 5, False)
    n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp64_peak_tflops = n_sms * 8 * 256 * 8192 * 8 * 2 / (fp64_ms * 1e-3) / 1e12
    clocks = sampler.stop() if rank == 0 else None

    nnz = {b: dev.nnz(b) for b in (B.BLOCK_F, B.BLOCK_BT, B.BLOCK_B, B.BLOCK_MP)}
    sb = dev.sweep_blocks(B.BLOCK_F)
    sweep_blocks = len(sb[0]) - 1 if sb else 0
    nnz_j = nnz[B.BLOCK_F] + nnz[B.BLOCK_BT] + nnz[B.BLOCK_B]
    peak, peak_kind = load_peaks()
    n_u_own, ncells_own = dev.n_u, d.ncells
    bts = spmv_bytes(nnz_j, n)
    spmv_gbs = bts / (k_ms["block_spmv"] * 1e-3) / 1e9
    spmv_f_bytes = 12 * nnz[B.BLOCK_F] + 20 * n_u_own
    # one SGS application = a lower and an upper sweep over the rank-local block: every value + column once, row pointers,
    # x in, intermediate out + in, y out
    sgs_bytes = 12 * nnz[B.BLOCK_F] + 8 * (n_u_own + 1) + 32 * n_u_own
    sgs_gbs = sgs_bytes / (k_ms["sgs_F"] * 1e-3) / 1e9
    asm_bytes = 8 * (nnz_j + nnz[B.BLOCK_MP]) + 16 * n + ncells_own * (64 + 4 * 41)
    kernels = {
        "block_spmv": {"ms": k_ms["block_spmv"], "GBps": spmv_gbs, "frac_hbm": spmv_gbs / peak, "algorithmic_bytes": bts},
        "spmv_F": {"ms": k_ms["spmv_F"], "GBps": spmv_f_bytes / (k_ms["spmv_F"] * 1e-3) / 1e9,
                   "frac_hbm": spmv_f_bytes / (k_ms["spmv_F"] * 1e-3) / 1e9 / peak, "algorithmic_bytes": spmv_f_bytes,
                   "stored_bytes": dev.stat("SPMV_BYTES_F") + 20 * n_u_own, "stored_GBps": (dev.stat("SPMV_BYTES_F") + 20 * n_u_own) / (k_ms["spmv_F"] * 1e-3) / 1e9},
        "assembly_newton": {"ms": k_ms["assembly_newton"], "GFLOPs_fp64": 1.206e5 * ncells_own / (k_ms["assembly_newton"] * 1e-3) / 1e9,
                            "fp64_peak_measured_TFLOPs": fp64_peak_tflops,
                            "fp64_utilisation": 1.206e5 * ncells_own / (k_ms["assembly_newton"] * 1e-3) / 1e12 / fp64_peak_tflops,
                            "GBps_min_bytes": asm_bytes / (k_ms["assembly_newton"] * 1e-3) / 1e9},
        "dot": {"ms": k_ms["dot"], "GBps": 16 * n / (k_ms["dot"] * 1e-3) / 1e9},
        "axpy": {"ms": k_ms["axpy"], "GBps": 24 * n / (k_ms["axpy"] * 1e-3) / 1e9},
        "sgs_F": {"ms": k_ms["sgs_F"], "levels": dev.stat("LEVELS_F"), "GBps": sgs_gbs, "frac_hbm": sgs_gbs / peak, "algorithmic_bytes": sgs_bytes,
                  "stored_bytes": dev.stat("SWEEP_BYTES_F"), "stored_GBps": dev.stat("SWEEP_BYTES_F") / (k_ms["sgs_F"] * 1e-3) / 1e9, "view_of_F": dev.stat("F_DECOUPLED")},
        "ilu_apply_F": {"ms": k_ms["ilu_apply_F"]},
        "ilu_factor_F": {"ms": k_ms["ilu_factor_F"]},
    }
    # share of the step spent in the two candidates for "dominant kernel" (launch counts x launch time)
    inner_f = stats["inner_F"]
    share = {"sgs_F": inner_f * k_ms["sgs_F"] * 1e-3 / value if args.prec == 0 else 0.0,
             "spmv_F": inner_f * k_ms["spmv_F"] * 1e-3 / value,
             "block_spmv": stats["outer"] * k_ms["block_spmv"] * 1e-3 / value}
    for k, v in share.items():
        kernels[k]["est_share_of_step"] = v
        kernels[k]["ms_l2_flushed_single_launch"] = k_flushed[k]
        kernels[k]["ms_back_to_back"] = k_b2b[k]
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` capture of
    # this configuration on one B200 (profiles/r01_prof_r1_top_kernels.md); null for any other mesh / partition
    ncu_traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written from an `ncu --set full` capture by tools/ncu_summary.py --traffic
    if os.path.exists(tp) and world == 1:
        with open(tp) as f:
            rec = json.load(f)
        if rec.get("mesh") == args.mesh and rec.get("ordering") == (2 if args.ordering is None else args.ordering):
            ncu_traffic = rec.get("bytes_per_launch", {})
    dom = max(share, key=share.get)
    sweep_kernel = {0: "k_tri_level / k_tri_chain", 1: "k_sweep_phased<SGS>", 2: "k_sweep_block<SGS>", 3: "k_sweep_block<SGS>"}[2 if args.ordering is None else args.ordering]
    dom_kernel = {"sgs_F": sweep_kernel + " (symmetric Gauss-Seidel sweeps on F, inner preconditioner of the inner FGMRES)",
                  "spmv_F": "k_spmv_tma (F SpMV of the inner FGMRES)", "block_spmv": "k_spmv_tma (Jacobian block SpMV)"}[dom]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, nx, ny), "cells": g.ncells, "dofs": g.n, "nnz_J_rank0": nnz_j,
                   "partition": f"{world} strips of cells, owned rows per rank (rank 0: {n} dofs)" if world > 1 else "one rank",
                   "elimination_order": ["natural (Ifpack)", "multicolour over the owned range", f"multicolour inside {sweep_blocks} CTA-local blocks (Ifpack overlap 0 at one rank per block)",
                                         f"natural inside {sweep_blocks} CTA-local blocks"][2 if args.ordering is None else args.ordering],
                   "ortho_option": args.ortho, "budget_guard": guard,
                   "outer_iterations": stats["outer"], "inner_F_iterations": stats["inner_F"], "inner_Mp_or_S_iterations": stats["inner_S"],
                   "final_residual": stats["final_res"],
                   "l2_policy": "step: working set (0.5 GB matrix + 60 Krylov vectors) exceeds the 126 MB L2; kernel timings: SpMV / sweep inputs (330-510 MB) exceed L2 and are timed back to back in one event pair (the L2-flushed single-launch times are listed beside them), vector kernels flush L2 (512 MiB) before every launch",
                   "setup_s": setup_s, "wall_s_per_step": wall_s},
        "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_all,
        "spmv_launches_rank0": spmv_calls,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                     "frac": kernels[dom]["GBps"] / peak, "traffic": ncu_traffic.get(dom), "peak_kind": peak_kind,
                     "algorithmic_bytes": kernels[dom].get("algorithmic_bytes", spmv_f_bytes), "est_share_of_step": share[dom],
                     "stored_bytes": kernels[dom].get("stored_bytes"), "frac_of_stored_bytes": (kernels[dom]["stored_GBps"] / peak) if kernels[dom].get("stored_GBps") else None,
                     "note": "achieved = SURVEY 8(d) bytes over the FULL pattern of F per launch / launch time; in the Stokes-type branches the kernel streams an exact, smaller view of F (DESIGN.md section 3), so achieved can exceed the HBM peak: stored_bytes is what it really moves"},
        "spmv_roofline": {"kernel": "k_spmv_tma (Jacobian block SpMV)", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                          "algorithmic_bytes": bts, "traffic": ncu_traffic.get("block_spmv")},
        "kernels": kernels,
    }
    if rank == 0 and not args.no_cpu and world == 1:
        smp = CpuSampler(nx, ny, args.solver, args.prec, args.tol, nu).sample(args.cpu_outer_cap, target_s=args.cpu_sample_s)
        fixture = outer_total_fixture(args.mesh, args.solver, args.prec, smp["threads"])
        scale, total, how = extrapolation(smp, fixture, args.cpu_outer_total, stats["outer"], "this run's GPU solve")
        extrapolated = not smp["converged"]
        cpu_val = 2 * smp["t_asm"] + smp["t_solve"] * scale
        cores = smp["threads"]
        line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": extrapolated,
                                "measured_s": 2 * smp["t_asm"] + smp["t_solve"], "outer_done": smp["outer_done"], "outer_total": total, "scale": scale,
                                "sample": f"oracle port, mesh cut into {cores} rank-local blocks on {cores} host threads (as mpirun -n {cores}): 2 full assemblies ({smp['t_asm']:.2f} s each) + solve " +
                                          (f"capped at {smp['outer_done']} outer iterations ({smp['t_solve']:.2f} s measured), extrapolated {how}"
                                           if extrapolated else f"run to convergence ({smp['outer_done']} outer iterations, {smp['t_solve']:.2f} s)")}
    if rank == 0:
        emit(line)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
