"""Host-side loop logic of the executables on CPU: StationaryNSSolver / NSSolver are linked against a test-only shim that
implements the device C ABI over the oracle (tests/shim/nsx_over_oracle.cpp).  The Newton / Reynolds-continuation /
inlet-ladder / line-search / time loops of navier_stokes_solver_b200/apps are then one restatement of the reference's
control flow (NSSolverStationary.cpp:649-758, NSSolver.cpp:674-837) and the oracle's drivers (oracle_solve.inc) another;
with the same arithmetic underneath they must print the same history: Newton residuals, Krylov iteration counts,
line-search trials, lift and drag."""
import os
import re
import subprocess

import numpy as np
import pytest

import nsxlib as N

CSRC = os.path.join(N.ROOT, "navier_stokes_solver_b200", "csrc")
APPS = os.path.join(N.ROOT, "navier_stokes_solver_b200", "apps")
ORC = os.path.join(N.ROOT, "oracle")


@pytest.fixture(scope="module")
def cpu_apps(tmp_path_factory):
    t = str(tmp_path_factory.mktemp("shim"))
    gxx = "/usr/bin/g++"
    subprocess.run([gxx, "-O3", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", f"{t}/libnsx_shim.so",
                    os.path.join(N.ROOT, "tests", "shim", "nsx_over_oracle.cpp"), f"{CSRC}/hostsetup.cpp", f"{CSRC}/capi_host.cpp",
                    "-L", ORC, "-loracle", f"-Wl,-rpath,{ORC}"], check=True)
    for src, exe in (("stationary_main.cpp", "StationaryNSSolver_cpu"), ("unsteady_main.cpp", "NSSolver_cpu")):
        subprocess.run([gxx, "-O2", "-std=c++17", "-Wno-reorder", "-pthread", f"{APPS}/{src}", "-o", f"{t}/{exe}", "-L", t, "-lnsx_shim", f"-Wl,-rpath,{t}"], check=True)
    return t


def history(stdout):
    newton = [float(x) for x in re.findall(r"Newton iteration \d+/\d+ - \|\|r\|\| = ([0-9.e+-]+)", stdout)]
    krylov = [int(x) for x in re.findall(r"   (\d+) (?:solver )?iterations", stdout)]
    trials = [(float(a), float(r)) for a, r in re.findall(r"Evaluating alpha=([0-9.e+-]+), \|\|r\|\|=([0-9.e+-]+)", stdout)]
    return newton, krylov, trials


def close(a, b, rel=2e-6, floor=1e-7):
    """7 printed digits; residuals far below the Krylov tolerance's reach are rounding noise (they depend on e.g. whether the
    compiler contracts `eval + alpha * delta` into an FMA) and are compared against `floor` instead of their own size."""
    return len(a) == len(b) and all(abs(x - y) <= rel * max(abs(y), floor) for x, y in zip(a, b))


def test_stationary_loop_prints_the_oracle_history(cpu_apps, tmp_path):
    """-m 16,6 -r 30: Stokes stage with the inlet ladder 0.1 -> 1.0 (never converging by design), then one
    Navier-Stokes stage."""
    r = subprocess.run([f"{cpu_apps}/StationaryNSSolver_cpu", "-m", "16,6", "-r", "30", "-s", "1", "-p", "2", "-t", "1e-10"], capture_output=True,
                       text=True, timeout=900, cwd=tmp_path, env=dict(os.environ, NSX_NO_OUTPUT="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    newton, krylov, trials = history(r.stdout)
    o = N.Oracle(N.Disc.generate(16, 6))
    rc, log, nu, u = o.newton_stationary(30.0, 1, 2, 1e-10)
    assert rc == 0
    assert krylov == [int(row[1]) for row in log if row[0] == 2]
    assert close(newton, [row[2] for row in log if row[0] == 1])
    assert close([t[1] for t in trials], [row[2] for row in log if row[0] == 3]) and close([t[0] for t in trials], [row[1] for row in log if row[0] == 3])
    # stage banners: the inlet is imposed once, Stokes mode lasts for the whole first Reynolds stage
    assert r.stdout.count("Solving Stokes adding BCs") == 1
    assert r.stdout.count("Solving Stokes without adding BCs") == sum(1 for row in log if row[0] == 0 and row[3] == 1)
    assert r.stdout.count("Solving NS") == sum(1 for row in log if row[0] == 0 and row[3] == 2)
    drag, lift = o.lift_drag(nu)
    U_avg = 2 * (4 * u * 0.205 * (0.41 - 0.205) / 0.41 ** 2) / 3
    cd = float(re.findall(r"Drag coefficient: ([0-9.e+-]+)", r.stdout)[-1])
    assert abs(cd - 2 * drag / (U_avg ** 2 * 0.1)) <= 2e-6 * abs(cd)


def test_unsteady_loop_prints_the_oracle_history(cpu_apps, tmp_path):
    """-m 16,6 -r 21 -T 0.03,0.01: two time steps, three Reynolds stages inside each, inlet only in the first assembly."""
    r = subprocess.run([f"{cpu_apps}/NSSolver_cpu", "-m", "16,6", "-r", "21", "-T", "0.03,0.01", "-s", "1", "-p", "2", "-t", "1e-8"], capture_output=True,
                       text=True, timeout=900, cwd=tmp_path, env=dict(os.environ, NSX_NO_OUTPUT="1", NSX_MAX_TIME_STEPS="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    newton, krylov, trials = history(r.stdout)
    o = N.Oracle(N.Disc.generate(16, 6), inlet_amplitude=0.3)
    rc, log, nu = o.run_unsteady(21.0, 0.03, 0.01, 1, 2, 1e-8, n_steps_max=2)
    assert rc == 0
    assert krylov == [int(row[1]) for row in log if row[0] == 2]
    assert close(newton, [row[2] for row in log if row[0] == 1])
    assert close([t[1] for t in trials], [row[2] for row in log if row[0] == 3])
    coeffs = [row for row in log if row[0] == 6]
    cl = [float(x) for x in re.findall(r"Lift coefficient: ([0-9.e+-]+)", r.stdout)]
    cd = [float(x) for x in re.findall(r"Drag coefficient: ([0-9.e+-]+)", r.stdout)]
    assert len(cd) == len(coeffs) == 2
    for k, row in enumerate(coeffs):
        scale = np.hypot(row[2], row[3])
        assert abs(cd[k] - row[3]) <= 2e-6 * scale and abs(cl[k] - row[2]) <= 2e-6 * scale
    assert "n =   1, t = 0.010000" in r.stdout and "n =   2, t = 0.020000" in r.stdout


def test_bad_preconditioner_ends_like_the_reference(cpu_apps, tmp_path):
    """-p 7: std::invalid_argument at the first solve_system (NSSolverStationary.cpp:641-643), uncaught -> abort."""
    r = subprocess.run([f"{cpu_apps}/StationaryNSSolver_cpu", "-m", "8,4", "-p", "7"], capture_output=True, text=True, timeout=300, cwd=tmp_path,
                       env=dict(os.environ, NSX_NO_OUTPUT="1"))
    assert r.returncode != 0
    assert "Invalid preconditioner type. Use 0: blockDiagonal, 1: blockTriangular, 2: aSIMPLE." in r.stderr
    assert "Newton iteration 0/15" in r.stdout


def test_output_files_written_behind_the_solver(cpu_apps, tmp_path):
    """The VTU stand-in (NSSolverStationary.cpp:765-800, NSSolver.cpp:788-793) is written by a worker thread from a snapshot of
    the solution: same bytes as the synchronous writer (NSX_SYNC_OUTPUT=1), well-formed, one piece per rank with the owned cells,
    and the unsteady executable leaves one file per time step."""
    import xml.etree.ElementTree as ET
    outs = {}
    for mode in ("async", "sync"):
        d = tmp_path / mode
        d.mkdir()
        env = dict(os.environ)
        env.pop("NSX_NO_OUTPUT", None)
        if mode == "sync":
            env["NSX_SYNC_OUTPUT"] = "1"
        r = subprocess.run([f"{cpu_apps}/StationaryNSSolver_cpu", "-m", "8,4", "-r", "10", "-s", "1", "-p", "2", "-t", "1e-8"], capture_output=True,
                           text=True, timeout=900, cwd=d, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Output written to output-stokes" in r.stdout
        outs[mode] = (d / "output-stokes_0.0.vtu").read_bytes()
    assert outs["async"] == outs["sync"] and len(outs["sync"]) > 1000
    root = ET.fromstring(outs["async"])
    piece = root.find("UnstructuredGrid/Piece")
    disc = N.Disc.generate(8, 4)
    assert int(piece.get("NumberOfCells")) == disc.ncells and int(piece.get("NumberOfPoints")) == 4 * disc.ncells
    vel = np.array(piece.find("PointData/DataArray[@Name='velocity']").text.split(), dtype=float).reshape(-1, 3)
    assert vel.shape[0] == 4 * disc.ncells and np.abs(vel[:, 0]).max() > 0.01 and (vel[:, 2] == 0).all()
    d = tmp_path / "unsteady"
    d.mkdir()
    env = dict(os.environ, NSX_MAX_TIME_STEPS="2")
    env.pop("NSX_NO_OUTPUT", None)
    r = subprocess.run([f"{cpu_apps}/NSSolver_cpu", "-m", "8,4", "-r", "11", "-T", "0.03,0.01", "-s", "1", "-p", "2", "-t", "1e-6"], capture_output=True,
                       text=True, timeout=900, cwd=d, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted(p.name for p in d.glob("output_*.vtu"))
    assert len(files) >= 2, files
    for name in files:
        ET.fromstring((d / name).read_bytes())
