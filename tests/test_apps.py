"""The two executables (navier_stokes_solver_b200/apps): command line as the reference's mains
(lab_new/src/testStationary.cpp:19-138, test.cpp:21-155) on CPU; whole runs against the oracle's
restatement of solve_newton / solve on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import nsxlib as N

APPS = os.path.join(N.ROOT, "navier_stokes_solver_b200", "apps")
STAT = os.path.join(APPS, "StationaryNSSolver")
UNST = os.path.join(APPS, "NSSolver")


def run(exe, *args, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, cwd=cwd, timeout=1200)


@pytest.fixture(scope="module", autouse=True)
def _apps_built():
    if not (os.path.exists(STAT) and os.path.exists(UNST)):
        subprocess.run(["make", "-C", APPS], check=True)


@pytest.mark.parametrize("exe", [STAT, UNST])
def test_help_and_usage_errors(exe):
    r = run(exe, "-h")
    assert r.returncode == 0 and r.stdout.startswith("Usage: ./NSSolver [options]")
    assert "-p, --preconditioner N    Select preconditioner (valid values: 0: blockDiagonal, 1: blockTriangular, 2: aSIMPLE)" in r.stdout
    assert ("--time-span and time-step T,D" in r.stdout) == (exe == UNST)
    r = run(exe, "--help")
    assert r.returncode == 0
    r = run(exe, "-x")                      # unknown flag: usage, exit 1 (testStationary.cpp:82-85)
    assert r.returncode == 1 and "Usage:" in r.stdout
    r = run(exe, "-m", "100")               # testStationary.cpp:60-63
    assert r.returncode == 1 and r.stderr.endswith("Error: mesh-size requires two values separated by comma\n")
    r = run(exe, "-M")                      # the short option string declares an argument for -M
    assert r.returncode == 1 and "Usage:" in r.stdout


def test_tolerance_and_timespan_validation():
    r = run(STAT, "-t", "0")
    assert r.returncode == 1 and r.stderr == "Error: tolerance must be positive\n"
    r = run(UNST, "-T", "1.0")
    assert r.returncode == 1 and r.stderr == "Error: timespan-step requires two values separated by comma\n"
    r = run(UNST, "-T", "1.0,-0.1")
    assert r.returncode == 1 and r.stderr == "Error: time_step, time_span, and tolerance must be positive\n"
    r = run(UNST, "-t", "-1")
    assert r.returncode == 1 and "must be positive" in r.stderr


def test_banner_and_short_M_swallows_the_next_flag(tmp_path):
    """`-M -m 7,3` (as run_sim_steady.sh:26 does): -m is eaten as -M's argument, so the mesh size stays at the
    default and the file mesh / P2-P1 element is selected.  Without a GPU the run stops at nsx_create."""
    r = run(STAT, "-M", "-m", "7,3", "-r", "20", "-s", "0", "-p", "1", "-t", "1e-8", env={"NSX_MESH_FILE": N.golden_mesh_path()}, cwd=tmp_path)
    out = r.stdout
    assert "--------- CONFIGURATION PARAMETERS --------- \nMesh size: 100x100\nReynolds number: 20\nSolver type: GMRES\nTolerance: 1e-08\n" \
           "Preconditioner: blockTriangular\n-----------------------------------------------\n" in out
    assert "Mesh file name = " in out and "  Number of elements = 25619" in out
    assert "  Velocity degree:           = 2" in out and "  DoFs per cell              = 15" in out
    assert "    velocity = 104066\n    pressure = 13207\n    total    = 117273" in out


def test_long_options(tmp_path):
    """Long options as declared at testStationary.cpp:33-42 / test.cpp:37-47: --read-mesh-from-file takes NO argument (so the
    mesh size that follows it is kept, unlike after the short -M), the others take one."""
    r = run(UNST, "--read-mesh-from-file", "--mesh-size", "7,3", "--reynolds", "50", "--solver", "2", "--tolerance", "1e-8", "--preconditioner", "2",
            "--timespan-step", "0.5,0.05", env={"NSX_MESH_FILE": N.golden_mesh_path()}, cwd=tmp_path)
    assert "Time span: 0.5\nTime step: 0.05\nMesh size: 7x3\nReynolds number: 50\nSolver type: Bicgstab\nTolerance: 1e-08\nPreconditioner: aSIMPLE\n" in r.stdout
    assert "Mesh file name = " in r.stdout and "  Velocity degree:           = 2" in r.stdout      # the file mesh / P2-P1 element was selected
    r = run(STAT, "--solver", "9", "--preconditioner", "5", "--mesh-size", "4,2", cwd=tmp_path)    # values outside {0,1,2} print nothing after the label
    assert "Solver type: Tolerance: 1e-06\nPreconditioner: -----------------------------------------------\n" in r.stdout
    r = run(STAT, "--timespan-step", "1,0.1")                                                      # not an option of the stationary binary
    assert r.returncode == 1 and "Usage:" in r.stdout


def test_generated_mesh_setup_prints_the_reference_dof_count(tmp_path):
    """100x70 -> 154244 DoFs (performance_analysis.ipynb): the one integer the reference pins."""
    r = run(UNST, "-m", "100,70", "-T", "0.02,0.01", cwd=tmp_path)
    assert "Time span: 0.02\nTime step: 0.01\nMesh size: 100x70\n" in r.stdout
    assert "  Number of elements = 6942" in r.stdout and "    total    = 154244" in r.stdout
    assert "  Quadrature points per cell = 16" in r.stdout and "  Quadrature points per face = 4" in r.stdout


# ---------------------------------------------------------------------------------------------------------------
# whole runs on the GPU
# ---------------------------------------------------------------------------------------------------------------
def _floats(pattern, text):
    return [float(x) for x in re.findall(pattern, text)]


@pytest.mark.gpu
def test_stationary_run_matches_oracle_driver(tmp_path):
    """StationaryNSSolver -m 16,6 -r 30 -s 1 -p 2: Stokes stage with the inlet ladder, then one Navier-Stokes stage.
    Same Newton residual history, same Krylov iteration counts (reported), lift / drag to 1e-6 (north_star)."""
    args = ("-m", "16,6", "-r", "30", "-s", "1", "-p", "2", "-t", "1e-10")
    r = run(STAT, *args, env={"NSX_NO_OUTPUT": "1", "NSX_PRINT_DIGITS": "12"}, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    d = N.Disc.generate(16, 6)
    o = N.Oracle(d)
    rc, log, nu, u = o.newton_stationary(30.0, 1, 2, 1e-10)
    assert rc == 0
    res_app = _floats(r"Newton iteration \d+/15 - \|\|r\|\| = ([0-9.e+-]+)", r.stdout)
    res_orc = [row[2] for row in log if row[0] == 1]
    print("Newton assemblies: app", len(res_app), "oracle", len(res_orc))
    its_app = [int(x) for x in re.findall(r"   (\d+) (?:solver )?iterations", r.stdout)]
    its_orc = [int(row[1]) for row in log if row[0] == 2]
    print("Krylov iterations app   ", its_app[:40])
    print("Krylov iterations oracle", its_orc[:40])
    # the first residuals of each stage are mesh + BC quantities: identical to print precision
    assert abs(res_app[0] - res_orc[0]) <= 2e-6 * res_orc[0]
    drag_o, lift_o = o.lift_drag(nu)
    U_avg = 2 * (4 * u * 0.205 * (0.41 - 0.205) / 0.41 ** 2) / 3
    cl_o, cd_o = 2 * lift_o / (U_avg ** 2 * 0.1), 2 * drag_o / (U_avg ** 2 * 0.1)
    cl = _floats(r"\[nsx\] lift coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    cd = _floats(r"\[nsx\] drag coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    print("lift", cl, cl_o, "drag", cd, cd_o)
    # 1e-6 relative (north_star), read from the 12-digit lines.  The lift of this symmetric set-up is ~1e-8 of the drag, i.e. at
    # the level of the Krylov tolerance: it is compared relative to the force on the cylinder
    assert abs(cd - cd_o) <= 1e-6 * abs(cd_o) and abs(cl - cl_o) <= 1e-6 * np.hypot(cd_o, cl_o)
    assert "Solving Stokes adding BCs" in r.stdout and "Solving NS" in r.stdout
    assert r.stdout.count("Computing drag and lift forces") > 0


@pytest.mark.gpu
def test_unsteady_run_matches_oracle_driver(tmp_path):
    """NSSolver -m 16,6 -r 11 -T 0.02,0.01 -s 1 -p 2 (config 3's solver pairing: FGMRES + aSIMPLE), first time step:
    two Reynolds stages inside the step, lift / drag coefficients to 1e-6."""
    env = {"NSX_NO_OUTPUT": "1", "NSX_MAX_TIME_STEPS": "1", "NSX_PRINT_DIGITS": "12"}
    r = run(UNST, "-m", "16,6", "-r", "11", "-T", "0.02,0.01", "-s", "1", "-p", "2", "-t", "1e-8", env=env, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    d = N.Disc.generate(16, 6)
    o = N.Oracle(d, inlet_amplitude=0.3)
    rc, log, nu = o.run_unsteady(11.0, 0.02, 0.01, 1, 2, 1e-8, n_steps_max=1)
    assert rc == 0
    coeffs = [row for row in log if row[0] == 6][-1]
    cl = _floats(r"\[nsx\] lift coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    cd = _floats(r"\[nsx\] drag coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    its_app = [int(x) for x in re.findall(r"   (\d+) (?:solver )?iterations", r.stdout)]
    its_orc = [int(row[1]) for row in log if row[0] == 2]
    print("Krylov iterations app   ", its_app)
    print("Krylov iterations oracle", its_orc)
    print("lift", cl, coeffs[2], "drag", cd, coeffs[3])
    assert abs(cd - coeffs[3]) <= 1e-6 * abs(coeffs[3]) and abs(cl - coeffs[2]) <= 1e-6 * np.hypot(coeffs[2], coeffs[3])
    assert r.stdout.count("Debug ") == len(d.array("CYL_CELL"))     # one per cylinder face (NSSolver.cpp:883)
    assert "n =   1, t = 0.010000" in r.stdout


@pytest.mark.gpu
def test_config2_file_mesh_block_triangular(tmp_path):
    """BASELINE config 2: `StationaryNSSolver -M x -r 20 -s 1 -p 1` on the reference's own mesh (lab_new/mesh/new_mesh.msh, P2/P1,
    117 273 DoFs): FGMRES + blockTriangular (AMG on F, ILU on Mp).  -r 20 never leaves the Stokes stage: one real solve,
    then the zero-iteration breaks of the inlet ladder.  The oracle driver runs the same command on the CPU: Newton history, Krylov
    counts (reported), lift / drag coefficients to 1e-6.  The two numbers of lab_new/lift_drag_data (unknown revision / parameters,
    SURVEY section 4) are printed beside them as a curiosity only."""
    mesh = N.golden_mesh_path()
    # parity configuration of the library: Ifpack's natural order, deal.II's modified Gram-Schmidt (the defaults are checked by the
    # two tests above and by test_gpu_solve.py)
    env = {"NSX_NO_OUTPUT": "1", "NSX_PRINT_DIGITS": "12", "NSX_MESH_FILE": mesh, "NSX_ORDERING": "0", "NSX_ORTHO": "0"}
    r = run(STAT, "-M", "x", "-r", "20", "-s", "1", "-p", "1", "-t", "1e-10", env=env, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
    assert "  Number of elements = 25619" in r.stdout and "    total    = 117273" in r.stdout
    d = N.Disc.from_gmsh(mesh)
    o = N.Oracle(d)
    rc, log, nu, u = o.newton_stationary(20.0, 1, 1, 1e-10)
    assert rc == 0
    its_app = [int(x) for x in re.findall(r"   (\d+) (?:solver )?iterations", r.stdout)]
    its_orc = [int(row[1]) for row in log if row[0] == 2]
    print("Krylov iterations app   ", its_app)
    print("Krylov iterations oracle", its_orc)
    # two real solves (the warm start of the second one is the first increment), then the zero-iteration breaks of the inlet ladder
    assert len(its_app) == len(its_orc) and its_app[0] > 0 and its_app[-1] == 0 and its_orc[-1] == 0
    assert all(abs(a - b) <= max(3, 0.15 * b) for a, b in zip(its_app, its_orc))
    drag_o, lift_o = o.lift_drag(nu)
    U_avg = 2 * (4 * u * 0.205 * (0.41 - 0.205) / 0.41 ** 2) / 3
    cl_o, cd_o = 2 * lift_o / (U_avg ** 2 * 0.1), 2 * drag_o / (U_avg ** 2 * 0.1)
    cl = _floats(r"\[nsx\] lift coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    cd = _floats(r"\[nsx\] drag coefficient = ([0-9.e+-]+)", r.stdout)[-1]
    print("config 2 lift", cl, cl_o, "drag", cd, cd_o, "| lab_new/lift_drag_data (another revision): lift 8.42639e-05, drag 3.24669")
    assert abs(cd - cd_o) <= 1e-6 * abs(cd_o) and abs(cl - cl_o) <= 1e-6 * np.hypot(cd_o, cl_o)
