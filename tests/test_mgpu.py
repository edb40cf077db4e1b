"""Row-partitioned runs over NCCL on the GPUs of one box (needs >= 2 GPUs: `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

import nsxlib as N

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("elem,prec", [("tri", 0), ("quad", 1), ("quad", 0), ("quad", 2), ("tri", 2)])
def test_two_gpu_run_matches_global_oracle(elem, prec):
    """blockDiagonal (SGS), blockTriangular (AMG + ILU) and aSIMPLE on two ranks.  aSIMPLE: the ghost rows of Bt are assembled
    redundantly from the ghost-layer cells, S = B diag(F)^-1 Bt is formed for the rank-local block (ILU) and applied as
    B (diag(F)^-1 (Bt x)) with two ghost imports in the CG solve."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    _run_worker(elem, prec)


@pytest.mark.parametrize("elem,prec", [("quad", 0), ("tri", 2)])
def test_two_gpu_run_with_library_defaults(elem, prec):
    """The same two-rank runs with the library's defaults (CTA-local sweep blocks, node view of F in the Stokes branch with ghost
    pairs and the node-layout ghost import, device-driven inner FGMRES with one reduction per iteration) against the global oracle in
    its natural order: another preconditioner, the same converged increment (1e-8), residual and forces."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    _run_worker(elem, prec, "defaults")


def _run_worker(elem, prec, *extra):
    worker = os.path.join(N.ROOT, "tests", "mgpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", worker, elem, str(prec), *extra], capture_output=True, text=True, timeout=200)
    print(r.stdout[-2000:])
    logdir = os.path.join(N.ROOT, "gpurun_out")
    if os.path.isdir(logdir):
        with open(os.path.join(logdir, f"mgpu_{elem}_{prec}{'_' + extra[0] if extra else ''}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_WORKER_OK" in r.stdout


def test_two_process_executable_matches_one_process(tmp_path):
    """StationaryNSSolver as two processes (one per GPU, launcher environment + NCCL id file) against the same run on one
    GPU: same Newton history length, drag coefficient to 1e-6."""
    import re
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(N.ROOT, "navier_stokes_solver_b200", "apps", "StationaryNSSolver")
    args = ["-m", "24,10", "-r", "10", "-s", "1", "-p", "0", "-t", "1e-10"]
    env = dict(os.environ, NSX_NO_OUTPUT="1")
    one = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300, env=env, cwd=tmp_path)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29743", exe, *args], capture_output=True, text=True, timeout=300, env=env, cwd=tmp_path)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-3000:]
    cd1 = float(re.findall(r"Drag coefficient: ([0-9.e+-]+)", one.stdout)[-1])
    cd2 = float(re.findall(r"Drag coefficient: ([0-9.e+-]+)", two.stdout)[-1])
    it1 = [int(x) for x in re.findall(r"   (\d+) solver iterations", one.stdout)]
    it2 = [int(x) for x in re.findall(r"   (\d+) solver iterations", two.stdout)]
    print("Krylov iterations 1 GPU ", it1[:20])
    print("Krylov iterations 2 GPUs", it2[:20])
    assert abs(cd1 - cd2) <= 2e-6 * abs(cd1)
    assert two.stdout.count("CONFIGURATION PARAMETERS") == 1     # rank 0 alone prints
