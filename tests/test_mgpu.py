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


@pytest.mark.parametrize("elem,prec", [("tri", 0), ("quad", 1), ("quad", 0)])
def test_two_gpu_run_matches_global_oracle(elem, prec):
    """blockDiagonal (SGS) and blockTriangular (AMG + ILU) on two ranks; aSIMPLE on a partitioned system needs the ghost rows of
    Bt for the Schur product and answers NSX_E_STATE in this round."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    worker = os.path.join(N.ROOT, "tests", "mgpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", worker, elem, str(prec)], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    logdir = os.path.join(N.ROOT, "gpurun_out")
    if os.path.isdir(logdir):
        with open(os.path.join(logdir, f"mgpu_{elem}_{prec}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_WORKER_OK" in r.stdout
