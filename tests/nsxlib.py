"""Helpers shared by the tests: the product binding (navier_stokes_solver_b200.binding) and the CPU
oracle wrapper (oracle/pyoracle.py, test infrastructure only)."""
import gzip
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from navier_stokes_solver_b200.binding import *  # noqa: E402,F401,F403
from navier_stokes_solver_b200.binding import Device, Disc, NsxError, nsx, nsx_host, ptr, synthetic_state  # noqa: E402,F401
from oracle.pyoracle import Oracle, orc  # noqa: E402,F401


def golden_mesh_path():
    """tests/golden/new_mesh.msh.gz (the reference's lab_new/mesh/new_mesh.msh input) unpacked to a temp file."""
    src = os.path.join(ROOT, "tests", "golden", "new_mesh.msh.gz")
    dst = os.path.join(tempfile.gettempdir(), "nsx_new_mesh.msh")
    if not os.path.exists(dst):
        with gzip.open(src, "rb") as f, open(dst + ".tmp", "wb") as g:
            shutil.copyfileobj(f, g)
        os.replace(dst + ".tmp", dst)
    return dst


