"""The C-ABI boundary: libnsx.so loads without a GPU, exports every symbol include/nsx.h and
include/nsx_host.h declare, and refuses to create a context when no CUDA device exists (the product
has no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import nsxlib as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L, H = N.nsx(), N.nsx_host()
    names, host_names = declared("nsx.h"), declared("nsx_host.h")
    assert len(names) > 40 and len(host_names) >= 8
    missing = [n for n in names if not hasattr(L, n)] + [n for n in host_names if not hasattr(H, n)]
    assert not missing, missing
    assert sorted(N.NSX_SYMBOLS) == names
    assert sorted(N.NSX_HOST_SYMBOLS) == host_names
    # the host set-up stand-in is its own library: the CPU legs of bench.py and the oracle tests never map the CUDA library
    import subprocess
    out = subprocess.run(["ldd", N.LIBNSX_HOST], stdout=subprocess.PIPE, text=True).stdout
    assert "libcudart" not in out and "libnsx.so" not in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = N.nsx().nsx_create(0, 1, 0, None, C.byref(h))
    assert rc == N.NSX_E_CUDA and not h.value
    with pytest.raises(N.NsxError):
        N.Device(N.Disc.generate(6, 3))


def test_bad_arguments_without_a_context():
    L = N.nsx()
    assert L.nsx_create(3, 2, 0, None, C.byref(C.c_void_p())) == N.NSX_E_BADARG
    assert L.nsx_destroy(None) == N.NSX_OK
    assert L.nsx_assemble(None, 0, 0, 0.1, 0.0, 1.0, None) == N.NSX_E_BADARG
    assert L.nsx_last_error(None) == b"null context"


def test_product_never_touches_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may load oracle/ : the package sources must not mention it"""
    pkg = os.path.join(ROOT, "navier_stokes_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and f != "smoke.py":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "pyoracle" not in txt and "orc_" not in txt, f
    import subprocess
    out = subprocess.run(["ldd", N.LIBNSX], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out
