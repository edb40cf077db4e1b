"""nsx -- B200-native (sm_100a) assembly + Krylov hot path of a Taylor-Hood Navier-Stokes solver.

The product is the C-ABI shared library libnsx.so (include/nsx.h, include/nsx_host.h) and the two
host executables under apps/; this Python package is only the ctypes binding used by the tests,
bench.py and smoke()."""
from .binding import *  # noqa: F401,F403
from .binding import Device, Disc, NsxError, nsx  # noqa: F401
