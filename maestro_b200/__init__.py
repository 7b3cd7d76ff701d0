"""maestro_b200 -- B200-native implementation of MAESTRO's advective hot path.

csrc/   hand-written fp64 CUDA kernels (sm_100a) + the C ABI (include/maestro_b200.h)
abi.py  ctypes mirror of the ABI
fab.py  miniature multifab (one box, Fortran layout) and BC tables
operators.py  host mirror of the reference's L3/L4 operators (make_edge_scal, mk_rhoX_flux, ...)
lib.py  loader of the CUDA library; no CPU fallback
"""
from . import abi  # noqa: F401
from .fab import Fab, Geom, face_fabs, make_adv_bc, make_params, nbc_comps  # noqa: F401
from .operators import MaestroError, Operators  # noqa: F401
from . import slab  # noqa: F401,E402
