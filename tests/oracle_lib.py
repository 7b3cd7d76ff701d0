"""Loads the CPU oracle (oracle/_build/liboracle.so), building it on first use.  Tests only."""
import ctypes as C
import os
import subprocess

from maestro_b200 import Operators

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_cache = {}


def build(target="_build/liboracle.so"):
    subprocess.check_call(["make", "-s", "-C", ODIR, target])
    return os.path.join(ODIR, target)


def load(debug=False):
    key = "dbg" if debug else "opt"
    if key not in _cache:
        path = build("_build/liboracle_dbg.so" if debug else "_build/liboracle.so")
        lib = C.CDLL(path)
        ops = Operators(lib, "mo_")
        lib.mo_test_advect.argtypes = [C.c_int] * 5 + [C.c_double] * 2 + [C.POINTER(C.c_double)] * 3
        lib.mo_test_advect.restype = C.c_int
        lib.mo_test_advect_ex.argtypes = [C.c_int] * 5 + [C.c_double] * 2 + [C.POINTER(C.c_double)] * 5
        lib.mo_test_advect_ex.restype = C.c_int
        lib.mo_cell_to_edge.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        ops.raw = lib
        _cache[key] = ops
    return _cache[key]


def test_advect(ops, dm, n, ppm_type, bds_type, direction, cfl=0.7, stop_time=1.0, want_rho=False):
    import numpy as np

    a, r = C.c_double(), C.c_double()
    rho = np.zeros((n,) * dm) if want_rho else None
    rp = rho.ctypes.data_as(C.POINTER(C.c_double)) if want_rho else None
    rc = ops.raw.mo_test_advect(dm, n, ppm_type, bds_type, direction, cfl, stop_time, C.byref(a), C.byref(r), rp)
    if rc != 0:
        raise RuntimeError(ops.raw.mo_last_error().decode())
    return (a.value, r.value, rho) if want_rho else (a.value, r.value)


def test_advect_fields(ops, dm, n, ppm_type, direction=1, cfl=0.7, stop_time=1.0, W=-1.0, floor=1e-10, cutoff=1e-10,
                       species_pred_type=1, fixed_dt=0, slope_order=4, bds_type=0):
    """The restated unit test with its unrecorded parameters exposed (oracle/mo_driver.cpp test_advect_run_ex);
    returns the initial and the final density on the valid cells."""
    import numpy as np

    P = C.POINTER(C.c_double)
    opt = (C.c_double * 6)(W, floor, cutoff, float(species_pred_type), float(fixed_dt), float(slope_order))
    a, r = C.c_double(), C.c_double()
    rf, ri = np.zeros((n,) * dm), np.zeros((n,) * dm)
    rc = ops.raw.mo_test_advect_ex(dm, n, ppm_type, bds_type, direction, cfl, stop_time, opt, C.byref(a), C.byref(r),
                                   rf.ctypes.data_as(P), ri.ctypes.data_as(P))
    if rc != 0:
        raise RuntimeError(ops.raw.mo_last_error().decode())
    return ri, rf
