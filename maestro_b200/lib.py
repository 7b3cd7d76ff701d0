"""Loader of the CUDA C-ABI library (maestro_b200/lib/libmaestro_b200.so, built in-tree for sm_100a).

There is no CPU path: if the library is missing, or no CUDA device is present when `init()` is
called, this raises -- it never falls back to the oracle or to PyTorch ops.
"""
import ctypes as C
import os

from . import abi
from .operators import MaestroError, Operators

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmaestro_b200.so")

_lib = None
_ops = None


def load():
    """dlopen the library and declare every prototype (works without a GPU; no compute is done)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MaestroError(
                "CUDA library %s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        abi.declare(_lib, "mgpu_")
        for name, (res, args) in list(abi.LIFECYCLE.items()) + list(abi.MULTIBOX.items()):
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def init(device=0, use_torch_stream=False):
    """Bind this process to one GPU.  With use_torch_stream the library launches on torch's current
    stream so torch.cuda.Event timing brackets its kernels."""
    lib = load()
    if lib.mgpu_init(int(device)) != 0:
        raise MaestroError(lib.mgpu_last_error().decode())
    if use_torch_stream:
        import torch

        s = torch.cuda.current_stream(device).cuda_stream
        if lib.mgpu_set_stream(C.c_void_p(s)) != 0:
            raise MaestroError(lib.mgpu_last_error().decode())
    return ops()


def ops():
    global _ops
    if _ops is None:
        _ops = Operators(load(), "mgpu_")
    return _ops


def synchronize():
    if load().mgpu_synchronize() != 0:
        raise MaestroError(_lib.mgpu_last_error().decode())


def set_option(key, value):
    if load().mgpu_set_option(key.encode(), int(value)) != 0:
        raise MaestroError(_lib.mgpu_last_error().decode())


def profile(on=True):
    if load().mgpu_profile(1 if on else 0) != 0:
        raise MaestroError(_lib.mgpu_last_error().decode())


PROFILE_TAGS = ["edge_cell_states", "edge_simh", "edge_transverse", "edge_final", "flux", "update", "fill", "glue",
                "fused_edge", "velpred", "bds", "halo"]


def profile_get():
    """{kernel class: (total ms, launches)} since profile(True)"""
    out = {}
    for t, name in enumerate(PROFILE_TAGS):
        ms, n = C.c_double(), C.c_long()
        if load().mgpu_profile_get(t, C.byref(ms), C.byref(n)) != 0:
            raise MaestroError(_lib.mgpu_last_error().decode())
        if n.value:
            out[name] = (ms.value, n.value)
    return out


def launch_count(reset=False):
    return int(load().mgpu_launch_count(1 if reset else 0))


def copy_bytes(reset=False):
    """(host->device, device->host) bytes moved by host-pointer calls since the last reset"""
    a, b = C.c_long(), C.c_long()
    load().mgpu_copy_bytes(C.byref(a), C.byref(b), 1 if reset else 0)
    return a.value, b.value


def finalize():
    if _lib is not None:
        _lib.mgpu_finalize()


# ---- residency registry (include/maestro_b200.h): host fabs that keep a device mirror between calls -------------------
def _chk(rc):
    if rc != 0:
        raise RuntimeError(load().mgpu_last_error().decode())


def register(fab, pin=True):
    """keep a device mirror of a host Fab; host-pointer calls then move only its stale components"""
    _chk(load().mgpu_register(C.c_void_p(fab.ptr), fab.a.size, 1 if pin else 0))


def unregister(fab):
    _chk(load().mgpu_unregister(C.c_void_p(fab.ptr)))


def invalidate(fab, comp0=0, ncomp=-1):
    """the host wrote components comp0 .. comp0+ncomp-1 (0-based; all by default)"""
    _chk(load().mgpu_invalidate(C.c_void_p(fab.ptr), comp0, ncomp))


def download(fab, comp0=0, ncomp=-1):
    """bring the host copy up to date with what the device wrote"""
    _chk(load().mgpu_download(C.c_void_p(fab.ptr), comp0, ncomp))


def upload(fab, comp0=0, ncomp=-1):
    _chk(load().mgpu_upload(C.c_void_p(fab.ptr), comp0, ncomp))


# ---- several boxes per rank (multifab of nfabs boxes): the CUDA library's own entry points -----------------------------
def _fab_array(fabs):
    return (abi.mgpu_fab * len(fabs))(*[f.cfab() for f in fabs])


def _fab_arrays(per_dir):
    keep = [_fab_array(fs) for fs in per_dir]
    return (abi.F_ * len(per_dir))(*[C.cast(k, abi.F_) for k in keep]), keep


def fill_boundary_mf(p, fabs, scomp, bccomp, ncomp, adv_bc, pmask):
    """multifab_fill_boundary + multifab_physbc over the boxes `fabs` of one rank; adv_bc = the domain's table"""
    from .fab import as_int_p

    bc, bcp = as_int_p(adv_bc)
    pm, pmp = as_int_p(pmask)
    _chk(load().mgpu_fill_boundary_mf(C.byref(p), len(fabs), _fab_array(fabs), scomp, bccomp, ncomp, bcp, pmp))


def density_advance_mf(p, which_step, sold, snew, sedge, sflux, scal_force, umac, w0, etarhoflux, rho0_old, rho0_new,
                       p0_dummy, rho0_predicted_edge, adv_bc, pmask):
    """density_advance over lists of boxes (sedge / sflux / umac: dm lists of boxes); adv_bc = the domain's table"""
    from .fab import as_double_p, as_int_p

    keep = [as_double_p(x) for x in (w0, rho0_old, rho0_new, p0_dummy, rho0_predicted_edge)]
    bc, bcp = as_int_p(adv_bc)
    pm, pmp = as_int_p(pmask)
    se, k1 = _fab_arrays(sedge)
    sf, k2 = _fab_arrays(sflux)
    um, k3 = _fab_arrays(umac)
    _chk(load().mgpu_density_advance_mf(C.byref(p), which_step, len(sold), _fab_array(sold), _fab_array(snew), se, sf,
                                        _fab_array(scal_force), um, keep[0][1], _fab_array(etarhoflux), keep[1][1],
                                        keep[2][1], keep[3][1], keep[4][1], bcp, pmp))


def velocity_advance_mf(p, uold, unew, sold, rhohalf, umac, gpi, w0, w0_force, rho0_old, rho0_nph, grav_cell_old,
                        grav_cell_nph, sponge, adv_bc, pmask):
    """velocity_advance over lists of boxes (umac: dm lists of boxes); adv_bc = the domain's table"""
    from .fab import as_double_p, as_int_p

    keep = [as_double_p(x) for x in (w0, w0_force, rho0_old, rho0_nph, grav_cell_old, grav_cell_nph)]
    bc, bcp = as_int_p(adv_bc)
    pm, pmp = as_int_p(pmask)
    um, k1 = _fab_arrays(umac)
    _chk(load().mgpu_velocity_advance_mf(C.byref(p), len(uold), _fab_array(uold), _fab_array(unew), _fab_array(sold),
                                         _fab_array(rhohalf), um, _fab_array(gpi), *[k[1] for k in keep],
                                         _fab_array(sponge), bcp, pmp))


def advance_premac_mf(p, uold, sold, umac, gpi, w0, w0_force, rho0_old, grav_cell_old, adv_bc, phys_bc, pmask):
    """advance_premac over lists of boxes; adv_bc / phys_bc = the domain's tables"""
    from .fab import as_double_p, as_int_p

    keep = [as_double_p(x) for x in (w0, w0_force, rho0_old, grav_cell_old)]
    ints = [as_int_p(x) for x in (adv_bc, phys_bc, pmask)]
    um, k1 = _fab_arrays(umac)
    _chk(load().mgpu_advance_premac_mf(C.byref(p), len(uold), _fab_array(uold), _fab_array(sold), um, _fab_array(gpi),
                                       *[k[1] for k in keep], *[k[1] for k in ints]))


def enthalpy_advance_mf(p, which_step, sold, snew, sedge, sflux, scal_force, thermal, umac, w0, rho0_old, rhoh0_old,
                        rho0_new, rhoh0_new, p0_old, p0_new, psi, grav_old, grav_nph, adv_bc, pmask):
    """enthalpy_advance over lists of boxes (predict_rhoh / predict_rhohprime / predict_h)"""
    from .fab import as_double_p, as_int_p

    keep = [as_double_p(x) for x in (w0, rho0_old, rhoh0_old, rho0_new, rhoh0_new, p0_old, p0_new, psi, grav_old, grav_nph)]
    ints = [as_int_p(x) for x in (adv_bc, pmask)]
    se, k1 = _fab_arrays(sedge)
    sf, k2 = _fab_arrays(sflux)
    um, k3 = _fab_arrays(umac)
    _chk(load().mgpu_enthalpy_advance_mf(C.byref(p), which_step, len(sold), _fab_array(sold), _fab_array(snew), se, sf,
                                         _fab_array(scal_force), _fab_array(thermal), um, *[k[1] for k in keep],
                                         *[k[1] for k in ints]))
