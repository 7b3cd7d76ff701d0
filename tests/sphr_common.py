"""Shared inputs of the spherical-geometry tests (SURVEY config C5 in miniature): a cubic box whose centre is the
centre of the star, outlet on all sides, radial base state on dr = dx/drdxfac bins, w0mac / rho0mac built by the
path's own make_w0mac / make_s0mac."""
import numpy as np

from maestro_b200 import Fab, Geom, abi, face_fabs, make_adv_bc
from synth import make_state

OUTLET_3D = [[abi.OUTLET, abi.OUTLET]] * 3


def make_sphr_state(n=12, seed=4242, ops=None, **geom_kw):
    """state dict of synth.make_state (3-D, outlet BCs, spherical=1) + geometry + radial base state arrays +
    the face arrays the Fortran driver would build with make_w0mac / make_s0mac (built with `ops`, the oracle)"""
    shape = (n, n, n) if np.isscalar(n) else tuple(n)
    st = make_state(3, shape, phys_bc=OUTLET_3D, seed=seed)
    p = st["p"]
    p.spherical = 1
    rng = np.random.default_rng(seed + 1)
    g = Geom(p, **geom_kw)
    nr = g.nr_fine
    rc, re = g.r_cc_loc, g.r_edge_loc
    st["geom"] = g
    st["rad"] = dict(
        rho0_old=2.0 * np.exp(-(rc / 0.35) ** 2) + 0.1 + 0.01 * rng.uniform(-1, 1, nr),
        rho0_new=2.02 * np.exp(-(rc / 0.35) ** 2) + 0.1 + 0.01 * rng.uniform(-1, 1, nr),
        rhoh0_old=3.0 * np.exp(-(rc / 0.4) ** 2) + 0.2 + 0.01 * rng.uniform(-1, 1, nr),
        rhoh0_new=3.03 * np.exp(-(rc / 0.4) ** 2) + 0.2 + 0.01 * rng.uniform(-1, 1, nr),
        w0=0.3 * re * np.exp(-(re / 0.3) ** 2) + 0.002 * rng.uniform(-1, 1, nr + 1),
    )
    lo, hi = st["lo"], st["hi"]
    if ops is not None:
        def cart_of(arr, edge, vec, bccomp):
            c = Fab(lo, hi, 2, 3 if vec else 1, dm=3)
            ops.put_1d_array_on_cart(p, g, arr, c, edge, vec)
            # ghost cells: first-order extrapolation at the outlets (put_1d_array_on_cart fills them with
            # multifab_physbc on the foextrap component, fill_3d_data.f90:214-236)
            a = c.a
            for ax in (1, 2, 3):
                idx_lo = [slice(None)] * 4
                idx_hi = [slice(None)] * 4
                for gcell in range(2):
                    idx_lo[ax], idx_hi[ax] = gcell, a.shape[ax] - 1 - gcell
                    src_lo, src_hi = list(idx_lo), list(idx_hi)
                    src_lo[ax], src_hi[ax] = 2, a.shape[ax] - 3
                    a[tuple(idx_lo)] = a[tuple(src_lo)]
                    a[tuple(idx_hi)] = a[tuple(src_hi)]
            return c

        st["w0_cart"] = cart_of(st["rad"]["w0"], True, True, 1)
        st["w0mac"] = face_fabs(lo, hi, 1, 1, 3)
        ops.make_w0mac(p, g, st["rad"]["w0"], st["w0mac"], st["w0_cart"])
        for key in ("rho0_old", "rho0_new", "rhoh0_old", "rhoh0_new"):
            cart = cart_of(st["rad"][key], False, False, 1)
            st[key + "_cart"] = cart
            mac = face_fabs(lo, hi, 1, 1, 3)
            ops.make_s0mac(p, g, st["rad"][key], mac, cart)
            st[key.replace("0_", "0mac_")] = mac
    return st
