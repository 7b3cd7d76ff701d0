"""Miniature multifab: one box, Fortran layout, host (numpy) or device (torch.cuda) storage.

Mirrors what FBoxLib hands the reference kernels (`dataptr`, `get_box`, `nghost`; Docs/architecture/
architecture.tex:617-700): a contiguous fp64 block `(lo-ng:hi+ng [+1 if nodal], ..., 1:nc)` with x
fastest and the component index slowest.  numpy/torch see it as a C-ordered `(nc, nz, ny, nx)` array.
"""
import ctypes as C

import numpy as np

from . import abi


class Fab:
    def __init__(self, lo, hi, ng, nc=1, nodal=(0, 0, 0), dm=3, device=None, fill=0.0):
        self.dm = dm
        self.lo = [int(lo[d]) if d < dm else 0 for d in range(3)]
        self.hi = [int(hi[d]) if d < dm else 0 for d in range(3)]
        self.ng = int(ng)
        self.nc = int(nc)
        self.nodal = [int(nodal[d]) if d < dm else 0 for d in range(3)]
        n = [self.hi[d] - self.lo[d] + 1 + 2 * self.ng + self.nodal[d] if d < dm else 1 for d in range(3)]
        self.shape = (self.nc, n[2], n[1], n[0])
        self.device = device
        if device is None:
            self.a = np.full(self.shape, fill, dtype=np.float64)
        else:
            import torch

            self.a = torch.full(self.shape, fill, dtype=torch.float64, device=device)

    # ---- views -------------------------------------------------------------------------------
    @property
    def ptr(self):
        return self.a.ctypes.data if self.device is None else self.a.data_ptr()

    def cfab(self):
        f = abi.mgpu_fab()
        f.ptr = self.ptr
        for d in range(3):
            f.lo[d], f.hi[d], f.nodal[d] = self.lo[d], self.hi[d], self.nodal[d]
        f.ng, f.nc = self.ng, self.nc
        return f

    def valid(self, comp=None):
        """View of the valid region (faces lo..hi+1 in the nodal direction), shape (nc, nz, ny, nx)."""
        sl = [slice(None)]
        for d in (2, 1, 0):
            if d < self.dm:
                sl.append(slice(self.ng, self.shape[3 - d] - self.ng))
            else:
                sl.append(slice(None))
        v = self.a[tuple(sl)]
        return v if comp is None else v[comp]

    def numpy(self):
        return self.a if self.device is None else self.a.cpu().numpy()

    def to(self, device):
        out = Fab(self.lo, self.hi, self.ng, self.nc, self.nodal, self.dm, device=device)
        if device is None:
            out.a[...] = self.numpy()
        else:
            import torch

            src = self.a if self.device is not None else torch.from_numpy(self.a)
            out.a.copy_(src)
        return out

    def clone(self):
        return self.to(self.device)


def face_fabs(lo, hi, ng, nc, dm, device=None, fill=0.0):
    """dm face-centred fabs (umac, sedge, sflux ...): fab d is nodal in direction d."""
    return [Fab(lo, hi, ng, nc, nodal=tuple(1 if q == d else 0 for q in range(3)), dm=dm, device=device, fill=fill)
            for d in range(dm)]


def make_params(dm, n=None, lo=None, hi=None, nspec=3, ntrac=1, **kw):
    """mgpu_params with the reference's defaults (Source/_parameters, variables.f90:100-124)."""
    p = abi.mgpu_params()
    p.dm = dm
    p.mem_space = abi.HOST
    p.ppm_type = 1
    p.bds_type = 0
    p.slope_order = 4
    p.ppm_trace_forces = 0
    p.species_pred_type = abi.PREDICT_RHOPRIME_AND_X
    p.enthalpy_pred_type = abi.PREDICT_RHOHPRIME
    p.spherical = 0
    p.evolve_base_state = 1
    p.do_sponge = 0
    p.do_eos_h_above_cutoff = 0
    p.nspec, p.ntrac = nspec, ntrac
    p.rho_comp, p.rhoh_comp, p.spec_comp = 1, 2, 3
    p.temp_comp = p.spec_comp + nspec
    p.pi_comp = p.temp_comp + 1
    p.trac_comp = p.pi_comp + 1
    p.nscal = nspec + ntrac + 4
    if n is not None:
        lo = [0, 0, 0]
        hi = [n[d] - 1 if d < dm else 0 for d in range(3)]
    for d in range(3):
        p.domlo[d] = lo[d] if d < dm else 0
        p.domhi[d] = hi[d] if d < dm else 0
        p.dx[d] = 1.0 / (p.domhi[0] - p.domlo[0] + 1)
    p.nr = p.domhi[dm - 1] - p.domlo[dm - 1] + 1
    p.dt = 0.0
    p.rel_eps = 0.0
    p.base_cutoff_density = 1.0e-10
    p.base_cutoff_density_coord = p.nr  # no cutoff inside the domain
    p.buoyancy_cutoff_factor = 5.0
    p.omega, p.sin_theta, p.cos_theta, p.rotation_radius = 0.0, 0.0, 1.0, 1.0e6
    for k, v in kw.items():
        if k == "dx":
            for d in range(3):
                p.dx[d] = v[d] if d < len(v) else v[-1]
        else:
            setattr(p, k, v)
    return p


def nbc_comps(p):
    """number of BC columns: dm velocities + nscal + press, foextrap, hoextrap (variables.f90:117-122)."""
    return p.dm + p.nscal + 3


def make_adv_bc(p, phys_bc):
    """adv_bc table from physical BCs, restating adv_bc_level_build (Source/define_bc_tower.f90:199-294).

    phys_bc[d][side] in {PERIODIC, INTERIOR, INLET, OUTLET, SYMMETRY, SLIP_WALL, NO_SLIP_WALL}.
    Returns int32 array indexed [bccomp-1, side, d] == Fortran adv_bc(d+1, side+1, bccomp).
    """
    dm = p.dm
    nbc = nbc_comps(p)
    bc = np.full((nbc, 2, dm), abi.INTERIOR, dtype=np.int32)
    press = dm + p.nscal + 1
    fo, ho = press + 1, press + 2
    for d in range(dm):
        for side in range(2):
            pb = phys_bc[d][side]
            col = bc[:, side, d]
            scal = slice(dm, dm + p.nscal)  # all scalar columns (1-based dm+1 .. dm+nscal)
            if pb == abi.SLIP_WALL:
                col[0:dm] = abi.HOEXTRAP
                col[d] = abi.EXT_DIR
                col[scal] = abi.HOEXTRAP
                col[press - 1], col[fo - 1], col[ho - 1] = abi.FOEXTRAP, abi.FOEXTRAP, abi.HOEXTRAP
            elif pb == abi.NO_SLIP_WALL:
                col[0:dm] = abi.EXT_DIR
                col[scal] = abi.HOEXTRAP
                col[press - 1], col[fo - 1], col[ho - 1] = abi.FOEXTRAP, abi.FOEXTRAP, abi.HOEXTRAP
            elif pb == abi.INLET:
                col[0:dm] = abi.EXT_DIR
                col[scal] = abi.EXT_DIR
                col[press - 1], col[fo - 1], col[ho - 1] = abi.FOEXTRAP, abi.FOEXTRAP, abi.HOEXTRAP
            elif pb == abi.OUTLET:
                col[0:dm] = abi.FOEXTRAP
                col[scal] = abi.FOEXTRAP
                col[press - 1], col[fo - 1], col[ho - 1] = abi.EXT_DIR, abi.FOEXTRAP, abi.HOEXTRAP
            elif pb == abi.SYMMETRY:
                col[0:dm] = abi.REFLECT_EVEN
                col[d] = abi.REFLECT_ODD
                col[scal] = abi.REFLECT_EVEN
                col[press - 1], col[fo - 1], col[ho - 1] = abi.REFLECT_EVEN, abi.REFLECT_EVEN, abi.REFLECT_EVEN
            elif pb in (abi.PERIODIC, abi.INTERIOR):
                pass
            else:
                raise ValueError("unknown physical bc %r" % (pb,))
    return np.ascontiguousarray(bc)


def as_double_p(x):
    if x is None:
        return None
    x = np.ascontiguousarray(x, dtype=np.float64)
    return x, x.ctypes.data_as(abi.c_double_p)


def as_int_p(x):
    x = np.ascontiguousarray(x, dtype=np.int32)
    return x, x.ctypes.data_as(abi.c_int_p)


def fab_ptr(f):
    """pointer to a 1-element array of mgpu_fab (nfabs = 1)"""
    arr = (abi.mgpu_fab * 1)(f.cfab())
    return arr


def fab_pp(fabs):
    """array of dm pointers, each to a 1-element array of mgpu_fab"""
    keep = [fab_ptr(f) for f in fabs]
    pp = (abi.F_ * len(fabs))(*[C.cast(k, abi.F_) for k in keep])
    return pp, keep


class Geom:
    """Spherical geometry as MAESTRO's geometry module builds it (Source/geometry.f90: init_radial /
    init_spherical): dr = dx/drdxfac, nr_fine radial bins reaching the domain corner, r_cc_loc = (r+1/2) dr,
    r_edge_loc = r dr; centre = middle of the domain.  Owns the arrays the C struct points to."""

    def __init__(self, p, drdxfac=5, center=None, prob_lo=(0.0, 0.0, 0.0), s0_interp_type=3, w0_interp_type=2,
                 s0mac_interp_type=1, w0mac_interp_type=1, nr_fine=None):
        n = [p.domhi[d] - p.domlo[d] + 1 for d in range(3)]
        prob_hi = [prob_lo[d] + n[d] * p.dx[d] for d in range(3)]
        self.dr = p.dx[0] / drdxfac
        if center is None:
            center = [0.5 * (prob_lo[d] + prob_hi[d]) for d in range(3)]
        # far enough for every cell / face / node of the box grown by two ghost cells
        far = max(np.sqrt(sum((max(abs(prob_lo[d] - 3 * p.dx[d] - center[d]), abs(prob_hi[d] + 3 * p.dx[d] - center[d]))) ** 2
                              for d in range(3))), 0.0)
        self.nr_fine = int(far / self.dr) + 4 if nr_fine is None else nr_fine
        self.r_cc_loc = (np.arange(self.nr_fine) + 0.5) * self.dr
        self.r_edge_loc = np.arange(self.nr_fine + 1) * self.dr
        c = abi.mgpu_geom()
        for d in range(3):
            c.center[d] = center[d]
            c.prob_lo[d] = prob_lo[d]
        c.dr = self.dr
        c.nr_fine = self.nr_fine
        c.r_cc_loc = self.r_cc_loc.ctypes.data_as(abi.c_double_p)
        c.r_edge_loc = self.r_edge_loc.ctypes.data_as(abi.c_double_p)
        c.s0_interp_type, c.w0_interp_type = s0_interp_type, w0_interp_type
        c.s0mac_interp_type, c.w0mac_interp_type = s0mac_interp_type, w0mac_interp_type
        self.c = c
        self.center = list(center)
        self.prob_lo = list(prob_lo)
