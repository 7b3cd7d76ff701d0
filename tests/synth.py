"""Deterministic synthetic states for the parity tests (velocity sets A/B/C)."""
import numpy as np

from maestro_b200 import Fab, abi, face_fabs, make_adv_bc, make_params, nbc_comps


def cell_coords(f, p):
    """cell-centre coordinates (x,y,z) broadcastable to f.a[comp] for a cell-centred or nodal fab"""
    out = []
    for d in range(3):
        if d < f.dm:
            n = f.shape[3 - d]
            idx = np.arange(n) + f.lo[d] - f.ng
            x = (idx + (0.0 if f.nodal[d] else 0.5)) * p.dx[d]
        else:
            x = np.zeros(1)
        shape = [1, 1, 1]
        shape[2 - d] = x.size
        out.append(x.reshape(shape))
    return out


def make_state(dm, n, ng_s=4, ng_f=1, seed=12345, vel="C", phys_bc=None, nspec=3, noise=0.1, **pkw):
    """Returns dict with params, sold, force, umac, adv_bc, pmask, base state, all host fabs."""
    rng = np.random.default_rng(seed)
    nn = [n] * dm if np.isscalar(n) else list(n)
    p = make_params(dm, n=nn + [1] * (3 - dm), nspec=nspec, **pkw)
    lo, hi = [0, 0, 0], [nn[d] - 1 if d < dm else 0 for d in range(3)]
    if phys_bc is None:
        phys_bc = [[abi.PERIODIC, abi.PERIODIC]] * dm
    pmask = [1 if phys_bc[d][0] == abi.PERIODIC else 0 for d in range(dm)] + [0] * (3 - dm)
    adv_bc = make_adv_bc(p, phys_bc)
    s = Fab(lo, hi, ng_s, p.nscal, dm=dm)
    x, y, z = cell_coords(s, p)
    r2 = (x - 0.5) ** 2 + (y - 0.5 * p.dx[1] * nn[1]) ** 2 + ((z - 0.5 * p.dx[2] * nn[2]) ** 2 if dm == 3 else 0.0)
    rho = 1.0 + np.maximum(np.exp(-r2 / 0.05 ** 2), 1e-10) + 0.3 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    s.a[p.rho_comp - 1] = rho
    X = rng.uniform(0.1, 1.0, size=(nspec,) + s.shape[1:])
    X /= X.sum(axis=0, keepdims=True)
    for c in range(nspec):
        s.a[p.spec_comp - 1 + c] = rho * X[c]
    s.a[p.rhoh_comp - 1] = rho * (2.0 + np.cos(2 * np.pi * (x + y)))
    s.a[p.temp_comp - 1] = 1.0 + 0.1 * np.sin(2 * np.pi * x)
    s.a[p.trac_comp - 1] = np.sin(2 * np.pi * y) + (np.cos(2 * np.pi * z) if dm == 3 else 0.0)
    s.a[...] += noise * rng.uniform(-0.1, 0.1, size=s.shape) * (np.abs(s.a) > 0)
    umac = face_fabs(lo, hi, 1, 1, dm)
    for d, u in enumerate(umac):
        xx, yy, zz = cell_coords(u, p)
        if vel == "A":
            u.a[...] = 1.0 if d == 0 else 0.0
        else:
            o = [xx, yy, zz]
            a, b = o[(d + 1) % dm], o[(d + 2) % dm] if dm == 3 else o[(d + 1) % dm]
            u.a[0] = np.sin(2 * np.pi * b) + np.cos(2 * np.pi * a) + 0.0 * (xx + yy + zz)
            if vel == "C":
                u.a[...] += rng.uniform(-0.1, 0.1, size=u.shape)
    umax = max(np.abs(u.a).max() for u in umac)
    p.dt = 0.7 * p.dx[0] / umax
    force = Fab(lo, hi, ng_f, p.nscal, dm=dm)
    force.a[...] = rng.uniform(-1.0, 1.0, size=force.shape)
    nr = p.nr
    zr = (np.arange(nr) + 0.5) * p.dx[dm - 1]
    ze = np.arange(nr + 1) * p.dx[dm - 1]
    base = dict(
        rho0_old=1.0 + 0.5 * np.exp(-zr / 0.5), rho0_new=1.0 + 0.52 * np.exp(-zr / 0.5),
        rhoh0_old=2.0 + np.exp(-zr), rhoh0_new=2.0 + 1.01 * np.exp(-zr),
        w0=0.05 * np.sin(2 * np.pi * ze), rho0_predicted_edge=1.0 + 0.51 * np.exp(-ze / 0.5),
        p0=np.ones(nr),
    )
    return dict(p=p, lo=lo, hi=hi, s=s, umac=umac, force=force, adv_bc=adv_bc, pmask=pmask, base=base,
                phys_bc=phys_bc, dm=dm, ng_s=ng_s)


def relerr(a, b):
    """max-norm relative error of field a against reference b (per field, as the north star words it)"""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max()
    if den == 0.0:
        return np.abs(a - b).max()
    return np.abs(a - b).max() / den


def same(a, b):
    """bit-for-bit equality; NaNs (the reference's own 0/0 in the negative-species redistribution when no
    other species is positive, update_scal.f90:486-494) compare equal whatever their payload"""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def make_vel_state(dm, n, ng_u=4, ng_f=1, seed=2468, phys_bc=None, noise=0.1, w0amp=0.05, oracle=None, **pkw):
    """Inputs of mkutrans / velpred (Source/advance_premac.f90:75-116): utilde with ghost cells filled per adv_bc,
    ufull = utilde + w0 on cells (radial component), force (dm comps), w0 on edges.  Host fabs."""
    rng = np.random.default_rng(seed)
    nn = [n] * dm if np.isscalar(n) else list(n)
    p = make_params(dm, n=nn + [1] * (3 - dm), **pkw)
    lo, hi = [0, 0, 0], [nn[d] - 1 if d < dm else 0 for d in range(3)]
    if phys_bc is None:
        phys_bc = [[abi.PERIODIC, abi.PERIODIC]] * dm
    pmask = [1 if phys_bc[d][0] == abi.PERIODIC else 0 for d in range(dm)] + [0] * (3 - dm)
    adv_bc = make_adv_bc(p, phys_bc)
    pb = np.ascontiguousarray(np.array(phys_bc, dtype=np.int32).T)  # [side, d] == Fortran phys_bc(d+1, side+1)
    utilde = Fab(lo, hi, ng_u, dm, dm=dm)
    x, y, z = cell_coords(utilde, p)
    co = [x, y, z]
    for c in range(dm):
        a, b = co[(c + 1) % dm], co[(c + 2) % dm] if dm == 3 else co[(c + 1) % dm]
        utilde.a[c] = np.sin(2 * np.pi * b) + np.cos(2 * np.pi * a) + 0.3 * np.sin(4 * np.pi * co[c]) + 0.0 * (x + y + z)
    utilde.a[...] += noise * rng.uniform(-1.0, 1.0, size=utilde.shape)
    nr = p.nr
    ze = np.arange(nr + 1) * p.dx[dm - 1]
    w0 = w0amp * np.sin(2 * np.pi * ze)
    if oracle is not None:  # ghost cells as the reference's fills leave them (periodic wrap + multifab_physbc)
        oracle.fill_boundary(p, utilde, 1, 1, dm, adv_bc, pmask)
    ufull = utilde.clone()
    w0c = 0.5 * (w0[:-1] + w0[1:])  # put_1d_array_on_cart of the cell-centred w0 (advance_premac.f90:75-78)
    idx = np.clip(np.arange(-ng_u, nn[dm - 1] + ng_u), 0, nr - 1)
    shape = [1, 1, 1]
    shape[2 - (dm - 1)] = idx.size
    ufull.a[dm - 1] += w0c[idx].reshape(shape)
    umax = float(np.abs(ufull.a).max())
    p.dt = 0.7 * p.dx[0] / umax
    p.rel_eps = 1e-8 * umax
    force = Fab(lo, hi, ng_f, dm, dm=dm)
    force.a[...] = rng.uniform(-1.0, 1.0, size=force.shape)
    return dict(p=p, lo=lo, hi=hi, dm=dm, utilde=utilde, ufull=ufull, force=force, w0=w0, adv_bc=adv_bc, phys_bc=pb,
                pmask=pmask, phys=phys_bc)


def fill_face_ghosts(fabs, pmask, dm):
    """Ghost faces of utrans/umac as multifab_fill_boundary leaves them on a single periodic box (plain copies);
    at non-periodic sides the neighbouring valid face is copied (FBoxLib's multifab_physbc_edgevel is not in the
    reference tree; those values only reach states that the reference overwrites by a BC)."""
    for f in fabs:
        a = f.a
        for d in range(dm):
            ax = 3 - d
            n = f.hi[d] - f.lo[d] + 1
            ng = f.ng
            nod = f.nodal[d]
            ext = a.shape[ax]
            def sl(i0, i1):
                s = [slice(None)] * 4
                s[ax] = slice(i0, i1)
                return tuple(s)
            if pmask[d]:
                a[sl(0, ng)] = a[sl(n, n + ng)]
                a[sl(ng + n + nod, ext)] = a[sl(ng + nod, ng + nod + (ext - ng - n - nod))]
            else:
                a[sl(0, ng)] = a[sl(ng, ng + 1)]
                a[sl(ng + n + nod, ext)] = a[sl(ng + n + nod - 1, ng + n + nod)]


def make_episode_extras(st, seed=97):
    """Extra inputs of enthalpy_advance / velocity_advance / advance_premac on top of make_state / make_vel_state:
    gpi, rhohalf, sponge, thermal (fabs) and psi, grav, p0, w0_force (1-D).  Deterministic."""
    rng = np.random.default_rng(seed)
    p, dm, lo, hi = st["p"], st["dm"], st["lo"], st["hi"]
    nr = p.nr
    zr = (np.arange(nr) + 0.5) * p.dx[dm - 1]
    gpi = Fab(lo, hi, 1, dm, dm=dm)
    gpi.a[...] = rng.uniform(-0.5, 0.5, size=gpi.shape)
    rhohalf = Fab(lo, hi, 1, 1, dm=dm)
    rhohalf.a[...] = 1.5 + rng.uniform(-0.2, 0.2, size=rhohalf.shape)
    sponge = Fab(lo, hi, 0, 1, dm=dm)
    sponge.a[...] = rng.uniform(0.8, 1.0, size=sponge.shape)
    thermal = Fab(lo, hi, 1, 1, dm=dm)
    thermal.a[...] = rng.uniform(-0.3, 0.3, size=thermal.shape)
    return dict(gpi=gpi, rhohalf=rhohalf, sponge=sponge, thermal=thermal,
                psi=0.1 * np.cos(2 * np.pi * zr), grav_old=-1.0 - 0.2 * zr, grav_nph=-1.02 - 0.2 * zr,
                p0_old=2.0 * np.exp(-zr), p0_new=2.05 * np.exp(-zr), w0_force=0.03 * np.sin(2 * np.pi * zr),
                rho0_nph=1.0 + 0.51 * np.exp(-zr / 0.5))


def make_estdt_inputs(dm, n, seed=97, speed=1.0, force_amp=1.0, divu_amp=1.0, dsdt_amp=1.0):
    """Inputs of estdt (Source/estdt.f90:29) on one periodic box: u (ng 3), s (ng 3; density > 0), the velocity
    force (ng 1), divU, dSdt (ng 1), w0 on edges, p0 and gamma1bar on cells.  Host fabs, all random but smooth
    enough that every constraint (velocity, force, divU, dS/dt) is active somewhere."""
    rng = np.random.default_rng(seed)
    nn = [n] * dm if np.isscalar(n) else list(n)
    p = make_params(dm, n=nn + [1] * (3 - dm))
    lo, hi = [0, 0, 0], [nn[d] - 1 if d < dm else 0 for d in range(3)]
    u = Fab(lo, hi, 3, dm, dm=dm)
    u.a[...] = speed * rng.uniform(-1.0, 1.0, size=u.shape)
    s = Fab(lo, hi, 3, p.nscal, dm=dm)
    s.a[...] = rng.uniform(0.5, 2.0, size=s.shape)
    force = Fab(lo, hi, 1, dm, dm=dm)
    force.a[...] = force_amp * rng.uniform(-40.0, 40.0, size=force.shape)
    divU = Fab(lo, hi, 1, 1, dm=dm)
    divU.a[...] = divu_amp * rng.uniform(-30.0, 30.0, size=divU.shape)
    dSdt = Fab(lo, hi, 1, 1, dm=dm)
    dSdt.a[...] = dsdt_amp * rng.uniform(-50.0, 50.0, size=dSdt.shape)
    nr = p.nr
    zc = (np.arange(nr) + 0.5) * p.dx[dm - 1]
    w0 = 0.3 * speed * np.sin(2 * np.pi * np.arange(nr + 1) * p.dx[dm - 1])
    p0 = 10.0 * np.exp(-zc / 0.4)
    gamma1bar = 1.4 + 0.2 * np.cos(2 * np.pi * zc)
    return dict(p=p, lo=lo, hi=hi, dm=dm, u=u, s=s, force=force, divU=divU, dSdt=dSdt, w0=w0, p0=p0,
                gamma1bar=gamma1bar)


def python_test_advect(ops, dm, n, ppm_type, direction, cfl=0.7, stop_time=1.0, device=None, fill_ops=None):
    """The reference's unit test Exec/UNIT_TESTS/test_advect/varden.f90:16 driven through the operator interface:
    Gaussian density (test_advect.f90:58), unit velocity along `direction` (+-1, +-2, +-3), dt = cfl dx, repeated
    density_advance until stop_time, then |rho_final - rho_init|_2 and the relative norm (varden.f90:509-524).
    Returns (abs_norm, rel_norm, rho_final valid cells as numpy).  `ops` runs the episodes (device fabs when
    `device` is given); `fill_ops` fills the initial ghost cells (defaults to ops)."""
    import math

    p = make_params(dm, n=[n] * dm + [1] * (3 - dm), ppm_type=ppm_type)
    lo, hi = [0, 0, 0], [n - 1 if d < dm else 0 for d in range(3)]
    adv_bc = make_adv_bc(p, [[abi.PERIODIC, abi.PERIODIC]] * dm)
    pmask = [1] * dm + [0] * (3 - dm)
    W = float(np.float32(0.05))  # test_advect.f90:11: a dp parameter initialised from a single-precision literal
    x = (np.arange(n) + 0.5) * p.dx[0]
    if dm == 3:
        d2 = (x[None, None, :] - 0.5) ** 2 + (x[None, :, None] - 0.5) ** 2 + (x[:, None, None] - 0.5) ** 2
    else:
        d2 = ((x[None, :] - 0.5) ** 2 + (x[:, None] - 0.5) ** 2)[None]
    dist = np.sqrt(d2)
    # libm's exp, element by element: numpy's vectorised exp differs from it in the last bit for some arguments
    arg = -(dist * dist) / (W * W)
    rho = np.array([math.exp(v) for v in arg.ravel().tolist()]).reshape(arg.shape)
    rho = np.maximum(rho, p.base_cutoff_density)
    sold = Fab(lo, hi, 4, p.nscal, dm=dm)
    sold.valid()[p.rho_comp - 1] = rho
    sold.valid()[p.spec_comp - 1] = rho
    (fill_ops or ops).fill_boundary(p, sold, p.rho_comp, dm + p.rho_comp, p.nscal, adv_bc, pmask)
    dens_orig = sold.valid()[p.rho_comp - 1].copy()
    umac = face_fabs(lo, hi, 1, 1, dm)
    idim = abs(direction) - 1
    for d, u in enumerate(umac):
        u.a[...] = (1.0 if direction > 0 else -1.0) if d == idim else 0.0
    snew = Fab(lo, hi, 4, p.nscal, dm=dm)
    force = Fab(lo, hi, 1, p.nscal, dm=dm)
    sedge = face_fabs(lo, hi, 0, p.nscal, dm)
    sflux = face_fabs(lo, hi, 0, p.nscal, dm)
    nod = [0] * 3
    nod[dm - 1] = 1
    eta = Fab(lo, hi, 0, 1, nodal=nod, dm=dm)
    if device is not None:
        sold, snew, force, eta = sold.to(device), snew.to(device), force.to(device), eta.to(device)
        umac, sedge, sflux = [u.to(device) for u in umac], [f.to(device) for f in sedge], [f.to(device) for f in sflux]
        p.mem_space = abi.DEVICE
    zc, ze = np.zeros(n), np.zeros(n + 1)
    dt = cfl * p.dx[0] / 1.0
    t = 0.0
    try:
        while t < stop_time:
            p.dt = dt
            ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, ze, eta, zc, zc, zc, ze, adv_bc, pmask)
            if device is None:
                sold.a[...] = snew.a
            else:
                sold.a.copy_(snew.a)
            t = t + dt
            if t + dt > stop_time:
                dt = stop_time - t
    finally:
        p.mem_space = abi.HOST
    final = snew.valid()[p.rho_comp - 1]
    final = final if device is None else final.cpu().numpy()
    e = (final - dens_orig).ravel()
    sa = sr = 0.0
    for v, r0 in zip(e.tolist(), dens_orig.ravel().tolist()):  # the reference's summation order
        sa += v * v
        sr += (v / r0) * (v / r0)
    return math.sqrt(sa), math.sqrt(sr), np.array(final)
