"""Synthetic workloads of the BASELINE.json configurations, shared by bench.py and the tests.

Each builder returns a `Workload`: the multifabs of one rank (host numpy or device torch storage), a `step(ops)` that
runs the advective episodes of the configuration once through the operator interface, a `reset()` that restores the
inputs the episodes modify in place, and the bookkeeping the metric needs (zones, components advanced, algorithmic
bytes per zone).  `ops` is any object with the operator interface: the CUDA library or, for the CPU baseline, the
oracle.

  c2  test_advect 3-D n^3 per GPU, ppm_type 1, periodic: density_advance            (weak scaling, slabs in z)
  c3  reacting_bubble 3-D n^3, planar base state, periodic x/y, slip wall z-lo, outlet z-hi, ppm_type 2:
      density_advance + enthalpy_advance + velocity_advance                          (one GPU)
  c4  Rayleigh-Taylor 2-D n^2 (4096^2), periodic x, slip wall y-lo, outlet y-hi, ppm_type 2, nspec 2:
      density_advance + enthalpy_advance + velocity_advance                          (strong scaling, slabs in y)
  c5  wdconvect 3-D n^3 (512^3) spherical, outlets, ppm_type 1: density_advance_sphr  (strong scaling, slabs in z)
"""
import numpy as np

from . import abi, slab
from .fab import Fab, Geom, face_fabs, make_adv_bc, make_params


class Workload:
    def __init__(self, name, desc, p, zones, ncomp, bytes_per_zone, step, reset, outputs, scaling, extra=None):
        self.name, self.desc, self.p = name, desc, p
        self.zones, self.ncomp, self.bytes_per_zone = zones, ncomp, bytes_per_zone
        self.step, self.reset, self.outputs, self.scaling = step, reset, outputs, scaling
        self.extra = extra or {}

    @property
    def zone_updates(self):
        return self.zones * self.ncomp


def _dev(f, device):
    return f.to(device) if device is not None else f


def _take(fab_g, lo, hi, ng, nodal, dm, device):
    """slab [lo,hi] (+ghosts, + nodal face) cut out of a global host fab with the same ng"""
    out = Fab(lo, hi, ng, fab_g.nc, nodal=nodal, dm=dm)
    r = dm - 1
    o = lo[r] - fab_g.lo[r]
    sl = [slice(None)] * 4
    sl[3 - r] = slice(o, o + out.shape[3 - r])
    out.a[...] = fab_g.a[tuple(sl)]
    return _dev(out, device)


def _copy(dst, src):
    if isinstance(dst, np.ndarray):
        dst[...] = src
    else:
        dst.copy_(src)


def _clone(a):
    return a.copy() if isinstance(a, np.ndarray) else a.clone()


# ---- c2 ----------------------------------------------------------------------------------------------------------
def c2(n=256, device=None, rank=0, world=1, seed=67890):
    """test_advect initial data (Exec/UNIT_TESTS/test_advect/test_advect.f90:58) + velocity set C of SURVEY 8d
    (smooth solenoidal field + seeded noise) on this rank's n^3 slab of the periodic n x n x (n*world) domain."""
    p = make_params(3, n=[n, n, n * world], ppm_type=1)
    p.base_cutoff_density = 1e-10
    lo, hi = [0, 0, rank * n], [n - 1, n - 1, rank * n + n - 1]
    W = float(np.float32(0.05))
    x = (np.arange(-4, n + 4) + 0.5) / n
    xp = ((x % 1.0) - 0.5) ** 2
    r2 = xp[None, None, :] + xp[None, :, None] + xp[:, None, None]
    rho = np.maximum(np.exp(-r2 / W ** 2), 1e-10)
    sold = Fab(lo, hi, 4, p.nscal, dm=3)
    sold.a[p.rho_comp - 1] = rho
    sold.a[p.spec_comp - 1] = 0.6 * rho
    sold.a[p.spec_comp] = 0.3 * rho
    sold.a[p.spec_comp + 1] = 0.1 * rho
    sold.a[p.trac_comp - 1] = np.sin(2 * np.pi * x)[None, None, :] * np.ones_like(rho)
    rng = np.random.default_rng(seed)  # the same noise in every slab: the global field stays 1-periodic
    umac = face_fabs(lo, hi, 1, 1, 3)
    for d, u in enumerate(umac):
        c = [(np.arange(-1, u.shape[3 - q] - 1) + (0.0 if q == d else 0.5)) / n for q in range(3)]
        X, Y, Z = c[0][None, None, :], c[1][None, :, None], c[2][:, None, None]
        o = [X, Y, Z]
        a, b = o[(d + 1) % 3], o[(d + 2) % 3]
        u.a[0] = np.sin(2 * np.pi * b) + np.cos(2 * np.pi * a) + 0.05 * np.sin(2 * np.pi * 7 * (a + b))
        noise = rng.uniform(-0.1, 0.1, size=u.a[0].shape)
        # periodic-consistent noise: the ghost faces and the hi+1 face repeat the faces they are images of
        for ax in range(3):
            m = noise.shape[ax]
            nn = m - 2 - (1 if (2 - ax) == d else 0)
            idx = (np.arange(m) - 1) % nn + 1
            noise = np.take(noise, idx, axis=ax)
        u.a[0] += noise
    umax = max(np.abs(u.a).max() for u in umac)
    p.dt = 0.7 * p.dx[0] / umax
    p.rel_eps = 1e-8 * umax
    adv_bc = make_adv_bc(p, [[abi.PERIODIC, abi.PERIODIC]] * 3)
    zero_c, zero_e = np.zeros(n * world), np.zeros(n * world + 1)
    pmask = [1, 1, 1]
    e = dict(sold=_dev(sold, device), snew=Fab(lo, hi, 4, p.nscal, dm=3, device=device),
             umac=[_dev(u, device) for u in umac], sedge=face_fabs(lo, hi, 0, p.nscal, 3, device=device),
             sflux=face_fabs(lo, hi, 0, p.nscal, 3, device=device), force=Fab(lo, hi, 1, p.nscal, dm=3, device=device),
             eta=Fab(lo, hi, 0, 1, nodal=[0, 0, 1], dm=3, device=device))
    sold0, umac0 = _clone(e["sold"].a), [_clone(u.a) for u in e["umac"]]

    def reset():
        _copy(e["sold"].a, sold0)
        for u, u0 in zip(e["umac"], umac0):
            _copy(u.a, u0)

    def step(ops):
        ops.density_advance(p, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["umac"], zero_e, e["eta"],
                            zero_c, zero_c, zero_c, zero_e, adv_bc, pmask)

    def outputs():
        out = {"snew": e["snew"].valid()}
        for d in range(3):
            out["sedge%s" % "xyz"[d]] = e["sedge"][d].a
            out["sflux%s" % "xyz"[d]] = e["sflux"][d].a
        return out

    ncomp = p.nspec + 1 + p.ntrac
    desc = {"workload": "test_advect 3D %d^3 per GPU, ppm_type=1, density_advance episode (%d comps: %d species + rho' + "
                        "%d tracer), ng_s=4, periodic, velocity set C (smooth solenoidal + seeded noise)"
                        % (n, ncomp, p.nspec, p.ntrac),
            "zones_per_gpu": n ** 3, "components": ncomp}
    return Workload("c2", desc, p, n ** 3, ncomp, 368.0, step, reset, outputs, "weak", extra=dict(e=e, adv_bc=adv_bc))


# ---- c3 / c4: planar base state, walls, the three advance episodes ----------------------------------------------
def _planar_advective_step(name, dm, nn, ppm_type, nspec, device, rank, world, desc_text, bytes_per_zone):
    """density_advance + enthalpy_advance + velocity_advance on this rank's slab of an nn-zone planar domain with
    periodic sides, a slip wall at the bottom and an outlet at the top (reacting_bubble/inputs_3d, rt/inputs_2d).
    The fields come from the generators the parity tests use (tests/synth.py): smooth profiles + seeded noise."""
    from synth import make_episode_extras, make_state, make_vel_state  # test helpers (on sys.path in bench / tests)

    walls = [[abi.PERIODIC, abi.PERIODIC]] * (dm - 1) + [[abi.SLIP_WALL, abi.OUTLET]]
    st = make_state(dm, nn, phys_bc=walls, ppm_type=ppm_type, nspec=nspec)
    vs = make_vel_state(dm, nn, phys_bc=walls, ppm_type=ppm_type, nspec=nspec)
    p, q, b = st["p"], vs["p"], st["base"]
    ex, exv = make_episode_extras(st), make_episode_extras(vs)
    r = dm - 1
    klo, khi = slab.slab_bounds(nn[r], rank, world)
    lo, hi = list(st["lo"]), list(st["hi"])
    lo[r], hi[r] = klo, khi
    phys_r = slab.slab_phys_bc(walls, dm, rank, world)
    adv_bc, adv_bc_v = make_adv_bc(p, phys_r), make_adv_bc(q, phys_r)
    pb_v = np.ascontiguousarray(np.array(phys_r, dtype=np.int32).T)
    nod = lambda d: [1 if k == d else 0 for k in range(3)]
    T = lambda f, ng, nd=(0, 0, 0): _take(f, lo, hi, ng, nd, dm, device)
    umax = max(np.abs(u.a).max() for u in st["umac"])
    p.rel_eps = q.rel_eps = 1e-8 * umax
    e = dict(sold=T(st["s"], 4), snew=T(st["s"], 4), umac=[T(st["umac"][d], 1, nod(d)) for d in range(dm)],
             sedge=face_fabs(lo, hi, 0, p.nscal, dm, device=device), sflux=face_fabs(lo, hi, 0, p.nscal, dm, device=device),
             force=T(st["force"], 1), eta=Fab(lo, hi, 0, 1, nodal=nod(r), dm=dm, device=device),
             thermal=T(ex["thermal"], ex["thermal"].ng), ut=T(vs["utilde"], 4), unew=T(vs["utilde"], 4),
             gpi=T(exv["gpi"], 1), rhohalf=T(exv["rhohalf"], 1), sponge=T(exv["sponge"], 0))
    sold0, umac0 = _clone(e["sold"].a), [_clone(u.a) for u in e["umac"]]
    rho0 = 1.0 + 0.5 * np.exp(-(np.arange(q.nr) + 0.5) * q.dx[r] / 0.5)

    def reset():
        _copy(e["sold"].a, sold0)
        for u, u0 in zip(e["umac"], umac0):
            _copy(u.a, u0)

    def step(ops):
        ops.density_advance(p, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["umac"], b["w0"], e["eta"],
                            b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], adv_bc, st["pmask"])
        ops.enthalpy_advance(p, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["thermal"], e["umac"],
                             b["w0"], b["rho0_old"], b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], ex["p0_old"],
                             ex["p0_new"], ex["psi"], ex["grav_old"], ex["grav_nph"], adv_bc, st["pmask"])
        ops.velocity_advance(q, e["ut"], e["unew"], e["sold"], e["rhohalf"], e["umac"], e["gpi"], vs["w0"],
                             exv["w0_force"], rho0, exv["rho0_nph"], exv["grav_old"], exv["grav_nph"], e["sponge"],
                             adv_bc_v, vs["pmask"])

    def outputs():
        out = {"snew": e["snew"].valid(), "unew": e["unew"].valid()}
        for d in range(dm):
            out["sedge%s" % "xyz"[d]] = e["sedge"][d].a
        return out

    zones = 1
    for d in range(dm):
        zones *= hi[d] - lo[d] + 1
    ncomp = (p.nspec + 1 + p.ntrac) + 1 + dm
    desc = {"workload": desc_text, "zones_per_gpu": zones, "components": ncomp}
    return Workload(name, desc, p, zones, ncomp, bytes_per_zone, step, reset, outputs,
                    "strong" if world > 1 else "weak", extra=dict(e=e, pb_v=pb_v, params=[q]))


def c3(n=256, device=None, rank=0, world=1):
    if world != 1:
        raise ValueError("config c3 (reacting_bubble 256^3, advective step) is a one-GPU configuration")
    text = ("reacting_bubble 3D %d^3 planar base state, periodic x/y + slip wall z-lo + outlet z-hi, ppm_type=2, "
            "advective step: density_advance (5 comps) + enthalpy_advance (1) + velocity_advance (3)" % n)
    # density 368 + enthalpy (72 + 24) + velocity 176 B per zone
    return _planar_advective_step("c3", 3, [n, n, n], 2, 3, device, rank, world, text, 368.0 + 96.0 + 176.0)


def c4(n=4096, device=None, rank=0, world=1):
    text = ("rt Rayleigh-Taylor 2D %d^2, periodic x + slip wall y-lo + outlet y-hi, ppm_type=2, nspec=2, slab-partitioned "
            "in y over %d GPU(s), advective step: density_advance (4 comps) + enthalpy_advance (1) + velocity_advance (2)"
            % (n, world))
    # 2-D: density 4 x 56 + 16, enthalpy 56 + 16, velocity 104 B per zone
    return _planar_advective_step("c4", 2, [n, n], 2, 2, device, rank, world, text, 240.0 + 72.0 + 104.0)


# ---- c5: spherical star in a box, built on the device ---------------------------------------------------------------
def c5(n=512, device="cuda:0", rank=0, world=1, seed=4242):
    """wdconvect in miniature physics, at size: spherical base state rho0(r), w0(r) on dr = dx/5 bins, outlets on all
    sides, slabs in z.  Every large array is generated on the device from analytic profiles on cells / faces
    (bypassing the fill_3d_data interpolation choices) plus seeded noise."""
    import torch

    p = make_params(3, n=[n, n, n], ppm_type=1)
    p.spherical = 1
    g = Geom(p)
    outlet = [[abi.OUTLET, abi.OUTLET]] * 3
    klo, khi = slab.slab_bounds(n, rank, world)
    lo, hi = [0, 0, klo], [n - 1, n - 1, khi]
    adv_bc = make_adv_bc(p, slab.slab_phys_bc(outlet, 3, rank, world))
    pmask = [0, 0, 0]
    rc, re = g.r_cc_loc, g.r_edge_loc
    rad = dict(rho0_old=2.0 * np.exp(-(rc / 0.35) ** 2) + 0.1, rho0_new=2.02 * np.exp(-(rc / 0.35) ** 2) + 0.1,
               w0=0.3 * re * np.exp(-(re / 0.3) ** 2))
    gen = torch.Generator(device=device).manual_seed(seed + rank)
    dx = p.dx[0]

    def coords(f, face=None):
        """cell-centre (or face, along `face`) coordinates of every point of fab f relative to the centre"""
        out = []
        for d in range(3):
            m = f.shape[3 - d]
            off = 0.0 if d == face else 0.5
            c = (torch.arange(m, device=device, dtype=torch.float64) + (f.lo[d] - f.ng) + off) * dx - g.center[d]
            shp = [1, 1, 1]
            shp[2 - d] = m
            out.append(c.reshape(shp))
        return out

    sold = Fab(lo, hi, 4, p.nscal, dm=3, device=device)
    X, Y, Z = coords(sold)
    r = torch.sqrt(X * X + Y * Y + Z * Z)
    rho = (2.0 * torch.exp(-(r / 0.35) ** 2) + 0.1) * (1.0 + 0.01 * torch.sin(9.0 * X) * torch.cos(7.0 * Y) * torch.sin(5.0 * Z))
    frac = [0.5 + 0.2 * torch.tanh((r - 0.2) / 0.05), None, 0.1 + 0.0 * r]
    frac[1] = 1.0 - frac[0] - frac[2]
    sold.a[p.rho_comp - 1] = rho
    for k in range(3):
        sold.a[p.spec_comp - 1 + k] = rho * frac[k]
    sold.a[p.rhoh_comp - 1] = rho * (1.5 + 0.1 * torch.cos(4.0 * r))
    sold.a[p.trac_comp - 1] = torch.sin(6.0 * X) * torch.cos(6.0 * Z)
    del rho, frac, r
    umac = face_fabs(lo, hi, 1, 1, 3, device=device)
    w0mac = face_fabs(lo, hi, 1, 1, 3, device=device)
    umax = 0.0
    for d in range(3):
        c = coords(umac[d], face=d)
        a, b = c[(d + 1) % 3], c[(d + 2) % 3]
        u = 0.5 * (torch.sin(2 * np.pi * b) + torch.cos(2 * np.pi * a)) + 0.0 * c[d]
        u = u + 0.05 * (torch.rand(u.shape, generator=gen, device=device, dtype=torch.float64) - 0.5)
        umac[d].a[0] = u
        rr = torch.sqrt(c[0] ** 2 + c[1] ** 2 + c[2] ** 2)
        w0mac[d].a[0] = 0.3 * torch.exp(-(rr / 0.3) ** 2) * c[d]  # w0(r) r_d / r with w0 = 0.3 r exp(-(r/0.3)^2)
        umax = max(umax, float((umac[d].a.abs() + w0mac[d].a.abs()).max()))
        del u, rr
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([umax], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        umax = float(t.item())
    p.dt = 0.7 * dx / umax
    p.rel_eps = 1e-8 * umax
    e = dict(sold=sold, snew=Fab(lo, hi, 4, p.nscal, dm=3, device=device), umac=umac, w0mac=w0mac,
             sedge=face_fabs(lo, hi, 0, p.nscal, 3, device=device), sflux=face_fabs(lo, hi, 0, p.nscal, 3, device=device),
             force=Fab(lo, hi, 1, p.nscal, dm=3, device=device))
    # the enthalpy and velocity episodes: thermal, grad(pi), rho at the half time, sponge, the velocity itself, the
    # unit radial vector and the w0 force on the cell centres
    rad.update(rhoh0_old=rad["rho0_old"] * (1.5 + 0.1 * np.cos(4.0 * rc)), rhoh0_new=rad["rho0_new"] * (1.5 + 0.1 * np.cos(4.0 * rc)),
               p0_old=5.0 * np.exp(-(rc / 0.5) ** 2) + 0.3, p0_new=5.05 * np.exp(-(rc / 0.5) ** 2) + 0.3,
               psi=0.2 * np.sin(3.0 * rc), grav=-3.0 * rc / (0.05 + rc ** 2), grav_nph=-3.1 * rc / (0.05 + rc ** 2))
    for k in ("rho0_old", "rho0_new"):
        rad[k + "_nph"] = rad[k]
    e["thermal"] = Fab(lo, hi, 1, 1, dm=3, device=device)
    e["gpi"] = Fab(lo, hi, 1, 3, dm=3, device=device)
    e["rhohalf"] = Fab(lo, hi, 1, 1, dm=3, device=device)
    e["sponge"] = Fab(lo, hi, 0, 1, dm=3, device=device, fill=1.0)
    e["ut"] = Fab(lo, hi, 4, 3, dm=3, device=device)
    e["unew"] = Fab(lo, hi, 4, 3, dm=3, device=device)
    e["normal"] = Fab(lo, hi, 1, 3, dm=3, device=device)
    e["w0fc"] = Fab(lo, hi, 1, 3, dm=3, device=device)
    Xc, Yc, Zc = coords(e["gpi"])
    rr = torch.sqrt(Xc * Xc + Yc * Yc + Zc * Zc)
    e["thermal"].a[0] = 0.05 * torch.cos(5.0 * Xc) * torch.sin(4.0 * Yc) + 0.0 * Zc
    for d, Cd in enumerate((Xc, Yc, Zc)):
        e["gpi"].a[d] = 0.1 * torch.sin(3.0 * Cd) + 0.0 * rr
        e["normal"].a[d] = Cd / rr
        e["w0fc"].a[d] = 0.02 * torch.exp(-(rr / 0.3) ** 2) * Cd
    e["rhohalf"].a[0] = 2.01 * torch.exp(-(rr / 0.35) ** 2) + 0.1
    del rr
    Xu, Yu, Zu = coords(e["ut"])
    e["ut"].a[0] = 0.5 * (torch.sin(2 * np.pi * Zu) + torch.cos(2 * np.pi * Yu)) + 0.0 * Xu
    e["ut"].a[1] = 0.5 * (torch.sin(2 * np.pi * Xu) + torch.cos(2 * np.pi * Zu)) + 0.0 * Yu
    e["ut"].a[2] = 0.5 * (torch.sin(2 * np.pi * Yu) + torch.cos(2 * np.pi * Xu)) + 0.0 * Zu
    sold0, umac0 = e["sold"].a.clone(), [u.a.clone() for u in umac]

    def reset():
        _copy(e["sold"].a, sold0)
        for u, u0 in zip(e["umac"], umac0):
            _copy(u.a, u0)

    def step(ops):
        ops.density_advance_sphr(p, g, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["umac"], rad["w0"],
                                 e["w0mac"], rad["rho0_old"], rad["rho0_new"], adv_bc, pmask)
        ops.enthalpy_advance_sphr(p, g, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["thermal"], e["umac"],
                                  rad["w0"], e["w0mac"], rad["rho0_old"], rad["rhoh0_old"], rad["rho0_new"], rad["rhoh0_new"],
                                  rad["p0_old"], rad["p0_new"], rad["psi"], adv_bc, pmask)
        ops.velocity_advance_sphr(p, g, e["ut"], e["unew"], e["sold"], e["rhohalf"], e["umac"], e["gpi"], e["normal"],
                                  rad["w0"], e["w0mac"], e["w0fc"], rad["rho0_old"], rad["rho0_old_nph"], rad["grav"],
                                  rad["grav_nph"], e["sponge"], adv_bc, pmask)

    def outputs():
        return {"snew": e["snew"].valid(), "unew": e["unew"].valid()}

    zones = n * n * (khi - klo + 1)
    ncomp = (p.nspec + 1 + p.ntrac) + 1 + 3
    desc = {"workload": "wdconvect-like spherical 3D %d^3 (drdxfac 5), outlets, ppm_type=1, slab-partitioned in z over %d "
                        "GPU(s), advective step with the spherical base state: density_advance (5 comps) + "
                        "enthalpy_advance (1) + velocity_advance (3)" % (n, world),
            "zones_per_gpu": zones, "components": ncomp}
    # density 368 + enthalpy 96 + velocity 176 B per zone + w0mac 24 B read by each episode
    return Workload("c5", desc, p, zones, ncomp, 368.0 + 96.0 + 176.0 + 72.0, step, reset, outputs,
                    "strong" if world > 1 else "weak", extra=dict(e=e, geom=g))


BUILDERS = {"c2": c2, "c3": c3, "c4": c4, "c5": c5}
DEFAULT_N = {"c2": 256, "c3": 256, "c4": 4096, "c5": 512}
