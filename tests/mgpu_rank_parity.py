"""Multi-GPU parity worker (run under torchrun, one rank per GPU): density_advance on a domain partitioned into one
slab per rank, NCCL halo exchange inside every ghost fill, against the oracle on the single global box.
Prints 'RANK r OK <max rel err>' per rank; exits nonzero on mismatch."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np
import torch
import torch.distributed as dist

import oracle_lib
from maestro_b200 import Fab, abi, face_fabs, lib, make_adv_bc, slab
from synth import make_state


def take(fab_g, lo, hi, ng, nodal, dm):
    """slab [lo,hi] (+ghosts, + nodal face) cut out of a global fab with the same ng"""
    out = Fab(lo, hi, ng, fab_g.nc, nodal=nodal, dm=dm)
    r = dm - 1
    o = lo[r] - fab_g.lo[r]
    sl = [slice(None)] * 4
    sl[3 - r] = slice(o, o + out.shape[3 - r])
    out.a[...] = fab_g.a[tuple(sl)]
    return out


def finish(rank, local, worst):
    ok = worst <= 1e-12
    print("RANK %d %s %.3e" % (rank, "OK" if ok else "MISMATCH", worst), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda:%d" % local)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(int(t.item()) != 0)


def sphr_main(rank, world, local, ops, exact):
    """SURVEY config C5 in miniature: spherical star in a box with outlets on all sides, slabs in z, the device-resident
    spherical density_advance with NCCL halo exchange in every ghost fill, against the oracle on the whole box."""
    from sphr_common import make_sphr_state

    oracle = oracle_lib.load()
    n = [14, 12, 8 * world]
    st = make_sphr_state(n, ops=oracle)
    p, g = st["p"], st["geom"]
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    sold_g = st["s"].clone()
    oracle.fill_boundary(p, sold_g, 1, 4, p.nscal, st["adv_bc"], st["pmask"])
    ref_in = dict(sold=sold_g.clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone())
    snew_g = sold_g.clone()
    sedge_g = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3)
    sflux_g = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3)
    umac_g = [u.clone() for u in st["umac"]]
    force_g = st["force"].clone()
    rad = st["rad"]
    oracle.density_advance_sphr(p, g, 2, sold_g, snew_g, sedge_g, sflux_g, force_g, umac_g, rad["w0"], st["w0mac"],
                                rad["rho0_old"], rad["rho0_new"], st["adv_bc"], st["pmask"])
    klo, khi = slab.slab_bounds(n[2], rank, world)
    lo, hi = list(st["lo"]), list(st["hi"])
    lo[2], hi[2] = klo, khi
    adv_bc_r = make_adv_bc(p, slab.slab_phys_bc(st["phys_bc"], 3, rank, world))
    nod = lambda d: [1 if q == d else 0 for q in range(3)]
    sold = take(ref_in["sold"], lo, hi, 4, [0, 0, 0], 3)
    snew = sold.clone()
    umac = [take(ref_in["umac"][d], lo, hi, 1, nod(d), 3) for d in range(3)]
    w0mac = [take(st["w0mac"][d], lo, hi, 1, nod(d), 3) for d in range(3)]
    force = take(ref_in["force"], lo, hi, 1, [0, 0, 0], 3)
    sedge = face_fabs(lo, hi, 0, p.nscal, 3)
    sflux = face_fabs(lo, hi, 0, p.nscal, 3)
    ops.density_advance_sphr(p, g, 2, sold, snew, sedge, sflux, force, umac, rad["w0"], w0mac, rad["rho0_old"],
                             rad["rho0_new"], adv_bc_r, st["pmask"])
    comps = [p.rho_comp - 1] + list(range(p.spec_comp - 1, p.spec_comp - 1 + p.nspec)) + [p.trac_comp - 1]
    worst = 0.0
    pairs = [(snew, snew_g, 4, [0, 0, 0], comps)]
    pairs += [(sedge[d], sedge_g[d], 0, nod(d), comps) for d in range(3)]
    pairs += [(sflux[d], sflux_g[d], 0, nod(d), comps[1:]) for d in range(3)]
    for mine, glob, ng, nd, cc in pairs:
        want = take(glob, lo, hi, ng, nd, 3)
        a, w = mine.a[cc], want.a[cc]
        if ng:  # valid cells only: ghost values at the physical boundaries of a slab depend on the slab's own fills
            a, w = a[:, ng:-ng, ng:-ng, ng:-ng], w[:, ng:-ng, ng:-ng, ng:-ng]
        worst = max(worst, float(np.abs(a - w).max() / np.abs(w).max()))
    finish(rank, local, worst)


def compare(pairs, lo, hi, dm, interior_ghosts_only=True):
    """max relative error (max-norm per field) of this rank's fabs against the same region of the global ones"""
    worst = 0.0
    for mine, glob, ng, nd in pairs:
        want = take(glob, lo, hi, ng, nd, dm)
        a, w = mine.a, want.a
        if ng:  # valid cells + slab-direction ghost planes
            sl = [slice(None)] * 4
            for d in range(dm - 1):
                sl[3 - d] = slice(ng, a.shape[3 - d] - ng)
            a, w = a[tuple(sl)], w[tuple(sl)]
        den = max(np.abs(w[np.isfinite(w)]).max(), 1e-300)
        err = np.abs(a - w)
        worst = max(worst, err[np.isfinite(err)].max() / den)
    return worst


def episodes_main(rank, world, local, ops, dm, bcset, ppm_type, exact):
    """The other episodes and the reductions over slabs: velocity_advance, advance_premac, enthalpy_advance (NCCL halo
    exchange in their ghost fills), firstdt and average (NCCL min / max / sum), against the oracle on the global box."""
    from synth import fill_face_ghosts, make_episode_extras, make_vel_state
    import ctypes as C

    oracle = oracle_lib.load()
    r = dm - 1
    n = [16, 12, 10 * world] if dm == 3 else [24, 12 * world]
    walls = [[abi.PERIODIC, abi.PERIODIC]] * (dm - 1) + [[abi.SLIP_WALL, abi.OUTLET]]
    phys = None if bcset == "periodic" else walls
    vs = make_vel_state(dm, n, phys_bc=phys, ppm_type=ppm_type, do_sponge=1, oracle=oracle)
    st = make_state(dm, n, phys_bc=phys, ppm_type=ppm_type, enthalpy_pred_type=1)
    p, q, b = st["p"], vs["p"], st["base"]
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    ex = make_episode_extras(vs)
    sold_g = st["s"].clone()
    oracle.fill_boundary(p, sold_g, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
    rng = np.random.default_rng(8)
    umac0 = face_fabs(vs["lo"], vs["hi"], 1, 1, dm)
    for u in umac0:
        u.a[...] = rng.uniform(-1, 1, size=u.shape)
    fill_face_ghosts(umac0, vs["pmask"], dm)
    nr = q.nr
    zr = (np.arange(nr) + 0.5) * q.dx[r]
    rho0 = 1.0 + 0.5 * np.exp(-zr / 0.5)
    nod = lambda d: [1 if k == d else 0 for k in range(3)]
    # ---- the global references
    unew_g = vs["utilde"].clone()
    um_v = [u.clone() for u in umac0]
    oracle.velocity_advance(q, vs["utilde"], unew_g, sold_g, ex["rhohalf"], um_v, ex["gpi"], vs["w0"], ex["w0_force"], rho0,
                            ex["rho0_nph"], ex["grav_old"], ex["grav_nph"], ex["sponge"], vs["adv_bc"], vs["pmask"])
    um_p = face_fabs(vs["lo"], vs["hi"], 1, 1, dm, fill=-777.0)
    oracle.advance_premac(q, vs["utilde"], sold_g, um_p, ex["gpi"], vs["w0"], ex["w0_force"], rho0, ex["grav_old"],
                          vs["adv_bc"], vs["phys_bc"], vs["pmask"])
    e_in = dict(sold=sold_g.clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone())
    e_g = dict(sold=sold_g.clone(), snew=sold_g.clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone(),
               sedge=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm), sflux=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm))
    for f in e_g["sedge"]:
        f.a[p.rho_comp - 1] = 1.0 + rng.uniform(0.0, 0.5, size=f.a[0].shape)
    sedge_in = [f.clone() for f in e_g["sedge"]]
    eargs = (b["w0"], b["rho0_old"], b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], ex["p0_old"], ex["p0_new"], ex["psi"],
             ex["grav_old"], ex["grav_nph"])
    oracle.enthalpy_advance(p, 2, e_g["sold"], e_g["snew"], e_g["sedge"], e_g["sflux"], e_g["force"], ex["thermal"], e_g["umac"],
                            *eargs, st["adv_bc"], st["pmask"])
    avg_g = oracle.average(p, sold_g, p.rhoh_comp)
    # ---- my slab
    klo, khi = slab.slab_bounds(n[r], rank, world)
    lo, hi = list(st["lo"]), list(st["hi"])
    lo[r], hi[r] = klo, khi
    phys_r = slab.slab_phys_bc(st["phys_bc"], dm, rank, world)
    adv_s, adv_v = make_adv_bc(p, phys_r), make_adv_bc(q, phys_r)
    pb_r = np.ascontiguousarray(np.array(phys_r, dtype=np.int32).T)
    cut = lambda f, ng, nd=(0, 0, 0): take(f, lo, hi, ng, list(nd), dm)
    worst = 0.0
    ut, s_r = cut(vs["utilde"], vs["utilde"].ng), cut(sold_g, 4)
    gpi, rhohalf, sponge = cut(ex["gpi"], 1), cut(ex["rhohalf"], 1), cut(ex["sponge"], 0)
    unew = ut.clone()
    um = [cut(umac0[d], 1, nod(d)) for d in range(dm)]
    ops.velocity_advance(q, ut, unew, s_r, rhohalf, um, gpi, vs["w0"], ex["w0_force"], rho0, ex["rho0_nph"], ex["grav_old"],
                         ex["grav_nph"], sponge, adv_v, vs["pmask"])
    worst = max(worst, compare([(unew, unew_g, unew.ng, [0, 0, 0])] + [(um[d], um_v[d], 1, nod(d)) for d in range(dm)], lo, hi, dm))
    um2 = face_fabs(lo, hi, 1, 1, dm, fill=-777.0)
    ops.advance_premac(q, ut, s_r, um2, gpi, vs["w0"], ex["w0_force"], rho0, ex["grav_old"], adv_v, pb_r, vs["pmask"])
    worst = max(worst, compare([(Fab_valid(um2[d]), Fab_valid_of(um_p[d], lo, hi, nod(d), dm), 0, nod(d)) for d in range(dm)], lo, hi, dm))
    e = dict(sold=cut(e_in["sold"], 4), umac=[cut(e_in["umac"][d], 1, nod(d)) for d in range(dm)], force=cut(e_in["force"], 1),
             sedge=[cut(sedge_in[d], 0, nod(d)) for d in range(dm)], sflux=face_fabs(lo, hi, 0, p.nscal, dm),
             thermal=cut(ex["thermal"], 1))
    e["snew"] = e["sold"].clone()
    ops.enthalpy_advance(p, 2, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["thermal"], e["umac"], *eargs, adv_s,
                         st["pmask"])
    worst = max(worst, compare([(e["snew"], e_g["snew"], 4, [0, 0, 0])] +
                               [(e["sedge"][d], e_g["sedge"][d], 0, nod(d)) for d in range(dm)] +
                               [(e["sflux"][d], e_g["sflux"][d], 0, nod(d)) for d in range(dm)], lo, hi, dm))
    avg = ops.average(p, s_r, p.rhoh_comp)
    worst = max(worst, float(np.abs(avg - avg_g).max() / np.abs(avg_g).max()))
    # firstdt (gamma-law EOS set in both libraries): min / max over the ranks, exact
    e_ = abi.mgpu_eos()
    e_.kind, e_.assume_neutral, e_.nspec, e_.gamma, e_.k_B, e_.n_A = 1, 1, p.nspec, 5.0 / 3.0, 1.3806488e-16, 6.02214129e23
    for k in ("mintemp", "mindens", "mine", "minp", "minh"):
        setattr(e_, k, 1e-200)
    for k in ("maxtemp", "maxdens", "maxe", "maxp", "maxh"):
        setattr(e_, k, 1e200)
    for m in range(p.nspec):
        e_.aion[m], e_.zion[m] = 4.0 * (m + 1), 2.0 * (m + 1)
    sT_g = sold_g.clone()
    sT_g.a[p.temp_comp - 1] = 10.0 ** rng.uniform(3.0, 5.0, size=sT_g.a[0].shape)
    divU_g = Fab(st["lo"], st["hi"], 1, 1, dm=dm)
    divU_g.a[...] = rng.uniform(-30, 30, size=divU_g.shape)
    gamma1bar = 1.4 + 0.1 * np.cos(2 * np.pi * zr)
    p0 = 10.0 * np.exp(-zr / 0.4)
    for o in (oracle, ops):
        o.set_eos(e_)
    want = oracle.firstdt(q, vs["utilde"], ex["gpi"], sT_g, divU_g, rho0, p0, ex["grav_old"], gamma1bar, 0.5, 0.1, 1.0e20,
                          use_soundspeed_firstdt=True, use_divu_firstdt=True)
    got = ops.firstdt(q, ut, gpi, cut(sT_g, 4), cut(divU_g, 1), rho0, p0, ex["grav_old"], gamma1bar, 0.5, 0.1, 1.0e20,
                      use_soundspeed_firstdt=True, use_divu_firstdt=True)
    for o in (oracle, ops):
        o.set_eos(None)
    if got != want:
        print("RANK %d firstdt %r != %r" % (rank, got, want), flush=True)
        worst = 1.0
    finish(rank, local, worst)


def Fab_valid(f):
    """a ghost-free copy of a face fab (advance_premac writes the valid faces only)"""
    out = Fab(f.lo, f.hi, 0, f.nc, nodal=f.nodal, dm=f.dm)
    out.a[...] = f.valid()
    return out


def Fab_valid_of(glob, lo, hi, nodal, dm):
    """ghost-free global face fab (take() then cuts the rank's region out of it)"""
    out = Fab(glob.lo, glob.hi, 0, glob.nc, nodal=glob.nodal, dm=dm)
    out.a[...] = glob.valid()
    return out


def main():
    dm = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    bcset = sys.argv[2] if len(sys.argv) > 2 else "periodic"
    ppm_type = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    exact = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ops = lib.init(local)
    slab.comm_init_from_torch(lib.load(), "cuda:%d" % local)
    lib.set_option("exact", exact)  # 0: the FAST kernels (upwind-first fused edge kernel on interior slabs)
    r = dm - 1
    if bcset == "sphr":
        return sphr_main(rank, world, local, ops, exact)
    if bcset.startswith("episodes-"):
        return episodes_main(rank, world, local, ops, dm, bcset.split("-", 1)[1], ppm_type, exact)
    n = ([16, 12, 8 * world] if exact else [40, 12, 20 * world]) if dm == 3 else [24, 10 * world]  # FAST: >= 4 ng planes per slab, so the overlapped exchange of the updated boundary planes runs
    walls = [[abi.PERIODIC, abi.PERIODIC]] * (dm - 1) + [[abi.SLIP_WALL, abi.OUTLET]]
    phys = None if bcset == "periodic" else walls
    st = make_state(dm, n, phys_bc=phys, ppm_type=ppm_type)
    p, b = st["p"], st["base"]
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    oracle = oracle_lib.load()
    # ---- global single-box reference ------------------------------------------------------------
    sold_g = st["s"].clone()
    oracle.fill_boundary(p, sold_g, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
    ref_in = dict(sold=sold_g.clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone())
    snew_g = sold_g.clone()
    sedge_g = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
    sflux_g = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
    eta_g = Fab(st["lo"], st["hi"], 0, 1, nodal=[0] * (dm - 1) + [1], dm=dm)
    umac_g = [u.clone() for u in st["umac"]]
    force_g = st["force"].clone()
    oracle.density_advance(p, 1, sold_g, snew_g, sedge_g, sflux_g, force_g, umac_g, b["w0"], eta_g, b["rho0_old"],
                           b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
    # ---- my slab through the CUDA library ----------------------------------------------------------
    klo, khi = slab.slab_bounds(n[r], rank, world)
    lo, hi = list(st["lo"]), list(st["hi"])
    lo[r], hi[r] = klo, khi
    phys_r = slab.slab_phys_bc(st["phys_bc"], dm, rank, world)
    adv_bc_r = make_adv_bc(p, phys_r)
    nod = lambda d: [1 if q == d else 0 for q in range(3)]
    sold = take(ref_in["sold"], lo, hi, 4, [0, 0, 0], dm)
    snew = sold.clone()
    umac = [take(ref_in["umac"][d], lo, hi, 1, nod(d), dm) for d in range(dm)]
    force = take(ref_in["force"], lo, hi, 1, [0, 0, 0], dm)
    sedge = face_fabs(lo, hi, 0, p.nscal, dm)
    sflux = face_fabs(lo, hi, 0, p.nscal, dm)
    eta = Fab(lo, hi, 0, 1, nodal=[0] * (dm - 1) + [1], dm=dm)
    ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"], b["rho0_new"],
                        b["p0"], b["rho0_predicted_edge"], adv_bc_r, st["pmask"])
    worst = 0.0
    pairs = [(snew, snew_g, 4, [0, 0, 0]), (eta, eta_g, 0, nod(r))]
    pairs += [(sedge[d], sedge_g[d], 0, nod(d)) for d in range(dm)] + [(sflux[d], sflux_g[d], 0, nod(d)) for d in range(dm)]
    for mine, glob, ng, nd in pairs:
        want = take(glob, lo, hi, ng, nd, dm)
        a, w = mine.a, want.a
        if ng:  # compare valid cells + slab-direction ghost planes that have a source (all of them if periodic)
            sl = [slice(None)] * 4
            for d in range(dm - 1):
                sl[3 - d] = slice(ng, a.shape[3 - d] - ng)
            a, w = a[tuple(sl)], w[tuple(sl)]
        den = np.abs(w[np.isfinite(w)]).max()
        err = np.abs(a - w)
        err = err[np.isfinite(err)].max() / den
        worst = max(worst, err)
    ok = worst <= 1e-12
    print("RANK %d %s %.3e" % (rank, "OK" if ok else "MISMATCH", worst), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda:%d" % local)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(int(t.item()) != 0)


if __name__ == "__main__":
    main()
