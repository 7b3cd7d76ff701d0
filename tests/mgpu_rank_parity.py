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
