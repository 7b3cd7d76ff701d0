"""Several boxes per rank: the multifab ghost fill (box-to-box copies with periodic images + physical
BCs per box) and density_advance over a list of boxes, against the oracle on the single box that covers the domain.
The reference's unit test lays its 3-D domain out as 8 boxes (Exec/UNIT_TESTS/test_advect/gr0_3d: 64^3 as 8 x 32^3)."""
import itertools

import numpy as np
import pytest

import oracle_lib
from maestro_b200 import Fab, abi, face_fabs
from synth import make_state, same

pytestmark = pytest.mark.gpu

WALLS = {2: [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]],
         3: [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]}


def boxes_of(n, parts, dm):
    """regular decomposition of [0, n-1]^dm into parts^dm boxes (uneven cuts on purpose)"""
    cuts = [0] + [int(round(n * (q + 1) / parts)) + (1 if 0 < q + 1 < parts else 0) * (q % 2) for q in range(parts)]
    cuts[-1] = n
    rng = [(cuts[q], cuts[q + 1] - 1) for q in range(parts)]
    out = []
    for idx in itertools.product(range(parts), repeat=dm):
        lo = [rng[idx[d]][0] for d in range(dm)] + [0] * (3 - dm)
        hi = [rng[idx[d]][1] for d in range(dm)] + [0] * (3 - dm)
        out.append((lo, hi))
    return out


def take(g, lo, hi, dm):
    """box [lo, hi] (+ghost cells, + nodal faces) cut out of the global fab g"""
    out = Fab(lo, hi, g.ng, g.nc, nodal=g.nodal, dm=dm)
    sl = [slice(None)]
    for d in (2, 1, 0):
        if d < dm:
            o = lo[d] - g.lo[d]
            sl.append(slice(o, o + out.shape[3 - d]))
        else:
            sl.append(slice(None))
    out.a[...] = g.a[tuple(sl)]
    return out


def valid_equal(box, glob, dm, comps=None):
    """the valid region (nodal faces included) of a box against the same region of the global fab"""
    ng = box.ng
    sl_b, sl_g = [slice(None)], [slice(None)]
    for d in (2, 1, 0):
        if d < dm:
            m = box.shape[3 - d] - 2 * ng
            o = box.lo[d] - glob.lo[d] + glob.ng
            sl_b.append(slice(ng, ng + m))
            sl_g.append(slice(o, o + m))
        else:
            sl_b.append(slice(None))
            sl_g.append(slice(None))
    a, b = box.a[tuple(sl_b)], glob.a[tuple(sl_g)]
    if comps is not None:
        a, b = a[comps], b[comps]
    return same(a, b)


@pytest.mark.parametrize("dm,n,parts", [(2, 24, 2), (3, 16, 2), (2, 30, 3)])
@pytest.mark.parametrize("bcset", ["periodic", "walls"])
def test_multifab_fill_boundary(gpu_ops, oracle, dm, n, parts, bcset):
    from maestro_b200 import lib

    st = make_state(dm, n, phys_bc=None if bcset == "periodic" else WALLS[dm])
    p = st["p"]
    ref = st["s"].clone()
    oracle.fill_boundary(p, ref, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
    boxes = [take(st["s"], lo, hi, dm) for lo, hi in boxes_of(n, parts, dm)]
    for b in boxes:  # poison the ghost cells: every one the reference fills must be overwritten
        v = b.valid().copy()
        b.a[...] = -7.0e77
        b.valid()[...] = v
    lib.fill_boundary_mf(p, boxes, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
    for b in boxes:
        want = take(ref, b.lo, b.hi, dm)
        # ghost corners next to a physical wall that the single-box fill leaves alone keep whatever they held
        # (multifab_physbc.f90:165-175): compare where the reference wrote
        wrote = want.a != take(st["s"], b.lo, b.hi, dm).a
        assert same(np.where(wrote, b.a, 0.0), np.where(wrote, want.a, 0.0))
        assert valid_equal(b, ref, dm)


@pytest.mark.parametrize("dm,n,parts", [(2, 24, 2), (3, 16, 2)])
@pytest.mark.parametrize("bcset,ppm_type,spt", [("periodic", 1, 1), ("walls", 2, 1), ("walls", 0, 3), ("periodic", 1, 2)])
def test_density_advance_over_several_boxes(gpu_ops, oracle, dm, n, parts, bcset, ppm_type, spt):
    """exact build: every valid zone / face of every box is bit-identical to the single-box oracle"""
    from maestro_b200 import lib

    lib.set_option("exact", 1)
    try:
        st = make_state(dm, n, phys_bc=None if bcset == "periodic" else WALLS[dm], ppm_type=ppm_type, species_pred_type=spt)
        p, b = st["p"], st["base"]
        p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
        nod = lambda d: [1 if q == d else 0 for q in range(3)]
        g = dict(sold=st["s"].clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone(),
                 sedge=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm), sflux=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm),
                 eta=Fab(st["lo"], st["hi"], 0, 1, nodal=nod(dm - 1), dm=dm))
        oracle.fill_boundary(p, g["sold"], 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        g["snew"] = g["sold"].clone()
        bx = boxes_of(n, parts, dm)
        m = dict(sold=[take(g["sold"], lo, hi, dm) for lo, hi in bx], snew=[take(g["snew"], lo, hi, dm) for lo, hi in bx],
                 force=[take(g["force"], lo, hi, dm) for lo, hi in bx], eta=[take(g["eta"], lo, hi, dm) for lo, hi in bx],
                 umac=[[take(g["umac"][d], lo, hi, dm) for lo, hi in bx] for d in range(dm)],
                 sedge=[[take(g["sedge"][d], lo, hi, dm) for lo, hi in bx] for d in range(dm)],
                 sflux=[[take(g["sflux"][d], lo, hi, dm) for lo, hi in bx] for d in range(dm)])
        oracle.density_advance(p, 1, g["sold"], g["snew"], g["sedge"], g["sflux"], g["force"], g["umac"], b["w0"], g["eta"],
                               b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
        lib.density_advance_mf(p, 1, m["sold"], m["snew"], m["sedge"], m["sflux"], m["force"], m["umac"], b["w0"], m["eta"],
                               b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
        adv = [p.rho_comp - 1] + list(range(p.spec_comp - 1, p.spec_comp - 1 + p.nspec)) + [p.trac_comp - 1]
        for i in range(len(bx)):
            assert valid_equal(m["snew"][i], g["snew"], dm, adv), ("snew", i)
            assert valid_equal(m["eta"][i], g["eta"], dm), ("eta", i)
            for d in range(dm):
                assert valid_equal(m["sedge"][d][i], g["sedge"][d], dm, adv), ("sedge", d, i)
                assert valid_equal(m["sflux"][d][i], g["sflux"][d], dm, adv[1:]), ("sflux", d, i)
    finally:
        lib.set_option("exact", 0)


def test_reference_unit_test_in_its_own_box_layout(gpu_ops, oracle):
    """test_advect 3-D as gr0_3d lays it out -- 8 boxes of 32^3 covering 64^3 -- for three steps of the +x run through
    the CUDA library (exact build), against the oracle driver's single box: bit-identical density."""
    from maestro_b200 import lib, make_adv_bc, make_params

    n, nsteps = 64, 3
    lib.set_option("exact", 1)
    try:
        p = make_params(3, n=[n, n, n], ppm_type=1)
        p.rel_eps = 0.0
        dx = 1.0 / n
        dt = 0.7 * dx
        # the driver's own initial density (libm exp) and its density after nsteps fixed steps
        rho, rho_ref = oracle_lib.test_advect_fields(oracle, 3, n, 1, 1, cfl=0.7, stop_time=nsteps * dt, fixed_dt=1)
        adv_bc = make_adv_bc(p, [[abi.PERIODIC, abi.PERIODIC]] * 3)
        pmask = [1, 1, 1]
        bx = boxes_of(n, 2, 3)
        assert sorted(hi[0] - lo[0] + 1 for lo, hi in bx) == [32] * 8
        glob = Fab([0, 0, 0], [n - 1] * 3, 4, p.nscal, dm=3)
        glob.valid()[p.rho_comp - 1] = rho
        glob.valid()[p.spec_comp - 1] = rho
        sold = [take(glob, lo, hi, 3) for lo, hi in bx]
        lib.fill_boundary_mf(p, sold, 1, 4, p.nscal, adv_bc, pmask)
        snew = [s.clone() for s in sold]
        force = [Fab(lo, hi, 1, p.nscal, dm=3) for lo, hi in bx]
        eta = [Fab(lo, hi, 0, 1, nodal=[0, 0, 1], dm=3) for lo, hi in bx]
        nod = lambda d: [1 if q == d else 0 for q in range(3)]
        umac = [[Fab(lo, hi, 1, 1, nodal=nod(d), dm=3, fill=1.0 if d == 0 else 0.0) for lo, hi in bx] for d in range(3)]
        sedge = [[Fab(lo, hi, 0, p.nscal, nodal=nod(d), dm=3) for lo, hi in bx] for d in range(3)]
        sflux = [[Fab(lo, hi, 0, p.nscal, nodal=nod(d), dm=3) for lo, hi in bx] for d in range(3)]
        zc, ze = np.zeros(n), np.zeros(n + 1)
        for _ in range(nsteps):
            p.dt = dt
            lib.density_advance_mf(p, 1, sold, snew, sedge, sflux, force, umac, ze, eta, zc, zc, zc, ze, adv_bc, pmask)
            for s, t in zip(sold, snew):
                s.a[...] = t.a
        for s in snew:
            lo, hi = s.lo, s.hi
            mine = s.valid()[p.rho_comp - 1]
            want = rho_ref[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
            assert same(mine, want)
    finally:
        lib.set_option("exact", 0)


# ---- the other three episodes over several boxes (velocity_advance, advance_premac, enthalpy_advance) ----------------
VWALLS = {2: [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]],
          3: [[abi.PERIODIC, abi.PERIODIC], [abi.NO_SLIP_WALL, abi.SLIP_WALL], [abi.INLET, abi.OUTLET]]}


def _vel_inputs(oracle, dm, n, bcset, ppm_type):
    from synth import fill_face_ghosts, make_episode_extras, make_vel_state

    phys = None if bcset == "periodic" else VWALLS[dm]
    st = make_vel_state(dm, [n] * dm, phys_bc=phys, ppm_type=ppm_type, do_sponge=1, oracle=oracle)
    ex = make_episode_extras(st)
    sc = make_state(dm, [n] * dm, phys_bc=phys, seed=5)
    oracle.fill_boundary(sc["p"], sc["s"], 1, dm + 1, sc["p"].nscal, sc["adv_bc"], sc["pmask"])
    nr = st["p"].nr
    zr = (np.arange(nr) + 0.5) * st["p"].dx[dm - 1]
    return st, ex, sc["s"], 1.0 + 0.5 * np.exp(-zr / 0.5)


@pytest.mark.parametrize("dm,n,parts", [(2, 24, 2), (3, 16, 2)])
@pytest.mark.parametrize("bcset,ppm_type", [("periodic", 1), ("walls", 2), ("walls", 0)])
def test_velocity_advance_over_several_boxes(gpu_ops, oracle, dm, n, parts, bcset, ppm_type):
    from maestro_b200 import lib
    from synth import fill_face_ghosts

    lib.set_option("exact", 1)
    try:
        st, ex, s, rho0 = _vel_inputs(oracle, dm, n, bcset, ppm_type)
        p = st["p"]
        rng = np.random.default_rng(8)
        umac = face_fabs(st["lo"], st["hi"], 1, 1, dm)
        for u in umac:
            u.a[...] = rng.uniform(-1, 1, size=u.shape)
        fill_face_ghosts(umac, st["pmask"], dm)
        bx = boxes_of(n, parts, dm)
        cut = lambda f: [take(f, lo, hi, dm) for lo, hi in bx]
        m = dict(uold=cut(st["utilde"]), unew=cut(st["utilde"]), s=cut(s), rhohalf=cut(ex["rhohalf"]), gpi=cut(ex["gpi"]),
                 sponge=cut(ex["sponge"]), umac=[cut(umac[d]) for d in range(dm)])
        unew = st["utilde"].clone()
        oracle.velocity_advance(p, st["utilde"], unew, s, ex["rhohalf"], umac, ex["gpi"], st["w0"], ex["w0_force"], rho0,
                                ex["rho0_nph"], ex["grav_old"], ex["grav_nph"], ex["sponge"], st["adv_bc"], st["pmask"])
        lib.velocity_advance_mf(p, m["uold"], m["unew"], m["s"], m["rhohalf"], m["umac"], m["gpi"], st["w0"], ex["w0_force"],
                                rho0, ex["rho0_nph"], ex["grav_old"], ex["grav_nph"], m["sponge"], st["adv_bc"], st["pmask"])
        for i in range(len(bx)):
            assert valid_equal(m["unew"][i], unew, dm), ("unew", i)
            for d in range(dm):
                assert valid_equal(m["umac"][d][i], umac[d], dm), ("umac", d, i)
    finally:
        lib.set_option("exact", 0)


@pytest.mark.parametrize("dm,n,parts", [(2, 24, 2), (3, 16, 2)])
@pytest.mark.parametrize("bcset,ppm_type", [("periodic", 1), ("walls", 2), ("walls", 0)])
def test_advance_premac_over_several_boxes(gpu_ops, oracle, dm, n, parts, bcset, ppm_type):
    from maestro_b200 import lib

    st, ex, s, rho0 = _vel_inputs(oracle, dm, n, bcset, ppm_type)
    p = st["p"]
    bx = boxes_of(n, parts, dm)
    cut = lambda f: [take(f, lo, hi, dm) for lo, hi in bx]
    umac = face_fabs(st["lo"], st["hi"], 1, 1, dm, fill=-777.0)
    m = dict(uold=cut(st["utilde"]), s=cut(s), gpi=cut(ex["gpi"]), umac=[cut(umac[d]) for d in range(dm)])
    oracle.advance_premac(p, st["utilde"], s, umac, ex["gpi"], st["w0"], ex["w0_force"], rho0, ex["grav_old"], st["adv_bc"],
                          st["phys_bc"], st["pmask"])
    lib.set_option("exact", 1)
    try:
        lib.advance_premac_mf(p, m["uold"], m["s"], m["umac"], m["gpi"], st["w0"], ex["w0_force"], rho0, ex["grav_old"],
                              st["adv_bc"], st["phys_bc"], st["pmask"])
    finally:
        lib.set_option("exact", 0)
    for i in range(len(bx)):
        for d in range(dm):
            assert valid_equal(m["umac"][d][i], umac[d], dm), ("umac", d, i)


@pytest.mark.parametrize("dm,n,parts", [(2, 24, 2), (3, 16, 2)])
@pytest.mark.parametrize("bcset,ppm_type,ept,which_step", [("periodic", 1, 1, 1), ("walls", 2, 0, 2), ("walls", 1, 2, 2)])
def test_enthalpy_advance_over_several_boxes(gpu_ops, oracle, dm, n, parts, bcset, ppm_type, ept, which_step):
    from maestro_b200 import lib
    from synth import make_episode_extras

    lib.set_option("exact", 1)
    try:
        st = make_state(dm, n, phys_bc=None if bcset == "periodic" else WALLS[dm], ppm_type=ppm_type, enthalpy_pred_type=ept)
        p, b = st["p"], st["base"]
        p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
        ex = make_episode_extras(st)
        rng = np.random.default_rng(21)
        g = dict(sold=st["s"].clone(), umac=[u.clone() for u in st["umac"]], force=st["force"].clone(),
                 sedge=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm), sflux=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm))
        for f in g["sedge"]:  # density edge states "left by density_advance"
            f.a[p.rho_comp - 1] = 1.0 + rng.uniform(0.0, 0.5, size=f.a[0].shape)
        oracle.fill_boundary(p, g["sold"], 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        g["snew"] = g["sold"].clone()
        bx = boxes_of(n, parts, dm)
        cut = lambda f: [take(f, lo, hi, dm) for lo, hi in bx]
        m = dict(sold=cut(g["sold"]), snew=cut(g["snew"]), force=cut(g["force"]), thermal=cut(ex["thermal"]),
                 umac=[cut(g["umac"][d]) for d in range(dm)], sedge=[cut(g["sedge"][d]) for d in range(dm)],
                 sflux=[cut(g["sflux"][d]) for d in range(dm)])
        args = (b["w0"], b["rho0_old"], b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], ex["p0_old"], ex["p0_new"], ex["psi"],
                ex["grav_old"], ex["grav_nph"], st["adv_bc"], st["pmask"])
        oracle.enthalpy_advance(p, which_step, g["sold"], g["snew"], g["sedge"], g["sflux"], g["force"], ex["thermal"], g["umac"],
                                *args)
        lib.enthalpy_advance_mf(p, which_step, m["sold"], m["snew"], m["sedge"], m["sflux"], m["force"], m["thermal"],
                                m["umac"], *args)
        c = [p.rhoh_comp - 1]
        for i in range(len(bx)):
            assert valid_equal(m["snew"][i], g["snew"], dm, c), ("snew", i)
            assert valid_equal(m["sold"][i], g["sold"], dm, c), ("sold", i)
            assert valid_equal(m["force"][i], g["force"], dm, c), ("force", i)
            for d in range(dm):
                assert valid_equal(m["sedge"][d][i], g["sedge"][d], dm, c), ("sedge", d, i)
                assert valid_equal(m["sflux"][d][i], g["sflux"][d], dm, c), ("sflux", d, i)
                assert valid_equal(m["umac"][d][i], g["umac"][d], dm), ("umac", d, i)
    finally:
        lib.set_option("exact", 0)
