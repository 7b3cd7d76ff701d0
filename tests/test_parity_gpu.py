"""GPU parity tests: the CUDA library (through the C ABI, host-pointer calls) against the CPU oracle on
identical seeded inputs.  Tolerance from the north star: 1e-12 relative (max-norm per field); the
parity build (-fmad=false) is expected to be bit-identical, which is asserted where noted."""
import numpy as np
import pytest

from maestro_b200 import Fab, abi, face_fabs
from synth import make_state, relerr, same

pytestmark = pytest.mark.gpu
TOL = 1e-12
FUSED_BY_DEFAULT = 1616  # rows per CTA of the upwind-first kernel the library starts with (1616: the 16x16 tile)

WALLS_3D = [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]
INOUT_3D = [[abi.INLET, abi.OUTLET], [abi.NO_SLIP_WALL, abi.SLIP_WALL], [abi.SYMMETRY, abi.SYMMETRY]]
WALLS_2D = [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]
INOUT_2D = [[abi.SYMMETRY, abi.INLET], [abi.OUTLET, abi.NO_SLIP_WALL]]


def edge_pair(ops, oracle, st, comps, is_vel=False, cons=False, bccomp0=None):
    p, dm = st["p"], st["dm"]
    out = []
    for o in (ops, oracle):
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm, fill=-777.0)
        scomp, ncomp = comps
        o.make_edge_scal(p, st["s"], sedge, st["umac"], st["force"], st["adv_bc"], is_vel, scomp,
                         (dm + scomp) if bccomp0 is None else bccomp0, ncomp, cons)
        out.append(sedge)
    return out


@pytest.fixture(autouse=True)
def exact_arithmetic(gpu_ops):
    """Parity tests run the bit-identical build (option exact=1) unless a test asks for the FAST build."""
    from maestro_b200 import lib

    lib.set_option("exact", 1)
    lib.set_option("fused", 1)
    yield
    lib.set_option("exact", 0)
    lib.set_option("fused", 1)
    lib.set_option("kchunk", -1)


def check(g, c, bitwise=True):
    """parity criterion: 1e-12 relative (max-norm per field); bit-identical for the exact builds"""
    assert relerr(g, c) <= TOL
    if bitwise:
        assert same(g, c), "exact build should be bit-identical"


@pytest.fixture(params=["fused-exact", "fused-fast", "staged"])
def fused(request):
    """3-D non-conservative edge states have three device variants: the fused single-launch kernel in its
    bit-identical build and in its FAST build (dt/dx folded, FMA; mgpu_fused_fast.cu, the default), and
    the staged general path (mgpu_edge.cu).  All must match the oracle."""
    from maestro_b200 import lib

    lib.set_option("fused", 0 if request.param == "staged" else 1)
    lib.set_option("exact", 0 if request.param == "fused-fast" else 1)
    return request.param


@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
@pytest.mark.parametrize("shape,kchunk", [((37, 9, 11), 4), ((8, 33, 5), 64), ((64, 16, 7), 3), ((33, 47, 40), -1)])
@pytest.mark.parametrize("exact", [1, 0, -1], ids=["exact", "fast", "fast-16x16"])
def test_fused_edge_ragged_boxes(gpu_ops, oracle, ppm_type, bcset, shape, kchunk, exact):
    """Fused kernel on boxes that are not multiples of the CTA tile, several CTAs in x/y and several
    z-chunks per column: tile seams, the last face hi+1 and chunk seams must be written exactly once
    and bit-identically."""
    from maestro_b200 import lib

    lib.set_option("fused", 1)
    lib.set_option("exact", max(exact, 0))
    lib.set_option("fused_by", 1616 if exact < 0 else 8)
    lib.set_option("kchunk", kchunk)
    phys = {"periodic": None, "walls": WALLS_3D, "inout": INOUT_3D}[bcset]
    st = make_state(3, shape, phys_bc=phys, ppm_type=ppm_type)
    st["p"].rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    try:
        g, c = edge_pair(gpu_ops, oracle, st, (1, 3))
        gv, cv = edge_pair(gpu_ops, oracle, st, (1, 3), is_vel=True, bccomp0=1)
    finally:
        lib.set_option("kchunk", -1)
        lib.set_option("fused_by", FUSED_BY_DEFAULT)
    for d in range(3):
        for c_ in range(3):
            check(g[d].a[c_], c[d].a[c_], bitwise=exact == 1)
            check(gv[d].a[c_], cv[d].a[c_], bitwise=exact == 1)


@pytest.mark.parametrize("ppm_type", [1, 2])
@pytest.mark.parametrize("bcset", ["walls", "inout"])
@pytest.mark.parametrize("shape", [(46, 44, 72), (44, 58, 131)])
@pytest.mark.parametrize("thin,pair", [(1, 1), (1, 0), (0, 1)])
def test_fused_edge_thin_end_chunks(gpu_ops, oracle, ppm_type, bcset, shape, thin, pair):
    """Boxes with a physical boundary in z that are large enough for the tile split (option thin_edge, the default): the
    boundary kernel runs 8-plane chunks at the two ends of every column and the plain kernel the planes between them in
    chunks of its own size -- the seams between the three kinds of chunk must be written exactly once, and the result
    must not depend on the chunking."""
    from maestro_b200 import lib

    lib.set_option("fused", 1)
    lib.set_option("exact", 0)
    lib.set_option("thin_edge", thin)
    lib.set_option("pair_streams", pair)
    phys = {"walls": WALLS_3D, "inout": INOUT_3D}[bcset]
    st = make_state(3, shape, phys_bc=phys, ppm_type=ppm_type)
    st["p"].rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    g, c = edge_pair(gpu_ops, oracle, st, (1, 3))
    gv, cv = edge_pair(gpu_ops, oracle, st, (1, 3), is_vel=True, bccomp0=1)
    for d in range(3):
        for c_ in range(3):
            check(g[d].a[c_], c[d].a[c_], bitwise=False)
            check(gv[d].a[c_], cv[d].a[c_], bitwise=False)


@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("variant", [0, 1, 2], ids=["literal", "upwind-first", "upwind-first-16x16"])
@pytest.mark.parametrize("shape,kchunk", [((37, 9, 11), 4), ((30, 6, 40), 16), ((70, 20, 9), 64), ((45, 33, 70), -1)])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
def test_fused_edge_slow_faces(gpu_ops, oracle, ppm_type, variant, shape, kchunk, bcset):
    """FAST fused kernels on periodic boxes where a large share of the faces has |u| <= rel_eps (the
    reference then averages the left and right states, make_edge_scal.f90:881-883), some faces have u == 0
    exactly, and the velocity changes sign inside warps.  The upwind-first kernel (mgpu_fused2.cu) takes its
    averaged-T/G branch there."""
    from maestro_b200 import lib

    phys = {"periodic": None, "walls": WALLS_3D, "inout": INOUT_3D}[bcset]
    st = make_state(3, shape, phys_bc=phys, ppm_type=ppm_type)
    umax = max(np.abs(u.a).max() for u in st["umac"])
    for u in st["umac"]:
        u.a[np.abs(u.a) < 0.15 * umax] = 0.0
    st["p"].rel_eps = 0.3 * umax
    lib.set_option("exact", 0)
    lib.set_option("fused_variant", 3 if variant else 0)  # 3: upwind-first kernel for every box and ppm_type
    lib.set_option("fused_by", 1616 if variant == 2 else 8)  # 16x16 tile (boxes without physical boundaries)
    lib.set_option("kchunk", kchunk)  # -1: chosen per launch from the resident CTA slots
    try:
        g, c = edge_pair(gpu_ops, oracle, st, (1, 3))
        gv, cv = edge_pair(gpu_ops, oracle, st, (1, 3), is_vel=True, bccomp0=1)
    finally:
        lib.set_option("kchunk", -1)
        lib.set_option("fused_variant", 1)
        lib.set_option("fused_by", FUSED_BY_DEFAULT)
    for d in range(3):
        for c_ in range(3):
            check(g[d].a[c_], c[d].a[c_], bitwise=False)
            check(gv[d].a[c_], cv[d].a[c_], bitwise=False)


@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
@pytest.mark.parametrize("shape", [(37, 9), (8, 33), (64, 16), (95, 70)])
@pytest.mark.parametrize("slow", [False, True], ids=["moving", "slow-faces"])
@pytest.mark.parametrize("tile", [2, 0, 1, 3], ids=["32x16", "32x8", "16x16", "march"])
def test_fused_edge_2d(gpu_ops, oracle, ppm_type, bcset, shape, slow, tile):
    """FAST fused 2-D kernel (k_fused_edge2d, the whole of make_edge_scal_2d in one launch) on boxes that are not
    multiples of the CTA tile, scalar and velocity components, periodic / wall / inflow-outflow boxes, with and
    without faces whose velocity is below rel_eps (or exactly zero)."""
    from maestro_b200 import lib

    phys = {"periodic": None, "walls": WALLS_2D, "inout": INOUT_2D}[bcset]
    st = make_state(2, shape, phys_bc=phys, ppm_type=ppm_type)
    umax = max(np.abs(u.a).max() for u in st["umac"])
    st["p"].rel_eps = 1e-8 * umax
    if slow:
        for u in st["umac"]:
            u.a[np.abs(u.a) < 0.15 * umax] = 0.0
        st["p"].rel_eps = 0.3 * umax
    lib.set_option("exact", 0)
    lib.set_option("tile2d", tile)
    try:
        g, c = edge_pair(gpu_ops, oracle, st, (1, 3))
        gv, cv = edge_pair(gpu_ops, oracle, st, (1, 2), is_vel=True, bccomp0=1)
    finally:
        lib.set_option("tile2d", 2)
    for d in range(2):
        check(g[d].a[:3], c[d].a[:3], bitwise=False)
        check(gv[d].a[:2], cv[d].a[:2], bitwise=False)


@pytest.mark.parametrize("dm,n", [(2, 24), (3, 16)])
@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
@pytest.mark.parametrize("cons", [False, True])
def test_make_edge_scal(gpu_ops, oracle, fused, dm, n, ppm_type, bcset, cons):
    if fused != "fused-exact" and (cons or (dm == 2 and fused == "staged")):
        pytest.skip("only one device path for this case")
    phys = {"periodic": None, "walls": WALLS_3D if dm == 3 else WALLS_2D,
            "inout": INOUT_3D if dm == 3 else INOUT_2D}[bcset]
    st = make_state(dm, n, phys_bc=phys, ppm_type=ppm_type)
    st["p"].rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    g, c = edge_pair(gpu_ops, oracle, st, (1, st["p"].nscal), cons=cons)
    for d in range(dm):
        check(g[d].a, c[d].a, bitwise=(fused != "fused-fast"))


@pytest.mark.parametrize("dm,n", [(2, 20), (3, 12)])
@pytest.mark.parametrize("ppm_type", [0, 1, 2])
def test_make_edge_scal_velocity(gpu_ops, oracle, dm, n, ppm_type):
    """is_vel=.true. (velocity_advance.f90:102): comps 1..dm with the inflow clamps at FOEXTRAP/HOEXTRAP"""
    st = make_state(dm, n, phys_bc=INOUT_3D if dm == 3 else INOUT_2D, ppm_type=ppm_type)
    g, c = edge_pair(gpu_ops, oracle, st, (1, dm), is_vel=True, bccomp0=1)
    for d in range(dm):
        assert same(g[d].a, c[d].a)


@pytest.mark.parametrize("slope_order", [0, 2])
def test_slope_orders(gpu_ops, oracle, slope_order):
    st = make_state(3, 12, phys_bc=WALLS_3D, ppm_type=0, slope_order=slope_order)
    g, c = edge_pair(gpu_ops, oracle, st, (1, 2))
    for d in range(3):
        assert same(g[d].a, c[d].a)


def test_ppm_trace_forces(gpu_ops, oracle):
    st = make_state(3, 12, ng_f=4, phys_bc=WALLS_3D, ppm_type=1, ppm_trace_forces=1)
    g, c = edge_pair(gpu_ops, oracle, st, (1, 3))
    for d in range(3):
        assert same(g[d].a, c[d].a)


def test_ppm2_needs_4_ghosts(gpu_ops):
    st = make_state(3, 8, ng_s=3, ppm_type=2)
    with pytest.raises(RuntimeError, match="4 ghost"):
        edge_pair(gpu_ops, gpu_ops, st, (1, 1))


def test_invalid_bc_is_an_error(gpu_ops):
    st = make_state(2, 8)
    st["adv_bc"][:] = 99
    with pytest.raises(RuntimeError, match="invalid boundary"):
        edge_pair(gpu_ops, gpu_ops, st, (1, 1))


@pytest.mark.parametrize("dm,n", [(2, 24), (3, 16)])
@pytest.mark.parametrize("spt", [1, 2, 3])
def test_flux_update(gpu_ops, oracle, dm, n, spt):
    st = make_state(dm, n, species_pred_type=spt)
    p, b = st["p"], st["base"]
    rng = np.random.default_rng(7)
    sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
    for f in sedge:
        f.a[...] = rng.uniform(0.5, 1.5, size=f.shape)
    e_old = np.linspace(1.0, 2.0, p.nr + 1)
    e_new = e_old * 1.01
    res = []
    for o in (gpu_ops, oracle):
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[1 if d == dm - 1 else 0 for d in range(3)], dm=dm)
        o.mk_rhoX_flux(p, sflux, eta, sedge, st["umac"], b["w0"], b["rho0_old"], e_old, b["rho0_new"], e_new,
                       b["rho0_predicted_edge"], p.spec_comp, p.spec_comp + p.nspec - 1)
        o.mk_rhoX_flux(p, sflux, eta, sedge, st["umac"], b["w0"], b["rho0_old"], e_old, b["rho0_new"], e_new,
                       b["rho0_predicted_edge"], p.trac_comp, p.trac_comp)
        for ept in (0, 1, 2):
            p.enthalpy_pred_type = ept
            o.mk_rhoh_flux(p, sflux, sedge, st["umac"], b["w0"], b["rho0_old"], e_old, b["rho0_new"], e_new,
                           b["rhoh0_old"], e_old, b["rhoh0_new"], e_new)
        snew = st["s"].clone()
        snew.a[...] = -5.0
        o.update_scal(p, p.spec_comp, p.spec_comp + p.nspec - 1, st["s"], snew, sflux, st["force"])
        o.update_scal(p, p.rhoh_comp, p.rhoh_comp, st["s"], snew, sflux, st["force"])
        o.update_scal(p, p.trac_comp, p.trac_comp, st["s"], snew, sflux, st["force"])
        res.append((sflux, eta, snew))
    (gf, ge, gs), (cf, ce, cs) = res
    for d in range(dm):
        assert same(gf[d].a, cf[d].a)
    assert same(ge.a, ce.a)
    assert same(gs.a, cs.a)


def test_update_scal_floor_and_negative_species(gpu_ops, oracle):
    """density floor + negative-species redistribution (update_scal.f90:453-505)"""
    st = make_state(3, 8)
    p = st["p"]
    p.base_cutoff_density = 3.0  # everything is below the floor
    rng = np.random.default_rng(3)
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3)
    for f in sflux:
        f.a[...] = rng.uniform(-30, 30, size=f.shape)  # large fluxes => negative species
    res = []
    for o in (gpu_ops, oracle):
        snew = st["s"].clone()
        o.update_scal(p, p.spec_comp, p.spec_comp + p.nspec - 1, st["s"], snew, sflux, st["force"])
        res.append(snew.a)
    assert (res[1][p.spec_comp - 1] == 0.0).any()
    assert same(res[0], res[1])


@pytest.mark.parametrize("dm,n", [(2, 16), (3, 12)])
def test_update_velocity_addw0_glue(gpu_ops, oracle, dm, n):
    st = make_state(dm, n)
    p, b = st["p"], st["base"]
    p.do_sponge = 1
    rng = np.random.default_rng(11)
    uold = Fab(st["lo"], st["hi"], 3, dm, dm=dm)
    uold.a[...] = rng.uniform(-1, 1, size=uold.shape)
    uedge = face_fabs(st["lo"], st["hi"], 0, dm, dm)
    for f in uedge:
        f.a[...] = rng.uniform(-1, 1, size=f.shape)
    force = Fab(st["lo"], st["hi"], 1, dm, dm=dm)
    force.a[...] = rng.uniform(-1, 1, size=force.shape)
    sponge = Fab(st["lo"], st["hi"], 0, 1, dm=dm)
    sponge.a[...] = rng.uniform(0.5, 1, size=sponge.shape)
    res = []
    for o in (gpu_ops, oracle):
        unew = uold.clone()
        o.update_velocity(p, uold, unew, st["umac"], uedge, force, sponge, b["w0"])
        um = [u.clone() for u in st["umac"]]
        o.addw0(p, um, b["w0"], 1.0)
        o.addw0(p, um, b["w0"], -1.0)
        f2, s2 = st["force"].clone(), st["s"].clone()
        e_old = np.linspace(1.0, 2.0, p.nr + 1)
        o.modify_scal_force(p, f2, s2, st["umac"], b["rho0_old"], e_old, b["w0"], p.rho_comp, False)
        o.modify_scal_force(p, f2, s2, st["umac"], b["rho0_old"], e_old, b["w0"], p.rhoh_comp, True)
        o.convert_rhoX_to_X(p, s2, True)
        o.put_in_pert_form(p, s2, b["rho0_old"], p.rho_comp, True)
        o.put_in_pert_form(p, s2, b["rho0_old"], p.rho_comp, False)
        o.convert_rhoX_to_X(p, s2, False)
        res.append((unew.a, um[dm - 1].a, f2.a, s2.a))
    for x, y in zip(*res):
        assert same(x, y)


@pytest.mark.parametrize("dm,n", [(2, 12), (3, 10)])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
def test_fill_boundary(gpu_ops, oracle, dm, n, bcset):
    phys = {"periodic": None, "walls": WALLS_3D if dm == 3 else WALLS_2D,
            "inout": INOUT_3D if dm == 3 else INOUT_2D}[bcset]
    st = make_state(dm, n, phys_bc=phys)
    p = st["p"]
    res = []
    for o in (gpu_ops, oracle):
        s = st["s"].clone()
        o.fill_boundary(p, s, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        um = [u.clone() for u in st["umac"]]
        for u in um:
            o.fill_boundary(p, u, 1, 1, 1, st["adv_bc"], st["pmask"])
        res.append([s.a] + [u.a for u in um])
    for x, y in zip(*res):
        assert same(x, y)


@pytest.mark.parametrize("dm,n", [(2, 32), (3, 16)])
@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("spt,which_step,bcset", [(1, 1, "periodic"), (1, 2, "walls"), (2, 2, "periodic"),
                                                 (3, 1, "walls")])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_density_advance(gpu_ops, oracle, dm, n, ppm_type, spt, which_step, bcset, exact):
    """Whole L4 episode (density_advance.f90:20) through the C ABI with host buffers."""
    phys = None if bcset == "periodic" else (WALLS_3D if dm == 3 else WALLS_2D)
    st = make_state(dm, n, phys_bc=phys, ppm_type=ppm_type, species_pred_type=spt)
    p, b = st["p"], st["base"]
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    from maestro_b200 import lib

    lib.set_option("exact", exact)
    res = []
    for o in (gpu_ops, oracle):
        sold = st["s"].clone()
        oracle.fill_boundary(p, sold, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        snew = sold.clone()
        snew.a[...] = 0.0
        umac = [u.clone() for u in st["umac"]]
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        force = st["force"].clone()
        eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[1 if d == dm - 1 else 0 for d in range(3)], dm=dm)
        o.density_advance(p, which_step, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"],
                          b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
        res.append(dict(sold=sold.a, snew=snew.a, eta=eta.a, force=force.a,
                        **{"sedge%d" % d: sedge[d].a for d in range(dm)},
                        **{"sflux%d" % d: sflux[d].a for d in range(dm)},
                        **{"umac%d" % d: umac[d].valid() for d in range(dm)}))
    for k in res[0]:
        assert relerr(res[0][k], res[1][k]) <= TOL, k
        if exact:
            assert same(res[0][k], res[1][k]), k


# ---- mkutrans / velpred ---------------------------------------------------------------------------------
VP_WALLS = {2: [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]],
            3: [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]}
VP_INOUT = {2: [[abi.OUTLET, abi.INLET], [abi.NO_SLIP_WALL, abi.SYMMETRY]],
            3: [[abi.OUTLET, abi.INLET], [abi.NO_SLIP_WALL, abi.SLIP_WALL], [abi.SYMMETRY, abi.OUTLET]]}
VP_INOUT2 = {2: [[abi.INLET, abi.OUTLET], [abi.SYMMETRY, abi.NO_SLIP_WALL]],
             3: [[abi.INLET, abi.OUTLET], [abi.SLIP_WALL, abi.NO_SLIP_WALL], [abi.OUTLET, abi.SYMMETRY]]}


@pytest.mark.parametrize("dm,n", [(2, (24, 17)), (3, (16, 9, 12))])
@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout", "inout2"])
@pytest.mark.parametrize("trace", [0, 1])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_mkutrans_velpred(gpu_ops, oracle, dm, n, ppm_type, bcset, trace, exact):
    """advance_premac.f90:90-116: mkutrans -> ghost fill -> velpred (utrans on every face direction, umac on every face
    direction): bit-identical to the oracle in the exact build, 1e-12 in the FAST build (FMA, reciprocals)."""
    from maestro_b200 import lib
    from synth import fill_face_ghosts, make_vel_state

    if trace and ppm_type == 0:
        pytest.skip("ppm_trace_forces needs ppm_type >= 1")
    lib.set_option("exact", exact)
    phys = {"periodic": None, "walls": VP_WALLS[dm], "inout": VP_INOUT[dm], "inout2": VP_INOUT2[dm]}[bcset]
    st = make_vel_state(dm, list(n), phys_bc=phys, ppm_type=ppm_type, ppm_trace_forces=trace, ng_f=4 if trace else 1,
                        oracle=oracle)
    p = st["p"]
    res = []
    for o in (gpu_ops, oracle):
        utrans = face_fabs(st["lo"], st["hi"], 1, 1, dm, fill=-777.0)
        o.mkutrans(p, st["utilde"], st["ufull"], utrans, st["w0"], st["adv_bc"], st["phys_bc"])
        fill_face_ghosts(utrans, st["pmask"], dm)
        umac = face_fabs(st["lo"], st["hi"], 1, 1, dm, fill=-777.0)
        o.velpred(p, st["utilde"], st["ufull"], umac, utrans, st["force"], st["w0"], st["adv_bc"], st["phys_bc"])
        res.append(utrans + umac)
    for g, c in zip(*res):
        check(g.a, c.a, bitwise=bool(exact))


@pytest.mark.parametrize("slope_order", [0, 2])
def test_velpred_slope_orders(gpu_ops, oracle, slope_order):
    from synth import fill_face_ghosts, make_vel_state

    st = make_vel_state(3, 10, phys_bc=VP_WALLS[3], ppm_type=0, slope_order=slope_order, oracle=oracle)
    res = []
    for o in (gpu_ops, oracle):
        utrans = face_fabs(st["lo"], st["hi"], 1, 1, 3)
        o.mkutrans(st["p"], st["utilde"], st["ufull"], utrans, st["w0"], st["adv_bc"], st["phys_bc"])
        fill_face_ghosts(utrans, st["pmask"], 3)
        umac = face_fabs(st["lo"], st["hi"], 1, 1, 3)
        o.velpred(st["p"], st["utilde"], st["ufull"], umac, utrans, st["force"], st["w0"], st["adv_bc"], st["phys_bc"])
        res.append(utrans + umac)
    for g, c in zip(*res):
        assert same(g.a, c.a)


def test_velpred_invalid_phys_bc_is_an_error(gpu_ops):
    from synth import make_vel_state

    st = make_vel_state(2, 8)
    st["phys_bc"][:] = 99
    utrans = face_fabs(st["lo"], st["hi"], 1, 1, 2)
    with pytest.raises(RuntimeError, match="invalid boundary"):
        gpu_ops.mkutrans(st["p"], st["utilde"], st["ufull"], utrans, st["w0"], st["adv_bc"], st["phys_bc"])


# ---- BDS --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dm,n", [(2, (24, 17)), (3, (14, 9, 11))])
@pytest.mark.parametrize("cons", [False, True])
@pytest.mark.parametrize("vel", ["A", "C"])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_bds(gpu_ops, oracle, dm, n, cons, vel, exact):
    """bds (Source/bds.f90:16): bdsslope + bdsconc with sheared, sign-changing velocities (set C exercises the
    z-face corner quirk): bit-identical to the oracle in the exact build, 1e-12 in the FAST build (FMA, reciprocals)."""
    from maestro_b200 import lib

    st = make_state(dm, list(n), bds_type=1, vel=vel)
    p = st["p"]
    out = []
    lib.set_option("exact", exact)
    try:
        for o in (gpu_ops, oracle):
            sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm, fill=-777.0)
            o.bds(p, st["s"], sedge, st["umac"], st["force"], st["adv_bc"], False, 1, dm + 1, p.nscal, cons)
            out.append(sedge)
    finally:
        lib.set_option("exact", 0)
    for d in range(dm):
        check(out[0][d].a, out[1][d].a, bitwise=bool(exact))


@pytest.mark.parametrize("dm,n", [(2, 20), (3, 12)])
@pytest.mark.parametrize("spt", [1, 2])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_density_advance_bds(gpu_ops, oracle, dm, n, spt, exact):
    from maestro_b200 import lib

    st = make_state(dm, n, bds_type=1, species_pred_type=spt)
    p, b = st["p"], st["base"]
    res = []
    lib.set_option("exact", exact)
    for o in (gpu_ops, oracle):
        sold = st["s"].clone()
        oracle.fill_boundary(p, sold, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        snew = sold.clone()
        umac = [u.clone() for u in st["umac"]]
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        force = st["force"].clone()
        eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[0] * (dm - 1) + [1], dm=dm)
        o.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"], b["rho0_new"],
                          b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
        res.append([snew.a, eta.a] + [f.a for f in sedge] + [f.a for f in sflux])
    lib.set_option("exact", 0)
    for g, c in zip(*res):
        check(g, c, bitwise=bool(exact))


# ---- multi-GPU: slab partition + NCCL halo exchange (needs >= 2 GPUs; run with gpurun --gpus 2) -------------
@pytest.mark.parametrize("dm,bcset,ppm_type,exact", [(3, "periodic", 1, 1), (3, "walls", 2, 1), (2, "walls", 2, 1),
                                                      (3, "periodic", 1, 0), (3, "periodic", 2, 0),
                                                      (3, "sphr", 1, 1), (3, "sphr", 1, 0),
                                                      # velocity_advance, advance_premac, enthalpy_advance, average
                                                      (3, "episodes-periodic", 1, 1), (3, "episodes-walls", 2, 1),
                                                      (2, "episodes-walls", 1, 0), (3, "episodes-periodic", 2, 0)])
def test_multi_gpu_density_advance(gpu_ops, dm, bcset, ppm_type, exact):
    import os
    import subprocess
    import sys

    import torch

    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ng < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(here, "mgpu_rank_parity.py"), str(dm),
           bcset, str(ppm_type), str(exact)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK ") == world


# ---- force builders and the other L4 episodes -------------------------------------------------------------
def _base_for_vel(st):
    nr, dm = st["p"].nr, st["dm"]
    zr = (np.arange(nr) + 0.5) * st["p"].dx[dm - 1]
    return dict(rho0_old=1.0 + 0.5 * np.exp(-zr / 0.5))


def _scalar_like(st, oracle, seed=5):
    """a scalar state (for the density used by mk_vel_force) on the box of a velocity state"""
    sc = make_state(st["dm"], [st["hi"][d] + 1 for d in range(st["dm"])], phys_bc=st["phys"], seed=seed)
    oracle.fill_boundary(sc["p"], sc["s"], 1, st["dm"] + 1, sc["p"].nscal, sc["adv_bc"], sc["pmask"])
    return sc["s"]


@pytest.mark.parametrize("dm,n", [(2, (20, 14)), (3, (12, 9, 10))])
@pytest.mark.parametrize("final", [False, True])
@pytest.mark.parametrize("omega", [0.0, 0.3])
def test_mk_vel_force(gpu_ops, oracle, dm, n, final, omega):
    from synth import make_episode_extras, make_vel_state

    st = make_vel_state(dm, list(n), oracle=oracle)
    p = st["p"]
    p.omega, p.sin_theta, p.cos_theta = omega, 0.6, 0.8
    p.base_cutoff_density = 0.25  # buoyancy cutoff 5 * 0.25 = 1.25 cuts part of the density range
    ex = make_episode_extras(st)
    s = _scalar_like(st, oracle)
    uedge = face_fabs(st["lo"], st["hi"], 1, 1, dm)
    rng = np.random.default_rng(3)
    for u in uedge:
        u.a[...] = rng.uniform(-1, 1, size=u.shape)
    out = []
    for o in (gpu_ops, oracle):
        f = Fab(st["lo"], st["hi"], 1, dm, dm=dm, fill=-777.0)
        o.mk_vel_force(p, f, final, st["utilde"], uedge, st["w0"], ex["gpi"], s, p.rho_comp, _base_for_vel(st)["rho0_old"],
                       ex["grav_old"], ex["w0_force"], True)
        out.append(f)
    check(out[0].a, out[1].a)


@pytest.mark.parametrize("dm,n", [(2, (20, 14)), (3, (12, 9, 10))])
@pytest.mark.parametrize("ept,pred,therm", [(1, True, True), (2, True, False), (0, False, True), (1, False, False)])
def test_mkrhohforce(gpu_ops, oracle, dm, n, ept, pred, therm):
    from synth import make_episode_extras

    st = make_state(dm, list(n), enthalpy_pred_type=ept)
    p, b = st["p"], st["base"]
    p.base_cutoff_density_coord = p.nr - 3
    ex = make_episode_extras(st)
    out = []
    for o in (gpu_ops, oracle):
        f = st["force"].clone()
        o.mkrhohforce(p, f, pred, ex["thermal"], st["umac"], ex["p0_old"], ex["p0_new"], b["rho0_old"], b["rho0_new"],
                      ex["grav_nph"], ex["psi"], therm)
        out.append(f)
    check(out[0].a, out[1].a)


@pytest.mark.parametrize("dm,n", [(2, (22, 15)), (3, (14, 9, 11))])
@pytest.mark.parametrize("ppm_type", [0, 1, 2])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_advance_premac(gpu_ops, oracle, dm, n, ppm_type, bcset, exact):
    """advance_premac.f90:21: ufull, mkutrans, mk_vel_force, addw0, velpred in one device-resident episode"""
    from maestro_b200 import lib
    from synth import make_episode_extras, make_vel_state

    lib.set_option("exact", exact)

    phys = {"periodic": None, "walls": VP_WALLS[dm], "inout": VP_INOUT[dm]}[bcset]
    st = make_vel_state(dm, list(n), phys_bc=phys, ppm_type=ppm_type, oracle=oracle)
    p = st["p"]
    ex = make_episode_extras(st)
    s = _scalar_like(st, oracle)
    out = []
    for o in (gpu_ops, oracle):
        umac = face_fabs(st["lo"], st["hi"], 1, 1, dm, fill=-777.0)
        o.advance_premac(p, st["utilde"], s, umac, ex["gpi"], st["w0"], ex["w0_force"], _base_for_vel(st)["rho0_old"],
                         ex["grav_old"], st["adv_bc"], st["phys_bc"], st["pmask"])
        out.append(umac)
    for g, c in zip(*out):
        check(g.a, c.a, bitwise=bool(exact))


@pytest.mark.parametrize("dm,n", [(2, (22, 15)), (3, (14, 9, 11)), (3, (40, 33, 19))])
@pytest.mark.parametrize("ppm_type", [1, 2])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_advance_premac_fused_and_staged_agree(gpu_ops, oracle, dm, n, ppm_type, exact):
    """On a periodic box advance_premac forms utrans inside velpred's face kernel (option premac_fuse, the default):
    the result equals the staged sequence (k_mkutrans, ghost fill, addw0, ghost fill, velpred) and the oracle bit for
    bit in the exact build, and within the FAST tolerance otherwise."""
    from maestro_b200 import lib
    from synth import make_episode_extras, make_vel_state

    lib.set_option("exact", exact)
    st = make_vel_state(dm, list(n), phys_bc=None, ppm_type=ppm_type, oracle=oracle)
    p = st["p"]
    ex = make_episode_extras(st)
    s = _scalar_like(st, oracle)
    out = []
    for o, fuse in ((gpu_ops, 1), (gpu_ops, 0), (oracle, 0)):
        lib.set_option("premac_fuse", fuse)
        umac = face_fabs(st["lo"], st["hi"], 1, 1, dm, fill=-777.0)
        o.advance_premac(p, st["utilde"], s, umac, ex["gpi"], st["w0"], ex["w0_force"], _base_for_vel(st)["rho0_old"],
                         ex["grav_old"], st["adv_bc"], st["phys_bc"], st["pmask"])
        out.append(umac)
    for f, g, c in zip(*out):
        check(f.a, g.a, bitwise=bool(exact))
        check(f.a, c.a, bitwise=bool(exact))


@pytest.mark.parametrize("dm,n", [(2, (22, 15)), (3, (14, 9, 11))])
@pytest.mark.parametrize("ppm_type,bds", [(0, 0), (1, 0), (2, 0), (1, 1)])
@pytest.mark.parametrize("bcset", ["periodic", "walls", "inout"])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_velocity_advance(gpu_ops, oracle, dm, n, ppm_type, bds, bcset, exact):
    """velocity_advance.f90:16: forces, is_vel edge states of all dm components, update_velocity, ghost fill"""
    from maestro_b200 import lib
    from synth import fill_face_ghosts, make_episode_extras, make_vel_state

    lib.set_option("exact", exact)
    phys = {"periodic": None, "walls": VP_WALLS[dm], "inout": VP_INOUT[dm]}[bcset]
    st = make_vel_state(dm, list(n), phys_bc=phys, ppm_type=ppm_type, bds_type=bds, do_sponge=1, oracle=oracle)
    p = st["p"]
    ex = make_episode_extras(st)
    s = _scalar_like(st, oracle)
    rng = np.random.default_rng(8)
    umac0 = face_fabs(st["lo"], st["hi"], 1, 1, dm)
    for u in umac0:
        u.a[...] = rng.uniform(-1, 1, size=u.shape)
    fill_face_ghosts(umac0, st["pmask"], dm)
    out = []
    for o in (gpu_ops, oracle):
        umac = [u.clone() for u in umac0]
        unew = st["utilde"].clone()
        o.velocity_advance(p, st["utilde"], unew, s, ex["rhohalf"], umac, ex["gpi"], st["w0"], ex["w0_force"],
                           _base_for_vel(st)["rho0_old"], ex["rho0_nph"], ex["grav_old"], ex["grav_nph"], ex["sponge"],
                           st["adv_bc"], st["pmask"])
        out.append([unew] + umac)
    for g, c in zip(*out):
        check(g.a, c.a, bitwise=bool(exact))


@pytest.mark.parametrize("dm,n", [(2, (22, 15)), (3, (14, 9, 11))])
@pytest.mark.parametrize("ept", [0, 1, 2])
@pytest.mark.parametrize("which_step", [1, 2])
@pytest.mark.parametrize("bcset,bds", [("periodic", 0), ("walls", 0), ("periodic", 1)])
def test_enthalpy_advance(gpu_ops, oracle, dm, n, ept, which_step, bcset, bds):
    """enthalpy_advance.f90:16 for predict_rhoh / predict_rhohprime / predict_h"""
    from synth import make_episode_extras

    phys = {"periodic": None, "walls": WALLS_3D if dm == 3 else WALLS_2D}[bcset]
    st = make_state(dm, list(n), phys_bc=phys, enthalpy_pred_type=ept, bds_type=bds, ppm_type=2 if bcset == "walls" else 1)
    p, b = st["p"], st["base"]
    ex = make_episode_extras(st)
    rng = np.random.default_rng(21)
    sedge0 = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
    for f in sedge0:  # density edge states "left by density_advance"
        f.a[p.rho_comp - 1] = 1.0 + rng.uniform(0.0, 0.5, size=f.a[0].shape)
    out = []
    for o in (gpu_ops, oracle):
        sold = st["s"].clone()
        oracle.fill_boundary(p, sold, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
        snew = sold.clone()
        umac = [u.clone() for u in st["umac"]]
        sedge = [f.clone() for f in sedge0]
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
        force = st["force"].clone()
        o.enthalpy_advance(p, which_step, sold, snew, sedge, sflux, force, ex["thermal"], umac, b["w0"], b["rho0_old"],
                           b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], ex["p0_old"], ex["p0_new"], ex["psi"],
                           ex["grav_old"], ex["grav_nph"], st["adv_bc"], st["pmask"])
        out.append([sold, snew, force] + sedge + sflux + umac)
    for g, c in zip(*out):
        check(g.a, c.a)


# ---- spherical geometry (SURVEY config C5 in miniature): every *_3d_sphr operator against the oracle ------------------
@pytest.mark.parametrize("s0t,w0t", [(3, 2), (1, 1), (2, 3)])
@pytest.mark.parametrize("vec", [False, True])
@pytest.mark.parametrize("edge", [False, True])
def test_sphr_put_1d_array_on_cart(gpu_ops, oracle, s0t, w0t, vec, edge):
    from sphr_common import make_sphr_state

    st = make_sphr_state((12, 10, 14), s0_interp_type=s0t, w0_interp_type=w0t)
    p, g = st["p"], st["geom"]
    prof = st["rad"]["w0"] if edge else st["rad"]["rho0_old"]
    out = []
    for o in (gpu_ops, oracle):
        c = Fab(st["lo"], st["hi"], 2, 3 if vec else 1, dm=3)
        c.a[...] = -5.0
        o.put_1d_array_on_cart(p, g, prof, c, edge, vec)
        out.append(c)
    assert same(out[0].a, out[1].a)


@pytest.mark.parametrize("itype", [1, 2, 3, 4])
def test_sphr_make_w0mac_s0mac(gpu_ops, oracle, itype):
    from sphr_common import make_sphr_state

    st = make_sphr_state((12, 10, 14), ops=oracle, w0mac_interp_type=itype, s0mac_interp_type=min(itype, 3),
                         center=[0.503, 0.41, 0.58])
    p, g = st["p"], st["geom"]
    res = []
    for o in (gpu_ops, oracle):
        wm = face_fabs(st["lo"], st["hi"], 1, 1, 3, fill=-3.0)
        sm = face_fabs(st["lo"], st["hi"], 1, 1, 3, fill=-3.0)
        o.make_w0mac(p, g, st["rad"]["w0"], wm, st["w0_cart"])
        o.make_s0mac(p, g, st["rad"]["rho0_old"], sm, st["rho0_old_cart"])
        res.append(wm + sm)
    for a, b in zip(*res):
        assert same(a.a, b.a)


@pytest.mark.parametrize("spt", [abi.PREDICT_RHOPRIME_AND_X, abi.PREDICT_RHOX, abi.PREDICT_RHO_AND_X])
def test_sphr_fluxes_addw0(gpu_ops, oracle, spt):
    from sphr_common import make_sphr_state

    st = make_sphr_state((12, 10, 14), ops=oracle)
    p, lo, hi = st["p"], st["lo"], st["hi"]
    p.species_pred_type = spt
    rng = np.random.default_rng(11)
    sedge = face_fabs(lo, hi, 0, p.nscal, 3)
    for f in sedge:
        f.a[...] = rng.uniform(0.5, 1.5, f.shape)
    epts = [abi.PREDICT_RHOH, abi.PREDICT_RHOHPRIME, abi.PREDICT_H, abi.PREDICT_T_THEN_H]
    if spt == abi.PREDICT_RHOPRIME_AND_X:
        epts.append(abi.PREDICT_HPRIME)
    res = []
    for o in (gpu_ops, oracle):
        um = [u.clone() for u in st["umac"]]
        o.addw0_sphr(p, um, st["w0mac"], 1.0)
        sflux = face_fabs(lo, hi, 0, p.nscal, 3, fill=-9.0)
        o.mk_rhoX_flux_sphr(p, sflux, sedge, st["umac"], st["w0mac"], st["rho0mac_old"], st["rho0mac_new"], p.spec_comp,
                            p.spec_comp + p.nspec - 1)
        o.mk_rhoX_flux_sphr(p, sflux, sedge, st["umac"], st["w0mac"], st["rho0mac_old"], st["rho0mac_new"], p.trac_comp,
                            p.trac_comp + p.ntrac - 1)
        outs = um + sflux
        for ept in epts:
            p.enthalpy_pred_type = ept
            hf = face_fabs(lo, hi, 0, p.nscal, 3, fill=-9.0)
            o.mk_rhoh_flux_sphr(p, hf, sedge, st["umac"], st["w0mac"], st["rho0mac_old"], st["rho0mac_new"],
                                st["rhoh0mac_old"], st["rhoh0mac_new"])
            outs += hf
        res.append(outs)
    for a, b in zip(*res):
        assert same(a.a, b.a)
    if spt != abi.PREDICT_RHOPRIME_AND_X:  # mkflux.f90:1406-1412
        p.enthalpy_pred_type = abi.PREDICT_HPRIME
        hf = face_fabs(lo, hi, 0, p.nscal, 3)
        for o in (gpu_ops, oracle):
            with pytest.raises(Exception, match="not supported together"):
                o.mk_rhoh_flux_sphr(p, hf, sedge, st["umac"], st["w0mac"], st["rho0mac_old"], st["rho0mac_new"],
                                    st["rhoh0mac_old"], st["rhoh0mac_new"])


@pytest.mark.parametrize("do_sponge", [0, 1])
def test_sphr_update_velocity(gpu_ops, oracle, do_sponge):
    from sphr_common import make_sphr_state

    st = make_sphr_state((12, 10, 14), ops=oracle)
    p, lo, hi = st["p"], st["lo"], st["hi"]
    p.do_sponge = do_sponge
    rng = np.random.default_rng(21)
    uold = Fab(lo, hi, 3, 3, dm=3)
    uold.a[...] = rng.uniform(-1, 1, uold.shape)
    force = Fab(lo, hi, 1, 3, dm=3)
    force.a[...] = rng.uniform(-1, 1, force.shape)
    uedge = face_fabs(lo, hi, 0, 3, 3)
    for f in uedge:
        f.a[...] = rng.uniform(-1, 1, f.shape)
    sponge = Fab(lo, hi, 0, 1, dm=3)
    sponge.a[...] = rng.uniform(0.5, 1.0, sponge.shape)
    res = []
    for o in (gpu_ops, oracle):
        unew = uold.clone()
        o.update_velocity_sphr(p, uold, unew, st["umac"], uedge, force, sponge, st["w0mac"])
        res.append(unew)
    assert same(res[0].a, res[1].a)


@pytest.mark.parametrize("fullform", [False, True])
@pytest.mark.parametrize("s0t", [1, 2, 3])
def test_sphr_modify_scal_force_pert_form(gpu_ops, oracle, fullform, s0t):
    from sphr_common import make_sphr_state

    st = make_sphr_state((12, 10, 14), ops=oracle, s0_interp_type=s0t)
    p, g, lo, hi = st["p"], st["geom"], st["lo"], st["hi"]
    res = []
    for o in (gpu_ops, oracle):
        force = st["force"].clone()
        o.modify_scal_force_sphr(p, g, force, st["s"], st["umac"], st["rho0_old_cart"], st["rad"]["w0"], p.rho_comp,
                                 fullform)
        s2 = st["s"].clone()
        o.put_in_pert_form_sphr(p, g, s2, st["rad"]["rho0_old"], p.rho_comp, True)
        s3 = s2.clone()
        o.put_in_pert_form_sphr(p, g, s3, st["rad"]["rho0_old"], p.rho_comp, False)
        res.append([force, s2, s3])
    for a, b in zip(*res):
        assert same(a.a, b.a)


@pytest.mark.parametrize("ppm_type", [0, 1, 2])
def test_sphr_mkutrans_velpred(gpu_ops, oracle, ppm_type):
    """advance_premac's two operators with spherical == 1: w0mac enters every Riemann problem"""
    from sphr_common import make_sphr_state
    from synth import fill_face_ghosts, make_vel_state

    vs = make_vel_state(3, (12, 10, 14), phys_bc=[[abi.OUTLET, abi.OUTLET]] * 3, w0amp=0.0, oracle=oracle,
                        ppm_type=ppm_type)
    st = make_sphr_state((12, 10, 14), ops=oracle)
    p, lo, hi = vs["p"], vs["lo"], vs["hi"]
    p.spherical = 1
    w0mac = st["w0mac"]
    for m in w0mac:  # comparable in size to the velocities so that the upwinding decisions change
        m.a[...] *= 20.0
    res = []
    for o in (gpu_ops, oracle):
        ut = face_fabs(lo, hi, 1, 1, 3)
        o.mkutrans_sphr(p, vs["utilde"], vs["ufull"], ut, w0mac, vs["adv_bc"], vs["phys_bc"])
        fill_face_ghosts(ut, vs["pmask"], 3)
        um = face_fabs(lo, hi, 1, 1, 3)
        o.velpred_sphr(p, vs["utilde"], vs["ufull"], um, ut, vs["force"], w0mac, vs["adv_bc"], vs["phys_bc"])
        res.append(ut + um)
    for a, b in zip(*res):
        assert same(a.a, b.a)
    # the spherical Riemann problems must differ from the planar ones for this w0mac
    p.spherical = 0
    ut0 = face_fabs(lo, hi, 1, 1, 3)
    oracle.mkutrans(p, vs["utilde"], vs["ufull"], ut0, np.zeros(p.nr + 1), vs["adv_bc"], vs["phys_bc"])
    assert not same(ut0[0].a, res[1][0].a)


@pytest.mark.parametrize("spt", [abi.PREDICT_RHOPRIME_AND_X, abi.PREDICT_RHOX, abi.PREDICT_RHO_AND_X])
@pytest.mark.parametrize("which_step,s0mac_t", [(1, 1), (2, 1), (2, 3)])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_sphr_density_advance(gpu_ops, oracle, spt, which_step, s0mac_t, exact):
    """density_advance with spherical == 1 (SURVEY config C5 in miniature: outlet on all sides, radial base state):
    rho0_old_cart, spherical perturbational form, addw0 with w0mac, rho0mac by make_s0mac, spherical fluxes"""
    from maestro_b200 import lib
    from sphr_common import make_sphr_state

    st = make_sphr_state((14, 12, 10), ops=oracle, s0mac_interp_type=s0mac_t)
    p, g, lo, hi = st["p"], st["geom"], st["lo"], st["hi"]
    p.species_pred_type = spt
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    lib.set_option("exact", exact)
    res = []
    for o in (gpu_ops, oracle):
        sold = st["s"].clone()
        oracle.fill_boundary(p, sold, 1, 4, p.nscal, st["adv_bc"], st["pmask"])
        snew = sold.clone()
        umac = [u.clone() for u in st["umac"]]
        sedge = face_fabs(lo, hi, 0, p.nscal, 3)
        sflux = face_fabs(lo, hi, 0, p.nscal, 3)
        force = st["force"].clone()
        o.density_advance_sphr(p, g, which_step, sold, snew, sedge, sflux, force, umac, st["rad"]["w0"], st["w0mac"],
                               st["rad"]["rho0_old"], st["rad"]["rho0_new"], st["adv_bc"], st["pmask"])
        comps = [p.rho_comp - 1] + list(range(p.spec_comp - 1, p.spec_comp - 1 + p.nspec)) + [p.trac_comp - 1]
        res.append([sold.a[comps], snew.a[comps], force.a] + [f.a[comps] for f in sedge] +
                   [f.a[comps[1:]] for f in sflux] + [u.a for u in umac])
    for a, b in zip(*res):
        check(a, b, bitwise=bool(exact))


# ---- reductions next to the path ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (70, 45)), (3, (33, 20, 27)), (3, (64, 64, 48))])
@pytest.mark.parametrize("active", ["all", "velocity", "force", "divU", "dSdt", "none"])
@pytest.mark.parametrize("space", ["host", "device"])
def test_estdt(gpu_ops, oracle, dm, n, active, space):
    """estdt.f90:29: max / min reductions are exact, so dt and umax are bit-identical to the restated reference
    whichever constraint decides (each is made the binding one in turn)."""
    import torch
    from synth import make_estdt_inputs

    amp = dict(speed=1.0, force_amp=1.0, divu_amp=1.0, dsdt_amp=1.0)
    if active != "all":
        amp = dict(speed=1e-3, force_amp=1e-6, divu_amp=1e-6, dsdt_amp=1e-30)
        key = {"velocity": "speed", "force": "force_amp", "divU": "divu_amp", "dSdt": "dsdt_amp"}.get(active)
        if key:
            amp[key] = 1.0
        if active == "dSdt":
            amp["divu_amp"] = 1e-3
        if active == "none":
            amp = dict(speed=0.0, force_amp=0.0, divu_amp=0.0, dsdt_amp=0.0)
    e = make_estdt_inputs(dm, n, **amp)
    p = e["p"]
    w0 = e["w0"] if active != "none" else np.zeros_like(e["w0"])
    want = oracle.estdt(p, e["u"], e["s"], e["force"], e["divU"], e["dSdt"], w0, e["p0"], e["gamma1bar"], 0.7, 1e30)
    fabs = [e[k] for k in ("u", "s", "force", "divU", "dSdt")]
    if space == "device":
        fabs = [f.to("cuda:0") for f in fabs]
        p.mem_space = abi.DEVICE
    try:
        got = gpu_ops.estdt(p, *fabs, w0, e["p0"], e["gamma1bar"], 0.7, 1e30)
    finally:
        p.mem_space = abi.HOST
    assert got == want, (got, want)
    assert np.isfinite(got[0]) and got[0] > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (70, 45)), (3, (33, 20, 27)), (3, (128, 96, 40))])
def test_make_etarho_planar(gpu_ops, oracle, dm, n):
    """make_eta.f90:36: plane averages of etarhoflux (floating-point sums in a different order: 1e-12 relative)."""
    from synth import make_estdt_inputs

    e = make_estdt_inputs(dm, n)
    p = e["p"]
    nod = [0] * 3
    nod[dm - 1] = 1
    eta = Fab(e["lo"], e["hi"], 0, 1, nodal=nod, dm=dm)
    eta.a[...] = np.random.default_rng(5).uniform(-1.0, 3.0, size=eta.shape)
    ec_w, cc_w = oracle.make_etarho_planar(p, eta)
    ec_g, cc_g = gpu_ops.make_etarho_planar(p, eta)
    assert relerr(ec_g, ec_w) <= TOL and relerr(cc_g, cc_w) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (70, 45)), (3, (33, 20, 27)), (3, (1, 1, 1))])
@pytest.mark.parametrize("space", ["host", "device"])
def test_minmax_verbose_report(gpu_ops, oracle, dm, n, space):
    """multifab_min_c / multifab_max_c behind the `verbose >= 1` lines (density_advance.f90:374-402, enthalpy_advance
    .f90:440-449, velocity_advance.f90:142-160): min and max are exact in any order, so the device reduction equals
    numpy's over the valid zones bit for bit -- ghost cells (set to huge values here) must not enter, the species are
    reported as rhoX / rho and the state is left untouched."""
    from maestro_b200.operators import fortran_e

    st = make_state(dm, list(n))
    p = st["p"]
    s = st["s"].clone()
    v = np.array(s.valid(), copy=True)
    s.a[...] = 1.0e300 * np.where(np.arange(s.a.size).reshape(s.a.shape) % 2 == 0, 1.0, -1.0)
    s.valid()[...] = v
    before = np.array(s.a, copy=True)
    f = s.to("cuda:0") if space == "device" else s
    if space == "device":
        p.mem_space = abi.DEVICE
    try:
        for comp in range(1, p.nscal + 1):
            got = gpu_ops.minmax(p, f, comp)
            want = (float(v[comp - 1].min()), float(v[comp - 1].max()))
            assert got == want, (comp, got, want)
            assert oracle.minmax(p, s, comp) == want
        for n_ in range(p.nspec):
            x = v[p.spec_comp - 1 + n_] / v[p.rho_comp - 1]
            got = gpu_ops.minmax(p, f, p.spec_comp + n_, p.rho_comp)
            assert got == (float(x.min()), float(x.max())), (n_, got)
            assert oracle.minmax(p, s, p.spec_comp + n_, p.rho_comp) == got
        lines = gpu_ops.verbose_report(p, "density_advance", f, spec_names=["helium-4", "carbon-12", "oxygen-16"])
        lines += gpu_ops.verbose_report(p, "enthalpy_advance", f)
        with pytest.raises(Exception):
            gpu_ops.minmax(p, f, p.nscal + 1)
    finally:
        p.mem_space = abi.HOST
    assert lines[0] == "... Level 1 update:"
    x = v[p.spec_comp - 1] / v[p.rho_comp - 1]
    assert lines[1] == "... new min/max : helium-4          %s  %s" % (fortran_e(float(x.min())), fortran_e(float(x.max())))
    assert lines[1 + p.nspec].startswith("... new min/max : density           ")
    assert lines[-2].startswith("... new min/max : rho * H           ") and lines[-1] == " "
    assert np.array_equal(f.numpy() if space == "device" else s.a, before)  # nothing written


@pytest.mark.gpu
@pytest.mark.parametrize("n", [12, (20, 14, 17)])
@pytest.mark.parametrize("active", ["all", "velocity", "divU", "dSdt"])
@pytest.mark.parametrize("w0_interp_type", [1, 2, 3])
def test_estdt_sphr(gpu_ops, oracle, n, active, w0_interp_type):
    """estdt_3d_sphr (estdt.f90:620): w0 on the faces (make_w0mac), grad p0 through put_1d_array_on_cart; max / min
    reductions of bit-identical operands, so dt and umax are bit-identical to the restated reference."""
    from sphr_common import make_sphr_state
    from synth import make_estdt_inputs

    st = make_sphr_state(n=n, ops=oracle, w0_interp_type=w0_interp_type)
    p, g = st["p"], st["geom"]
    shape = (n, n, n) if np.isscalar(n) else n
    amp = dict(speed=1.0, force_amp=1.0, divu_amp=1.0, dsdt_amp=1.0)
    if active != "all":
        amp = dict(speed=1e-3, force_amp=1e-6, divu_amp=1e-6, dsdt_amp=1e-30)
        amp[{"velocity": "speed", "divU": "divu_amp", "dSdt": "dsdt_amp"}[active]] = 1.0
        if active == "dSdt":
            amp["divu_amp"] = 1e-3
    e = make_estdt_inputs(3, list(shape), **amp)
    nr = g.nr_fine
    rc = g.r_cc_loc
    p0 = 10.0 * np.exp(-rc / 0.4)
    g1 = 1.4 + 0.2 * np.cos(2 * np.pi * rc)
    w0 = st["rad"]["w0"] * (1.0 if active in ("all", "velocity") else 1e-4)
    args = (p, g, e["u"], e["s"], e["force"], e["divU"], e["dSdt"], st["w0mac"], w0, p0, g1, 0.7, 1e30)
    want = oracle.estdt_sphr(*args)
    got = gpu_ops.estdt_sphr(*args)
    assert got == want, (got, want)
    assert np.isfinite(got[0]) and got[0] > 0.0


# ---- the reference's own unit test through the CUDA path ------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key", ["dm2_n64_ppm0_dir-2", "dm2_n64_ppm1_dir+1", "dm2_n64_ppm2_dir+1", "dm3_n24_ppm1_dir-3"])
def test_reference_unit_test_through_the_cuda_path(gpu_ops, oracle, key):
    """Exec/UNIT_TESTS/test_advect/varden.f90 (Gaussian advected by a unit velocity, repeated density_advance with host
    multifabs, error norms at the end) with the CUDA library doing every step, exact build: the final density is bit
    for bit the oracle driver's, and the norms are the committed golden values (tests/golden/test_advect_norms.json)."""
    import json
    import os

    import oracle_lib
    from synth import python_test_advect

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "test_advect_norms.json")))
    c = gold["cases"][key]
    a, r, rho = python_test_advect(gpu_ops, c["dm"], c["n"], c["ppm_type"], c["dir"], stop_time=c["stop_time"])
    a0, r0, rho0 = oracle_lib.test_advect(oracle, c["dm"], c["n"], c["ppm_type"], 0, c["dir"], stop_time=c["stop_time"],
                                          want_rho=True)
    assert np.array_equal(rho.reshape(rho0.shape), rho0)
    assert a == a0 and r == r0
    assert abs(a - c["abs"]) <= 1e-13 * c["abs"] and abs(r - c["rel"]) <= 1e-13 * c["rel"]
