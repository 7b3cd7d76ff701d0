"""Spherical velocity path: mk_vel_force_3d_sphr (mkforce.f90:484), make_normal
(fill_3d_data.f90:1280) and the advance_premac / velocity_advance episodes with spherical == 1, CUDA against the oracle."""
import numpy as np
import pytest

from maestro_b200 import Fab, abi, face_fabs
from synth import make_episode_extras, make_vel_state, relerr
from synth import same as _same_bits

pytestmark = pytest.mark.gpu

_EXACT = [True]


def same(a, b):
    """bit-identical in the exact build; in the FAST build (the default: reciprocals instead of divisions, e.g. in the
    interpolation of the base state onto the grid) the north-star tolerance, 1e-12 relative"""
    return _same_bits(a, b) if _EXACT[0] else relerr(a, b) <= 1e-12


@pytest.fixture(autouse=True, params=[1, 0], ids=["exact", "fast"])
def build(request):
    """every test of this file runs against both builds of the library, unless it chooses the build itself"""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch

    if not torch.cuda.is_available():
        yield
        return
    from maestro_b200 import lib

    if "exact" in request.fixturenames:
        if request.param == 0:
            pytest.skip("the test sets the build itself")
        yield
        return
    lib.set_option("exact", request.param)
    _EXACT[0] = bool(request.param)
    yield
    _EXACT[0] = True
    lib.set_option("exact", 0)
OUTLET = [[abi.OUTLET, abi.OUTLET]] * 3
SHAPE = (12, 10, 14)


def sphr_velocity_state(oracle, ppm_type=1, omega=0.0, seed=77):
    """velocity-state generator of the planar tests + the spherical extras: geometry, w0mac (scaled up so that it
    matters in the upwinding), normal, w0_force_cart, radial rho0 / grav"""
    from sphr_common import make_sphr_state

    vs = make_vel_state(3, SHAPE, phys_bc=OUTLET, w0amp=0.0, oracle=oracle, ppm_type=ppm_type)
    st = make_sphr_state(SHAPE, ops=oracle)
    p, g = vs["p"], st["geom"]
    p.spherical = 1
    p.omega = omega
    p.rel_eps = 1e-8
    rng = np.random.default_rng(seed)
    lo, hi = vs["lo"], vs["hi"]
    w0 = st["rad"]["w0"] * 20.0
    w0mac = [m.clone() for m in st["w0mac"]]
    for m in w0mac:
        m.a[...] *= 20.0
    normal = Fab(lo, hi, 1, 3, dm=3)
    oracle.make_normal(p, g, normal)
    w0fc = Fab(lo, hi, 1, 3, dm=3)
    w0fc.a[...] = rng.uniform(-0.2, 0.2, size=w0fc.shape)
    ex = make_episode_extras(vs)
    rc = g.r_cc_loc
    rad = dict(w0=w0, rho0=1.5 * np.exp(-(rc / 0.4) ** 2) + 0.2, rho0_nph=1.52 * np.exp(-(rc / 0.4) ** 2) + 0.2,
               grav=-3.0 * rc / (0.05 + rc ** 2), grav_nph=-3.1 * rc / (0.05 + rc ** 2))
    sold = st["s"].clone()
    sold.a[p.rho_comp - 1] = np.abs(sold.a[p.rho_comp - 1]) + 0.5
    return dict(vs=vs, st=st, p=p, g=g, lo=lo, hi=hi, w0mac=w0mac, normal=normal, w0fc=w0fc, ex=ex, rad=rad, sold=sold)


def test_make_normal_is_the_unit_radial_vector(gpu_ops, oracle):
    S = sphr_velocity_state(oracle)
    p, g = S["p"], S["g"]
    n_o = S["normal"]
    n_g = Fab(S["lo"], S["hi"], 1, 3, dm=3)
    gpu_ops.make_normal(p, g, n_g)
    assert same(n_g.a, n_o.a)
    assert np.abs((n_o.a ** 2).sum(0) - 1.0).max() < 1e-14


@pytest.mark.parametrize("is_final", [False, True])
@pytest.mark.parametrize("omega,add_utilde", [(0.0, True), (0.7, True), (0.7, False)])
def test_mk_vel_force_sphr(gpu_ops, oracle, is_final, omega, add_utilde):
    S = sphr_velocity_state(oracle, omega=omega)
    p, g, vs, ex, rad = S["p"], S["g"], S["vs"], S["ex"], S["rad"]
    res = []
    for o in (gpu_ops, oracle):
        f = Fab(S["lo"], S["hi"], 1, 3, dm=3)
        o.mk_vel_force_sphr(p, g, f, is_final, vs["utilde"], S["st"]["umac"], rad["w0"], S["w0mac"], ex["gpi"], S["sold"],
                            p.rho_comp, S["normal"], rad["rho0"], rad["grav"], S["w0fc"], add_utilde)
        res.append(f.valid())
    assert same(res[0], res[1])
    assert np.isfinite(res[1]).all()


def test_mk_vel_force_sphr_known_answers(oracle):
    """no rotation, no w0, no pressure gradient: the force is the buoyancy rhopert g(r) e_r / rho, radial, and it
    vanishes where rho equals the base state"""
    S = sphr_velocity_state(oracle)
    p, g, vs, rad = S["p"], S["g"], S["vs"], S["rad"]
    zero3 = Fab(S["lo"], S["hi"], 1, 3, dm=3)
    w0z = np.zeros_like(rad["w0"])
    w0mz = face_fabs(S["lo"], S["hi"], 1, 1, 3)
    f = Fab(S["lo"], S["hi"], 1, 3, dm=3)
    oracle.mk_vel_force_sphr(p, g, f, False, vs["utilde"], S["st"]["umac"], w0z, w0mz, zero3, S["sold"], p.rho_comp,
                             S["normal"], rad["rho0"], rad["grav"], zero3, True)
    fv, nv = f.valid(), S["normal"].valid()
    cross = np.cross(np.moveaxis(fv, 0, -1), np.moveaxis(nv, 0, -1))
    assert np.abs(cross).max() <= 1e-12 * np.abs(fv).max()
    rho0_cart = Fab(S["lo"], S["hi"], 0, 1, dm=3)
    oracle.put_1d_array_on_cart(p, g, rad["rho0"], rho0_cart, False, False)
    s2 = S["sold"].clone()
    s2.valid()[p.rho_comp - 1] = rho0_cart.a[0]
    oracle.mk_vel_force_sphr(p, g, f, False, vs["utilde"], S["st"]["umac"], w0z, w0mz, zero3, s2, p.rho_comp, S["normal"],
                             rad["rho0"], rad["grav"], zero3, True)
    assert np.abs(f.valid()).max() == 0.0


@pytest.mark.parametrize("ppm_type", [1, 2])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_advance_premac_sphr(gpu_ops, oracle, ppm_type, exact):
    from maestro_b200 import lib

    S = sphr_velocity_state(oracle, ppm_type=ppm_type, omega=0.3)
    p, g, vs, ex, rad = S["p"], S["g"], S["vs"], S["ex"], S["rad"]
    lib.set_option("exact", exact)
    try:
        res = []
        for o in (gpu_ops, oracle):
            um = face_fabs(S["lo"], S["hi"], 1, 1, 3)
            o.advance_premac_sphr(p, g, vs["utilde"], S["sold"], um, ex["gpi"], S["normal"], rad["w0"], S["w0mac"], S["w0fc"],
                                  rad["rho0"], rad["grav"], vs["adv_bc"], vs["phys_bc"], vs["pmask"])
            res.append([u.valid() for u in um])
        for a, b in zip(*res):
            if exact:
                assert same(a, b)
            assert relerr(a, b) <= 1e-12
    finally:
        lib.set_option("exact", 0)


@pytest.mark.parametrize("ppm_type,do_sponge", [(1, 0), (2, 1)])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_velocity_advance_sphr(gpu_ops, oracle, ppm_type, do_sponge, exact):
    from maestro_b200 import lib

    S = sphr_velocity_state(oracle, ppm_type=ppm_type, omega=0.3)
    p, g, vs, ex, rad = S["p"], S["g"], S["vs"], S["ex"], S["rad"]
    p.do_sponge = do_sponge
    p.dt = 0.4 * p.dx[0] / max(np.abs(u.a).max() + np.abs(w.a).max() for u, w in zip(S["st"]["umac"], S["w0mac"]))
    lib.set_option("exact", exact)
    try:
        res = []
        for o in (gpu_ops, oracle):
            um = [u.clone() for u in S["st"]["umac"]]
            unew = vs["utilde"].clone()
            o.velocity_advance_sphr(p, g, vs["utilde"], unew, S["sold"], ex["rhohalf"], um, ex["gpi"], S["normal"], rad["w0"],
                                    S["w0mac"], S["w0fc"], rad["rho0"], rad["rho0_nph"], rad["grav"], rad["grav_nph"],
                                    ex["sponge"], vs["adv_bc"], vs["pmask"])
            res.append([unew.valid()] + [u.valid() for u in um])
        for a, b in zip(*res):
            if exact:
                assert same(a, b)
            assert relerr(a, b) <= 1e-12
    finally:
        lib.set_option("exact", 0)


# ---- enthalpy path, spherical: mkrhohforce_3d_sphr (mkscalforce.f90:388) and the enthalpy_advance episode -----------
def sphr_enthalpy_state(oracle, ept, ppm_type=1):
    from sphr_common import make_sphr_state

    st = make_sphr_state((14, 12, 10), ops=oracle)
    p, g = st["p"], st["geom"]
    p.enthalpy_pred_type = ept
    p.ppm_type = ppm_type
    p.rel_eps = 1e-8 * max(np.abs(u.a).max() for u in st["umac"])
    rc = g.r_cc_loc
    rad = dict(st["rad"])
    rad.update(p0_old=5.0 * np.exp(-(rc / 0.5) ** 2) + 0.3, p0_new=5.05 * np.exp(-(rc / 0.5) ** 2) + 0.3,
               psi=0.2 * np.sin(3.0 * rc))
    ex = make_episode_extras(st)
    sold = st["s"].clone()
    sold.a[p.rho_comp - 1] = np.abs(sold.a[p.rho_comp - 1]) + 0.5
    oracle.fill_boundary(p, sold, 1, 4, p.nscal, st["adv_bc"], st["pmask"])
    return st, p, g, rad, ex, sold


@pytest.mark.parametrize("is_prediction,ept,add_thermal", [(True, abi.PREDICT_RHOHPRIME, True), (True, abi.PREDICT_H, False),
                                                           (False, abi.PREDICT_RHOHPRIME, False)])
def test_mkrhohforce_sphr(gpu_ops, oracle, is_prediction, ept, add_thermal):
    st, p, g, rad, ex, sold = sphr_enthalpy_state(oracle, ept)
    res = []
    for o in (gpu_ops, oracle):
        f = st["force"].clone()
        o.mkrhohforce_sphr(p, g, f, is_prediction, ex["thermal"], st["umac"], rad["p0_old"], rad["p0_new"], rad["psi"],
                           add_thermal, st["adv_bc"], st["pmask"])
        res.append(f.a[p.rhoh_comp - 1])
    assert same(res[0], res[1])
    assert np.abs(res[1]).max() > 0.0


def test_mkrhohforce_sphr_error(gpu_ops, oracle):
    st, p, g, rad, ex, sold = sphr_enthalpy_state(oracle, abi.PREDICT_T_THEN_H)
    for o in (gpu_ops, oracle):
        with pytest.raises(Exception, match="should only call mkrhohforce"):
            o.mkrhohforce_sphr(p, g, st["force"].clone(), True, ex["thermal"], st["umac"], rad["p0_old"], rad["p0_new"],
                               rad["psi"], True, st["adv_bc"], st["pmask"])


@pytest.mark.parametrize("ept,which_step,ppm_type", [(abi.PREDICT_RHOHPRIME, 1, 1), (abi.PREDICT_RHOHPRIME, 2, 2),
                                                     (abi.PREDICT_H, 2, 1), (abi.PREDICT_RHOH, 1, 1)])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_enthalpy_advance_sphr(gpu_ops, oracle, ept, which_step, ppm_type, exact):
    from maestro_b200 import lib

    st, p, g, rad, ex, sold0 = sphr_enthalpy_state(oracle, ept, ppm_type)
    lo, hi = st["lo"], st["hi"]
    lib.set_option("exact", exact)
    try:
        res = []
        for o in (gpu_ops, oracle):
            sold = sold0.clone()
            snew = sold.clone()
            umac = [u.clone() for u in st["umac"]]
            sedge = face_fabs(lo, hi, 0, p.nscal, 3)
            for d in range(3):  # the density edge states density_advance leaves behind
                sedge[d].a[p.rho_comp - 1] = 0.4 + 0.1 * np.cos(np.arange(sedge[d].a[0].size).reshape(sedge[d].a[0].shape) * 0.01)
            sflux = face_fabs(lo, hi, 0, p.nscal, 3)
            force = st["force"].clone()
            o.enthalpy_advance_sphr(p, g, which_step, sold, snew, sedge, sflux, force, ex["thermal"], umac, rad["w0"],
                                    st["w0mac"], rad["rho0_old"], rad["rhoh0_old"], rad["rho0_new"], rad["rhoh0_new"],
                                    rad["p0_old"], rad["p0_new"], rad["psi"], st["adv_bc"], st["pmask"])
            c = p.rhoh_comp - 1
            res.append([sold.a[c], snew.a[c], force.a[c]] + [f.a[c] for f in sedge] + [f.a[c] for f in sflux] +
                       [u.a for u in umac])
        for a, b in zip(*res):
            if exact:
                assert same(a, b)
            assert relerr(a, b) <= 1e-12
    finally:
        lib.set_option("exact", 0)
