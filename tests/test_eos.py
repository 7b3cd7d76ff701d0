"""The EOS pieces next to the advective path: gamma_law_general behind the generic eos() front
end, the reset of rhoh below the cutoff in update_scal, makeHfromRhoT_edge, mktempforce, firstdt, makeTfromRhoH/P and
the temperature-based predictions of enthalpy_advance.

CPU tests: the restated EOS (oracle/mo_eos.cpp) against the closed-form ideal-gas identities -- the reference holds no
golden numbers for its EOS.  GPU tests: the CUDA library against the oracle, bit for bit (this code is built without FMA
contraction; the only FAST piece is the fused edge kernel inside the episodes, where the bound is 1e-12).
"""
import ctypes as C

import numpy as np
import pytest

from maestro_b200 import Fab, abi, face_fabs
from synth import make_episode_extras, make_state, relerr
from synth import same as _same_bits

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

K_B, N_A = 1.3806488e-16, 6.02214129e23  # Source/constants_cgs.f90:15,24
AION = [4.0, 12.0, 16.0]
ZION = [2.0, 6.0, 8.0]


def make_eos(nspec=3, gamma=5.0 / 3.0, neutral=True, small_temp=5.0e6, **lim):
    e = abi.mgpu_eos()
    e.kind, e.assume_neutral, e.nspec = abi.EOS_GAMMA_LAW, int(neutral), nspec
    e.gamma, e.k_B, e.n_A = gamma, K_B, N_A
    for k in ("mintemp", "mindens", "mine", "minp", "minh"):  # eos_type.f90:42-57
        setattr(e, k, lim.get(k, 1.0e-200))
    for k in ("maxtemp", "maxdens", "maxe", "maxp", "maxh"):
        setattr(e, k, lim.get(k, 1.0e200))
    e.small_temp = small_temp
    for n in range(nspec):
        e.aion[n], e.zion[n] = AION[n % 3], ZION[n % 3]
    return e


@pytest.fixture
def eos_on(oracle):
    """EOS set in the oracle (and in the CUDA library when there is one) for the duration of a test"""
    libs = [oracle]
    try:
        import torch

        if torch.cuda.is_available():
            from maestro_b200 import lib

            libs.append(lib.ops())
    except Exception:
        pass

    def setter(e):
        for o in libs:
            o.set_eos(e)

    yield setter
    for o in libs:
        o.set_eos(None)


def points(n=257, seed=3, nspec=3):
    rng = np.random.default_rng(seed)
    xn = rng.uniform(0.05, 1.0, size=(n, nspec))
    xn /= xn.sum(axis=1, keepdims=True)
    return dict(rho=10.0 ** rng.uniform(-3, 7, n), T=10.0 ** rng.uniform(5, 9.5, n)), xn


# ---- CPU: the oracle against the closed forms -------------------------------------------------------------------------
@pytest.mark.parametrize("neutral", [True, False])
@pytest.mark.parametrize("gamma", [5.0 / 3.0, 1.4])
def test_oracle_eos_ideal_gas_identities(oracle, eos_on, neutral, gamma):
    eos_on(make_eos(neutral=neutral, gamma=gamma))
    st, xn = points()
    r = oracle.eos_eval(abi.EOS_INPUT_RT, st, xn)
    a, z = np.array(AION), np.array(ZION)
    abar = 1.0 / (xn / a).sum(axis=1)
    mu = abar if neutral else 1.0 / ((1.0 + z) * xn / a).sum(axis=1)
    p = st["rho"] * K_B * st["T"] * N_A / mu
    assert relerr(r["abar"], abar) < 1e-14 and relerr(r["mu"], mu) < 1e-14
    assert relerr(r["zbar"], abar * (xn * z / a).sum(axis=1)) < 1e-14
    assert relerr(r["p"], p) < 1e-13
    assert relerr(r["e"], p / ((gamma - 1.0) * st["rho"])) < 1e-13
    assert relerr(r["h"], gamma / (gamma - 1.0) * p / st["rho"]) < 1e-13
    assert relerr(r["cs"], np.sqrt(gamma * p / st["rho"])) < 1e-13
    assert relerr(r["cp"], gamma * r["cv"]) < 1e-15 and relerr(r["cv"], r["e"] / st["T"]) < 1e-15
    assert relerr(r["dpdr"], p / st["rho"]) < 1e-13 and np.all(r["dedr"] == 0.0)
    # hydrogen at 1 g/cc, 1e6 K: p = R T = 8.3144621e13 erg/cc (k_B * N_A of CODATA 2010)
    e1 = make_eos(nspec=1)
    e1.aion[0], e1.zion[0] = 1.0, 1.0
    eos_on(e1)
    r1 = oracle.eos_eval(abi.EOS_INPUT_RT, dict(rho=np.ones(1), T=np.full(1, 1.0e6)), np.ones((1, 1)))
    assert abs(r1["p"][0] / 8.3144621e13 - 1.0) < 1e-8  # R is quoted to 8 digits


@pytest.mark.parametrize("mode", ["rh", "tp", "rp", "re"])
def test_oracle_eos_inverse_modes_round_trip(oracle, eos_on, mode):
    eos_on(make_eos(neutral=False))
    st, xn = points(seed=5)
    f = oracle.eos_eval(abi.EOS_INPUT_RT, st, xn)
    inp = dict(rh=abi.EOS_INPUT_RH, tp=abi.EOS_INPUT_TP, rp=abi.EOS_INPUT_RP, re=abi.EOS_INPUT_RE)[mode]
    given = {"rh": dict(rho=f["rho"], h=f["h"]), "tp": dict(T=f["T"], p=f["p"]), "rp": dict(rho=f["rho"], p=f["p"]),
             "re": dict(rho=f["rho"], e=f["e"])}[mode]
    if "T" not in given:
        given["T"] = np.full_like(f["T"], 1.0e7)  # the initial guess the callers pass; unused by a gamma law
    if "rho" not in given:
        given["rho"] = np.ones_like(f["rho"])
    b = oracle.eos_eval(inp, given, xn)
    for k in ("rho", "T", "p", "e", "h", "cs"):
        assert relerr(b[k], f[k]) < 1e-13, k


def test_oracle_eos_reset_inputs(oracle, eos_on):
    """eos.F90:129-183: rho and T are clamped; an out-of-range p / h / e falls back to (clamped rho, T) (eos_reset :291)"""
    eos_on(make_eos(mintemp=1.0e4, maxtemp=1.0e9, mindens=1.0e-2, maxdens=1.0e6, minp=1.0e10, maxp=1.0e20))
    xn = np.array([[0.5, 0.3, 0.2]] * 3)
    r = oracle.eos_eval(abi.EOS_INPUT_RT, dict(rho=np.array([1e-5, 1.0, 1e9]), T=np.array([1.0, 1e7, 1e12])), xn)
    assert list(r["rho"]) == [1e-2, 1.0, 1e6] and list(r["T"]) == [1e4, 1e7, 1e9]
    bad = oracle.eos_eval(abi.EOS_INPUT_RP, dict(rho=np.array([1.0] * 3), T=np.array([2e7] * 3), p=np.array([1.0, 1e15, 1e30])), xn)
    ref = oracle.eos_eval(abi.EOS_INPUT_RT, dict(rho=np.array([1.0]), T=np.array([2e7])), xn[:1])
    assert bad["T"][0] == 2e7 and bad["T"][2] == 2e7 and bad["p"][0] == ref["p"][0] and bad["p"][2] == ref["p"][0]
    assert abs(bad["p"][1] / 1e15 - 1.0) < 1e-14


def test_oracle_eos_errors(oracle, eos_on):
    st, xn = points(n=4)
    with pytest.raises(Exception, match="EOS: not initialized"):
        oracle.eos_eval(abi.EOS_INPUT_RT, st, xn)
    eos_on(make_eos())
    for inp, msg in ((abi.EOS_INPUT_PH, "eos_input_ph not implemented"), (abi.EOS_INPUT_TH, "not a valid input"),
                     (abi.EOS_INPUT_PS, "entropy"), (9, "invalid input")):
        with pytest.raises(Exception, match=msg):
            oracle.eos_eval(inp, st, xn)
    bad = make_eos()
    bad.gamma = -1.0
    with pytest.raises(Exception, match="eos_gamma cannot be < 0"):
        oracle.set_eos(bad)


def eos_state(dm, n, **kw):
    """synth.make_state with a physical temperature field and a density that dips below the cutoff in places"""
    st = make_state(dm, list(n), **kw)
    p, s = st["p"], st["s"]
    rng = np.random.default_rng(77)
    s.a[p.temp_comp - 1] = 10.0 ** rng.uniform(6.5, 8.5, size=s.a[0].shape)
    p.base_cutoff_density = 1.2  # the synthetic density spans 0.7 .. 2.3
    return st


def test_oracle_update_scal_reset_matches_pointwise_eos(oracle, eos_on):
    """update_scal.f90:421-447 restated on top of the pointwise EOS: rhoh = rho * h(rho, p0_new(k), X) below the cutoff"""
    eos_on(make_eos())
    st = eos_state(3, (10, 8, 9), do_eos_h_above_cutoff=1)
    p = st["p"]
    sold, snew = st["s"].clone(), st["s"].clone()
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3)
    force = st["force"].clone()
    force.a[...] = 0.0
    p0_new = 1.0e16 * np.exp(-np.arange(p.nr) * 0.1)
    oracle.update_scal(p, p.rhoh_comp, p.rhoh_comp, sold, snew, sflux, force, p0_new=p0_new)
    v, vo = snew.valid(), sold.valid()
    rho = v[p.rho_comp - 1]
    below = rho <= p.base_cutoff_density
    assert 0 < below.sum() < below.size
    assert same(v[p.rhoh_comp - 1][~below], vo[p.rhoh_comp - 1][~below])
    kk = np.broadcast_to(np.arange(rho.shape[0])[:, None, None], rho.shape)[below]
    xn = np.stack([v[p.spec_comp - 1 + c][below] / rho[below] for c in range(p.nspec)], axis=1)
    r = oracle.eos_eval(abi.EOS_INPUT_RP, dict(rho=rho[below], T=vo[p.temp_comp - 1][below], p=p0_new[kk]), xn)
    assert same(v[p.rhoh_comp - 1][below], rho[below] * r["h"])


def test_update_scal_without_eos_fails_below_cutoff(oracle):
    st = eos_state(2, (12, 9), do_eos_h_above_cutoff=1)
    p = st["p"]
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2)
    with pytest.raises(Exception, match="below base_cutoff_density"):
        oracle.update_scal(p, p.rhoh_comp, p.rhoh_comp, st["s"].clone(), st["s"].clone(), sflux, st["force"].clone())


# ---- GPU: the CUDA library against the oracle ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("inp", [abi.EOS_INPUT_RT, abi.EOS_INPUT_RH, abi.EOS_INPUT_TP, abi.EOS_INPUT_RP, abi.EOS_INPUT_RE])
@pytest.mark.parametrize("neutral", [True, False])
def test_eos_eval(gpu_ops, oracle, eos_on, inp, neutral):
    eos_on(make_eos(neutral=neutral, mindens=1.0e-2, maxdens=1.0e6, minp=1.0e8, maxp=1.0e24, minh=1.0e12, maxh=1.0e18,
                    mine=1.0e12, maxe=1.0e18))
    st, xn = points(n=1000, seed=11)
    full = oracle.eos_eval(abi.EOS_INPUT_RT, st, xn)
    given = {k: full[k] for k in ("rho", "T", "p", "e", "h")}
    given["T"] = np.full_like(full["T"], 3.0e7)
    if inp == abi.EOS_INPUT_RT:
        given = st
    a, b = gpu_ops.eos_eval(inp, given, xn), oracle.eos_eval(inp, given, xn)
    for k in abi.EOS_Q:
        assert same(a[k], b[k]), k


@pytest.mark.gpu
def test_eos_errors_gpu(gpu_ops, eos_on):
    st, xn = points(n=4)
    gpu_ops.set_eos(None)
    with pytest.raises(Exception, match="EOS: not initialized"):
        gpu_ops.eos_eval(abi.EOS_INPUT_RT, st, xn)
    eos_on(make_eos())
    for inp, msg in ((abi.EOS_INPUT_PH, "eos_input_ph not implemented"), (abi.EOS_INPUT_TH, "not a valid input"),
                     (abi.EOS_INPUT_PS, "entropy"), (9, "invalid input")):
        with pytest.raises(Exception, match=msg):
            gpu_ops.eos_eval(inp, st, xn)


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (23, 14)), (3, (13, 9, 11))])
def test_update_scal_eos_reset(gpu_ops, oracle, eos_on, dm, n):
    st = eos_state(dm, n, do_eos_h_above_cutoff=1)
    p = st["p"]
    rng = np.random.default_rng(5)
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, dm)
    for f in sflux:
        f.a[...] = rng.uniform(-1, 1, size=f.shape)
    p0_new = 1.0e16 * np.exp(-np.arange(p.nr) * 0.1)
    # without an EOS the call fails instead of skipping the reset
    with pytest.raises(Exception, match="needs an EOS"):
        gpu_ops.update_scal(p, p.rhoh_comp, p.rhoh_comp, st["s"].clone(), st["s"].clone(), sflux, st["force"].clone())
    eos_on(make_eos())
    with pytest.raises(Exception, match="needs p0_new"):
        gpu_ops.update_scal(p, p.rhoh_comp, p.rhoh_comp, st["s"].clone(), st["s"].clone(), sflux, st["force"].clone())
    out = []
    for o in (gpu_ops, oracle):
        sold, snew = st["s"].clone(), st["s"].clone()
        o.update_scal(p, p.rhoh_comp, p.rhoh_comp, sold, snew, sflux, st["force"].clone(), p0_new=p0_new)
        out.append(snew)
    assert same(out[0].a, out[1].a)
    assert not same(out[1].valid()[p.rhoh_comp - 1], st["s"].valid()[p.rhoh_comp - 1])


def _edge_states(st, rng, ng=1):
    p, dm = st["p"], st["dm"]
    sedge = face_fabs(st["lo"], st["hi"], ng, p.nscal, dm)
    for f in sedge:
        f.a[...] = rng.uniform(0.2, 1.0, size=f.shape)
        f.a[p.temp_comp - 1] = 10.0 ** rng.uniform(6.0, 8.5, size=f.a[0].shape)  # some below small_temp
    return sedge


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (21, 13)), (3, (12, 9, 10))])
@pytest.mark.parametrize("ept", [abi.PREDICT_T_THEN_RHOHPRIME, abi.PREDICT_T_THEN_H, abi.PREDICT_TPRIME_THEN_H])
@pytest.mark.parametrize("spt", [1, 2, 3])
def test_make_h_from_rhot_edge(gpu_ops, oracle, eos_on, dm, n, ept, spt):
    eos_on(make_eos())
    st = make_state(dm, list(n), enthalpy_pred_type=ept, species_pred_type=spt)
    p, b = st["p"], st["base"]
    rng = np.random.default_rng(31)
    sedge0 = _edge_states(st, rng)
    nr = p.nr
    t0 = dict(c_old=1.0e7 + 1.0e6 * rng.uniform(size=nr), c_new=1.0e7 + 1.0e6 * rng.uniform(size=nr),
              e_old=1.0e7 + 1.0e6 * rng.uniform(size=nr + 1), e_new=1.0e7 + 1.0e6 * rng.uniform(size=nr + 1))
    edge = {k: rng.uniform(0.5, 1.5, size=nr + 1) for k in ("r_old", "r_new", "rh_old", "rh_new")}
    out = []
    for o in (gpu_ops, oracle):
        sedge = [f.clone() for f in sedge0]
        o.make_h_from_rhot_edge(p, sedge, b["rho0_old"], b["rhoh0_old"], t0["c_old"], edge["r_old"], edge["rh_old"],
                                t0["e_old"], b["rho0_new"], b["rhoh0_new"], t0["c_new"], edge["r_new"], edge["rh_new"],
                                t0["e_new"])
        out.append(sedge)
    for g, c, f0 in zip(out[0], out[1], sedge0):
        assert same(g.a, c.a)
        assert not same(c.a[p.rhoh_comp - 1], f0.a[p.rhoh_comp - 1])


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (21, 13)), (3, (12, 9, 10))])
def test_mktempforce(gpu_ops, oracle, eos_on, dm, n):
    eos_on(make_eos(neutral=False))
    st = eos_state(dm, n, enthalpy_pred_type=abi.PREDICT_T_THEN_H, phys_bc=None)
    p = st["p"]
    ex = make_episode_extras(st)
    out = []
    for o in (gpu_ops, oracle):
        f = st["force"].clone()
        o.mktempforce(p, f, st["umac"], st["s"], ex["thermal"], 1.0e16 * ex["p0_old"], ex["psi"], st["adv_bc"], st["pmask"])
        out.append(f)
    assert same(out[0].a, out[1].a)
    assert not same(out[1].a[p.temp_comp - 1], st["force"].a[p.temp_comp - 1])
    p.enthalpy_pred_type = abi.PREDICT_RHOH
    for o in (gpu_ops, oracle):
        with pytest.raises(Exception, match="should only call mkrhohforce when predicting T"):
            o.mktempforce(p, st["force"].clone(), st["umac"], st["s"], ex["thermal"], ex["p0_old"], ex["psi"], st["adv_bc"],
                          st["pmask"])


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (33, 20)), (3, (17, 12, 14))])
@pytest.mark.parametrize("case", ["velocity", "sound", "force", "divu", "still"])
def test_firstdt(gpu_ops, oracle, eos_on, dm, n, case):
    from synth import make_estdt_inputs

    eos_on(make_eos())
    speed = 0.0 if case in ("sound", "still") else 1.0
    I = make_estdt_inputs(dm, n, speed=speed, force_amp=1.0, divu_amp=1.0 if case == "divu" else 0.0)
    p = I["p"]
    rng = np.random.default_rng(13)
    s = I["s"]
    s.a[p.temp_comp - 1] = (0.0 if case == "still" else 1.0) * 10.0 ** rng.uniform(3.0, 5.0, size=s.a[0].shape)
    gpi = Fab(I["lo"], I["hi"], 1, dm, dm=dm)
    gpi.a[...] = (1.0e3 if case == "force" else 1.0e-3) * rng.uniform(-1, 1, size=gpi.shape)
    nr = p.nr
    rho0 = 1.0 + 0.1 * rng.uniform(size=nr)
    grav = -1.0e-3 * np.ones(nr)
    res = []
    for o in (gpu_ops, oracle):
        res.append(o.firstdt(p, I["u"], gpi, s, I["divU"], rho0, I["p0"], grav, I["gamma1bar"], 0.5, 0.1, 1.0e20,
                             use_soundspeed_firstdt=(case == "sound"), use_divu_firstdt=(case == "divu")))
    assert res[0] == res[1]
    assert 0.0 < res[1][0] < 1.0e20


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (21, 13)), (3, (12, 9, 10))])
@pytest.mark.parametrize("which", ["rhoh", "rhoh_e", "rhop", "rhop_update", "rhop_pprime"])
def test_make_t(gpu_ops, oracle, eos_on, dm, n, which):
    eos_on(make_eos(neutral=False))
    st = eos_state(dm, n, phys_bc=WALLS[dm])
    p = st["p"]
    st["s"].a[p.rhoh_comp - 1] *= 1.0e15
    st["s"].a[p.pi_comp - 1] = 1.0e13
    p0 = 1.0e16 * np.exp(-np.arange(p.nr) * 0.1)
    out = []
    for o in (gpu_ops, oracle):
        s = st["s"].clone()
        if which.startswith("rhoh"):
            o.make_t_from_rhoh(p, s, p0, st["adv_bc"], st["pmask"], use_eos_e_instead_of_h=(which == "rhoh_e"))
        else:
            o.make_t_from_rhop(p, s, p0, st["adv_bc"], st["pmask"], update_rhoh=(which == "rhop_update"),
                               use_pprime_in_tfromp=(which == "rhop_pprime"))
        out.append(s)
    assert same(out[0].a, out[1].a)
    assert not same(out[1].a[p.temp_comp - 1], st["s"].a[p.temp_comp - 1])


WALLS = {2: [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]],
         3: [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.SLIP_WALL], [abi.INLET, abi.OUTLET]]}


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (22, 15)), (3, (14, 9, 11))])
@pytest.mark.parametrize("ept", [abi.PREDICT_T_THEN_RHOHPRIME, abi.PREDICT_T_THEN_H, abi.PREDICT_TPRIME_THEN_H])
@pytest.mark.parametrize("which_step,spt,bcset", [(1, 1, "periodic"), (2, 2, "walls"), (2, 3, "periodic")])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_enthalpy_advance_temperature_predictions(gpu_ops, oracle, eos_on, dm, n, ept, which_step, spt, bcset, exact):
    """enthalpy_advance.f90:16 with the temperature-based predictions (mktempforce, T / T' edge states,
    makeHfromRhoT_edge) and the EOS reset below the cutoff in its update_scal"""
    from maestro_b200 import lib

    eos_on(make_eos())
    phys = None if bcset == "periodic" else WALLS[dm]
    st = eos_state(dm, n, phys_bc=phys, enthalpy_pred_type=ept, species_pred_type=spt, do_eos_h_above_cutoff=1,
                   ppm_type=2 if bcset == "walls" else 1)
    p, b = st["p"], st["base"]
    ex = make_episode_extras(st)
    rng = np.random.default_rng(21)
    sedge0 = _edge_states(st, rng)  # density / species edge states "left by density_advance"; one ghost layer (QUIRK :487)
    tempbar = 3.0e7 + 1.0e6 * rng.uniform(size=p.nr)
    p0_old, p0_new = 1.0e16 * ex["p0_old"], 1.0e16 * ex["p0_new"]
    lib.set_option("exact", exact)
    try:
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
                               b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], p0_old, p0_new, ex["psi"], ex["grav_old"],
                               ex["grav_nph"], st["adv_bc"], st["pmask"], tempbar=tempbar)
            out.append([sold, snew, force] + sedge + sflux + umac)
        names = ["sold", "snew", "force"] + ["sedge%d" % d for d in range(dm)] + ["sflux%d" % d for d in range(dm)] + \
                ["umac%d" % d for d in range(dm)]
        for nm, g, c in zip(names, *out):
            if exact:
                assert same(g.a, c.a), nm
            assert relerr(g.a, c.a) <= 1e-12, nm
        snew_v = out[1][1].valid()
        assert (snew_v[p.rho_comp - 1] <= p.base_cutoff_density).any()
    finally:
        lib.set_option("exact", 0)


@pytest.mark.gpu
def test_enthalpy_advance_temperature_prediction_needs_eos(gpu_ops, oracle):
    st = eos_state(2, (12, 9), enthalpy_pred_type=abi.PREDICT_T_THEN_H)
    p, b = st["p"], st["base"]
    ex = make_episode_extras(st)
    gpu_ops.set_eos(None)
    sedge = face_fabs(st["lo"], st["hi"], 1, p.nscal, 2)
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2)
    with pytest.raises(Exception, match="EOS: not initialized"):
        gpu_ops.enthalpy_advance(p, 1, st["s"].clone(), st["s"].clone(), sedge, sflux, st["force"].clone(), ex["thermal"],
                                 [u.clone() for u in st["umac"]], b["w0"], b["rho0_old"], b["rhoh0_old"], b["rho0_new"],
                                 b["rhoh0_new"], ex["p0_old"], ex["p0_new"], ex["psi"], ex["grav_old"], ex["grav_nph"],
                                 st["adv_bc"], st["pmask"], tempbar=np.ones(p.nr))


# ---- GPU, spherical ---------------------------------------------------------------------------------------------------------
def _sphr(oracle, ept, spt=1, ppm_type=1):
    from test_sphr_velocity_gpu import sphr_enthalpy_state

    st, p, g, rad, ex, sold = sphr_enthalpy_state(oracle, ept, ppm_type)
    p.species_pred_type = spt
    rng = np.random.default_rng(9)
    sold.a[p.temp_comp - 1] = 10.0 ** rng.uniform(6.5, 8.5, size=sold.a[0].shape)
    rad["tempbar"] = 3.0e7 + 1.0e6 * rng.uniform(size=g.c.nr_fine)
    rad["p0_old"] = 1.0e16 * rad["p0_old"]
    rad["p0_new"] = 1.0e16 * rad["p0_new"]
    return st, p, g, rad, ex, sold


@pytest.mark.gpu
@pytest.mark.parametrize("ept", [abi.PREDICT_T_THEN_RHOHPRIME, abi.PREDICT_T_THEN_H, abi.PREDICT_TPRIME_THEN_H])
@pytest.mark.parametrize("spt", [1, 2, 3])
def test_make_h_from_rhot_edge_sphr(gpu_ops, oracle, eos_on, ept, spt):
    eos_on(make_eos())
    st, p, g, rad, ex, sold = _sphr(oracle, ept, spt)
    sedge0 = _edge_states(st, np.random.default_rng(3), ng=0)
    out = []
    for o in (gpu_ops, oracle):
        sedge = [f.clone() for f in sedge0]
        o.make_h_from_rhot_edge_sphr(p, g, sedge, rad["rho0_old"], rad["rhoh0_old"], rad["tempbar"], rad["rho0_new"],
                                     rad["rhoh0_new"], 1.01 * rad["tempbar"], st["adv_bc"], st["pmask"])
        out.append(sedge)
    for a, b, f0 in zip(out[0], out[1], sedge0):
        assert same(a.a, b.a)
        assert not same(b.a[p.rhoh_comp - 1], f0.a[p.rhoh_comp - 1])


@pytest.mark.gpu
def test_mktempforce_firstdt_make_t_sphr(gpu_ops, oracle, eos_on):
    eos_on(make_eos())
    st, p, g, rad, ex, sold = _sphr(oracle, abi.PREDICT_T_THEN_H)
    res = []
    for o in (gpu_ops, oracle):
        f = st["force"].clone()
        o.mktempforce(p, f, st["umac"], sold, ex["thermal"], rad["p0_old"], rad["psi"], st["adv_bc"], st["pmask"], geom=g)
        res.append(f.a)
    assert same(res[0], res[1]) and np.abs(res[1][p.temp_comp - 1]).max() > 0.0
    rng = np.random.default_rng(2)
    u = Fab(st["lo"], st["hi"], 3, 3, dm=3)
    u.a[...] = rng.uniform(-1, 1, size=u.shape)
    gpi = Fab(st["lo"], st["hi"], 1, 3, dm=3)
    gpi.a[...] = rng.uniform(-1, 1, size=gpi.shape)
    divU = Fab(st["lo"], st["hi"], 1, 1, dm=3)
    divU.a[...] = rng.uniform(-30, 30, size=divU.shape)
    nr = g.c.nr_fine
    grav, gamma1bar = -1.0 - 0.1 * rng.uniform(size=nr), 1.4 + 0.1 * rng.uniform(size=nr)
    for divu in (False, True):
        r = [o.firstdt(p, u, gpi, sold, divU, rad["rho0_old"], rad["p0_old"], grav, gamma1bar, 0.5, 0.1, 1.0e20,
                       use_soundspeed_firstdt=True, use_divu_firstdt=divu, geom=g) for o in (gpu_ops, oracle)]
        assert r[0] == r[1] and 0.0 < r[1][0] < 1.0e20
    for kind in ("rhoh", "rhop"):
        res = []
        for o in (gpu_ops, oracle):
            s = sold.clone()
            s.a[p.rhoh_comp - 1] = np.abs(s.a[p.rhoh_comp - 1]) * 1.0e15
            if kind == "rhoh":
                o.make_t_from_rhoh(p, s, rad["p0_old"], st["adv_bc"], st["pmask"], geom=g)
            else:
                o.make_t_from_rhop(p, s, rad["p0_old"], st["adv_bc"], st["pmask"], update_rhoh=True, geom=g)
            res.append(s.a)
        assert same(res[0], res[1])


@pytest.mark.gpu
@pytest.mark.parametrize("ept,which_step,spt", [(abi.PREDICT_T_THEN_RHOHPRIME, 1, 1), (abi.PREDICT_T_THEN_H, 2, 2),
                                                (abi.PREDICT_TPRIME_THEN_H, 2, 3)])
@pytest.mark.parametrize("exact", [1, 0], ids=["exact", "fast"])
def test_enthalpy_advance_sphr_temperature_predictions(gpu_ops, oracle, eos_on, ept, which_step, spt, exact):
    from maestro_b200 import lib

    eos_on(make_eos())
    st, p, g, rad, ex, sold0 = _sphr(oracle, ept, spt)
    p.do_eos_h_above_cutoff = 1
    p.base_cutoff_density = float(np.median(sold0.valid()[p.rho_comp - 1]))
    lo, hi = st["lo"], st["hi"]
    sedge0 = _edge_states(st, np.random.default_rng(3), ng=0)
    lib.set_option("exact", exact)
    try:
        res = []
        for o in (gpu_ops, oracle):
            sold = sold0.clone()
            snew = sold.clone()
            umac = [u.clone() for u in st["umac"]]
            sedge = [f.clone() for f in sedge0]
            sflux = face_fabs(lo, hi, 0, p.nscal, 3)
            force = st["force"].clone()
            o.enthalpy_advance_sphr(p, g, which_step, sold, snew, sedge, sflux, force, ex["thermal"], umac, rad["w0"],
                                    st["w0mac"], rad["rho0_old"], rad["rhoh0_old"], rad["rho0_new"], rad["rhoh0_new"],
                                    rad["p0_old"], rad["p0_new"], rad["psi"], st["adv_bc"], st["pmask"],
                                    tempbar=rad["tempbar"])
            res.append([sold.a, snew.a, force.a] + [f.a for f in sedge] + [f.a for f in sflux] + [u.a for u in umac])
        names = ["sold", "snew", "force"] + ["sedge%d" % d for d in range(3)] + ["sflux%d" % d for d in range(3)] + \
                ["umac%d" % d for d in range(3)]
        for nm, a, b in zip(names, *res):
            if exact:
                assert same(a, b), nm
            assert relerr(a, b) <= 1e-12, nm
    finally:
        lib.set_option("exact", 0)
