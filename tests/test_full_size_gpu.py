"""Size-independent properties at BASELINE.json's full size (config C2: test_advect 3-D 256^3, ppm_type 1), where
the CPU oracle would take minutes: the CUDA path is checked against identities the algorithm guarantees.

  * conservation on the periodic domain: sum(snew) == sum(sold) for every conservatively updated component
    (the fluxes telescope, update_scal.f90:401-415);
  * the update is the divergence of the returned fluxes: snew == sold - dt * div(sflux) from the arrays the episode
    hands back (mkflux.f90:370 / update_scal.f90:370 chained through HBM);
  * density is the sum of the partial densities (update_scal.f90:453-460);
  * the fluxes are the edge states times the MAC velocity (mkflux.f90:415-431, species_pred_type 1);
  * translation symmetry: shifting every input by a number of cells that is not a multiple of any tile or chunk
    size shifts every output by the same amount, BIT FOR BIT (each zone's arithmetic does not depend on where its
    CTA tile or z chunk begins) -- this is what exercises tile seams and chunk seams at full size;
  * a uniform state in a discretely divergence-free flow stays uniform.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
N = 256
TOL = 1e-12


def _episode(gpu_ops, st, dev, shift=(0, 0, 0), uniform=False):
    """one device-resident density_advance at N^3 with the inputs rolled by `shift` cells (x, y, z); returns the
    episode's multifabs (torch tensors on the device)"""
    import torch

    import bench
    from maestro_b200 import abi

    p = st["p"]
    p.mem_space = abi.DEVICE
    e = bench.alloc_episode(st, dev)
    sx, sy, sz = shift
    if uniform:
        for c, v in ((p.rho_comp - 1, 1.5), (p.spec_comp - 1, 0.75), (p.spec_comp, 0.5), (p.spec_comp + 1, 0.25),
                     (p.trac_comp - 1, 2.0)):
            e["sold"].a[c] = v
        # u = u(y,z), v = v(x,z), w = w(x,y): every face difference along a face's own direction is exactly zero
        for d, u in enumerate(e["umac"]):
            u.a[...] = u.a.select(3 - d, 3).unsqueeze(3 - d)
    if any(shift):
        g = e["sold"].ng
        v = e["sold"].a[:, g:-g, g:-g, g:-g]
        v.copy_(torch.roll(v, shifts=(sz, sy, sx), dims=(1, 2, 3)))
        for d, u in enumerate(e["umac"]):
            per = [slice(None), slice(1, 1 + N), slice(1, 1 + N), slice(1, 1 + N)]  # the N periodic faces / cells
            w = u.a[tuple(per)]
            w.copy_(torch.roll(w, shifts=(sz, sy, sx), dims=(1, 2, 3)))
            hi = list(per)
            hi[3 - d] = slice(1 + N, 2 + N)  # face hi+1 is face lo
            lo = list(per)
            lo[3 - d] = slice(1, 2)
            u.a[tuple(hi)] = u.a[tuple(lo)]
    # the library runs on its own stream: the torch work above must have finished before it reads the fabs
    torch.cuda.synchronize()
    # ghost cells of every component as the reference's fills leave them on entry (periodic wrap)
    gpu_ops.fill_boundary(p, e["sold"], 1, 3 + 1, p.nscal, st["adv_bc"], st["pmask"])
    e["sold0"] = e["sold"].a.clone()
    bench.run_episode(gpu_ops, st, e)
    torch.cuda.synchronize()
    p.mem_space = abi.HOST
    return e


@pytest.fixture(scope="module")
def c2_state():
    import bench

    return bench.test_advect_state(N)


@pytest.fixture(scope="module")
def c2_run(gpu_ops, c2_state):
    from maestro_b200 import lib

    lib.set_option("exact", 0)
    return _episode(gpu_ops, c2_state, "cuda:0")


def _comps(p):
    return list(range(p.spec_comp - 1, p.spec_comp - 1 + p.nspec)) + [p.trac_comp - 1]


def test_c2_conservation_and_density_sum(c2_state, c2_run):
    p, e = c2_state["p"], c2_run
    g = 4
    so = e["sold0"][:, g:-g, g:-g, g:-g]
    sn = e["snew"].a[:, g:-g, g:-g, g:-g]
    for c in _comps(p) + [p.rho_comp - 1]:
        scale = float(so[c].abs().sum())
        assert abs(float(sn[c].sum() - so[c].sum())) <= TOL * scale, c
    r = p.rho_comp - 1
    spec = slice(p.spec_comp - 1, p.spec_comp - 1 + p.nspec)
    lhs = sn[r] - so[r]
    rhs = (sn[spec] - so[spec]).sum(0)
    assert float((lhs - rhs).abs().max()) <= TOL * float(so[r].abs().max())


def test_c2_update_is_the_divergence_of_the_returned_fluxes(c2_state, c2_run):
    p, e = c2_state["p"], c2_run
    g = 4
    so = e["sold0"][:, g:-g, g:-g, g:-g]
    sn = e["snew"].a[:, g:-g, g:-g, g:-g]
    fx, fy, fz = (f.a for f in e["sflux"])
    for c in _comps(p):
        div = (fx[c][:, :, 1:] - fx[c][:, :, :-1]) / p.dx[0] + (fy[c][:, 1:, :] - fy[c][:, :-1, :]) / p.dx[1] + \
              (fz[c][1:, :, :] - fz[c][:-1, :, :]) / p.dx[2]
        want = so[c] - p.dt * div
        assert float((sn[c] - want).abs().max()) <= TOL * float(so[c].abs().max()), c


def test_c2_fluxes_are_edge_states_times_velocity(c2_state, c2_run):
    """species_pred_type 1 with a zero base state: F = umac * (rho0_edge + rho'_edge) * X_edge (mkflux.f90:415-421);
    the tracers go through the same routine (density_advance.f90:287-292)"""
    p, e = c2_state["p"], c2_run
    r = p.rho_comp - 1
    for d in range(3):
        um = e["umac"][d].valid(0)
        se, sf = e["sedge"][d].a, e["sflux"][d].a
        for c in _comps(p):
            want = um * (se[r] * se[c])
            assert float((sf[c] - want).abs().max()) <= TOL * float(want.abs().max()), (d, c)


def test_c2_translation_symmetry_is_bitwise(gpu_ops, c2_state, c2_run):
    import torch

    shift = (37, 11, 5)  # not a multiple of the 14-cell tile interior, of 32, or of the z chunk
    e2 = _episode(gpu_ops, c2_state, "cuda:0", shift=shift)
    sx, sy, sz = shift
    g = 4
    a = torch.roll(c2_run["snew"].a[:, g:-g, g:-g, g:-g], shifts=(sz, sy, sx), dims=(1, 2, 3))
    b = e2["snew"].a[:, g:-g, g:-g, g:-g]
    p = c2_state["p"]
    for c in _comps(p) + [p.rho_comp - 1]:
        assert torch.equal(a[c], b[c]), "snew comp %d" % c
    for d in range(3):
        for name in ("sedge", "sflux"):
            x = c2_run[name][d].a[:, :N, :N, :N]  # the N periodic faces of each direction
            y = e2[name][d].a[:, :N, :N, :N]
            x = torch.roll(x, shifts=(sz, sy, sx), dims=(1, 2, 3))
            for c in _comps(p) + ([p.rho_comp - 1] if name == "sedge" else []):
                assert torch.equal(x[c], y[c]), "%s%d comp %d" % (name, d, c)


def test_c2_uniform_state_in_divergence_free_flow_is_preserved(gpu_ops, c2_state):
    e = _episode(gpu_ops, c2_state, "cuda:0", uniform=True)
    p = c2_state["p"]
    g = 4
    for c in _comps(p) + [p.rho_comp - 1]:
        v0 = float(e["sold0"][c, g, g, g])
        assert float((e["snew"].a[c, g:-g, g:-g, g:-g] - v0).abs().max()) <= 1e-13 * abs(v0), c
        for d in range(3):
            want = v0 if c != p.rho_comp - 1 else v0  # rho' = rho - 0
            if p.spec_comp - 1 <= c < p.spec_comp - 1 + p.nspec:
                want = v0 / 1.5  # the species are predicted as X = rhoX / rho
            assert float((e["sedge"][d].a[c] - want).abs().max()) <= 1e-13 * abs(want), (c, d)


def _max_rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-300)


def test_c2_fast_build_agrees_with_the_bit_identical_build(gpu_ops, c2_state, c2_run):
    """The FAST episode (upwind-first kernels, FMA, on-the-fly transforms) against the exact build (literal kernels,
    bit-identical to the oracle in the small-size parity tests) at 256^3: 1e-12 relative, max-norm per field."""
    from maestro_b200 import lib

    lib.set_option("exact", 1)
    try:
        ex = _episode(gpu_ops, c2_state, "cuda:0")
    finally:
        lib.set_option("exact", 0)
    p = c2_state["p"]
    comps = _comps(p) + [p.rho_comp - 1]
    for c in comps:
        assert _max_rel(c2_run["snew"].a[c], ex["snew"].a[c]) <= TOL, ("snew", c)
        for d in range(3):
            assert _max_rel(c2_run["sedge"][d].a[c], ex["sedge"][d].a[c]) <= TOL, ("sedge", d, c)
    for c in _comps(p):
        for d in range(3):
            assert _max_rel(c2_run["sflux"][d].a[c], ex["sflux"][d].a[c]) <= TOL, ("sflux", d, c)
    assert _max_rel(c2_run["eta"].a, ex["eta"].a) <= TOL


# ---- config C4: rt, 2-D 4096^2, ppm_type 2, periodic x, slip wall / outlet in y, nspec = 2 -----------------------
N4 = 4096


@pytest.fixture(scope="module")
def c4_state():
    from maestro_b200 import abi
    from synth import make_state

    walls = [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]
    return make_state(2, N4, phys_bc=walls, ppm_type=2, nspec=2, noise=0.0)


def _episode_2d(gpu_ops, st, exact, shift_x=0):
    import torch

    from maestro_b200 import Fab, abi, face_fabs, lib

    dev = "cuda:0"
    p, b = st["p"], st["base"]
    sold, snew = st["s"].to(dev), st["s"].to(dev)
    umac = [u.to(dev) for u in st["umac"]]
    if shift_x:
        g = sold.ng
        v = sold.a[:, :, g:-g, g:-g]
        v.copy_(torch.roll(v, shifts=shift_x, dims=3))
        u = umac[0].a[:, :, 1:-1, 1:1 + N4]  # periodic x-faces
        u.copy_(torch.roll(u, shifts=shift_x, dims=3))
        w = umac[1].a[:, :, 1:-1, 1:-1]
        w.copy_(torch.roll(w, shifts=shift_x, dims=3))
    umac[0].a[:, :, 1:-1, 1 + N4] = umac[0].a[:, :, 1:-1, 1]  # periodic in x: face hi+1 is face lo (always)
    sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2, device=dev)
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2, device=dev)
    force = st["force"].to(dev)
    eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[0, 1, 0], dm=2, device=dev)
    torch.cuda.synchronize()  # the library runs on its own stream
    p.mem_space = abi.DEVICE
    lib.set_option("exact", exact)
    try:
        gpu_ops.fill_boundary(p, sold, 1, 2 + 1, p.nscal, st["adv_bc"], st["pmask"])
        for d in range(2):
            gpu_ops.fill_boundary(p, umac[d], 1, 1, 1, st["adv_bc"], st["pmask"])
        sold0 = sold.a.clone()
        gpu_ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"], b["rho0_new"],
                                b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
        torch.cuda.synchronize()
    finally:
        p.mem_space = abi.HOST
        lib.set_option("exact", 0)
    return dict(sold0=sold0, snew=snew, sedge=sedge, sflux=sflux, eta=eta)


@pytest.fixture(scope="module")
def c4_run(gpu_ops, c4_state):
    return _episode_2d(gpu_ops, c4_state, exact=0)


def test_c4_fused_2d_kernel_agrees_with_the_bit_identical_build(gpu_ops, c4_state, c4_run):
    """config C4 at full size: the FAST episode (fused 2-D kernel) against the exact build (staged path,
    bit-identical to the oracle in the small-size parity tests): 1e-12 relative, max-norm per field."""
    ex = _episode_2d(gpu_ops, c4_state, exact=1)
    p = c4_state["p"]
    comps = _comps(p) + [p.rho_comp - 1]
    for c in comps:
        assert _max_rel(c4_run["snew"].a[c], ex["snew"].a[c]) <= TOL, ("snew", c)
        for d in range(2):
            assert _max_rel(c4_run["sedge"][d].a[c], ex["sedge"][d].a[c]) <= TOL, ("sedge", d, c)
    for c in _comps(p):
        for d in range(2):
            assert _max_rel(c4_run["sflux"][d].a[c], ex["sflux"][d].a[c]) <= TOL, ("sflux", d, c)


def test_c4_update_is_the_divergence_of_the_returned_fluxes(c4_state, c4_run):
    """walls and an outlet: the boundary fluxes are part of the returned arrays, so the identity holds cell by cell"""
    p = c4_state["p"]
    g = 4
    so = c4_run["sold0"][:, 0, g:-g, g:-g]
    sn = c4_run["snew"].a[:, 0, g:-g, g:-g]
    fx, fy = (f.a[:, 0] for f in c4_run["sflux"])
    for c in _comps(p):
        div = (fx[c][:, 1:] - fx[c][:, :-1]) / p.dx[0] + (fy[c][1:, :] - fy[c][:-1, :]) / p.dx[1]
        want = so[c] - p.dt * div
        assert float((sn[c] - want).abs().max()) <= TOL * float(so[c].abs().max()), c


def test_c4_translation_symmetry_in_x_is_bitwise(gpu_ops, c4_state, c4_run):
    import torch

    sh = 45  # not a multiple of the 30-cell tile interior
    e2 = _episode_2d(gpu_ops, c4_state, exact=0, shift_x=sh)
    p = c4_state["p"]
    g = 4
    for c in _comps(p) + [p.rho_comp - 1]:
        a = torch.roll(c4_run["snew"].a[c, 0, g:-g, g:-g], shifts=sh, dims=1)
        assert torch.equal(a, e2["snew"].a[c, 0, g:-g, g:-g]), ("snew", c)
        for d in range(2):
            a = torch.roll(c4_run["sedge"][d].a[c, 0, :, :N4], shifts=sh, dims=1)
            assert torch.equal(a, e2["sedge"][d].a[c, 0, :, :N4]), ("sedge", d, c)


# ---- config C3: reacting_bubble, 3-D 256^3, planar base state, periodic x/y, slip wall z-lo, outlet z-hi, ppm_type 2 --
def test_c3_fast_episodes_agree_with_the_bit_identical_build(gpu_ops):
    """density_advance and velocity_advance at C3's size and boundary conditions: FAST build (upwind-first kernel with
    the boundary-face rules, ppm_type 2) against the exact build, 1e-12 relative, max-norm per field."""
    import torch

    from maestro_b200 import Fab, abi, face_fabs, lib
    from synth import make_episode_extras, make_state, make_vel_state

    dev = "cuda:0"
    walls = [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]
    st = make_state(3, N, phys_bc=walls, ppm_type=2, noise=0.0)
    vs = make_vel_state(3, N, phys_bc=walls, ppm_type=2, noise=0.0)
    exv = make_episode_extras(vs)
    p, b, q = st["p"], st["base"], vs["p"]
    rho0 = 1.0 + 0.5 * np.exp(-(np.arange(q.nr) + 0.5) * q.dx[2] / 0.5)
    out = {}
    for exact in (0, 1):
        sold, snew = st["s"].to(dev), st["s"].to(dev)
        umac = [u.to(dev) for u in st["umac"]]
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3, device=dev)
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3, device=dev)
        force = st["force"].to(dev)
        eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[0, 0, 1], dm=3, device=dev)
        ut, unew, sv = vs["utilde"].to(dev), vs["utilde"].to(dev), st["s"].to(dev)
        gpi, rhohalf, sponge = exv["gpi"].to(dev), exv["rhohalf"].to(dev), exv["sponge"].to(dev)
        um2 = [u.to(dev) for u in st["umac"]]
        torch.cuda.synchronize()  # the library runs on its own stream
        p.mem_space = q.mem_space = abi.DEVICE
        lib.set_option("exact", exact)
        try:
            gpu_ops.fill_boundary(p, sold, 1, 3 + 1, p.nscal, st["adv_bc"], st["pmask"])
            gpu_ops.fill_boundary(p, sv, 1, 3 + 1, p.nscal, st["adv_bc"], st["pmask"])
            gpu_ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"],
                                    b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])
            gpu_ops.velocity_advance(q, ut, unew, sv, rhohalf, um2, gpi, vs["w0"], exv["w0_force"], rho0,
                                     exv["rho0_nph"], exv["grav_old"], exv["grav_nph"], sponge, vs["adv_bc"],
                                     vs["pmask"])
            torch.cuda.synchronize()
        finally:
            p.mem_space = q.mem_space = abi.HOST
            lib.set_option("exact", 0)
        g = 4
        out[exact] = dict(snew=snew.a[:, g:-g, g:-g, g:-g].clone(), unew=unew.valid().clone(),
                          **{"sedge%d" % d: sedge[d].a.clone() for d in range(3)},
                          **{"sflux%d" % d: sflux[d].a.clone() for d in range(3)})
        del sold, snew, umac, sedge, sflux, force, eta, ut, unew, sv, gpi, rhohalf, sponge, um2
        torch.cuda.empty_cache()
    adv = _comps(p) + [p.rho_comp - 1]
    for k in out[0]:
        a, e = out[0][k], out[1][k]
        comps = range(a.shape[0]) if k == "unew" else (adv if not k.startswith("sflux") else _comps(p))
        for c in comps:
            assert _max_rel(a[c], e[c]) <= TOL, (k, c)
