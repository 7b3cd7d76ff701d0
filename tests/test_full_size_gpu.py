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
