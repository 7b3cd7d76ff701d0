"""average (Source/average.f90:24) and make_etarho_spherical (Source/make_eta.f90:256), SURVEY 8 f2.

CPU: the restated average against what it must return for fields whose average is known.  GPU: the CUDA library
against the oracle; the sums run in a different order (atomics / tree), so the bound is 1e-12 relative."""
import numpy as np
import pytest

from maestro_b200 import Fab, face_fabs
from synth import make_state, relerr


def nr_irreg_of(n):
    """initialize.f90:1269-1272 (full star, not an octant; domhi = upb(domain)+1): integer division, then truncation"""
    return int((3 * (n // 2 - 0.5) ** 2 - 0.75) / 2.0)


def sphr_state(n=(16, 16, 16)):
    from sphr_common import make_sphr_state

    return make_sphr_state(n)


def radius_of(f, p, g):
    from synth import cell_coords

    x, y, z = cell_coords(f, p)
    return np.sqrt((x + g.c.prob_lo[0] - g.c.center[0]) ** 2 + (y + g.c.prob_lo[1] - g.c.center[1]) ** 2 +
                   (z + g.c.prob_lo[2] - g.c.center[2]) ** 2)


@pytest.mark.parametrize("dm,n", [(2, (12, 9)), (3, (8, 7, 10))])
def test_oracle_average_planar(oracle, dm, n):
    st = make_state(dm, list(n))
    p, s = st["p"], st["s"]
    got = oracle.average(p, s, p.rhoh_comp)
    v = s.valid()[p.rhoh_comp - 1]
    want = v.mean(axis=(1, 2)) if dm == 3 else v[0].mean(axis=1)
    assert relerr(got, want) < 1e-14


def test_oracle_average_spherical_known_answers(oracle):
    st = sphr_state()
    p, g = st["p"], st["geom"]
    n = st["hi"][0] + 1
    phi = Fab(st["lo"], st["hi"], 1, 2, dm=3)
    r = radius_of(phi, p, g)
    phi.a[0] = 3.25
    phi.a[1] = 2.0 + 0.5 * r * r  # a function of the radius alone: the binning reproduces it on the bins, the
    got0 = oracle.average(p, phi, 1, geom=g, nr_irreg=nr_irreg_of(n), drdxfac=5)  # quadratic interpolation exactly
    got1 = oracle.average(p, phi, 2, geom=g, nr_irreg=nr_irreg_of(n), drdxfac=5)
    rc = g.r_cc_loc
    # radii inside the cube along the axes, away from the centre point (an extrapolation, average.f90:213)
    inside = (rc < 0.5 * n * p.dx[0]) & (rc > 2.0 * p.dx[0])
    assert np.all(got0[inside] == 3.25)
    assert relerr(got1[inside], (2.0 + 0.5 * rc * rc)[inside]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dm,n", [(2, (33, 21)), (3, (17, 12, 19))])
def test_average_planar(gpu_ops, oracle, dm, n):
    st = make_state(dm, list(n))
    p, s = st["p"], st["s"]
    a, b = gpu_ops.average(p, s, p.spec_comp + 1), oracle.average(p, s, p.spec_comp + 1)
    assert relerr(a, b) <= 1e-12 and np.abs(b).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,drdxfac", [((16, 16, 16), 5), ((20, 20, 20), 1)])
def test_average_spherical(gpu_ops, oracle, n, drdxfac):
    from sphr_common import make_sphr_state

    st = make_sphr_state(n, drdxfac=drdxfac)
    p, g = st["p"], st["geom"]
    s = st["s"]
    for comp in (p.rho_comp, p.trac_comp):
        a = gpu_ops.average(p, s, comp, geom=g, nr_irreg=nr_irreg_of(n[0]), drdxfac=drdxfac)
        b = oracle.average(p, s, comp, geom=g, nr_irreg=nr_irreg_of(n[0]), drdxfac=drdxfac)
        assert relerr(a, b) <= 1e-12 and np.abs(b).max() > 0
    for o in (gpu_ops, oracle):
        with pytest.raises(Exception, match="incomp out of range|outside"):
            o.average(p, s, p.nscal + 1, geom=g, nr_irreg=nr_irreg_of(n[0]))


@pytest.mark.gpu
def test_make_etarho_spherical(gpu_ops, oracle):
    st = sphr_state((18, 18, 18))
    p, g, rad = st["p"], st["geom"], st["rad"]
    lo, hi = st["lo"], st["hi"]
    rng = np.random.default_rng(6)
    normal = Fab(lo, hi, 1, 3, dm=3)
    oracle.make_normal(p, g, normal)
    snew = st["s"].clone()
    snew.a[...] *= 1.0 + 0.01 * rng.uniform(-1, 1, size=snew.shape)
    w0mac = face_fabs(lo, hi, 1, 1, 3)
    for f in w0mac:
        f.a[...] = 0.05 * rng.uniform(-1, 1, size=f.shape)
    res = [o.make_etarho_spherical(p, g, st["s"], snew, st["umac"], w0mac, rad["rho0_old"], rad["rho0_new"], normal,
                                   nr_irreg_of(18), drdxfac=5) for o in (gpu_ops, oracle)]
    for a, b in zip(*res):
        assert relerr(a, b) <= 1e-12 and np.abs(b).max() > 0
    assert res[1][0][0] == 0.0 and res[1][0][-1] == res[1][1][-1]
    p.spherical = 0
    for o in (gpu_ops, oracle):
        with pytest.raises(Exception, match="should not be called for plane-parallel"):
            o.make_etarho_spherical(p, g, st["s"], snew, st["umac"], w0mac, rad["rho0_old"], rad["rho0_new"], normal,
                                    nr_irreg_of(18))
