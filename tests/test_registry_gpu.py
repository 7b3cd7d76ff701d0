"""Residency registry (include/maestro_b200.h): registered host multifabs keep a device mirror between
calls; only stale components cross PCIe, outputs stay on the device until they are downloaded."""
import numpy as np
import pytest

from maestro_b200 import Fab, abi, face_fabs
from synth import make_state, same

pytestmark = pytest.mark.gpu

WALLS_3D = [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]


def _fabs(st, oracle):
    p, dm = st["p"], st["dm"]
    sold = st["s"].clone()
    oracle.fill_boundary(p, sold, 1, dm + 1, p.nscal, st["adv_bc"], st["pmask"])
    f = dict(sold=sold, snew=sold.clone(), umac=[u.clone() for u in st["umac"]],
             sedge=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm), sflux=face_fabs(st["lo"], st["hi"], 0, p.nscal, dm),
             force=st["force"].clone(), eta=Fab(st["lo"], st["hi"], 0, 1, nodal=[0] * (dm - 1) + [1], dm=dm))
    f["snew"].a[...] = 0.0
    return f


def _all(f):
    return [f["sold"], f["snew"], f["force"], f["eta"]] + f["umac"] + f["sedge"] + f["sflux"]


def _run(ops, st, f, step):
    p, b = st["p"], st["base"]
    ops.density_advance(p, step, f["sold"], f["snew"], f["sedge"], f["sflux"], f["force"], f["umac"], b["w0"], f["eta"],
                        b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])


@pytest.mark.parametrize("bcset,exact", [("periodic", 0), ("walls", 1)])
def test_registered_fabs_give_the_same_results_and_stay_on_the_device(gpu_ops, oracle, bcset, exact):
    from maestro_b200 import lib

    lib.set_option("exact", exact)
    st = make_state(3, 24, phys_bc=None if bcset == "periodic" else WALLS_3D, ppm_type=1)
    st["p"].rel_eps = 1e-8
    try:
        # two steps with plain host pointers: every output comes back after every call
        ref = _fabs(st, oracle)
        _run(gpu_ops, st, ref, 1)
        ref["sold"].a[...] = ref["snew"].a
        _run(gpu_ops, st, ref, 2)
        # the same two steps with every multifab registered
        f = _fabs(st, oracle)
        for x in _all(f):
            lib.register(x, pin=False)
        lib.copy_bytes(reset=True)
        _run(gpu_ops, st, f, 1)
        h2d1, d2h1 = lib.copy_bytes(reset=True)
        assert d2h1 == 0                       # nothing came back on its own
        assert np.all(f["snew"].a == 0.0)      # the host copy is untouched until it is asked for
        lib.download(f["snew"])
        f["sold"].a[...] = f["snew"].a         # the host advances the state ...
        lib.invalidate(f["sold"])              # ... and says so
        lib.copy_bytes(reset=True)
        _run(gpu_ops, st, f, 2)
        h2d2, d2h2 = lib.copy_bytes(reset=True)
        assert d2h2 == 0
        # second step: only the state the host changed travels (umac, force, sedge, ... are current on the device)
        assert 0 < h2d2 <= f["sold"].a.nbytes
        assert h2d2 < h2d1
        for x in _all(f):
            lib.download(x)
        for k in ("sold", "snew", "force", "eta"):
            assert same(f[k].a, ref[k].a), k
        for k in ("umac", "sedge", "sflux"):
            for d in range(3):
                assert same(f[k][d].a, ref[k][d].a), (k, d)
        # a second download moves nothing
        lib.copy_bytes(reset=True)
        lib.download(f["snew"])
        assert lib.copy_bytes()[1] == 0
    finally:
        lib.set_option("exact", 0)
        for x in _all(f):
            try:
                lib.unregister(x)
            except Exception:
                pass


def test_registry_errors(gpu_ops):
    from maestro_b200 import lib

    a = Fab([0, 0, 0], [7, 7, 7], 1, 2, dm=3)
    with pytest.raises(RuntimeError, match="not registered"):
        lib.download(a)
    lib.register(a, pin=False)
    try:
        with pytest.raises(RuntimeError, match="already registered"):
            lib.register(a, pin=False)
    finally:
        lib.unregister(a)
    with pytest.raises(RuntimeError, match="not registered"):
        lib.unregister(a)
