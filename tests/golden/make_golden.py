"""Generates tests/golden/test_advect_norms.json from the CPU oracle (run once; committed).

The reference cannot be built in this environment (no Fortran compiler, FBoxLib not vendored), so
these vectors are oracle outputs pinned for regression, next to the archived known answers of
Exec/UNIT_TESTS/test_advect/advect_3d_report_example.out (recorded, soft)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402

ops = oracle_lib.load()
cases = {}
for dm, n, stop in [(2, 64, 0.5), (3, 24, 0.25)]:
    for t in (0, 1, 2):
        for dr in (1, -dm):
            a, r = oracle_lib.test_advect(ops, dm, n, t, 0, dr, stop_time=stop)
            cases["dm%d_n%d_ppm%d_dir%+d" % (dm, n, t, dr)] = dict(dm=dm, n=n, ppm_type=t, dir=dr, stop_time=stop,
                                                                 abs=a, rel=r)
full = {}
if "--full" in sys.argv:  # the reference's own configuration: 64^3, cfl 0.7, t = 1 (minutes of CPU)
    for t in (0, 1, 2):
        a, r = oracle_lib.test_advect(ops, 3, 64, t, 0, 1, stop_time=1.0)
        full["ppm%d" % t] = dict(abs=a, rel=r)
else:
    old = json.load(open(os.path.join(HERE, "test_advect_norms.json")))
    full = old.get("oracle_3d_64", {})
old = json.load(open(os.path.join(HERE, "test_advect_norms.json")))
maxnorm = old.get("oracle_3d_64_maxnorm", {})
if "--full" in sys.argv:
    import numpy as np
    for t in (0, 1, 2):
        ri, rf = oracle_lib.test_advect_fields(ops, 3, 64, t, 1)
        maxnorm["ppm%d" % t] = float(np.abs(rf - ri).max())
out = dict(
    cases=cases,
    oracle_3d_64_maxnorm=maxnorm,
    oracle_3d_64=full,
    archived_3d_64=dict(ppm0=0.135411700899960, ppm1=0.105604113268602, ppm2=4.140496304475560e-2,
                        source="Exec/UNIT_TESTS/test_advect/advect_3d_report_example.out: fcompare max-norm of rho_final - rho_init; ppm0 and ppm2 reproduced to 1e-10 (tests/test_oracle_cpu.py), ppm1 is from an older reconstruction"),
)
json.dump(out, open(os.path.join(HERE, "test_advect_norms.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(cases), "cases")
