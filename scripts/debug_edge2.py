"""Compare the upwind-first fused kernel with the oracle and print where they differ. GPU box only."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from maestro_b200 import face_fabs, lib
from synth import make_state

ops = lib.init(0)
oracle = oracle_lib.load()
shape = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (37, 9, 11)
kchunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4
slow = len(sys.argv) > 3 and sys.argv[3] == "slow"
for ppm in (0, 1, 2):
    st = make_state(3, shape, ppm_type=ppm)
    p = st["p"]
    umax = max(np.abs(u.a).max() for u in st["umac"])
    p.rel_eps = 1e-8 * umax
    if slow:
        for u in st["umac"]:
            u.a[np.abs(u.a) < 0.15 * umax] = 0.0
        p.rel_eps = 0.3 * umax
    lib.set_option("exact", 0); lib.set_option("fused_variant", 1); lib.set_option("kchunk", kchunk)
    out = []
    for o in (ops, oracle):
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3, fill=-777.0)
        o.make_edge_scal(p, st["s"], sedge, st["umac"], st["force"], st["adv_bc"], False, 1, 4, 1, False)
        out.append(sedge)
    for d in range(3):
        g, c = out[0][d].a[0], out[1][d].a[0]
        err = np.abs(g - c) / np.abs(c).max()
        bad = np.argwhere(err > 1e-12)
        print("ppm %d dir %d: max rel err %.3e, bad %d of %d" % (ppm, d, err.max(), len(bad), err.size))
        if len(bad):
            print("   k range", bad[:, 0].min(), bad[:, 0].max(), " j range", bad[:, 1].min(), bad[:, 1].max(),
                  " i range", bad[:, 2].min(), bad[:, 2].max())
            print("   first:", bad[:8].tolist())
    if ppm == 1 and len(sys.argv) > 4:
        for d in range(3):
            g, c = out[0][d].a[0], out[1][d].a[0]
            err = np.abs(g - c) / np.abs(c).max()
            for k in (0, 5):
                print("dir", d, "k", k, " rows j, cols i; X = bad, sign of umac[d] shown as +/-")
                for j in range(err.shape[1]):
                    print("".join("X" if err[k, j, i] > 1e-12 else "." for i in range(err.shape[2])))
        ng = 1
        for d in range(3):
            u = st["umac"][d].a[0]
            k = 5
            print("sign umac", d, "k", k, u.shape)
            for j in range(ng, u.shape[1] - ng):
                print("".join("+" if u[k + ng, j, i] > 0 else "-" for i in range(ng, u.shape[2] - ng)))
