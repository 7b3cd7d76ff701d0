"""Device-resident timing of density_advance in 2-D at n^2 (SURVEY config C4: rt, ppm_type 2, periodic x, slip wall /
outlet in y). GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from maestro_b200 import Fab, abi, face_fabs, lib
from synth import make_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ops = lib.init(0, use_torch_stream=True)
if len(sys.argv) > 2:
    lib.set_option("tile2d", int(sys.argv[2]))
dev = "cuda:0"
WALLS = [[abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]
for ppm in (1, 2):
    st = make_state(2, n, phys_bc=WALLS, ppm_type=ppm, nspec=2, noise=0.0)
    p, b = st["p"], st["base"]
    p.mem_space = abi.DEVICE
    sold0 = st["s"].to(dev); sold = st["s"].to(dev); snew = st["s"].to(dev)
    umac0 = [u.to(dev) for u in st["umac"]]; umac = [u.to(dev) for u in st["umac"]]
    sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2, device=dev)
    sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 2, device=dev)
    force = st["force"].to(dev)
    eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[0, 1, 0], dm=2, device=dev)

    def run():
        ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"], b["rho0_new"],
                            b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])

    tot = 0.0
    for it in range(4):
        sold.a.copy_(sold0.a)
        for u, u0 in zip(umac, umac0):
            u.a.copy_(u0.a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        if it:
            tot += e0.elapsed_time(e1)
    t = tot / 3
    lib.profile(True)
    run()
    torch.cuda.synchronize()
    prof = lib.profile_get()
    lib.profile(False)
    print("   kernel classes (ms, launches) of one profiled episode:", {k: (round(v[0], 3), v[1]) for k, v in prof.items()})
    print("2-D walls ppm%d n=%d density_advance %8.3f ms  %.2e zone-updates/s (4 comps), launches %d"
          % (ppm, n, t, 4 * n * n / t * 1e3, lib.launch_count(reset=True) // 4), flush=True)
