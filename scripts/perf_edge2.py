"""Device-resident timing of the fused make_edge_scal kernels (literal vs upwind-first) at n^3. GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maestro_b200 import Fab, abi, face_fabs, lib, make_adv_bc, make_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops = lib.init(0, use_torch_stream=True)
dev = "cuda:0"
lo, hi = [0, 0, 0], [n - 1] * 3


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for ppm in (1, 2):
    p = make_params(3, n=[n, n, n], ppm_type=ppm)
    p.mem_space = abi.DEVICE
    g = torch.Generator(device=dev).manual_seed(1)
    s = Fab(lo, hi, 4, p.nscal, dm=3, device=dev)
    s.a.copy_(1.0 + torch.rand(s.shape, generator=g, device=dev, dtype=torch.float64))
    force = Fab(lo, hi, 1, p.nscal, dm=3, device=dev)
    umac = face_fabs(lo, hi, 1, 1, 3, device=dev)
    for u in umac:
        u.a.copy_(torch.rand(u.shape, generator=g, device=dev, dtype=torch.float64) - 0.5)
    sedge = face_fabs(lo, hi, 0, p.nscal, 3, device=dev)
    adv_bc = make_adv_bc(p, [[-1, -1]] * 3)
    p.dt = 0.7 / n
    p.rel_eps = 1e-8
    zones = n ** 3
    for variant, by in ((1, 8), (1, 1616)):
        lib.set_option("fused_variant", variant)
        lib.set_option("fused_by", by)
        for kchunk in (-1, 64):
            lib.set_option("kchunk", kchunk)
            t = timeit(lambda: ops.make_edge_scal(p, s, sedge, umac, force, adv_bc, False, 1, 4, 1, False))
            print("variant %d by %d ppm%d n=%d kchunk=%3d: %.3f ms/comp -> %.2f Gzone/s, %.0f GB/s (64 B/zone algorithmic)"
                  % (variant, by, ppm, n, kchunk, t, zones / t / 1e6, 64 * zones / t / 1e6), flush=True)
