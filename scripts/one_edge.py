"""One fused make_edge_scal call at n^3 (for ncu captures). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maestro_b200 import Fab, abi, face_fabs, lib, make_adv_bc, make_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ppm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ops = lib.init(0, use_torch_stream=True)
dev = "cuda:0"
p = make_params(3, n=[n, n, n], ppm_type=ppm); p.mem_space = abi.DEVICE
lo, hi = [0, 0, 0], [n - 1] * 3
g = torch.Generator(device=dev).manual_seed(1)
s = Fab(lo, hi, 4, p.nscal, dm=3, device=dev)
s.a.copy_(1.0 + torch.rand(s.shape, generator=g, device=dev, dtype=torch.float64))
force = Fab(lo, hi, 1, p.nscal, dm=3, device=dev)
umac = face_fabs(lo, hi, 1, 1, 3, device=dev)
for u in umac:
    u.a.copy_(torch.rand(u.shape, generator=g, device=dev, dtype=torch.float64) - 0.5)
sedge = face_fabs(lo, hi, 0, p.nscal, 3, device=dev)
adv_bc = make_adv_bc(p, [[-1, -1]] * 3)
p.dt = 0.7 / n; p.rel_eps = 1e-8
if len(sys.argv) > 4: lib.set_option("kchunk", int(sys.argv[4]))
for _ in range(reps):
    ops.make_edge_scal(p, s, sedge, umac, force, adv_bc, False, 1, 4, 1, False)
torch.cuda.synchronize()
