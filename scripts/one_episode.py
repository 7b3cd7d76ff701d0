"""N device-resident density_advance episodes of the bench workload at n^3 (for ncu launch lists). GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import bench
from maestro_b200 import abi, lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
config = "c2"
ops = lib.init(0, use_torch_stream=True)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "config":
        config = v
    else:
        lib.set_option(k, int(v))
w = bench.build_workload(config, n, "cuda:0")
w.p.mem_space = abi.DEVICE
for q in w.extra.get("params", []):
    q.mem_space = abi.DEVICE
for _ in range(reps):
    w.reset()
    w.step(ops)
torch.cuda.synchronize()
print("launches per step:", lib.launch_count() // reps)
