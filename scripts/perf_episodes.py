"""Device-resident timing of the four L4 episodes at n^3 (periodic and wall BCs, ppm_type 1 and 2). GPU box only.
Prints ms per episode and zone-updates/s (zones x components advanced)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from maestro_b200 import Fab, abi, face_fabs, lib
from synth import fill_face_ghosts, make_episode_extras, make_state, make_vel_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ops = lib.init(0, use_torch_stream=True)
dev = "cuda:0"
WALLS = [[abi.PERIODIC, abi.PERIODIC], [abi.PERIODIC, abi.PERIODIC], [abi.SLIP_WALL, abi.OUTLET]]


if os.environ.get("PERF_VARIANT"):  # 3: upwind-first kernel for every box and ppm_type, 0: literal kernel everywhere
    lib.set_option("fused_variant", int(os.environ["PERF_VARIANT"]))
for kv in os.environ.get("PERF_OPTS", "").split(","):  # e.g. "premac_fuse=0,exact=1"
    if kv:
        lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
EPISODES = os.environ.get("PERF_EPISODES", "density,enthalpy,premac,velocity").split(",")
ONLY = os.environ.get("PERF_ONLY")  # e.g. "periodic,1": one configuration only (for ncu launch lists)


def timeit(fn, reset, reps=3):
    lib.launch_count(reset=True)
    reset(); fn(); torch.cuda.synchronize()
    print("   (%d launches per call)" % lib.launch_count(reset=True), flush=True)
    tot = 0.0
    for _ in range(reps):
        reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def dv(f):
    return f.to(dev)


for bcname, phys in (("periodic", None), ("walls", WALLS)):
    for ppm in (1, 2):
        if ONLY and ONLY != "%s,%d" % (bcname, ppm):
            continue
        st = make_state(3, n, phys_bc=phys, ppm_type=ppm, noise=0.0)
        vs = make_vel_state(3, n, phys_bc=phys, ppm_type=ppm, noise=0.0)
        p, b = st["p"], st["base"]
        ex = make_episode_extras(st)
        p.mem_space = abi.DEVICE
        vs["p"].mem_space = abi.DEVICE
        zones = n ** 3
        # density_advance
        sold0 = dv(st["s"]); sold = dv(st["s"]); snew = dv(st["s"])
        umac0 = [dv(u) for u in st["umac"]]; umac = [dv(u) for u in st["umac"]]
        sedge = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3, device=dev)
        sflux = face_fabs(st["lo"], st["hi"], 0, p.nscal, 3, device=dev)
        force = dv(st["force"])
        eta = Fab(st["lo"], st["hi"], 0, 1, nodal=[0, 0, 1], dm=3, device=dev)

        def reset():
            sold.a.copy_(sold0.a)
            for u, u0 in zip(umac, umac0):
                u.a.copy_(u0.a)

        t = timeit(lambda: ops.density_advance(p, 1, sold, snew, sedge, sflux, force, umac, b["w0"], eta, b["rho0_old"],
                                               b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"]), reset)
        print("%-9s ppm%d n=%d density_advance  %8.3f ms  %.2e zone-updates/s (5 comps)" % (bcname, ppm, n, t, 5 * zones / t * 1e3), flush=True)
        # enthalpy_advance (density edge states left in sedge by the call above)
        thermal = dv(ex["thermal"])
        t = timeit(lambda: ops.enthalpy_advance(p, 1, sold, snew, sedge, sflux, force, thermal, umac, b["w0"], b["rho0_old"],
                                                b["rhoh0_old"], b["rho0_new"], b["rhoh0_new"], ex["p0_old"], ex["p0_new"],
                                                ex["psi"], ex["grav_old"], ex["grav_nph"], st["adv_bc"], st["pmask"]), reset)
        print("%-9s ppm%d n=%d enthalpy_advance %8.3f ms  %.2e zone-updates/s (1 comp)" % (bcname, ppm, n, t, zones / t * 1e3), flush=True)
        # advance_premac + velocity_advance
        q = vs["p"]
        exv = make_episode_extras(vs)
        ut = dv(vs["utilde"]); unew = dv(vs["utilde"]); sv = dv(st["s"])
        gpi = dv(exv["gpi"]); rhohalf = dv(exv["rhohalf"]); sponge = dv(exv["sponge"])
        um2 = face_fabs(vs["lo"], vs["hi"], 1, 1, 3, device=dev)
        rho0 = 1.0 + 0.5 * np.exp(-(np.arange(q.nr) + 0.5) * q.dx[2] / 0.5)
        t = timeit(lambda: ops.advance_premac(q, ut, sv, um2, gpi, vs["w0"], exv["w0_force"], rho0, exv["grav_old"],
                                              vs["adv_bc"], vs["phys_bc"], vs["pmask"]), lambda: None)
        print("%-9s ppm%d n=%d advance_premac   %8.3f ms  %.2e zone-updates/s (3 comps)" % (bcname, ppm, n, t, 3 * zones / t * 1e3), flush=True)
        t = timeit(lambda: ops.velocity_advance(q, ut, unew, sv, rhohalf, um2, gpi, vs["w0"], exv["w0_force"], rho0,
                                                exv["rho0_nph"], exv["grav_old"], exv["grav_nph"], sponge, vs["adv_bc"],
                                                vs["pmask"]), lambda: None)
        print("%-9s ppm%d n=%d velocity_advance %8.3f ms  %.2e zone-updates/s (3 comps)" % (bcname, ppm, n, t, 3 * zones / t * 1e3), flush=True)
        del sold0, sold, snew, umac0, umac, sedge, sflux, force, eta, thermal, ut, unew, sv, gpi, rhohalf, sponge, um2
        torch.cuda.empty_cache()
