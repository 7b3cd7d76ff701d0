#!/usr/bin/env python
"""bench.py -- advective zone-updates/s of the B200-native MAESTRO hot path (BASELINE.json metric).

A "step" is one density_advance episode (Source/density_advance.f90:20: edge-state prediction of
nspec species + rho' + tracer, fluxes, conservative update) over one synthetic box per GPU:
configs[1] of BASELINE.json, "test_advect 3D 256^3 uniform grid, ppm_type=1" (Gaussian density of
Exec/UNIT_TESTS/test_advect/test_advect.f90:58, 3 species + 1 tracer, ng_s=4, periodic).

    value : zones x advanced components / device time, inputs resident in HBM (CUDA events, max over ranks)
    e2e   : the same episode through the C ABI with HOST (pinned) buffers: H2D of the inputs and D2H of
            every output the frozen Fortran signature exposes are inside the timed region
    roofline : dominant kernel (fused edge-state kernel), algorithmic bytes / CUDA-event kernel time
    cpu_baseline / --impl reference : the restated reference algorithm (oracle/, C++ + OpenMP) on the
            host cores over a bounded sample box of the same workload.

Launch: `python bench.py --gpus 1 --steps K --warmup W`, or under torchrun for N > 1 (one rank per GPU,
weak scaling: one n^3 box per rank).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "advective zone-updates/sec (3D PPM)"
UNIT = "zone-updates/s"


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def test_advect_state(n, device=None, seed=67890, rank=0, world=1):
    """test_advect initial data + velocity set B/C of SURVEY 8d on this rank's n^3 slab of the periodic
    n x n x (n*world) domain (world = 1: the n^3 unit box of the reference test)."""
    from maestro_b200 import Fab, abi, face_fabs, make_adv_bc, make_params

    p = make_params(3, n=[n, n, n * world], ppm_type=1)
    p.base_cutoff_density = 1e-10
    lo, hi = [0, 0, rank * n], [n - 1, n - 1, rank * n + n - 1]
    W = float(np.float32(0.05))
    x = (np.arange(-4, n + 4) + 0.5) / n
    xp = ((x % 1.0) - 0.5) ** 2
    r2 = xp[None, None, :] + xp[None, :, None] + xp[:, None, None]
    rho = np.maximum(np.exp(-r2 / W ** 2), 1e-10)
    sold = Fab(lo, hi, 4, p.nscal, dm=3)
    sold.a[p.rho_comp - 1] = rho
    sold.a[p.spec_comp - 1] = 0.6 * rho
    sold.a[p.spec_comp] = 0.3 * rho
    sold.a[p.spec_comp + 1] = 0.1 * rho
    sold.a[p.trac_comp - 1] = np.sin(2 * np.pi * x)[None, None, :] * np.ones_like(rho)
    rng = np.random.default_rng(seed)
    umac = face_fabs(lo, hi, 1, 1, 3)
    for d, u in enumerate(umac):
        c = [(np.arange(-1, u.shape[3 - q] - 1) + (0.0 if q == d else 0.5)) / n for q in range(3)]
        # the fields are 1-periodic, so every slab of the taller domain sees the same (globally periodic) data
        X, Y, Z = c[0][None, None, :], c[1][None, :, None], c[2][:, None, None]
        o = [X, Y, Z]
        a, b = o[(d + 1) % 3], o[(d + 2) % 3]
        u.a[0] = np.sin(2 * np.pi * b) + np.cos(2 * np.pi * a)
        u.a[0] += rng.uniform(-0.1, 0.1, size=u.a[0].shape) * 0.0 + 0.05 * np.sin(2 * np.pi * 7 * (a + b))
    umax = max(np.abs(u.a).max() for u in umac)
    p.dt = 0.7 * p.dx[0] / umax
    p.rel_eps = 1e-8 * umax
    adv_bc = make_adv_bc(p, [[abi.PERIODIC, abi.PERIODIC]] * 3)
    zero_c, zero_e = np.zeros(n * world), np.zeros(n * world + 1)
    st = dict(p=p, lo=lo, hi=hi, sold=sold, umac=umac, adv_bc=adv_bc, pmask=[1, 1, 1],
              base=dict(w0=zero_e, rho0_old=zero_c, rho0_new=zero_c, p0=zero_c, rho0_predicted_edge=zero_e))
    return st


def alloc_episode(st, device):
    """all multifabs density_advance touches, on `device` (None = host numpy)"""
    from maestro_b200 import Fab, face_fabs

    p, lo, hi = st["p"], st["lo"], st["hi"]
    e = dict(sold=st["sold"].to(device), snew=Fab(lo, hi, 4, p.nscal, dm=3, device=device),
             umac=[u.to(device) for u in st["umac"]], sedge=face_fabs(lo, hi, 0, p.nscal, 3, device=device),
             sflux=face_fabs(lo, hi, 0, p.nscal, 3, device=device),
             force=Fab(lo, hi, 1, p.nscal, dm=3, device=device),
             eta=Fab(lo, hi, 0, 1, nodal=[0, 0, 1], dm=3, device=device))
    return e


def run_episode(ops, st, e):
    p, b = st["p"], st["base"]
    ops.density_advance(p, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["umac"], b["w0"], e["eta"],
                        b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])


def ncomp_advanced(p):
    return p.nspec + 1 + p.ntrac  # species + density + tracers are predicted to edges and updated


def cpu_reference(n_sample, steps, warmup):
    """restated reference algorithm (oracle) on the host cores over an n_sample^3 box of the workload"""
    import ctypes

    import oracle_lib

    oracle = oracle_lib.load()
    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    gomp = ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL)
    gomp.omp_set_num_threads(len(os.sched_getaffinity(0)))
    st = test_advect_state(n_sample)
    e = alloc_episode(st, None)
    for _ in range(warmup):
        run_episode(oracle, st, e)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_episode(oracle, st, e)
    dt = (time.perf_counter() - t0) / steps
    zu = n_sample ** 3 * ncomp_advanced(st["p"])
    return zu / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="zones per side of the per-GPU box")
    ap.add_argument("--e2e-steps", type=lambda v: max(1, int(v)), default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (mgpu_set_option), e.g. fused_by=1616")
    args = ap.parse_args()
    if os.environ.get("BENCH_WATCHDOG"):  # debugging aid: dump every thread's Python stack if the run stalls
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) or 1
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        # the reference's own CPU implementation of the path cannot be built here (no Fortran compiler,
        # FBoxLib not vendored): the arm times the restated algorithm (oracle/), OpenMP over all host cores.
        if rank != 0:
            return
        n_s = 96
        zups, dt = cpu_reference(n_s, max(args.steps, 1), max(args.warmup, 1))
        line = {"metric": METRIC, "value": zups, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "test_advect 3D %d^3 per GPU, ppm_type=1, density_advance episode" % args.n,
                           "sample": "%d^3 box of the same workload per step" % n_s},
                "cpu_baseline": {"value": zups, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "density_advance on a %d^3 box, OpenMP over %d threads" % (n_s, cores)},
                "e2e": {"value": zups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from maestro_b200 import abi, lib

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    ops = lib.init(local_rank, use_torch_stream=True)
    for kv in args.opt:
        key, val = kv.split("=")
        lib.set_option(key, int(val))
    n = args.n
    if world > 1:  # NCCL communicator of the library: halo exchange inside every ghost fill of the episode
        from maestro_b200 import slab

        slab.comm_init_from_torch(lib.load(), dev)
    st = test_advect_state(n, rank=rank, world=world)
    p = st["p"]
    ncomp = ncomp_advanced(p)
    zone_updates = n ** 3 * ncomp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident episode (value) --------------------------------------------
    p.mem_space = abi.DEVICE
    e = alloc_episode(st, dev)
    sold0 = e["sold"].a.clone()
    umac0 = [u.a.clone() for u in e["umac"]]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    def reset_inputs():
        e["sold"].a.copy_(sold0)
        for u, u0 in zip(e["umac"], umac0):
            u.a.copy_(u0)

    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~0.5 s to produce its first sample: start it before the warm-up
    for _ in range(W):
        reset_inputs()
        run_episode(ops, st, e)
    # keep the GPU under the same load for about a second, until the sampler is running.  Every rank must run the
    # SAME number of episodes (each one exchanges halos with its neighbours): the count is agreed on first
    torch.cuda.synchronize()
    t_spin = time.time()
    run_episode(ops, st, e)
    torch.cuda.synchronize()
    n_spin = torch.tensor([max(1, int(1.0 / max(time.time() - t_spin, 1e-4)))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_spin, op=dist.ReduceOp.MAX)
    for _ in range(min(int(n_spin.item()), 2000)):
        run_episode(ops, st, e)
    torch.cuda.synchronize()
    sampler.lines.clear()
    barrier()
    lib.launch_count(reset=True)
    lib.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        reset_inputs()      # untimed: restores the in-place-modified inputs
        flush.zero_()       # untimed: evicts L2 between timed iterations
        ev[k][0].record()
        run_episode(ops, st, e)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop()
    launches = lib.launch_count()
    prof = lib.profile_get()
    lib.profile(False)
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3  # seconds for K steps on this rank
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    value = world * zone_updates * args.steps / t_max

    # roofline of the dominant kernel class
    hbm, peak_src = peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
    roof = None
    if dom[0] is not None and dom[1][1] > 0:
        name, (ms, nl) = dom
        # algorithmic bytes per launch of one make_edge_scal component (3-D): read s 8 + force 8 + umac 24,
        # write sedge 24 = 64 B per zone (kernel table); other classes report their own figure
        bytes_per_zone = {"fused_edge": 64.0, "edge_transverse": 64.0, "edge_final": 64.0}.get(name, 64.0)
        achieved = bytes_per_zone * n ** 3 / (ms / nl * 1e-3) / 1e9
        traffic = None
        ncu_static = None
        try:  # DRAM bytes per launch of this kernel class from the committed ncu --set full capture (n = 256 only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if n == 256 and name in tj:
                traffic = tj[name]["bytes_per_launch"]
                ncu_static = tj[name].get("ncu")  # issue / fp64-pipe / DRAM utilisation of the same capture
        except Exception:
            traffic = None
        others = {}
        if "update" in prof and prof["update"][1] > 0:  # flux + update + density kernel: 360 B per zone
            ms_u = prof["update"][0] / args.steps
            others["update"] = {"achieved": 360.0 * n ** 3 / (ms_u * 1e-3) / 1e9, "frac": 360.0 * n ** 3 / (ms_u * 1e-3) / 1e9 / hbm,
                                "ms_per_step": ms_u, "algorithmic_bytes_per_zone": 360.0,
                                "note": "k_copy + k_flux_update3_fast, all species/tracers in one launch"}
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src, "other_kernels": others,
                "kernel_ms_per_launch": ms / nl, "kernel_share_of_step": (ms / args.steps) / (t_dev / args.steps * 1e3),
                "algorithmic_bytes_per_launch": bytes_per_zone * n ** 3,
                "episode_algorithmic_GBs": 368.0 * n ** 3 * args.steps / t_dev / 1e9,
                "episode_frac": 368.0 * n ** 3 * args.steps / t_dev / 1e9 / hbm,
                "kernel_classes_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "ncu_capture": ncu_static}
    del flush

    # ---------------- end to end through the C ABI with host buffers (e2e) ------------------------
    p.mem_space = abi.HOST
    del e
    torch.cuda.empty_cache()
    eh = alloc_episode(st, None)
    import ctypes as C

    def pin(f):  # pinned host memory so the copies run at full PCIe rate
        t = torch.from_numpy(f.a)
        lib.load().mgpu_host_register(C.c_void_p(f.ptr), f.a.size)
        return t

    keep = [pin(f) for f in [eh["sold"], eh["snew"], eh["force"], eh["eta"]] + eh["umac"] + eh["sedge"] + eh["sflux"]]
    sold_h0 = eh["sold"].a.copy()
    umac_h0 = [u.a.copy() for u in eh["umac"]]
    run_episode(ops, st, eh)  # warm-up (allocates the staging pool)
    barrier()
    lib.copy_bytes(reset=True)  # the library counts the bytes of every cudaMemcpyAsync it issues
    t_e2e = 0.0
    for _ in range(args.e2e_steps):
        eh["sold"].a[...] = sold_h0  # untimed: restore the in-place-modified host inputs
        for u, u0 in zip(eh["umac"], umac_h0):
            u.a[...] = u0
        barrier()
        t0 = time.perf_counter()
        run_episode(ops, st, eh)  # synchronous: H2D of inputs, kernels, D2H of outputs, stream sync
        t_e2e += time.perf_counter() - t0
    in_bytes, out_bytes = [b // args.e2e_steps for b in lib.copy_bytes()]
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * zone_updates * args.e2e_steps / float(te.item())
    snew_sum = float(eh["snew"].valid(0).sum())

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "test_advect 3D %d^3 per GPU, ppm_type=1, density_advance episode "
                                   "(%d comps: %d species + rho' + %d tracer), ng_s=4, periodic" % (n, ncomp, p.nspec,
                                                                                                  p.ntrac),
                       "zones_per_gpu": n ** 3, "components": ncomp, "l2": "256 MB flush between timed iterations",
                       "multi_gpu": ("periodic %d x %d x %d domain, one %d^3 slab per rank, NCCL send/recv halo exchange "
                                     "in every ghost fill (timed)" % (n, n, n * world, n)) if world > 1 else "single box"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "steps": args.e2e_steps, "check_sum_rho_new": snew_sum},
            "roofline": roof}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        zups, dt = cpu_reference(96, 2, 1)
        line["cpu_baseline"] = {"value": zups, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "density_advance on a 96^3 box of the same workload, 2 steps, OpenMP over "
                                          "%d threads (restated reference algorithm; gfortran+FBoxLib build "
                                          "impossible here)" % cores}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
