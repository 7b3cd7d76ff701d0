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


PARAM_OVERRIDES = {}  # --bds: every leg (device, e2e, CPU arm) runs the same variant of the workload


def build_workload(config, n, device, rank=0, world=1):
    from maestro_b200 import workloads

    w = workloads.BUILDERS[config](n=n or workloads.DEFAULT_N[config], device=device, rank=rank, world=world)
    for k, v in PARAM_OVERRIDES.items():
        setattr(w.p, k, v)
    if PARAM_OVERRIDES.get("bds_type") == 1:
        w.desc["workload"] = w.desc["workload"].replace("ppm_type=1", "bds_type=1 (BDS edge states, Source/bds.f90)")
    return w


# ---- the c2 state as plain dicts (tests/test_full_size_gpu.py drives the episode on rolled / modified inputs) ---------
def test_advect_state(n, rank=0, world=1):
    import numpy as np

    w = build_workload("c2", n, None, rank=rank, world=world)
    e, p = w.extra["e"], w.p
    nz = n * world
    zero_c, zero_e = np.zeros(nz), np.zeros(nz + 1)
    return dict(p=p, lo=e["sold"].lo, hi=e["sold"].hi, sold=e["sold"], umac=e["umac"], adv_bc=w.extra["adv_bc"],
                pmask=[1, 1, 1], base=dict(w0=zero_e, rho0_old=zero_c, rho0_new=zero_c, p0=zero_c,
                                           rho0_predicted_edge=zero_e))


def alloc_episode(st, device):
    """all multifabs density_advance touches, on `device` (None = host numpy)"""
    from maestro_b200 import Fab, face_fabs

    p, lo, hi = st["p"], st["lo"], st["hi"]
    return dict(sold=st["sold"].to(device), snew=Fab(lo, hi, 4, p.nscal, dm=3, device=device),
                umac=[u.to(device) for u in st["umac"]], sedge=face_fabs(lo, hi, 0, p.nscal, 3, device=device),
                sflux=face_fabs(lo, hi, 0, p.nscal, 3, device=device), force=Fab(lo, hi, 1, p.nscal, dm=3, device=device),
                eta=Fab(lo, hi, 0, 1, nodal=[0, 0, 1], dm=3, device=device))


def run_episode(ops, st, e):
    p, b = st["p"], st["base"]
    ops.density_advance(p, 1, e["sold"], e["snew"], e["sedge"], e["sflux"], e["force"], e["umac"], b["w0"], e["eta"],
                        b["rho0_old"], b["rho0_new"], b["p0"], b["rho0_predicted_edge"], st["adv_bc"], st["pmask"])


def cpu_reference(config, n_sample, steps, warmup, budget_s=150.0):
    """restated reference algorithm (oracle) on the host cores over an n_sample box of the workload"""
    import ctypes

    import oracle_lib

    oracle = oracle_lib.load()
    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    gomp = ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL)
    gomp.omp_set_num_threads(len(os.sched_getaffinity(0)))
    if config == "c5":  # built with torch on the host for the CPU arm
        w = build_workload(config, n_sample, "cpu")
        ee = w.extra["e"]
        for f in [v for v in ee.values() if not isinstance(v, list)] + sum((v for v in ee.values() if isinstance(v, list)), []):
            f.device, f.a = None, f.a.numpy()  # same memory, numpy view: the oracle takes host pointers
    else:
        w = build_workload(config, n_sample, None)
    # bounded: a 256^3 step of the restated reference takes ~17 s on 16 threads, so the whole run (warm-up included) is
    # kept within `budget_s` by timing fewer steps than asked for once the first step has shown what one costs
    t_start = time.perf_counter()
    t, done, first = 0.0, 0, None
    k_warm, k_timed = warmup, steps
    i = 0
    while i < k_warm + k_timed:
        w.reset()
        t0 = time.perf_counter()
        w.step(oracle)
        d = time.perf_counter() - t0
        if first is None:
            first = d
            if first * (warmup + steps) > budget_s:
                k_warm = min(warmup, 1)
                k_timed = max(1, min(steps, int((budget_s - first) / first) - k_warm + 1))
        if i >= k_warm:
            t += d
            done += 1
        i += 1
    dt = t / done
    cpu_reference.steps_timed, cpu_reference.wall_s = done, time.perf_counter() - t_start
    return w.zone_updates / dt, dt, w


def per_zone_parity(ref, got):
    """per output field: max over zones of |got - ref| / max(|ref|, floor) with floor = 1e-10 x the field's max-norm
    (per-zone relative error; sign-changing fields reach it at their zero crossings), and the max-norm relative error
    max|got - ref| / max|ref| the tests assert"""
    import torch

    zone, norm = {}, {}
    for k in ref:
        a, b = got[k].double(), ref[k].double()
        mx = float(b.abs().max())
        err = (a - b).abs()
        norm[k] = float(err.max()) / mx if mx > 0.0 else float(err.max())
        zone[k] = float((err / torch.clamp(b.abs(), min=1e-10 * mx)).max()) if mx > 0.0 else float(err.max())
    return zone, norm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configuration (maestro_b200/workloads.py); c2 carries the headline metric")
    ap.add_argument("--n", type=int, default=0, help="zones per side (default: the configuration's own size)")
    ap.add_argument("--e2e-steps", type=lambda v: max(1, int(v)), default=2)
    ap.add_argument("--bds", action="store_true",
                    help="bds_type = 1: the second half of BASELINE configs[1] (BDS edge states instead of PPM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--selftest", action="store_true",
                    help="multi-rank parity of the NCCL path against the single-box oracle instead of the timing run")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (mgpu_set_option), e.g. exact=1")
    args = ap.parse_args()
    if args.bds:
        if args.config != "c2":
            ap.error("--bds applies to the test_advect configuration (c2)")
        PARAM_OVERRIDES["bds_type"] = 1
    if os.environ.get("BENCH_WATCHDOG"):  # debugging aid: dump every thread's Python stack if the run stalls
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) or 1
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from maestro_b200 import workloads

    n = args.n or workloads.DEFAULT_N[args.config]
    L2_NOTE = "256 MB flush between timed iterations"

    def config_of(w, nworld):
        multi = "single box"
        if nworld > 1:
            multi = ("one slab per rank (%s scaling), NCCL send/recv halo exchange in every ghost fill (timed)" % w.scaling)
        c = dict(w.desc)
        c.update({"l2": L2_NOTE, "multi_gpu": multi})
        return c

    if args.impl == "reference":
        # the reference's own CPU implementation of the path cannot be built here (no Fortran compiler,
        # FBoxLib not vendored): the arm times the restated algorithm (oracle/), OpenMP over all host cores, on the
        # SAME box as the GPU arm (c5: a bounded 96^3 sample -- 512^3 of oracle temporaries do not fit the host).
        if rank != 0:
            return
        n_s = 96 if args.config == "c5" else n
        zups, dt, w = cpu_reference(args.config, n_s, max(args.steps, 1), max(args.warmup, 0))
        sample = "the full %s box, every step" % args.config if n_s == n else "%d^3 box of the same workload per step" % n_s
        if cpu_reference.steps_timed != args.steps:
            sample += " (%d of the %d steps timed: the run is bounded to ~150 s)" % (cpu_reference.steps_timed, args.steps)
        cfg = config_of(build_desc_only(args.config, n, args.gpus) if n_s != n else w, args.gpus)
        line = {"metric": METRIC, "value": zups, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "steps_timed": cpu_reference.steps_timed,
                "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": w.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": zups, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%s, OpenMP over %d threads (restated reference algorithm, oracle pinned on the "
                                           "reference's test_advect figures)" % (sample, cores)},
                "e2e": {"value": zups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from maestro_b200 import abi, lib

    host_affinity = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    ops = lib.init(local_rank, use_torch_stream=True)
    for kv in args.opt:
        key, val = kv.split("=")
        lib.set_option(key, int(val))
    if world > 1:  # NCCL communicator of the library: halo exchange inside every ghost fill of the episode
        from maestro_b200 import slab

        slab.comm_init_from_torch(lib.load(), dev)
    if args.selftest:
        return selftest(ops, rank, world, local_rank, dev)
    w = build_workload(args.config, n, dev, rank=rank, world=world)
    p = w.p
    zone_updates = w.zone_updates

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident episode (value) --------------------------------------------
    p.mem_space = abi.DEVICE
    for q in w.extra.get("params", []):
        q.mem_space = abi.DEVICE
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~0.5 s to produce its first sample: start it before the warm-up
    for _ in range(W):
        w.reset()
        w.step(ops)
    # keep the GPU under the same load for about a second, until the sampler is running.  Every rank must run the
    # SAME number of episodes (each one exchanges halos with its neighbours): the count is agreed on first
    torch.cuda.synchronize()
    t_spin = time.time()
    w.step(ops)
    torch.cuda.synchronize()
    n_spin = torch.tensor([max(1, int(1.0 / max(time.time() - t_spin, 1e-4)))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_spin, op=dist.ReduceOp.MAX)
    for _ in range(min(int(n_spin.item()), 2000)):
        w.step(ops)
    torch.cuda.synchronize()
    sampler.lines.clear()
    barrier()
    lib.launch_count(reset=True)
    lib.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        w.reset()           # untimed: restores the in-place-modified inputs
        flush.zero_()       # untimed: evicts L2 between timed iterations
        ev[k][0].record()
        w.step(ops)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop()
    launches = lib.launch_count()
    prof = lib.profile_get()
    lib.profile(False)
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3  # seconds for K steps on this rank
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    zz = torch.tensor([float(zone_updates)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(zz, op=dist.ReduceOp.SUM)
    t_max = float(tt.item())
    value = float(zz.item()) * args.steps / t_max

    # roofline of the dominant kernel class
    hbm, peak_src = peaks()
    # ("glue" is a bag of a dozen single-purpose streaming kernels, not a kernel: the roofline line is about the largest
    # class that is ONE kernel -- on c5 that is still the edge-state kernel although the bag as a whole takes longer)
    single = {k: v for k, v in prof.items() if k != "glue"} or prof
    dom = max(single.items(), key=lambda kv: kv[1][0]) if single else (None, (0.0, 0))
    roof = None
    if dom[0] is not None and dom[1][1] > 0:
        name, (ms, nl) = dom
        # algorithmic bytes per launch of one make_edge_scal component: 3-D read s 8 + force 8 + umac 24, write
        # sedge 24 = 64 B per zone; 2-D 8 + 8 + 16 + 16 = 48 B per zone (kernel table)
        bytes_per_zone = 48.0 if p.dm == 2 else 64.0
        achieved = bytes_per_zone * w.zones / (ms / nl * 1e-3) / 1e9
        if name == "bds":  # staged: several launches per component -> per component, not per launch
            achieved = bytes_per_zone * w.zones * w.ncomp / (ms / args.steps * 1e-3) / 1e9
        traffic = None
        ncu_static = None
        try:  # DRAM bytes per launch of this kernel class from the committed ncu --set full capture (c2, n = 256 only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.config == "c2" and n == 256 and name in tj:
                traffic = tj[name]["bytes_per_launch"]
                ncu_static = tj[name].get("ncu")  # issue / fp64-pipe / DRAM utilisation of the same capture
        except Exception:
            traffic = None
        others = {}
        if args.config == "c2" and "update" in prof and prof["update"][1] > 0:  # flux + update + density kernel: 360 B per zone
            ms_u = prof["update"][0] / args.steps
            others["update"] = {"achieved": 360.0 * w.zones / (ms_u * 1e-3) / 1e9, "frac": 360.0 * w.zones / (ms_u * 1e-3) / 1e9 / hbm,
                                "ms_per_step": ms_u, "algorithmic_bytes_per_zone": 360.0,
                                "note": "k_flux_update3_fast, all species/tracers in one launch"}
        ep_gbs = w.bytes_per_zone * w.zones * args.steps / t_dev / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src, "other_kernels": others,
                "kernel_ms_per_launch": ms / nl, "kernel_share_of_step": (ms / args.steps) / (t_dev / args.steps * 1e3),
                "algorithmic_bytes_per_launch": bytes_per_zone * w.zones,
                "episode_algorithmic_bytes_per_zone": w.bytes_per_zone,
                "episode_algorithmic_GBs": ep_gbs, "episode_frac": ep_gbs / hbm,
                "kernel_classes_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "ncu_capture": ncu_static}
    del flush

    # ---------------- the bit-identical build beside it, and the per-zone distance between the two ------------------
    parity = None
    exact_line = None
    if world == 1 and not args.no_parity and args.config in ("c2", "c3", "c4"):
        w.reset()
        w.step(ops)
        fast = {k: v.clone() for k, v in w.outputs().items()}
        lib.set_option("exact", 1)
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w.reset(); w.step(ops)  # warm-up (first launch of the exact kernels)
        ms_exact = 0.0
        for _ in range(3):
            w.reset()
            te0.record(); w.step(ops); te1.record()
            torch.cuda.synchronize()
            ms_exact += te0.elapsed_time(te1) / 3
        zone, norm = per_zone_parity(w.outputs(), fast)
        parity = {"what": "FAST build (timed above) against the exact build (bit-identical to the oracle) per output field: "
                          "per_zone_rel = max over zones of |a-b| / max(|b|, 1e-10 max|b|), max_norm_rel = max|a-b| / max|b|",
                  "per_zone_rel": zone, "max_norm_rel": norm, "worst_per_zone": max(zone.values()),
                  "worst_max_norm": max(norm.values())}
        exact_line = {"ms_per_step": ms_exact, "value": zone_updates / (ms_exact * 1e-3),
                      "note": "exact build (--opt exact=1): no FMA contraction, reference expression order, bit-identical to the oracle"}
        lib.set_option("exact", 0)
        del fast

    # ---------------- end to end through the C ABI with host buffers (e2e) ------------------------
    e2e = None
    if args.config == "c2":
        p.mem_space = abi.HOST
        del w
        torch.cuda.empty_cache()
        wh = build_workload(args.config, n, None, rank=rank, world=world)
        wh.p.dt, wh.p.rel_eps = p.dt, p.rel_eps
        eh = wh.extra["e"]
        import ctypes as C

        def pin(f):  # pinned host memory so the copies run at full PCIe rate
            t = torch.from_numpy(f.a)
            lib.load().mgpu_host_register(C.c_void_p(f.ptr), f.a.size)
            return t

        keep = [pin(f) for f in [eh["sold"], eh["snew"], eh["force"], eh["eta"]] + eh["umac"] + eh["sedge"] + eh["sflux"]]
        wh.step(ops)  # warm-up (allocates the staging pool)
        barrier()
        lib.copy_bytes(reset=True)  # the library counts the bytes of every cudaMemcpyAsync it issues
        t_e2e = 0.0
        for _ in range(args.e2e_steps):
            wh.reset()  # untimed: restore the in-place-modified host inputs
            barrier()
            t0 = time.perf_counter()
            wh.step(ops)  # synchronous: H2D of inputs, kernels, D2H of outputs, stream sync
            t_e2e += time.perf_counter() - t0
        in_bytes, out_bytes = [b // args.e2e_steps for b in lib.copy_bytes()]
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        full = {"value": world * zone_updates * args.e2e_steps / float(te.item()), "unit": UNIT,
                "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes, "steps": args.e2e_steps,
                "check_sum_rho_new": float(eh["snew"].valid(0).sum()),
                "note": "plain host pointers: every output the frozen Fortran signature exposes is copied back"}
        # the integration INTEGRATION.md recommends: the multifabs are registered (mgpu_register), so the edge states,
        # fluxes, force and etarhoflux stay on the device for the episodes that consume them (enthalpy_advance,
        # make_etarho); every step the host still uploads the inputs it owns (state + MAC velocity) and reads the new state
        allf = [eh["sold"], eh["snew"], eh["force"], eh["eta"]] + eh["umac"] + eh["sedge"] + eh["sflux"]
        for f in allf:
            lib.register(f, pin=False)  # (already pinned above)
        wh.reset()
        wh.step(ops)  # warm-up: first upload of everything
        lib.download(eh["snew"])
        barrier()
        lib.copy_bytes(reset=True)
        t_res = 0.0
        for _ in range(args.e2e_steps):
            wh.reset()  # untimed: the host restores / advances its inputs ...
            barrier()
            t0 = time.perf_counter()
            lib.invalidate(eh["sold"])  # ... and says so: state and MAC velocity are uploaded again
            for u in eh["umac"]:
                lib.invalidate(u)
            wh.step(ops)
            lib.download(eh["snew"])  # the step's result
            t_res += time.perf_counter() - t0
        in_r, out_r = [b // args.e2e_steps for b in lib.copy_bytes()]
        tr = torch.tensor([t_res], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        e2e = {"value": world * zone_updates * args.e2e_steps / float(tr.item()), "unit": UNIT,
               "h2d_bytes_per_step": in_r, "d2h_bytes_per_step": out_r, "steps": args.e2e_steps,
               "check_sum_rho_new": float(eh["snew"].valid(0).sum()),
               "note": "host buffers through the C ABI, multifabs registered (mgpu_register): inputs (state, MAC velocity) "
                       "uploaded and the new state downloaded every step, edge states / fluxes stay on the device",
               "every_output_copied_back": full}
        for f in allf:
            lib.unregister(f)
        w = wh

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": w.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(w, world),
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "host_affinity": host_affinity}
    if parity is not None:
        line["parity"] = parity
        line["exact_build"] = exact_line
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = {"c2": 128, "c3": 96, "c4": 1024, "c5": 64}[args.config]
        zups, dt, _ = cpu_reference(args.config, n_s, 2, 1)
        line["cpu_baseline"] = {"value": zups, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "the same episode(s) on a %d-zone-per-side box of the same workload, 2 steps, "
                                          "OpenMP over %d threads (restated reference algorithm, oracle pinned on the "
                                          "reference's test_advect figures; gfortran+FBoxLib build impossible here); "
                                          "`--impl reference` times the full-size box" % (n_s, cores)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def build_desc_only(config, n, world):
    """the configuration text of the full-size workload without building it (reference arm on a sample box)"""
    class D:
        pass

    d = D()
    w = None
    from maestro_b200 import workloads

    # a tiny build of the same configuration gives the text; sizes are patched to the full ones
    w = workloads.BUILDERS[config](n=16 if config != "c4" else 32, device="cpu" if config == "c5" else None)
    w.desc["workload"] = w.desc["workload"].replace("16^3", "%d^3" % n).replace("32^2", "%d^2" % n)
    w.desc["zones_per_gpu"] = n ** w.p.dm // max(1, world if w.scaling == "strong" else 1)
    return w


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs next to its GPU (NVML's CPU affinity of the device, restricted to the CPUs the
    container allows) before any host buffer is allocated, so that the pinned staging memory of the e2e leg is
    first-touched on the GPU's NUMA node and the ranks of an 8-GPU run do not all stream through one socket.
    Best effort: returns what was done for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = near & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return {"cpus": len(use), "of_allowed": len(allowed), "bound": True}
        return {"cpus": len(allowed), "of_allowed": len(allowed), "bound": False}
    except Exception as e:  # no NVML, no permission: run unbound
        return {"bound": False, "why": str(e)[:80]}


def selftest(ops, rank, world, local_rank, dev):
    """`bench.py --gpus N --selftest` (under torchrun): the multi-rank parity cases of tests/mgpu_rank_parity.py, run
    inside the same launch the driver uses for the scaling bench, so that the NCCL halo path is checked on the
    driver's hardware.  Prints one JSON line; exits nonzero on mismatch."""
    import subprocess

    import torch.distributed as dist

    dist.destroy_process_group()
    if rank != 0:
        return
    cases = [("3", "periodic", "1", "1"), ("3", "walls", "2", "1"), ("2", "walls", "1", "1"), ("3", "periodic", "1", "0"),
             ("3", "walls", "2", "0"), ("2", "periodic", "2", "0"), ("3", "sphr", "1", "1"), ("3", "sphr", "1", "0"),
             # velocity_advance, advance_premac, enthalpy_advance, average over slabs (exact and FAST builds)
             ("3", "episodes-periodic", "1", "1"), ("3", "episodes-walls", "2", "1"), ("2", "episodes-walls", "1", "0")]
    res = {}
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_PORT", "TORCHELASTIC_RUN_ID")}
    for i, c in enumerate(cases):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(29730 + i), os.path.join(ROOT, "tests", "mgpu_rank_parity.py")] + list(c)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        oks = [ln for ln in r.stdout.splitlines() if ln.startswith("RANK")]
        res["dm%s_%s_ppm%s_%s" % (c[0], c[1], c[2], "exact" if c[3] == "1" else "fast")] = {"rc": r.returncode, "ranks": oks}
    ok = all(v["rc"] == 0 for v in res.values())
    print(json.dumps({"selftest": "multi-rank episodes (density_advance; episodes-*: velocity_advance, advance_premac, "
                                  "enthalpy_advance, average, firstdt) against the single-box oracle, NCCL halo exchange "
                                  "in every ghost fill, NCCL min / max / sum in the reductions",
                      "n_gpus": world, "ok": ok, "cases": res}))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
