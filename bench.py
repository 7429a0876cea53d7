#!/usr/bin/env python
"""bench.py — PTv2 pointops fwd+bwd throughput (Mpoints/s) on B200, the metric of BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W]            the sm_100a kernels (this repo)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's pure-torch CPU path
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU (weak scaling)

A "step" is one pass of the PTv2m2 point-operator schedule (ao_b200.schedule: every kNN / gather /
GVA aggregate / GridPool / interpolation call of one forward of semseg-pt-v2m2-0-base and all their
backward passes) over one S3DIS-shaped batch of 4 rooms x 80k points per GPU (BASELINE.json
configs[1]).  `value` = level-0 points of all ranks / device time, inputs resident in HBM;
`e2e` = the same schedule driven from pinned HOST buffers (H2D of coord/feat/offset and a D2H read of
the result scalar inside the timed region, wall clock).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ptv2_pointops_fwd_bwd_throughput"
UNIT = "Mpoints/s"
ROOMS_PER_GPU = 4
POINTS_PER_ROOM = 80000
DDP_GRAD_BYTES = 3908641 * 4   # S3DIS-cfg PTv2m2 parameters, fp32
TRACE_STEPS = int(os.environ.get("AOPT_BENCH_TRACE_STEPS", "1"))   # timed steps that also carry per-call CUDA events (roofline table)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# algorithmic bytes per C-ABI call (SURVEY.md §8d formulas; DESIGN.md §4 lists them per kernel)
# ---------------------------------------------------------------------------------------------------
def call_bytes(name, a, sizes, k):
    """a = the ctypes argument tuple of the call; sizes = level point counts of this step."""
    def finer(n_coarse):      # level size one above a coarse size
        for i in range(1, len(sizes)):
            if sizes[i] == n_coarse:
                return sizes[i - 1]
        return n_coarse
    if name == "aopt_knn_query":
        m, ns, n = a[0], a[1], a[2]
        return 12.0 * n + 12.0 * m + 8.0 * m * ns
    if name == "aopt_csr_build":
        n_src, e = a[0], a[1]
        return 4.0 * e + 4.0 * (n_src + 1) + 4.0 * e
    if name == "aopt_group_xyz":
        m, ns = a[0], a[1]
        return 24.0 * m + 4.0 * m * ns + 12.0 * m * ns
    if name == "aopt_gather_sub_forward":
        m, ns, c = a[0], a[1], a[2]
        return 4.0 * m * ns + 8.0 * m * c + 4.0 * m * ns * c
    if name == "aopt_grouping_forward":
        m, ns, c = a[0], a[1], a[2]
        return 4.0 * m * ns + 4.0 * m * c + 4.0 * m * ns * c
    if name == "aopt_grouping_backward":          # relation backward: E = n*k entries
        n, c = a[0], a[1]
        return 4.0 * n * k * c + 4.0 * n * k + 4.0 * (n + 1) + 4.0 * n * c
    if name == "aopt_relation_backward":          # one read of the (n,k,c) gradient, two (n,c) outputs
        n, ns, c = a[0], a[1], a[2]
        return 4.0 * n * ns * c + 4.0 * n * ns + 4.0 * (n + 1) + 8.0 * n * c
    if name == "aopt_sum_over_k":
        m, ns, c = a[0], a[1], a[2]
        return 4.0 * m * ns * c + 4.0 * m * c
    if name == "aopt_gva_forward":
        n, ns, c, g = a[0], a[1], a[2], a[3]
        return 4.0 * n * c + 4.0 * n * ns * c + 8.0 * n * ns * g + 4.0 * n * ns + 4.0 * n * c
    if name == "aopt_gva_backward_query":
        n, ns, c, g = a[0], a[1], a[2], a[3]
        return 8.0 * n * c + 4.0 * n * ns * c + 4.0 * n * ns * g + 4.0 * n * ns + 4.0 * n * ns * c + 4.0 * n * ns * g
    if name == "aopt_gva_backward_value":
        n, ns, c, g = a[0], a[1], a[2], a[3]
        return 4.0 * n * ns * g + 4.0 * n * c + 4.0 * (n + 1) + 4.0 * n * ns + 4.0 * n * c
    if name == "aopt_pool_forward":
        nv, c = a[0], a[1]
        n = finer(nv)
        return 4.0 * n * c + 12.0 * n + 4.0 * n + 4.0 * (nv + 1) + 8.0 * nv * c + 12.0 * nv
    if name == "aopt_pool_backward":
        n, c = a[0], a[1]
        nv = sizes[sizes.index(n) + 1] if n in sizes and sizes.index(n) + 1 < len(sizes) else n
        return 8.0 * nv * c + 4.0 * n + 4.0 * n * c
    if name == "aopt_interp_weights":
        n, kk = a[0], a[1]
        return 8.0 * n * kk
    if name == "aopt_interpolation_forward":
        n, c, kk, m = a[0], a[1], a[2], a[3]
        return 8.0 * n * kk + 4.0 * m * c + 4.0 * n * c
    if name == "aopt_interpolation_backward":
        m, c, kk = a[0], a[1], a[2]
        n = finer(m)
        return 4.0 * n * c + 8.0 * n * kk + 4.0 * (m + 1) + 4.0 * m * c
    if name in ("aopt_segment_min3", "aopt_voxel_keys"):
        n = a[0]
        return 12.0 * n + (8.0 * n if name == "aopt_voxel_keys" else 0.0)
    return 0.0


HBM_KERNELS = {"aopt_group_xyz", "aopt_gather_sub_forward", "aopt_grouping_forward", "aopt_grouping_backward",
               "aopt_relation_backward",
               "aopt_sum_over_k", "aopt_gva_forward", "aopt_gva_backward_query", "aopt_gva_backward_value",
               "aopt_pool_forward", "aopt_pool_backward", "aopt_interpolation_forward",
               "aopt_interpolation_backward"}


def summarise_trace(trace, sizes, k, step_ms_total, peak):
    """Per entry point: all launches of the traced steps, and separately its LARGEST launch shape (the
    level-0 launches) — small levels are launch-latency bound and say little about the kernel."""
    per = {}
    for name, args, s, e in trace:
        ms = s.elapsed_time(e)
        nb = call_bytes(name, args, sizes, k)
        d = per.setdefault(name, dict(ms=0.0, calls=0, bytes=0.0, shapes={}))
        d["ms"] += ms
        d["calls"] += 1
        d["bytes"] += nb
        sh = d["shapes"].setdefault(nb, [0.0, 0])
        sh[0] += ms
        sh[1] += 1
    out = []
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        big = max(d["shapes"])
        big_ms, big_calls = d["shapes"][big]
        big_gbs = big * big_calls / (big_ms * 1e-3) / 1e9 if big_ms > 0 else 0.0
        hbm = name in HBM_KERNELS
        out.append(dict(kernel=name, calls=d["calls"], ms=round(d["ms"], 4), share=round(d["ms"] / step_ms_total, 4),
                        alg_gb=round(d["bytes"] / 1e9, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 4) if hbm else None,
                        largest=dict(alg_bytes=big, launches=big_calls, us_per_launch=round(big_ms / big_calls * 1e3, 2),
                                     gbs=round(big_gbs, 1), frac=round(big_gbs / peak, 4) if hbm else None)))
    return out


# ---------------------------------------------------------------------------------------------------
# CPU reference path (oracle/cpu_path.py) — cpu_baseline leg and the --impl reference arm
# ---------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, budget_s, room_id=0):
    import torch

    from ao_b200 import scenes          # numpy scene generator only (no kernels)
    from oracle import cpu_path

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    coord, _ = scenes.indoor_room(room_id, POINTS_PER_ROOM)
    room = cpu_path.CpuRoom(coord)
    per_step = max(1.0, budget_s / max(1, steps + warmup + 1))
    f = max(0.02, cpu_path.calibrate_fraction(room, per_step, f0=0.02))
    for _ in range(warmup):
        room.step(f)
    tot_s, tot_pts, fam = 0.0, 0.0, {}
    for _ in range(steps):
        r = room.step(f)
        tot_s += r["total"]
        tot_pts += r["points"]
        for key in ("knn", "block", "pool", "interp"):
            fam[key] = fam.get(key, 0.0) + r[key]
    value = tot_pts / tot_s / 1e6
    sample = (f"one synthetic S3DIS room (80000 pts, k=16), reference op schedule fwd+bwd on the first "
              f"{f:.4f} of every level's query rows against the full level ({int(tot_pts / max(steps, 1))} "
              f"level-0 points per step); cdist+topk kNN, torch gather/index_put scatter")
    return dict(value=value, ms_per_step=tot_s / max(steps, 1) * 1e3, cores=cores, sample=sample,
                fraction=f, family_seconds={k2: round(v, 3) for k2, v in fam.items()})


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup, budget_s=float(os.environ.get("AOPT_BENCH_CPU_BUDGET", "150")))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "family_seconds": r["family_seconds"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_gpus):
    return {
        "workload": "PTv2m2 semseg-pt-v2m2-0-base point-operator schedule fwd+bwd, S3DIS-shaped batch of "
                    f"{ROOMS_PER_GPU} rooms x {POINTS_PER_ROOM} pts per GPU (BASELINE.json configs[1])",
        "rooms_per_gpu": ROOMS_PER_GPU, "points_per_room": POINTS_PER_ROOM, "k": 16,
        "channels": [48, 96, 192, 384], "groups": [6, 12, 24, 48], "blocks_per_level": [3, 3, 7, 2],
        "parallelism": f"scene-sharded x{n_gpus}" + (", fp32 grad all-reduce 14.9 MB/step (NCCL)" if n_gpus > 1 else ""),
        "l2": "per-step working set (>20 GB) exceeds the 126 MB L2; no explicit flush",
        "streams": "single stream (side-stream overlap of kNN / CSR walk / CSR build measured: no gain, left off)",
    }


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from ao_b200 import _lib, scenes
    from ao_b200.schedule import PointOpsSchedule, ScheduleConfig

    # torchrun exports RANK / LOCAL_RANK / WORLD_SIZE; a plain `python bench.py` is a single process
    under_torchrun = all(k in os.environ for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    world = int(os.environ["WORLD_SIZE"]) if under_torchrun else 1
    rank = int(os.environ["RANK"]) if under_torchrun else 0
    local = int(os.environ["LOCAL_RANK"]) if under_torchrun else 0
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- synthetic batch: rooms [rank*R, rank*R+R) --------------------------------------------------
    coord_np, feat_np, off_np = scenes.s3dis_batch(ROOMS_PER_GPU, POINTS_PER_ROOM, first_room=rank * ROOMS_PER_GPU)
    if args.presort:   # experiment: spatially coherent point order inside every room (Morton order of 0.1 m cells)
        import numpy as np
        order, s0 = [], 0
        for e0 in off_np:
            c = coord_np[s0:e0]
            cell = np.floor((c - c.min(0)) / 0.1).astype(np.int64)
            key = np.zeros(len(c), np.int64)
            for bit in range(10):
                for a in range(3):
                    key |= ((cell[:, a] >> bit) & 1) << (3 * bit + a)
            order.append(s0 + np.argsort(key, kind="stable"))
            s0 = e0
        order = np.concatenate(order)
        coord_np, feat_np = np.ascontiguousarray(coord_np[order]), np.ascontiguousarray(feat_np[order])
    coord_h = torch.from_numpy(coord_np).pin_memory()
    feat_h = torch.from_numpy(feat_np).pin_memory()
    off_h = torch.from_numpy(off_np).pin_memory()
    coord, feat, offset = coord_h.to(dev), feat_h.to(dev), off_h.to(dev)
    n0 = coord.shape[0]
    cfg = ScheduleConfig.s3dis()
    sched = PointOpsSchedule(cfg, device=dev, seed=rank)
    grads = torch.zeros(DDP_GRAD_BYTES // 4, device=dev) if world > 1 else None

    def one_step(c, o):
        acc = sched.step(c, o)
        if grads is not None:
            dist.all_reduce(grads)          # the DDP gradient all-reduce: the only collective
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, args.min_warmup)):
        one_step(coord, offset)
    barrier()

    # ---- timed region: resident inputs, CUDA events on the launching stream ----------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    # the GPU idled while the clock sampler started: one more untimed step brings clocks / power state back up
    # (without it the first two or three timed steps run 5-20 % slow), and the Python GC stays off while timing
    one_step(coord, offset)
    barrier()
    import gc
    gc.collect()
    gc.disable()
    _lib.trace_prepare(1024 * max(1, min(TRACE_STEPS, args.steps)))
    launches0 = _lib.kernel_launches()
    # per-call CUDA events are recorded on TRACE_STEPS of the timed steps (an event pair per call on all
    # ~200 calls of every step costs ~1 ms/step of host time, which would distort the step time)
    trace_steps = min(TRACE_STEPS, args.steps)
    trace = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]   # end of every step (diagnostics)
    wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        if i > 0:
            marks[i - 1].record()
        if i == args.steps - trace_steps:
            trace = _lib.trace_start()
            # per-kernel roofline = the kernel running ALONE: the traced steps issue everything on one stream
            # (they are still part of the timed region, so `value` is a slight under-estimate)
            overlap_was = _lib.overlap_mode()
            _lib.overlap(False)
        one_step(coord, offset)
    e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    _lib.trace_stop()
    if trace_steps:
        _lib.overlap_mode(overlap_was)
    launches = _lib.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    bounds = [e0] + marks[: args.steps - 1] + [e1]
    step_ms = [round(bounds[i].elapsed_time(bounds[i + 1]), 3) for i in range(args.steps)]
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n0 / (ms_step * 1e-3) / 1e6
    sizes = list(sched.last_sizes)

    # ---- e2e: host buffers → H2D → schedule → D2H of the result scalar, wall clock ---------------------
    # Every step copies its own inputs from pinned host memory and reads the result scalar back (a sync per
    # step, like a training loop that logs its loss).  As a data loader with pin_memory / non_blocking would, the
    # copy of step i+1 is issued on a copy stream while step i computes; all copies lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)

    def stage():
        with torch.cuda.stream(copy_stream):
            c = coord_h.to(dev, non_blocking=True)
            f = feat_h.to(dev, non_blocking=True)
            o = off_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return c, f, o, ev

    e2e_step_ms = []

    def e2e_loop(steps):
        last = 0.0
        nxt = stage()
        for i in range(steps):
            t_step = time.perf_counter()
            c, f, o, ev = nxt
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            for t_ in (c, f, o):
                t_.record_stream(cur)        # allocated on the copy stream, consumed on the compute stream
            if i + 1 < steps:
                nxt = stage()
            acc = one_step(c, o)
            del f
            last = float(acc.sum().item())   # D2H read of the step result (the loss stand-in)
            e2e_step_ms.append(round((time.perf_counter() - t_step) * 1e3, 3))
        return last

    e2e_steps = 0 if args.skip_e2e else args.steps
    if e2e_steps:
        # warm-up of this leg too: its copy-stream buffers are new allocations (the first two steps of a cold loop
        # took 37 and 16 ms in cudaMalloc, against 8.4 ms once the double buffers exist)
        e2e_loop(max(3, args.warmup))
    barrier()
    del e2e_step_ms[:]
    w0 = time.perf_counter()
    e2e_loop(e2e_steps)
    barrier()
    e2e_s = max(time.perf_counter() - w0, 1e-9)
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    gc.enable()
    e2e_value = world * n0 * e2e_steps / e2e_s / 1e6
    h2d = coord_h.numel() * 4 + feat_h.numel() * 4 + off_h.numel() * 4

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant HBM-bound kernel, measured live over the timed region ---------------
    peak, peak_src = hbm_peak()
    kernels = summarise_trace(trace, sizes, cfg.k, ms_step * trace_steps, peak)
    for kr in kernels:
        kr["ms_per_step"] = round(kr.pop("ms") / trace_steps, 4)
        kr["calls_per_step"] = kr.pop("calls") // trace_steps
        kr["alg_gb_per_step"] = round(kr.pop("alg_gb") / trace_steps, 4)
        kr["largest"]["launches_per_step"] = kr["largest"].pop("launches") // trace_steps
    hbm_rows = [kr for kr in kernels if kr["kernel"] in HBM_KERNELS]
    dom = hbm_rows[0] if hbm_rows else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom["kernel"])
        except Exception:
            traffic = None
    roofline = None
    if dom:
        # the dominant HBM-bound kernel at its level-0 launch shape: algorithmic bytes of one launch / its
        # CUDA-event duration averaged over the traced launches; `traffic` = dram read+write bytes of the
        # same launch shape from the committed ncu --set full capture (profiles/traffic.json)
        big = dom["largest"]
        info = traffic if isinstance(traffic, dict) else {}
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": big["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": big["frac"], "traffic": info.get("dram_bytes_per_launch"),
                    "traffic_source": info.get("source"), "peak_source": peak_src,
                    "alg_bytes_per_launch": big["alg_bytes"], "us_per_launch": big["us_per_launch"],
                    "launches_per_step": big["launches_per_step"], "launch_shape": "level 0: N=%d, k=%d, C=%d, G=%d" % (
                        sizes[0], cfg.k, cfg.channels[0], cfg.groups[0]),
                    "share_of_step_all_launches": dom["share"], "frac_all_launches": dom["frac"]}
    hbm_ms = sum(kr["ms_per_step"] for kr in hbm_rows)
    hbm_gb = sum(kr["alg_gb_per_step"] for kr in hbm_rows)
    knn_row = next((kr for kr in kernels if kr["kernel"] == "aopt_knn_query"), None)
    # brute-force-equivalent pair count of the searches (scenes of a level are near-equal in size)
    pairs = sum(float(a[0]) * float(a[2]) / max(a[3], 1) for nm, a, _, _ in trace if nm == "aopt_knn_query") / trace_steps

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, args.min_warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "level_sizes": sizes,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "timing": "wall clock incl. python launch overhead; H2D of step i+1 overlaps step i on a copy stream"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "hbm_kernels_total": {"ms_per_step": round(hbm_ms, 4), "alg_gb_per_step": round(hbm_gb, 3),
                              "gbs": round(hbm_gb / (hbm_ms * 1e-3), 1) if hbm_ms > 0 else None,
                              "frac": round(hbm_gb / (hbm_ms * 1e-3) / peak, 4) if hbm_ms > 0 else None},
        "knn": None if knn_row is None else {"ms_per_step": knn_row["ms_per_step"], "calls_per_step": knn_row["calls_per_step"],
                                             "brute_force_pairs_per_step": pairs,
                                             "equiv_pairs_per_s": pairs / (knn_row["ms_per_step"] * 1e-3)},
        "kernels": kernels,
        "host_wall_ms_per_step": wall / args.steps * 1e3,
        "step_ms": step_ms, "traced_steps": trace_steps, "e2e_step_ms": e2e_step_ms,
    }

    # ---- full PTv2m2 model step (information; the dense MLPs are cuBLAS, not part of the metric) -------
    if not args.no_model and world == 1:
        try:
            # the schedule's cached (N,k,C) blocks would make the model's different allocation pattern fall
            # back to synchronous cudaFree/cudaMalloc retries inside its first steps (measured 155 vs 69 ms)
            del sched
            trace = None
            torch.cuda.empty_cache()
            line["model_step"] = model_step(dev, coord, feat, offset)
        except Exception as ex:  # pragma: no cover
            line["model_step"] = {"error": repr(ex)[:200]}
    # ---- CPU baseline on the box's host cores (bounded sample) -----------------------------------------
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference(steps=1, warmup=0, budget_s=24.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": r["sample"], "family_seconds": r["family_seconds"]}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def model_step(dev, coord, feat, offset, steps=5):
    import torch

    from ao_b200 import ptv2

    torch.manual_seed(0)
    model = ptv2.PointTransformerV2(**ptv2.S3DIS_CFG).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    target = torch.randint(0, 13, (coord.shape[0],), device=dev)
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(dict(coord=coord, feat=feat, offset=offset))
        loss = torch.nn.functional.cross_entropy(logits.float(), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "full PTv2m2 (S3DIS cfg) training step: bf16 autocast GEMMs (cuBLAS) + these point ops + AdamW",
            "ms_per_step": ms, "mpoints_per_s": coord.shape[0] / (ms * 1e-3) / 1e6, "loss": float(loss.item()),
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints
    its version banner to fd 1 from C, torchrun children inherit it) and the line is written to the saved fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-model", action="store_true")
    ap.add_argument("--presort", action="store_true", help="(experiment) Morton-order the points of every room on the host")
    ap.add_argument("--skip-e2e", action="store_true", help="(profiling runs only) skip the host-buffer leg")
    ap.add_argument("--min-warmup", type=int, default=3, help="(profiling runs only) lower bound on warm-up steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
