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
# BASELINE.json configs[1..4].  rooms = scenes per GPU; params = fp32 parameters of the config's PTv2m2 (SURVEY.md §2.3:
# the DDP gradient all-reduce payload); variants = extra schedules timed after the headline one (fewer steps).
CONFIGS = {
    "s3dis4": dict(baseline=1, sched="s3dis", scenes="s3dis", rooms=4, points=80000, params=3908641, model="S3DIS_CFG",
                   what="S3DIS-shaped batch of {r} rooms x {p} pts per GPU (BASELINE.json configs[1])",
                   variants=[("fused", dict(variant="fused"))]),
    "s3dis8": dict(baseline=2, sched="s3dis", scenes="s3dis", rooms=8, points=80000, params=3908641, model="S3DIS_CFG",
                   what="S3DIS-shaped batch of {r} rooms x {p} pts per GPU, DDP training-step shape (BASELINE.json configs[2])",
                   variants=[("fused", dict(variant="fused"))]),
    "scannet150k": dict(baseline=3, sched="scannet", scenes="scannet", rooms=3, points=150000, params=11323948,
                        model="SCANNET_CFG",
                        what="ScanNet-shaped batch of {r} rooms x {p} pts per GPU (12 rooms / 4 GPUs in the reference "
                             "config), ScanNet cfg: patch k=8, 4 GridPool stages, k=16, `map` unpooling "
                             "(BASELINE.json configs[3])",
                        variants=[("interp_up", dict(unpool="interp")), ("k32", dict(k=32)), ("fused", dict(variant="fused"))]),
    "kitti120k": dict(baseline=4, sched="kitti", scenes="kitti", rooms=2, points=120000, params=11323659, model="KITTI_CFG",
                      what="SemanticKITTI-shaped batch of {r} scans x {p} pts per GPU (8 scans / 4 GPUs in the reference "
                           "config), KITTI cfg: patch k=8, 4 GridPool stages, k=16, `map` unpooling (BASELINE.json configs[4])",
                      variants=[("fused", dict(variant="fused"))]),
}
DEFAULT_STEPS = 250    # 250 x ~8 ms: a timed region of >= 2 s (>= 20 clock samples at 100 ms)
DEFAULT_STEPS_REFERENCE = 10
VARIANT_STEPS = 30
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
def call_bytes(name, a, sizes):
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
    if name == "aopt_grouping_backward":          # in the schedule: `map` unpool backward, E = points of the finer level
        n, c = a[0], a[1]
        e = finer(n)
        return 4.0 * e * c + 4.0 * e + 4.0 * (n + 1) + 4.0 * n * c
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
    if name == "aopt_gva_backward":               # fused pair: SURVEY §8d "Fused GVA bwd"
        n, ns, c, g = a[0], a[1], a[2], a[3]
        return (8.0 * n * c + 4.0 * n * ns * c + 4.0 * n * ns * g + 8.0 * n * ns + 4.0 * (n + 1)
                + 4.0 * n * ns * c + 4.0 * n * ns * g + 4.0 * n * c)
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
    if name == "aopt_pe_mlp_forward":             # rows x (12 B of pos in, 4C out, 4*ga aux out)
        rows, c, ga = a[0], a[1], a[16]
        return rows * (12.0 + 4.0 * c + 4.0 * ga)
    if name == "aopt_pe_mlp_backward":            # rows x (12 B of pos, 4C of grad_out, 4*ga of grad_aux) in
        rows, c, ga = a[0], a[1], a[15]
        return rows * (12.0 + 4.0 * c + 4.0 * ga)
    if name == "aopt_pos_moments":
        return 12.0 * a[0]
    return 0.0


HBM_KERNELS = {"aopt_group_xyz", "aopt_gather_sub_forward", "aopt_grouping_forward", "aopt_grouping_backward",
               "aopt_relation_backward",
               "aopt_sum_over_k", "aopt_gva_forward", "aopt_gva_backward_query", "aopt_gva_backward_value",
               "aopt_gva_backward",
               "aopt_pool_forward", "aopt_pool_backward", "aopt_interpolation_forward",
               "aopt_interpolation_backward"}
TENSOR_KERNELS = {"aopt_pe_mlp_forward", "aopt_pe_mlp_backward"}    # contraction kernels: HBM-bound too (2C flop / output byte)


def summarise_trace(trace, sizes, step_ms_total, peak):
    """Per entry point: all launches of the traced steps, and separately its LARGEST launch shape (the
    level-0 launches) — small levels are launch-latency bound and say little about the kernel."""
    per = {}
    for name, args, s, e in trace:
        ms = s.elapsed_time(e)
        nb = call_bytes(name, args, sizes)
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
        hbm = name in HBM_KERNELS or name in TENSOR_KERNELS
        out.append(dict(kernel=name, calls=d["calls"], ms=round(d["ms"], 4), share=round(d["ms"] / step_ms_total, 4),
                        alg_gb=round(d["bytes"] / 1e9, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 4) if hbm else None,
                        largest=dict(alg_bytes=big, launches=big_calls, us_per_launch=round(big_ms / big_calls * 1e3, 2),
                                     gbs=round(big_gbs, 1), frac=round(big_gbs / peak, 4) if hbm else None)))
    return out


def per_step(kernels, trace_steps):
    for kr in kernels:
        kr["ms_per_step"] = round(kr.pop("ms") / trace_steps, 4)
        kr["calls_per_step"] = kr.pop("calls") // trace_steps
        kr["alg_gb_per_step"] = round(kr.pop("alg_gb") / trace_steps, 4)
        kr["largest"]["launches_per_step"] = kr["largest"].pop("launches") // trace_steps
    return kernels


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def make_schedule_config(name, **over):
    from ao_b200.schedule import ScheduleConfig

    return getattr(ScheduleConfig, CONFIGS[name]["sched"])(**over)


def make_batch(name, rooms, rank):
    """Scenes [rank*rooms, rank*rooms+rooms) of the config's synthetic generator (numpy, host)."""
    from ao_b200 import scenes

    c = CONFIGS[name]
    if c["scenes"] == "s3dis":
        return scenes.s3dis_batch(rooms, c["points"], first_room=rank * rooms)
    if c["scenes"] == "scannet":
        return scenes.scannet_batch(rooms, c["points"], first_room=100 + rank * rooms)
    return scenes.kitti_batch(rooms, c["points"], first_scan=rank * rooms)


def workload_config(name, rooms, n_gpus, cfg=None):
    c = CONFIGS[name]
    if cfg is None:
        cfg = make_schedule_config(name)
    per = [cfg.patch_depth] + list(cfg.enc_depths)
    for i, d in enumerate(cfg.dec_depths):
        per[i] += d
    mb = c["params"] * 4 / 1e6
    return {
        "workload": "PTv2m2 point-operator schedule fwd+bwd (every kNN / gather / GVA aggregate / GridPool / unpool call of "
                    "one forward of the config's backbone and their backward passes), " + c["what"].format(r=rooms, p=c["points"])
                    + "; `value` = the materialised-relation variant (gva_relation at width C, what the model runs in "
                    "fp32); `variants.fused` = the schedule ptv2 runs under bf16 autocast at C in {48, 96}",
        "name": name, "rooms_per_gpu": rooms, "points_per_room": c["points"],
        "k": {"patch": cfg.k_patch(), "enc": [cfg.k_enc(i) for i in range(len(cfg.grid_sizes))],
              "dec": [cfg.k_dec(i) for i in range(len(cfg.grid_sizes))]},
        "channels": list(cfg.channels), "groups": list(cfg.groups), "blocks_per_level": per, "grid_sizes": list(cfg.grid_sizes),
        "unpool": cfg.unpool,
        "parallelism": f"scene-sharded x{n_gpus}" + (
            f", fp32 gradient all-reduce {mb:.1f} MB/step (NCCL) issued on its own stream under the backward pass "
            "(stand-in buffer of the model's gradient size; DDP's own bucketed all-reduce is what `model_step` runs)"
            if n_gpus > 1 else ""),
        "l2": "per-step working set (>20 GB) exceeds the 126 MB L2; no explicit flush",
        "streams": "single compute stream (side-stream overlap of kNN / CSR walk / CSR build measured: no gain, left off)",
    }


# ---------------------------------------------------------------------------------------------------
# CPU reference path (oracle/cpu_path.py) — cpu_baseline leg and the --impl reference arm
# ---------------------------------------------------------------------------------------------------
def cpu_reference(name, steps, warmup, fraction=None, room_id=0):
    import torch

    from oracle import cpu_path

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_schedule_config(name)
    coord, _, _ = make_batch(name, 1, room_id)
    room = cpu_path.CpuRoom(coord, cfg)
    f = cpu_path.FIXED_FRACTION if fraction is None else fraction
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
    sample = (f"one synthetic {CONFIGS[name]['scenes']} scene ({coord.shape[0]} pts), reference op schedule fwd+bwd on the "
              f"first {f:.4f} (fixed) of every level's query rows against the full level "
              f"({int(tot_pts / max(steps, 1))} level-0 points per step); cdist+topk kNN, torch gather/index_put scatter")
    return dict(value=value, ms_per_step=tot_s / max(steps, 1) * 1e3, cores=cores, sample=sample,
                fraction=f, family_seconds={k2: round(v, 3) for k2, v in fam.items()})


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = DEFAULT_STEPS_REFERENCE if args.steps is None else args.steps
    warmup = 1 if args.warmup is None else args.warmup
    # a step is a fixed 4 % row sample of one scene (~5 s on 16 cores); long runs shrink the sample, not the contract
    frac = None if steps + warmup <= 40 else 0.04 * 40.0 / (steps + warmup)
    r = cpu_reference(args.config, steps, warmup, fraction=frac)
    rooms = args.rooms_per_gpu or CONFIGS[args.config]["rooms"]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, rooms, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "family_seconds": r["family_seconds"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench run (rank, device, collective helpers)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        # torchrun exports RANK / LOCAL_RANK / WORLD_SIZE; a plain `python bench.py` is a single process
        under_torchrun = all(k in os.environ for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
        self.world = int(os.environ["WORLD_SIZE"]) if under_torchrun else 1
        self.rank = int(os.environ["RANK"]) if under_torchrun else 0
        self.local = int(os.environ["LOCAL_RANK"]) if under_torchrun else 0
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                             "(use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import datetime

            # a collective mismatch should fail in minutes, not in the default 10 (every rank holds a GPU meanwhile)
            dist.init_process_group("nccl", device_id=self.dev,
                                    timeout=datetime.timedelta(seconds=int(os.environ.get("AOPT_NCCL_TIMEOUT_S", "180"))))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class GradAllReduce:
    """The DDP gradient all-reduce of the op-schedule step at N > 1 (the only collective).  The schedule
    has no parameters, so the payload is a resident fp32 buffer of the model's gradient size.  It is issued between
    the forward and the backward pass on its own stream — DDP overlaps its buckets with the backward pass the same
    way (engines/defaults.py:38) — and joined at the end of the step."""

    def __init__(self, ctx, n_params):
        self.ctx = ctx
        self.buf = ctx.torch.zeros(n_params, device=ctx.dev)
        self.stream = ctx.torch.cuda.Stream(device=ctx.dev)
        self.pending = False

    def launch(self):
        torch = self.ctx.torch
        self.stream.wait_stream(torch.cuda.current_stream(self.ctx.dev))
        with torch.cuda.stream(self.stream):
            self.ctx.dist.all_reduce(self.buf)
        self.pending = True

    def join(self):
        if self.pending:
            self.ctx.torch.cuda.current_stream(self.ctx.dev).wait_stream(self.stream)
            self.pending = False


def time_schedule(ctx, sched, coord, offset, steps, warmup, trace_steps, allreduce=None, sampler=None, with_clocks=False):
    """W warm-up steps, then EXACTLY `steps` timed steps between barrier + synchronize, CUDA events on the launching
    stream, max over ranks.  The last `trace_steps` timed steps carry per-call CUDA events (roofline table)."""
    import gc

    torch = ctx.torch
    from ao_b200 import _lib

    def one_step(c, o):
        acc = sched.step(c, o)
        if allreduce is not None:
            allreduce.join()
        return acc

    sched.between = allreduce.launch if allreduce is not None else None
    for _ in range(warmup):
        one_step(coord, offset)
    ctx.barrier()
    if with_clocks:     # EVERY rank takes this branch (collectives inside); only rank 0 owns a sampler
        if sampler is not None:
            sampler.start()
            time.sleep(0.3)
        ctx.barrier()
        # the GPU idled while the clock sampler started: more untimed steps bring clocks / power state back up
        # (without them the first two or three timed steps run 5-20 % slow)
        for _ in range(2):
            one_step(coord, offset)
        ctx.barrier()
    gc.collect()
    gc.disable()      # the Python GC stays off while timing
    trace_steps = min(trace_steps, steps)
    _lib.trace_prepare(1024 * max(1, trace_steps))
    launches0 = _lib.kernel_launches()
    trace = []
    overlap_was = _lib.overlap_mode()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]   # end of every step (diagnostics)
    wall0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        if i > 0:
            marks[i - 1].record()
        if trace_steps and i == steps - trace_steps:
            # per-call CUDA events are recorded on the LAST trace_steps timed steps only (an event pair per call on
            # all ~250 calls of a step costs ~1 ms of host time).  Per-kernel roofline = the kernel running ALONE:
            # these steps issue everything on one stream; they are part of the timed region.
            trace = _lib.trace_start()
            _lib.overlap(False)
        one_step(coord, offset)
    e1.record()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    _lib.trace_stop()
    _lib.overlap_mode(overlap_was)
    gc.enable()
    launches = _lib.kernel_launches() - launches0
    clocks = sampler.stop() if sampler is not None else None
    ms_total = e0.elapsed_time(e1)
    bounds = [e0] + marks[: steps - 1] + [e1]
    step_ms = [round(bounds[i].elapsed_time(bounds[i + 1]), 3) for i in range(steps)]
    ms_step = ctx.max_over_ranks(ms_total) / steps
    return dict(ms_step=ms_step, step_ms=step_ms, trace=trace, trace_steps=trace_steps, launches=int(launches),
                clocks=clocks, wall_ms_per_step=wall / steps * 1e3, sizes=list(sched.last_sizes), one_step=one_step)


def brief(values):
    """min / median / max of a list of per-step times."""
    v = sorted(values)
    return {"min": v[0], "median": v[len(v) // 2], "max": v[-1]} if v else None


def run_b200_arm(args):
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    dev, world, rank = ctx.dev, ctx.world, ctx.rank

    from ao_b200 import _lib
    from ao_b200.schedule import PointOpsSchedule

    _lib.load()
    name = args.config
    conf = CONFIGS[name]
    rooms = args.rooms_per_gpu or conf["rooms"]
    steps = DEFAULT_STEPS if args.steps is None else args.steps
    warmup = max(5 if args.warmup is None else args.warmup, args.min_warmup)

    # ---- synthetic batch: scenes [rank*R, rank*R+R) -------------------------------------------------
    coord_np, feat_np, off_np = make_batch(name, rooms, rank)
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
    cfg = make_schedule_config(name)
    sched = PointOpsSchedule(cfg, device=dev, seed=rank)
    allreduce = GradAllReduce(ctx, conf["params"]) if world > 1 else None

    # ---- timed region: resident inputs, CUDA events on the launching stream ----------------------------
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    r = time_schedule(ctx, sched, coord, offset, steps, warmup, TRACE_STEPS, allreduce, sampler, with_clocks=True)
    ms_step, sizes, trace, trace_steps, one_step = r["ms_step"], r["sizes"], r["trace"], r["trace_steps"], r["one_step"]
    value = world * n0 / (ms_step * 1e-3) / 1e6

    # ---- e2e: host buffers → H2D → schedule → D2H of the result scalar, wall clock ---------------------
    # Every step copies its own inputs from pinned host memory and reads the result scalar back (a sync per
    # step, like a training loop that logs its loss).  As a data loader with pin_memory / non_blocking would, the
    # copy of step i+1 is issued on a copy stream while step i computes; all copies lie inside the timed region.
    import gc

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

    def e2e_loop(n_steps):
        last = 0.0
        nxt = stage()
        for i in range(n_steps):
            t_step = time.perf_counter()
            c, f, o, ev = nxt
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            for t_ in (c, f, o):
                t_.record_stream(cur)        # allocated on the copy stream, consumed on the compute stream
            if i + 1 < n_steps:
                nxt = stage()
            acc = one_step(c, o)
            del f
            last = float(acc.sum().item())   # D2H read of the step result (the loss stand-in)
            e2e_step_ms.append(round((time.perf_counter() - t_step) * 1e3, 3))
        return last

    e2e_steps = 0 if args.skip_e2e else steps
    e2e_value = None
    if e2e_steps:
        # warm-up of this leg too: its copy-stream buffers are new allocations (the first two steps of a cold loop
        # took 37 and 16 ms in cudaMalloc, against 8.4 ms once the double buffers exist)
        e2e_loop(max(3, min(warmup, 5)))
        ctx.barrier()
        del e2e_step_ms[:]
        gc.collect()
        gc.disable()
        w0 = time.perf_counter()
        e2e_loop(e2e_steps)
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(max(time.perf_counter() - w0, 1e-9))
        gc.enable()
        e2e_value = world * n0 * e2e_steps / e2e_s / 1e6
    h2d = coord_h.numel() * 4 + feat_h.numel() * 4 + off_h.numel() * 4

    # ---- roofline of the dominant HBM-bound kernel, measured live over the timed region ---------------
    peak, peak_src = hbm_peak()
    line = None
    if rank == 0:
        kernels = per_step(summarise_trace(trace, sizes, ms_step * max(trace_steps, 1), peak), max(trace_steps, 1))
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
            # same launch shape from the committed ncu --set full capture (profiles/traffic.json: s3dis4 shapes)
            big = dom["largest"]
            info = traffic if isinstance(traffic, dict) and name == "s3dis4" else {}
            roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": big["gbs"], "peak": peak, "unit": "GB/s",
                        "frac": big["frac"], "traffic": info.get("dram_bytes_per_launch"),
                        "traffic_source": info.get("source"), "peak_source": peak_src,
                        "alg_bytes_per_launch": big["alg_bytes"], "us_per_launch": big["us_per_launch"],
                        "launches_per_step": big["launches_per_step"],
                        "launch_shape": "level 0: N=%d, C=%d, G=%d, k per config.k" % (sizes[0], cfg.channels[0], cfg.groups[0]),
                        "share_of_step_all_launches": dom["share"], "frac_all_launches": dom["frac"]}
        hbm_ms = sum(kr["ms_per_step"] for kr in hbm_rows)
        hbm_gb = sum(kr["alg_gb_per_step"] for kr in hbm_rows)
        knn_row = next((kr for kr in kernels if kr["kernel"] == "aopt_knn_query"), None)
        # brute-force-equivalent pair count of the searches (scenes of a level are near-equal in size)
        pairs = sum(float(a[0]) * float(a[2]) / max(a[3], 1) for nm, a, _, _ in trace if nm == "aopt_knn_query") / max(trace_steps, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(name, rooms, world, cfg),
            "level_sizes": sizes,
            "e2e": None if e2e_value is None else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "timing": "wall clock incl. python launch overhead; H2D of step i+1 overlaps step i on a copy stream"},
            "gpu_launches": r["launches"],
            "clocks": r["clocks"],
            "roofline": roofline,
            "hbm_kernels_total": {"ms_per_step": round(hbm_ms, 4), "alg_gb_per_step": round(hbm_gb, 3),
                                  "gbs": round(hbm_gb / (hbm_ms * 1e-3), 1) if hbm_ms > 0 else None,
                                  "frac": round(hbm_gb / (hbm_ms * 1e-3) / peak, 4) if hbm_ms > 0 else None,
                                  "whole_step_frac": round(hbm_gb / (ms_step * 1e-3) / peak, 4)},
            "knn": None if knn_row is None else {"ms_per_step": knn_row["ms_per_step"], "calls_per_step": knn_row["calls_per_step"],
                                                 "brute_force_pairs_per_step": pairs,
                                                 "equiv_pairs_per_s": pairs / (knn_row["ms_per_step"] * 1e-3)},
            "kernels": kernels,
            "host_wall_ms_per_step": r["wall_ms_per_step"],
            "step_ms": brief(r["step_ms"]) if steps > 40 else r["step_ms"], "traced_steps": trace_steps,
            "e2e_step_ms": brief(e2e_step_ms) if len(e2e_step_ms) > 40 else e2e_step_ms,
        }

    # ---- other schedule variants (fewer steps; same timing rules) ---------------------------------------
    del sched, trace, one_step
    r = None
    torch.cuda.empty_cache()
    variants = {}
    if not args.no_variants:
        for label, over in conf["variants"]:
            vcfg = make_schedule_config(name, **over)
            vs = PointOpsSchedule(vcfg, device=dev, seed=rank)
            vr = time_schedule(ctx, vs, coord, offset, min(steps, VARIANT_STEPS), 3, 1, allreduce)
            if rank == 0:
                vk = per_step(summarise_trace(vr["trace"], vr["sizes"], vr["ms_step"], peak), 1)
                variants[label] = {"value": world * n0 / (vr["ms_step"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": vr["ms_step"],
                                   "steps": min(steps, VARIANT_STEPS), "overrides": over, "gpu_launches_per_step": vr["launches"] // min(steps, VARIANT_STEPS),
                                   "kernels": [{k2: kr[k2] for k2 in ("kernel", "calls_per_step", "ms_per_step", "frac")} |
                                               {"largest_us": kr["largest"]["us_per_launch"], "largest_frac": kr["largest"]["frac"]}
                                               for kr in vk[:8]]}
            del vs, vr
            torch.cuda.empty_cache()
    if line is not None and variants:
        line["variants"] = variants

    # ---- full PTv2m2 training step: bf16 autocast + AdamW, DDP-wrapped at N > 1 --------------------------
    if not args.no_model:
        try:
            ms_model = model_step(ctx, name, coord, feat, offset, args.bucket_cap_mb)
            if line is not None:
                line["model_step"] = ms_model
        except Exception as ex:  # pragma: no cover
            if line is not None:
                line["model_step"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    if rank == 0:
        # ---- the reference's own op schedule on this GPU (oracle/_ref kNN kernel + torch op chains) ----------
        if not args.no_gpu_reference and world == 1:
            try:
                line["gpu_reference"] = gpu_reference(ctx, name, cfg, coord, offset, ms_step)
            except Exception as ex:  # pragma: no cover
                line["gpu_reference"] = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()
        if name == "kitti120k" and not args.no_variants and world == 1:
            try:
                line["sweep"] = kitti_sweep(ctx, coord, offset, peak)
            except Exception as ex:  # pragma: no cover
                line["sweep"] = {"error": repr(ex)[:300]}
        # ---- CPU baseline on the box's host cores (bounded sample) -----------------------------------------
        if not args.no_cpu_baseline and world == 1:
            rc = cpu_reference(name, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": rc["value"], "unit": UNIT, "cores": rc["cores"], "kind": "port",
                                    "sample": rc["sample"], "family_seconds": rc["family_seconds"]}
        emit(line)
    ctx.finish()


def model_step(ctx, name, coord, feat, offset, bucket_cap_mb, steps=10):
    """Full training step of the config's PTv2m2 (ao_b200.ptv2): bf16 autocast forward, cross-entropy, backward,
    AdamW — the reference's Trainer.run_step (pointcept/engines/train.py:173-200).  At N > 1 the model is wrapped by
    ao_b200.sharding.ddp_wrap (engines/defaults.py:30-43), so the gradient all-reduce is DDP's own, bucketed and
    overlapped with the backward pass."""
    torch = ctx.torch
    from ao_b200 import ptv2, sharding

    dev = ctx.dev
    torch.manual_seed(0)
    mcfg = getattr(ptv2, CONFIGS[name]["model"])
    model = ptv2.PointTransformerV2(**mcfg).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    net = sharding.ddp_wrap(model, ctx.local, bucket_cap_mb=bucket_cap_mb) if ctx.world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, fused=True)   # one multi-tensor kernel pair instead of ~12 per step
    target = torch.randint(0, mcfg["num_classes"], (coord.shape[0],), device=dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(dict(coord=coord, feat=feat, offset=offset))
        loss = torch.nn.functional.cross_entropy(logits.float(), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    ctx.barrier()
    # the dense layers are compute-heavy: under the board's power cap their clocks (hence this number) depend on how
    # long the GPU has been loaded before — the SM clock during these steps is reported next to the time
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    ctx.barrier()
    clocks = sampler.stop() if sampler else None
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    out = {"what": f"full PTv2m2 ({CONFIGS[name]['model']}) training step: bf16 GEMMs (cuBLAS; q|k|v as one product, cached bf16 "
                   "weights) + the library's point operators, BatchNorm / ReLU / DropPath / residual kernels (bn_act) and "
                   "weight-encoding tail (we_tail) + cross-entropy + AdamW (fused=True)" + (f"; DistributedDataParallel over {ctx.world} ranks (broadcast_buffers=False, "
                   f"bucket_cap_mb={bucket_cap_mb}, gradient_as_bucket_view), NCCL all-reduce overlapped with backward"
                   if ctx.world > 1 else ""),
           "ms_per_step": ms, "mpoints_per_s": ctx.world * coord.shape[0] / (ms * 1e-3) / 1e6, "loss": float(loss.item()),
           "parameters": n_params, "grad_mb": round(n_params * 4 / 1e6, 2), "steps": steps, "clocks": clocks,
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del net, model, opt
    return out


def gpu_reference(ctx, name, cfg, coord, offset, ours_ms_step, steps=2):
    """The reference's op schedule with the reference's implementations on this B200 (oracle/gpu_path.py: unmodified
    reference kNN kernel from oracle/_ref + the reference's torch op chains on CUDA tensors).  Checker-side code,
    timed only here; never on the product path."""
    torch = ctx.torch
    from oracle import gpu_path, ref_cuda

    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/libpointops_ref.so not built"}
    n0 = coord.shape[0]
    out = {}
    # (1) the reference kNN kernel alone (knn_query_cuda_kernel.cu:60-104): k=16 self search on level 0
    ev = lambda: torch.cuda.Event(enable_timing=True)
    ref_cuda.knn_query(16, coord, offset)
    a, b = ev(), ev()
    a.record()
    ref_cuda.knn_query(16, coord, offset)
    b.record()
    torch.cuda.synchronize()
    out["knn_k16_self_level0_ms"] = round(a.elapsed_time(b), 3)
    from ao_b200 import pointops
    pointops.knn_query_raw(16, coord, offset)
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        pointops.knn_query_raw(16, coord, offset)
    b.record()
    torch.cuda.synchronize()
    out["knn_k16_self_level0_ms_this_repo"] = round(a.elapsed_time(b) / 5, 3)
    # (2) the whole reference op schedule
    sched = gpu_path.GpuReferenceSchedule(cfg, ctx.dev, seed=0)
    sched.step(coord, offset)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(steps):
        sched.step(coord, offset)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    out.update({"what": "reference op schedule on the same B200: unmodified reference kNN kernel (oracle/_ref; 7 self + "
                        "3 cross searches per step at S3DIS cfg, no neighbour-list sharing) + the reference's pure-torch "
                        "grouping / GVA tail / GridPool / interpolation on CUDA tensors, autograd backward; same "
                        "synthetic stand-ins for the dense layers", "ms_per_step": ms, "steps": steps,
                "value": n0 / (ms * 1e-3) / 1e6, "unit": UNIT, "speedup_of_this_repo": ms / ours_ms_step,
                "peak_mem_gb": torch.cuda.max_memory_allocated(ctx.dev) / 1e9})
    del sched
    return out


def kitti_sweep(ctx, coord, offset, peak):
    """BASELINE.json configs[4]: kNN over k in {8,16,32} and the fused GVA pair (gva_relation + gva_aggregate,
    forward + backward) over C in {48,96,192,384} on the level-0 scans."""
    torch = ctx.torch
    from ao_b200 import pointops

    ev = lambda: torch.cuda.Event(enable_timing=True)
    n = coord.shape[0]
    out = {"n": n, "knn": [], "gva": []}

    def timed(fn, reps=5):
        fn()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for k in (8, 16, 32):
        ms = timed(lambda: pointops.knn_query_raw(k, coord, offset))
        out["knn"].append({"k": k, "ms": round(ms, 4), "mpoints_per_s": round(n / ms / 1e3, 1)})
    idx, _ = pointops.knn_query(16, coord, offset)
    k = 16
    for c in (48, 96, 192, 384):
        g = c // 8
        key, query, value = (torch.randn(n, c, device=ctx.dev, requires_grad=True) for _ in range(3))
        peb = torch.randn(n, k, c, device=ctx.dev, requires_grad=True)
        logits = torch.randn(n, k, g, device=ctx.dev, requires_grad=True)
        g_rel, g_out = torch.randn(n, k, c, device=ctx.dev), torch.randn(n, c, device=ctx.dev)

        def fb():
            rel = pointops.gva_relation(key, query, idx)
            o = pointops.gva_aggregate(value, peb, logits, idx, g)
            torch.autograd.grad([rel, o], [key, query, value, peb, logits], [g_rel, g_out])

        ms = timed(fb, 3)
        nb = (4.0 * n * k + 8.0 * n * c + 4.0 * n * k * c) + (8.0 * n * c + 4.0 * n * k * c + 8.0 * n * k * g + 4.0 * n * k) \
            + (4.0 * n * k * c + 4.0 * n * k + 4.0 * (n + 1) + 8.0 * n * c) \
            + (8.0 * n * c + 8.0 * n * k * c + 8.0 * n * k * g + 4.0 * n * k) + (4.0 * n * k * g + 8.0 * n * c + 4.0 * (n + 1) + 4.0 * n * k)
        out["gva"].append({"c": c, "g": g, "k": k, "fwd_bwd_ms": round(ms, 4), "alg_gb": round(nb / 1e9, 3),
                           "gbs": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / peak, 4)})
        del key, query, value, peb, logits, g_rel, g_out
    return out


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
    ap.add_argument("--steps", type=int, default=None, help=f"timed steps (default {DEFAULT_STEPS}: a timed region >= 2 s; "
                    f"reference arm: {DEFAULT_STEPS_REFERENCE})")
    ap.add_argument("--warmup", type=int, default=None, help="untimed warm-up steps (default 5)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="s3dis4", choices=sorted(CONFIGS), help="BASELINE.json configs[1..4]")
    ap.add_argument("--rooms-per-gpu", type=int, default=None, help="override the config's scenes per GPU")
    ap.add_argument("--bucket-cap-mb", type=float, default=4.0, help="DDP bucket size of the N>1 model step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-model", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
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
