"""Per-step device time of the point-operator schedule with the side-stream overlap on / off (A/B)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ao_b200 import _lib, scenes
from ao_b200.schedule import PointOpsSchedule, ScheduleConfig

dev = torch.device("cuda", 0)
coord_np, feat_np, off_np = scenes.s3dis_batch(4, 80000)
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
sched = PointOpsSchedule(ScheduleConfig.s3dis(), device=dev, seed=0)
for _ in range(3):
    sched.step(coord, offset)
torch.cuda.synchronize()
for mode in sys.argv[1:] or ["0", "1", "0", "1"]:
    _lib.overlap_mode({"1": True, "0": False}.get(mode, None))   # "d" = per-role defaults (AOPT_OVERLAP_ROLES)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        sched.step(coord, offset)
        e1.record()
        host = (time.perf_counter() - w0) * 1e3
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), host))
    print("overlap", mode, "roles", os.environ.get("AOPT_OVERLAP_ROLES", "knn"), " device ms:", " ".join("%.2f" % a for a, _ in ts),
          " host-enqueue ms:", " ".join("%.2f" % b for _, b in ts), flush=True)
print("mem GB", torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_stats().get("num_alloc_retries"), "device mallocs", torch.cuda.memory_stats().get("num_device_alloc"))
