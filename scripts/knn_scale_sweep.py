"""Grid-kNN cell-edge sweep (AOPT_KNN_CELL_SCALE = cell edge / sampled k-th neighbour distance) and TILE vs GRID at
every level of the S3DIS pyramid.  Device time per call (CUDA events, best of 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ao_b200 import pointops, scenes

dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
levels = [(coord, offset)]
for gs in (0.1, 0.2, 0.4):
    c, o = levels[-1]
    (nc, _, no), _ = pointops.grid_pool(c, c.clone(), o, gs)
    levels.append((nc.contiguous(), no.int()))

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3)
    return best

scales = [float(x) for x in (sys.argv[1:] or ["0.8", "0.9", "1.0", "1.1", "1.25", "1.4", "1.6"])]
print("case".ljust(22) + "".join(f"  s={s:<5}" for s in scales) + "    tile")
for li, (c, o) in enumerate(levels):
    row = []
    for s in scales:
        os.environ["AOPT_KNN_CELL_SCALE"] = str(s)
        row.append(timeit(lambda: pointops.knn_query_raw(16, c, o, method="grid")))
    t = timeit(lambda: pointops.knn_query_raw(16, c, o, method="tile")) if c.shape[0] <= 60000 else float("nan")
    print(f"self L{li} n={c.shape[0]:<7d}".ljust(22) + "".join(f"  {x:7.1f}" for x in row) + f"  {t:7.1f}")
for li in range(len(levels) - 1):
    (fc, fo), (cc, co) = levels[li], levels[li + 1]
    row = []
    for s in scales:
        os.environ["AOPT_KNN_CELL_SCALE"] = str(s)
        row.append(timeit(lambda: pointops.knn_query_raw(3, cc, co, fc, fo, method="grid")))
    t = timeit(lambda: pointops.knn_query_raw(3, cc, co, fc, fo, method="tile")) if cc.shape[0] <= 60000 else float("nan")
    print(f"cross L{li+1}->L{li} k=3".ljust(22) + "".join(f"  {x:7.1f}" for x in row) + f"  {t:7.1f}")
