"""Times aopt_knn_query TILE vs GRID over the level shapes of the S3DIS pyramid (self k=16, cross k=3).
Run on the GPU box; prints a table (us per call, best of 5 after warm-up)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ao_b200 import pointops, scenes

dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
levels = [(coord, offset)]
for gs in (0.1, 0.2, 0.4, 0.8):
    c, o = levels[-1]
    (nc, _, no), _ = pointops.grid_pool(c, c.clone(), o, gs)
    levels.append((nc.contiguous(), no.int()))


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3)
    return best


print(f"{'case':34s} {'n':>8s} {'m':>8s} {'tile us':>10s} {'grid us':>10s}")
for li, (c, o) in enumerate(levels):
    for k in (16,):
        t = timeit(lambda: pointops.knn_query_raw(k, c, o, method="tile")) if c.shape[0] <= 60000 else float("nan")
        g = timeit(lambda: pointops.knn_query_raw(k, c, o, method="grid"))
        print(f"self  L{li} k={k:<3d}                     {c.shape[0]:8d} {c.shape[0]:8d} {t:10.1f} {g:10.1f}")
for li in range(len(levels) - 1):
    (fc, fo), (cc, co) = levels[li], levels[li + 1]
    for k in (3, 1):
        t = timeit(lambda: pointops.knn_query_raw(k, cc, co, fc, fo, method="tile")) if cc.shape[0] <= 60000 else float("nan")
        g = timeit(lambda: pointops.knn_query_raw(k, cc, co, fc, fo, method="grid"))
        print(f"cross L{li+1}->L{li} k={k:<3d}                 {cc.shape[0]:8d} {fc.shape[0]:8d} {t:10.1f} {g:10.1f}")
