"""A/B of the GRID kNN query kernel variants: top-k container (tuning "knn_topk": 1 = shared-memory heap,
2 = sorted list in registers) x insert schedule ("knn_pend": 1 = per-lane pending list drained per eight candidates,
2 = insert in place) on the BASELINE.json shapes: S3DIS levels 0..3 (4 rooms x 80k), one ScanNet room,
two KITTI scans; k in {8, 16, 32}.  Checks that both containers return identical bits, then prints CUDA-event times
(best of 7) of the whole search (grid build + query).
  python scripts/knn_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ao_b200 import _lib, pointops, scenes

dev = torch.device("cuda", 0)


def timeit(fn, reps=7):
    for _ in range(2):
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


def levels_of(coord_np, off_np, grids):
    out = [(torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev).int())]
    for gs in grids:
        c, o = out[-1]
        (nc, _, no), _ = pointops.grid_pool(c, c.clone(), o, gs)
        out.append((nc.contiguous(), no.int()))
    return out


cases = []
c, _, o = scenes.s3dis_batch(4, 80000)
for li, lv in enumerate(levels_of(c, o, (0.1, 0.2, 0.4))):
    cases.append((f"s3dis L{li}", lv))
c, _, o = scenes.scannet_batch(3, 150000)
cases.append(("scannet 3x150k", levels_of(c, o, ())[0]))
c, _, o = scenes.kitti_batch(2, 120000)
cases.append(("kitti 2 scans", levels_of(c, o, ())[0]))

VARIANTS = (("list+pend", 2, 1), ("list", 2, 2), ("heap+pend", 1, 1), ("heap", 1, 2))   # (label, knn_topk, knn_pend)
print(f"{'case':18s} {'n':>8s} {'k':>3s} " + " ".join(f"{v[0] + ' us':>12s}" for v in VARIANTS) + "  identical")
for name, (coord, offset) in cases:
    for k in (8, 16, 32):
        res, t = {}, {}
        for label, topk, pend in VARIANTS:
            _lib.set_tuning("knn_topk", topk)
            _lib.set_tuning("knn_pend", pend)
            res[label] = pointops.knn_query_raw(k, coord, offset, method="grid")
            t[label] = timeit(lambda: pointops.knn_query_raw(k, coord, offset, method="grid"))
        ref = res["list"]
        same = all(torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1]) for r in res.values())
        print(f"{name:18s} {coord.shape[0]:8d} {k:3d} " + " ".join(f"{t[v[0]]:12.1f}" for v in VARIANTS) + f"  {same}")
_lib.set_tuning("knn_topk", 0)
_lib.set_tuning("knn_pend", 0)
