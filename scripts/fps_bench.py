"""Farthest point sampling: csrc/fps.cu against the unmodified reference kernel on PTv1's TransitionDown
shapes (S3DIS batch of 4 rooms x 80k points, stride 4: 80k -> 20k -> 5k -> 1250).  CUDA events, best of 3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ao_b200 import pointops, scenes
from oracle import ref_cuda

dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
xyz, off = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


print(f"{'stage':28s} {'n':>8s} {'m':>8s} {'ours ms':>9s} {'ref ms':>9s} {'speed-up':>8s} {'us/iter':>8s} equal")
for stage in range(3):
    noff = (off // 4).int()
    ours = pointops.farthest_point_sampling(xyz, off, noff)
    t_o = timeit(lambda: pointops.farthest_point_sampling(xyz, off, noff))
    if ref_cuda.available():
        ref, _ = ref_cuda.farthest_point_sampling(xyz, off, noff)
        t_r = timeit(lambda: ref_cuda.farthest_point_sampling(xyz, off, noff), reps=2)
        eq = bool(torch.equal(ours, ref))
    else:
        t_r, eq = float("nan"), None
    m_scene = int(noff[0])
    print(f"TransitionDown {stage} (b=4)       {xyz.shape[0]:8d} {int(noff[-1]):8d} {t_o:9.2f} {t_r:9.2f} {t_r/t_o:8.1f} {t_o*1e3/m_scene:8.2f} {eq}")
    xyz, off = xyz[ours.long()].contiguous(), noff
for cl in (1, 2, 4, 8, 16):
    os.environ["AOPT_FPS_CLUSTER"] = str(cl)
    c, _, o = scenes.s3dis_batch(4, 80000)
    x, of = torch.from_numpy(c).to(dev), torch.from_numpy(o).to(dev)
    no = (of // 4).int()
    try:
        print(f"cluster {cl:2d}: {timeit(lambda: pointops.farthest_point_sampling(x, of, no), reps=2):8.2f} ms")
    except Exception as ex:
        print(f"cluster {cl:2d}: {ex!r}"[:200])
