"""S3DIS train transforms on one raw-room-sized cloud (800k points): CenterShift -> GridSample(0.04, fnv, train) ->
SphereCrop(80000) -> CenterShift -> NormalizeColor, ao_b200.datapipe (GPU, incl. the H2D upload of the raw arrays)
against the numpy restatement of the reference (oracle/datapipe_ref.py, one host core like one dataloader worker)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ao_b200 import datapipe
from oracle import datapipe_ref as dref

rng = np.random.default_rng(0)
n = 800000
ext = np.array([9.0, 7.0, 3.0])
p = rng.random((n, 3)) * ext
face = rng.integers(0, 6, n)
for a in range(3):
    p[face == 2 * a, a] = 0.0
    p[face == 2 * a + 1, a] = ext[a]
p += rng.normal(0, 0.003, p.shape)
data = dict(coord=p.astype(np.float32), color=rng.integers(0, 256, (n, 3)).astype(np.float32),
            segment=rng.integers(0, 13, n).astype(np.int64))
cfg = [dict(type="CenterShift", apply_z=True),
       dict(type="GridSample", grid_size=0.04, hash_type="fnv", mode="train", keys=("coord", "color", "segment")),
       dict(type="SphereCrop", point_max=80000, mode="random"), dict(type="CenterShift", apply_z=False),
       dict(type="NormalizeColor")]
pipe = datapipe.Compose(cfg)


def gpu_once():
    np.random.seed(1)
    out = pipe({k: v for k, v in data.items()})
    torch.cuda.synchronize()
    return out


def cpu_once():
    np.random.seed(1)
    c0 = dref.center_shift(data["coord"], True)
    vh = dref.voxel_hash(c0, 0.04)
    r = np.random.randint(0, vh["count"].max(), vh["count"].size)
    i1, _ = dref.grid_sample_train(c0, 0.04, r)
    c1, col1, seg1 = c0[i1], data["color"][i1], data["segment"][i1]
    i2, _ = dref.sphere_crop_index(c1, c1[np.random.randint(c1.shape[0])], 80000)
    return dict(coord=dref.center_shift(c1[i2], False), color=dref.normalize_color(col1[i2]), segment=seg1[i2])


for _ in range(2):
    g = gpu_once()
t0 = time.perf_counter()
for _ in range(5):
    g = gpu_once()
t_gpu = (time.perf_counter() - t0) / 5
t0 = time.perf_counter()
c = cpu_once()
t_cpu = time.perf_counter() - t0
same = all(np.array_equal(g[k].cpu().numpy(), c[k]) for k in ("coord", "color", "segment"))
print(f"raw points {n}, after GridSample+SphereCrop {g['coord'].shape[0]}")
print(f"GPU pipeline (H2D of {sum(v.nbytes for v in data.values())/1e6:.1f} MB included): {t_gpu*1e3:8.2f} ms / sample")
print(f"numpy restatement of the reference, 1 core:            {t_cpu*1e3:8.2f} ms / sample   ({t_cpu/t_gpu:.0f}x)")
print("outputs identical:", same)
