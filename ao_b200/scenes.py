"""Synthetic S3DIS / ScanNet / SemanticKITTI-shaped inputs.

There is no dataset in the reference checkout (data/ is a placeholder) and no network, so every
benchmark and parity input is synthetic but follows the reference's own data pipeline semantics
(/root/reference/pointcept/datasets/transform.py):

  GridSample(grid)        one ORIGINAL (un-snapped) point per voxel, random member     :792-830
  SphereCrop(point_max)   the point_max points nearest a random centre, in
                          distance-sorted order (ShufflePoint is commented out in the
                          S3DIS config)                                                :959-995
  CenterShift(apply_z=False)                                                            :129-142
  collate: concatenate scenes, offset = cumulative point counts    datasets/utils.py:29-37

numpy only (runs on the host, like the reference's DataLoader workers); seeds are explicit so that
CPU oracle, GPU kernels and committed golden vectors see identical bytes.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

SEED_BASE = 4242  # configs/s3dis/semseg-pt-v2m2-0-base.py:7


def _box_surface(rng, n, lo, hi):
    """n points uniform on the 6 faces of the axis-aligned box [lo, hi] (area-weighted)."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    ext = hi - lo
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2],
                      ext[0] * ext[1], ext[0] * ext[1]])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    p = lo + rng.random((n, 3)) * ext
    axis = face // 2
    side = face % 2
    p[np.arange(n), axis] = np.where(side == 0, lo[axis], hi[axis])
    return p


def _grid_sample(rng, coord, grid):
    """GridSample train mode: one random original point per occupied voxel."""
    cell = np.floor(coord / grid).astype(np.int64)
    cell -= cell.min(0)
    dims = cell.max(0) + 1
    key = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    count = np.diff(np.r_[first, ks.size])
    pick = first + rng.integers(0, count.max(), count.size) % count
    return order[pick]


def _sphere_crop(rng, coord, point_max):
    if coord.shape[0] <= point_max:
        return np.arange(coord.shape[0])
    center = coord[rng.integers(coord.shape[0])]
    return np.argsort(np.sum(np.square(coord - center), 1), kind="stable")[:point_max]


def indoor_room(room: int, n_points: int = 80000, grid: float = 0.04, raw: int = 600000) -> Tuple[np.ndarray, np.ndarray]:
    """One synthetic room → (coord (n,3) float32, color (n,3) float32 in [-1,1])."""
    rng = np.random.default_rng(SEED_BASE + room)
    dims = np.array([rng.uniform(6, 12), rng.uniform(5, 9), rng.uniform(2.7, 3.2)])
    for _attempt in range(8):
        n_wall = int(raw * 0.75)
        pts = [_box_surface(rng, n_wall, [0, 0, 0], dims)]
        n_f = (raw - n_wall) // 12
        for _ in range(12):  # furniture boxes standing on the floor
            size = np.array([rng.uniform(0.4, 2.0), rng.uniform(0.4, 2.0), rng.uniform(0.4, 1.8)])
            size = np.minimum(size, dims * 0.9)
            pos = np.array([rng.uniform(0, dims[0] - size[0]), rng.uniform(0, dims[1] - size[1]), 0.0])
            pts.append(_box_surface(rng, n_f, pos, pos + size))
        coord = np.concatenate(pts, 0)
        coord += rng.normal(0.0, 0.003, coord.shape)          # 3 mm jitter: distance ties have measure zero
        coord += rng.uniform(0.0, 0.37, 3)                     # random origin shift
        coord = coord[_grid_sample(rng, coord, grid)]
        if coord.shape[0] >= n_points:
            break
        dims[:2] *= 1.2                                        # room too small for n_points voxels
        raw = int(raw * 1.3)
    coord = coord[_sphere_crop(rng, coord, n_points)]
    mn, mx = coord.min(0), coord.max(0)
    coord = coord - np.array([(mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, 0.0])   # CenterShift(apply_z=False)
    color = rng.uniform(-1.0, 1.0, (coord.shape[0], 3))
    return coord.astype(np.float32), color.astype(np.float32)


def outdoor_scan(scan: int, n_points: int = 120000, grid: float = 0.05) -> Tuple[np.ndarray, np.ndarray]:
    """SemanticKITTI-shaped sweep: rings x azimuths rays on a ground plane with boxes, clipped to
    (-35.2,-35.2,-4)…(35.2,35.2,2), 0.05 m grid, cropped to n_points → (coord, strength (n,1))."""
    rng = np.random.default_rng(SEED_BASE + 1000 + scan)
    # several sweeps' worth of rays (the real scans are denser than one 64x2048 sweep after the
    # 0.05 m grid and the range clip leave 120k points): 192 rings x 4096 azimuths
    n_ring, n_az = 192, 4096
    az = np.tile(np.linspace(-np.pi, np.pi, n_az, endpoint=False), n_ring)
    elev = np.repeat(np.deg2rad(np.linspace(-24.8, 2.0, n_ring)), n_az)
    az = az + rng.normal(0, 2e-4, az.shape)
    elev = elev + rng.normal(0, 2e-4, elev.shape)
    sensor_h = 1.73
    d = np.stack([np.cos(elev) * np.cos(az), np.cos(elev) * np.sin(az), np.sin(elev)], 1)
    # ground hit
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d[:, 2] < -1e-3, sensor_h / -d[:, 2], 80.0)
    # boxes (cars / walls): axis-aligned slabs, keep the nearest hit
    for _ in range(40):
        c = np.array([rng.uniform(-35, 35), rng.uniform(-35, 35)])
        if np.linalg.norm(c) < 3:
            continue
        half = np.array([rng.uniform(0.8, 6.0), rng.uniform(0.8, 6.0)])
        top = rng.uniform(1.2, 3.5) - sensor_h
        lo = np.r_[c - half, -sensor_h]
        hi = np.r_[c + half, top]
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = lo / d
            t2 = hi / d
        tmin = np.nanmax(np.minimum(t1, t2), 1)
        tmax = np.nanmin(np.maximum(t1, t2), 1)
        hit = (tmax >= tmin) & (tmin > 0)
        t = np.where(hit & (tmin < t), tmin, t)
    keep = t < 79.0
    coord = d[keep] * t[keep, None] + rng.normal(0, 0.01, (keep.sum(), 3))
    clip = (np.abs(coord[:, 0]) < 35.2) & (np.abs(coord[:, 1]) < 35.2) & (coord[:, 2] > -4) & (coord[:, 2] < 2)
    coord = coord[clip]
    coord = coord[_grid_sample(rng, coord, grid)]
    coord = coord[_sphere_crop(rng, coord, n_points)]
    strength = rng.uniform(0, 1, (coord.shape[0], 1))
    return coord.astype(np.float32), strength.astype(np.float32)


def collate(scenes: List[Tuple[np.ndarray, np.ndarray]]):
    """Concatenate scenes into the offset-encoded batch layout: (coord, feat=[coord|extra], offset int32)."""
    coord = np.concatenate([s[0] for s in scenes], 0)
    extra = np.concatenate([s[1] for s in scenes], 0)
    feat = np.concatenate([coord, extra], 1).astype(np.float32)
    offset = np.cumsum([s[0].shape[0] for s in scenes]).astype(np.int32)
    return np.ascontiguousarray(coord), np.ascontiguousarray(feat), offset


def s3dis_batch(n_rooms: int = 4, n_points: int = 80000, first_room: int = 0):
    """BASELINE.json configs[1]: S3DIS-shaped batch (feat = [coord, color], 6 channels)."""
    return collate([indoor_room(first_room + r, n_points) for r in range(n_rooms)])


def scannet_batch(n_rooms: int = 1, n_points: int = 150000, first_room: int = 100):
    """BASELINE.json configs[3]: ScanNet-shaped rooms at 0.02 m voxels, 9 channels (coord+color+normal)."""
    scenes = []
    for r in range(n_rooms):
        coord, color = indoor_room(first_room + r, n_points, grid=0.02, raw=1200000)
        rng = np.random.default_rng(SEED_BASE + 5000 + r)
        normal = rng.normal(size=coord.shape)
        normal /= np.linalg.norm(normal, axis=1, keepdims=True)
        scenes.append((coord, np.concatenate([color, normal.astype(np.float32)], 1)))
    return collate(scenes)


def kitti_batch(n_scans: int = 1, n_points: int = 120000, first_scan: int = 0):
    """BASELINE.json configs[4]: SemanticKITTI-shaped scans, 4 channels (coord+strength)."""
    return collate([outdoor_scan(first_scan + s, n_points) for s in range(n_scans)])


def small_batch(seed: int, sizes=(700, 5, 1300, 257), dup: int = 0):
    """Tiny adversarial batch for parity tests: unequal scenes, one shorter than k, optional
    duplicated coordinates (tie rows)."""
    rng = np.random.default_rng(seed)
    scenes = []
    for s in sizes:
        c = rng.uniform(-2, 2, (s, 3)).astype(np.float32)
        if dup and s > 2 * dup:
            c[-dup:] = c[:dup]
        scenes.append((c, rng.uniform(-1, 1, (s, 3)).astype(np.float32)))
    return collate(scenes)
