#!/usr/bin/env python
"""Generates tests/golden/datapipe_ref.npz by RUNNING THE REFERENCE'S OWN TRANSFORMS
(/root/reference/pointcept/datasets/transform.py: GridSample, SphereCrop, CenterShift, NormalizeColor — the file is
loaded unmodified; only `pointcept.utils.registry.Registry` is replaced by a 5-line stub so that the dataset package
and its heavy imports are not pulled in) on seeded clouds.  Run in the build container (the GPU box has no
/root/reference); the fixture is committed.  numpy here is 2.x: `coord / np.array(grid)` evaluates in fp64."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/pointcept/datasets/transform.py"


def cloud(seed, n=6000):
    """Seeded raw cloud: points on the faces of a room-sized box plus clutter, ~cm spacing so that a 0.04-0.05 m grid
    holds several points per voxel; uint8-like colours and integer labels."""
    rng = np.random.default_rng(seed)
    ext = np.array([2.4, 1.8, 1.2])
    p = rng.random((n, 3)) * ext
    face = rng.integers(0, 6, n)
    for a in range(3):
        p[face == 2 * a, a] = 0.0
        p[face == 2 * a + 1, a] = ext[a]
    p += rng.normal(0, 0.004, p.shape)
    p -= np.array([1.3, -0.7, 0.2])                      # negative coordinates on purpose (floor, not truncation)
    coord = p.astype(np.float32)
    color = rng.integers(0, 256, (n, 3)).astype(np.float32)
    normal = rng.normal(size=(n, 3)).astype(np.float32)
    segment = rng.integers(0, 13, n).astype(np.int64)
    return dict(coord=coord, color=color, normal=normal, segment=segment)


def load_reference():
    class Registry:
        def __init__(self, name):
            self.name = name

        def register_module(self, *a, **k):
            return lambda cls: cls

    for name in ("pointcept", "pointcept.utils", "pointcept.utils.registry"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pointcept.utils.registry"].Registry = Registry
    spec = importlib.util.spec_from_file_location("ref_transform", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = {"a": dict(seed=11, grid=0.04, hash_type="fnv"), "b": dict(seed=12, grid=0.05, hash_type="ravel"),
         "c": dict(seed=13, grid=0.02, hash_type="fnv")}

if __name__ == "__main__":
    T = load_reference()
    out = {"numpy_version": np.array(np.__version__)}
    for name, cs in CASES.items():
        data = cloud(cs["seed"])
        # ---- GridSample train: the reference's pick with a seeded host generator
        np.random.seed(100 + cs["seed"])
        d = T.GridSample(grid_size=cs["grid"], hash_type=cs["hash_type"], mode="train", keys=("coord", "color", "segment"),
                         return_discrete_coord=True, return_min_coord=True)(dict(data))
        out[f"{name}_train_coord"] = d["coord"]
        out[f"{name}_train_segment"] = d["segment"]
        out[f"{name}_train_discrete"] = d["discrete_coord"]
        out[f"{name}_min_coord"] = d["min_coord"]
        # keys / unique / counts the way GridSample computes them (same statements, :806-812)
        gs = T.GridSample(grid_size=cs["grid"], hash_type=cs["hash_type"])
        scaled = data["coord"] / np.array(cs["grid"])
        disc = np.floor(scaled).astype(int)
        disc -= disc.min(0)
        key = gs.hash(disc)
        uniq, count = np.unique(key, return_counts=True)
        out[f"{name}_key"], out[f"{name}_uniq"], out[f"{name}_count"] = key, uniq, count
        out[f"{name}_scaled_dtype"] = np.array(str(scaled.dtype))
        # ---- GridSample test: parts
        parts = T.GridSample(grid_size=cs["grid"], hash_type=cs["hash_type"], mode="test", keys=("coord", "segment"))(dict(data))
        out[f"{name}_test_nparts"] = np.array(len(parts))
        out[f"{name}_test_index0"] = parts[0]["index"]
        out[f"{name}_test_union"] = np.unique(np.concatenate([p["index"] for p in parts]))
        # ---- SphereCrop (random) + CenterShift + NormalizeColor
        np.random.seed(200 + cs["seed"])
        c = T.SphereCrop(point_max=2500, mode="random")(dict(data))
        out[f"{name}_crop_coord"] = c["coord"]
        out[f"{name}_crop_segment"] = c["segment"]
        if name != "a":
            continue
        out[f"{name}_shift_z"] = T.CenterShift(apply_z=True)(dict(coord=data["coord"].copy()))["coord"]
        out[f"{name}_shift_noz"] = T.CenterShift(apply_z=False)(dict(coord=data["coord"].copy()))["coord"]
        out[f"{name}_color"] = T.NormalizeColor()(dict(color=data["color"].copy()))["color"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datapipe_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: out[k] for k in out if k.endswith("dtype") or k.endswith("nparts")})
