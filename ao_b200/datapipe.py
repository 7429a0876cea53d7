"""GPU versions of the per-sample transforms in front of the PTv2m2 hot path.

Same class names, constructor arguments and `data_dict` protocol as
/root/reference/pointcept/datasets/transform.py — GridSample (:770-896), SphereCrop (:899-993),
CenterShift (:128-142), NormalizeColor (:100-104), ToTensor (:69-97), Collect (:24-57), Compose (:1107-1117) and
the offset-encoded collate of pointcept/datasets/utils.py:29-37 — but the arrays are CUDA tensors and the per-point
work runs in csrc/datapipe.cu (voxel hash, voxel pick, squared distances, row gathers), the library radix sort
and aopt_voxel_partition.  numpy arrays found in `data_dict` are uploaded on first touch.

Determinism contract (differs from numpy where numpy itself is unspecified):
  * `np.argsort(key)` (transform.py:811, :975) is an UNSTABLE sort — the order of equal keys depends on the numpy
    build (AVX-512 / AVX2 / scalar introsort).  Here every sort is STABLE (ties keep ascending original index):
    the voxel set, the counts and every tie-free position equal numpy's; inside a voxel the r-th point is the r-th
    by original index.  oracle/datapipe_ref.py restates the reference with kind="stable" and is pinned against the
    reference classes themselves on everything the unstable sort leaves determined.
  * random numbers come from the SAME host generator calls as the reference (`np.random.randint(0, count.max(),
    count.size)`, `np.random.randint(n)`), so a seeded numpy state gives the same picks.  GridSample therefore has
    one device->host read (n_vox, count.max()) like np.unique has.
  * `coord / np.array(grid_size)` is fp64 under NumPy >= 2 and fp32 under NumPy 1.x; `division="float64"` (default,
    what this image's numpy does) or "float32".
No CPU fallback: CPU-only inputs are uploaded, a missing CUDA device or library raises.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Dict, List

import numpy as np
import torch

from . import _lib


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ao_b200.datapipe: needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _t(a, dev=None):
    """numpy / CPU tensor -> CUDA tensor (pinned staging for large arrays); CUDA tensors pass through."""
    if isinstance(a, torch.Tensor):
        return a if a.is_cuda else a.to(dev or _dev(), non_blocking=True)
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev or _dev(), non_blocking=True)
    return a


def select_rows(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """x[index] for a CUDA tensor of 4-byte elements (rows of any width) through aopt_select_rows; other dtypes
    (int64 labels, bool masks, fp64) go through torch indexing."""
    if not x.is_cuda:
        raise ValueError("select_rows: CUDA tensors only")
    index = index.long().contiguous()
    if x.element_size() != 4 or x.dim() == 0 or not x.is_contiguous():
        return x[index]
    rows = index.numel()
    w = int(np.prod(x.shape[1:])) if x.dim() > 1 else 1
    out = torch.empty((rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if rows and w:
        with _lib.on_device(x.device):
            _lib.check(_lib.load().aopt_select_rows(rows, w, _lib.ptr(x), _lib.ptr(index), _lib.ptr(out), _lib.stream()),
                       "select_rows")
    return out


class VoxelHash:
    """Result of the GridSample front half: stable argsort of the voxel keys and the voxel boundaries."""

    def __init__(self, cell, order, idx_ptr, n_vox, count_max, stats, grid):
        self.cell, self.order, self.idx_ptr, self.n_vox, self.count_max = cell, order, idx_ptr, n_vox, count_max
        self.stats, self.grid = stats, grid

    @property
    def count(self):
        return torch.diff(self.idx_ptr)

    @property
    def min_coord(self):                                  # transform.py:808
        return self.stats[:3].double() * torch.as_tensor(self.grid, dtype=torch.float64, device=self.stats.device)


def voxel_hash(coord: torch.Tensor, grid_size, hash_type: str = "fnv", division: str = "float64") -> VoxelHash:
    """transform.py:806-812: scaled = coord / grid; discrete = floor(scaled) - min; key = hash(discrete);
    idx_sort = argsort(key) (stable here); unique keys -> boundaries."""
    lib = _lib.load()
    coord = _t(coord)
    if coord.dtype != torch.float32:
        coord = coord.float()
    coord = coord.contiguous()
    dev = coord.device
    n = coord.shape[0]
    g = np.array(np.broadcast_to(np.asarray(grid_size, dtype=np.float64), (3,)))
    if division not in ("float64", "float32"):
        raise ValueError("division must be 'float64' (NumPy >= 2) or 'float32' (NumPy 1.x)")
    cell = torch.empty((n, 3), dtype=torch.int32, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    stats = torch.zeros(6, dtype=torch.int32, device=dev)
    if n == 0:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return VoxelHash(cell, z, torch.zeros(1, dtype=torch.int32, device=dev), 0, 0, stats, g)
    with _lib.on_device(dev):
        _lib.check(lib.aopt_grid_sample_keys(n, _lib.ptr(coord), float(g[0]), float(g[1]), float(g[2]),
                                             1 if division == "float64" else 0, 0 if hash_type == "fnv" else 1,
                                             _lib.ptr(cell), _lib.ptr(keys), _lib.ptr(stats), _lib.stream()),
                   "grid_sample_keys")
        sorted_keys, order64 = torch.sort(keys, stable=True)          # library radix sort (64-bit keys)
        off = torch.tensor([n], dtype=torch.int32, device=dev)
        order32 = torch.empty(n, dtype=torch.int32, device=dev)
        cluster32 = torch.empty(n, dtype=torch.int32, device=dev)
        cluster64 = torch.empty(n, dtype=torch.int64, device=dev)
        idx_ptr_full = torch.empty(n + 1, dtype=torch.int32, device=dev)
        new_offset = torch.empty(1, dtype=torch.int64, device=dev)
        meta = torch.zeros(2, dtype=torch.int32, device=dev)
        ws = _lib.workspace(lib.aopt_voxel_partition_workspace_bytes(n), dev)
        _lib.check(lib.aopt_voxel_partition(n, 1, _lib.ptr(sorted_keys), _lib.ptr(order64), _lib.ptr(off),
                                            _lib.ptr(order32), _lib.ptr(cluster32), _lib.ptr(cluster64),
                                            _lib.ptr(idx_ptr_full), _lib.ptr(new_offset), _lib.ptr(meta),
                                            _lib.ptr(ws), ws.numel(), _lib.stream()), "voxel_partition")
        # one host read: number of voxels and the largest voxel (np.unique(..., return_counts) + count.max())
        cmax = torch.diff(idx_ptr_full).clamp_(min=0)
        n_vox = int(meta[0].item())
        count_max = int(cmax[:n_vox].max().item()) if n_vox else 0
    return VoxelHash(cell, order32, idx_ptr_full[: n_vox + 1], n_vox, count_max, stats, g)


def voxel_pick(vh: VoxelHash, r=None, part: int = 0) -> torch.Tensor:
    """idx_sort[cumsum(count)[:-1] + r % count]  (transform.py:813-817); r = None: the same `part` for every voxel
    (test mode, :841-843).  Returns int64 point indices, one per voxel, voxels in ascending key order."""
    lib = _lib.load()
    dev = vh.order.device
    pick = torch.empty(vh.n_vox, dtype=torch.int64, device=dev)
    if vh.n_vox:
        rt = None if r is None else _t(np.ascontiguousarray(r, dtype=np.int64) if isinstance(r, np.ndarray) else r, dev).long().contiguous()
        with _lib.on_device(dev):
            _lib.check(lib.aopt_voxel_pick(vh.n_vox, _lib.ptr(vh.idx_ptr), _lib.ptr(vh.order),
                                           _lib.ptr(rt) if rt is not None else None, int(part), _lib.ptr(pick),
                                           _lib.stream()), "voxel_pick")
    return pick


class GridSample:
    """transform.py:770-862.  mode "train": one random point per voxel; "test": list of parts covering every point."""

    def __init__(self, grid_size=0.05, hash_type="fnv", mode="train", keys=("coord", "color", "normal", "segment"),
                 return_discrete_coord=False, return_min_coord=False, return_displacement=False,
                 project_displacement=False, division="float64"):
        assert mode in ["train", "test"]
        self.grid_size, self.hash_type, self.mode, self.keys = grid_size, hash_type, mode, keys
        self.return_discrete_coord, self.return_min_coord = return_discrete_coord, return_min_coord
        self.return_displacement, self.project_displacement = return_displacement, project_displacement
        self.division = division

    def _displacement(self, data_dict, vh, coord):
        # scaled_coord - discrete_coord - 0.5 with discrete_coord already shifted by its minimum (:826-833)
        scaled = coord.double() / torch.as_tensor(vh.grid, dtype=torch.float64, device=coord.device)
        if self.division == "float32":
            scaled = (coord / torch.as_tensor(vh.grid, dtype=torch.float32, device=coord.device)).double()
        disp = scaled - vh.cell.double() - 0.5
        if self.project_displacement:
            disp = torch.sum(disp * _t(data_dict["normal"]).double(), dim=-1, keepdim=True)
        return disp

    def __call__(self, data_dict):
        assert "coord" in data_dict.keys()
        coord = _t(data_dict["coord"])
        data_dict["coord"] = coord
        vh = voxel_hash(coord, self.grid_size, self.hash_type, self.division)
        if self.mode == "train":
            # the reference's host draw, bit for bit (:815): np.random.randint(0, count.max(), count.size)
            r = np.random.randint(0, vh.count_max, vh.n_vox) if vh.n_vox else np.zeros(0, np.int64)
            idx_unique = voxel_pick(vh, r)
            if "sampled_index" in data_dict:                        # :818-825 (ScanNet data-efficient)
                sampled = _t(data_dict["sampled_index"]).long()
                idx_unique = torch.unique(torch.cat([idx_unique, sampled]))
                mask = torch.zeros(coord.shape[0], dtype=torch.bool, device=coord.device)
                mask[sampled] = True
                data_dict["sampled_index"] = torch.where(mask[idx_unique])[0]
            if self.return_discrete_coord:
                data_dict["discrete_coord"] = select_rows(vh.cell, idx_unique).long()
            if self.return_min_coord:
                data_dict["min_coord"] = vh.min_coord.reshape(1, 3)
            if self.return_displacement:
                data_dict["displacement"] = self._displacement(data_dict, vh, coord)[idx_unique]
            for key in self.keys:
                data_dict[key] = select_rows(_t(data_dict[key]), idx_unique)
            return data_dict
        data_part_list = []
        for i in range(vh.count_max):                                # :838-860
            idx_part = voxel_pick(vh, None, part=i)
            data_part = dict(index=idx_part)
            if self.return_discrete_coord:
                data_part["discrete_coord"] = select_rows(vh.cell, idx_part).long()
            if self.return_min_coord:
                data_part["min_coord"] = vh.min_coord.reshape(1, 3)
            if self.return_displacement:
                data_dict["displacement"] = self._displacement(data_dict, vh, coord)[idx_part]
            for key in data_dict.keys():
                if key in self.keys:
                    data_part[key] = select_rows(_t(data_dict[key]), idx_part)
                else:
                    data_part[key] = data_dict[key]
            data_part_list.append(data_part)
        return data_part_list


_CROP_KEYS = ("coord", "origin_coord", "discrete_coord", "color", "normal", "segment", "instance", "displacement",
              "strength")                                          # transform.py:976-993


def sphere_crop_index(coord: torch.Tensor, center, point_max: int) -> torch.Tensor:
    """np.argsort(np.sum(np.square(coord - center), 1))[:point_max] (transform.py:973-975), stable."""
    lib = _lib.load()
    coord = _t(coord).float().contiguous()
    n = coord.shape[0]
    c = [float(np.float32(v)) for v in (center.tolist() if hasattr(center, "tolist") else center)]
    d2 = torch.empty(n, dtype=torch.float32, device=coord.device)
    with _lib.on_device(coord.device):
        _lib.check(lib.aopt_sphere_dist2(n, _lib.ptr(coord), c[0], c[1], c[2], _lib.ptr(d2), _lib.stream()), "sphere_dist2")
    return torch.sort(d2, stable=True).indices[:point_max]


class SphereCrop:
    """transform.py:899-993, modes "random" and "center" (mode "all" is the tester's fragment generator and stays on
    the host side of the reference; not built here)."""

    def __init__(self, point_max=80000, sample_rate=None, mode="random"):
        self.point_max, self.sample_rate = point_max, sample_rate
        assert mode in ["random", "center", "all"]
        if mode == "all":
            raise NotImplementedError("ao_b200.datapipe.SphereCrop(mode='all'): test-time fragment generator, not built")
        self.mode = mode

    def __call__(self, data_dict):
        assert "coord" in data_dict.keys()
        coord = _t(data_dict["coord"])
        data_dict["coord"] = coord
        n = coord.shape[0]
        point_max = int(self.sample_rate * n) if self.sample_rate is not None else self.point_max
        if n > point_max:
            ci = np.random.randint(n) if self.mode == "random" else n // 2     # :969-972
            center = coord[ci].float().cpu().numpy()
            idx_crop = sphere_crop_index(coord, center, point_max)
            for key in _CROP_KEYS:
                if key in data_dict.keys():
                    data_dict[key] = select_rows(_t(data_dict[key]), idx_crop)
        return data_dict


class CenterShift:
    """transform.py:128-142."""

    def __init__(self, apply_z=True):
        self.apply_z = apply_z

    def __call__(self, data_dict):
        if "coord" in data_dict.keys():
            coord = _t(data_dict["coord"])
            lo, hi = coord.min(dim=0).values, coord.max(dim=0).values
            shift = torch.stack([(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, lo[2] if self.apply_z else lo.new_zeros(())])
            data_dict["coord"] = coord - shift
        return data_dict


class NormalizeColor:
    """transform.py:100-104."""

    def __call__(self, data_dict):
        if "color" in data_dict.keys():
            color = _t(data_dict["color"])
            # tensor / tensor: a true IEEE division like numpy's (torch turns `x / python_float` into x * (1/f) on CUDA)
            data_dict["color"] = color / torch.tensor(127.5, dtype=color.dtype if color.is_floating_point() else torch.float32,
                                                      device=color.device) - 1
        return data_dict


class ToTensor:
    """transform.py:69-97, with CUDA tensors as the result (arrays are already on the device after the transforms
    above; anything still on the host is uploaded)."""

    def __call__(self, data):
        if isinstance(data, torch.Tensor):
            return _t(data)
        if isinstance(data, str):
            return data
        if isinstance(data, int):
            return torch.tensor([data], dtype=torch.long, device=_dev())
        if isinstance(data, float):
            return torch.tensor([data], dtype=torch.float32, device=_dev())
        if isinstance(data, np.ndarray) and np.issubdtype(data.dtype, bool):
            return _t(data)
        if isinstance(data, np.ndarray) and np.issubdtype(data.dtype, np.integer):
            return _t(data).long()
        if isinstance(data, np.ndarray) and np.issubdtype(data.dtype, np.floating):
            return _t(data).float()
        if isinstance(data, Mapping):
            return {k: self(v) for k, v in data.items()}
        if isinstance(data, Sequence):
            return [self(v) for v in data]
        raise TypeError(f"type {type(data)} cannot be converted to tensor.")


class Collect:
    """transform.py:24-57: keep `keys`, add offset tensors, concatenate `*_keys` groups into feature tensors."""

    def __init__(self, keys, offset_keys_dict=None, **kwargs):
        self.keys = keys
        self.offset_keys = dict(offset="coord") if offset_keys_dict is None else offset_keys_dict
        self.kwargs = kwargs

    def __call__(self, data_dict):
        data = dict()
        keys = [self.keys] if isinstance(self.keys, str) else self.keys
        for key in keys:
            data[key] = data_dict[key]
        for key, value in self.offset_keys.items():
            data[key] = torch.tensor([data_dict[value].shape[0]], device=_t(data_dict[value]).device)
        for name, group in self.kwargs.items():
            name = name.replace("_keys", "")
            assert isinstance(group, Sequence)
            data[name] = torch.cat([_t(data_dict[k]).float() for k in group], dim=1)
        return data


_TRANSFORMS = {}


class Compose:
    """transform.py:1107-1117 over the transforms of this module (cfg = list of dict(type=..., **kwargs))."""

    def __init__(self, cfg=None):
        self.transforms = [_TRANSFORMS[c["type"]](**{k: v for k, v in c.items() if k != "type"}) for c in (cfg or [])]

    def __call__(self, data_dict):
        for t in self.transforms:
            data_dict = t(data_dict)
        return data_dict


_TRANSFORMS.update(GridSample=GridSample, SphereCrop=SphereCrop, CenterShift=CenterShift, NormalizeColor=NormalizeColor,
                   ToTensor=ToTensor, Collect=Collect)


def collate_fn(batch: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """pointcept/datasets/utils.py:14-37 for dict samples: tensors are concatenated along dim 0 and every key
    containing "offset" becomes the cumulative END index vector — the offset-encoded batch layout."""
    out = {}
    for key in batch[0]:
        vals = [b[key] for b in batch]
        if isinstance(vals[0], torch.Tensor):
            out[key] = torch.cat(vals, dim=0)
        else:
            out[key] = vals
    for key in out:
        if "offset" in key:
            out[key] = torch.cumsum(out[key], dim=0)
    return out
