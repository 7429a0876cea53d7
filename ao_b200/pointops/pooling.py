"""GridPool operators (new names): voxel partition + per-voxel segment reduction, and the "map"
unpooling gather.  Successors of the third-party op chain in
/root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:244-269
(offset2batch loop + segment_csr(min) + voxel_grid + torch.unique + torch.sort + two permuted
copies + segment_csr(mean) + segment_csr(max)) and :308-309 (`proj(feat)[cluster]`).

Kernels: aopt_voxel_grid (per-scene bounding boxes, compact voxel keys, own stable radix sort, partition: all on
the device, ONE host synchronisation — the voxel count), aopt_pool_forward/backward, aopt_grouping_forward/backward
(map unpool).
"""
from __future__ import annotations

from typing import NamedTuple

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import NeighbourCSR, get_csr

_VOXEL_PASSES = 3   # 11-bit radix passes enqueued on first sight: covers 33 key bits (pool.cu aopt_voxel_grid)
_PASSES_NEEDED = {}  # (grid size, scenes) -> passes the last partition needed


class VoxelPartition(NamedTuple):
    order: torch.Tensor      # (n,) int32 point ids sorted by voxel (stable: ascending id inside a voxel)
    idx_ptr: torch.Tensor    # (n_vox+1,) int32
    cluster: torch.Tensor    # (n,) int64 voxel id of every point (the reference's `cluster`)
    cluster32: torch.Tensor  # (n,) int32 copy used by the kernels
    offset: torch.Tensor     # (b,) int64 cumulative voxel counts per scene (batch2offset)
    n_vox: int


def voxel_partition(coord, offset, grid_size, start=None) -> VoxelPartition:
    """Voxel id of every point, voxels numbered in ascending (scene, z, y, x) order — the order
    torch.unique(sorted=True) gives the keys of torch_cluster.grid_cluster (…v2m2_base.py:257-264)."""
    dev = _lib.require_cuda(coord, offset)
    lib = _lib.load()
    assert coord.is_contiguous() and coord.dtype == torch.float32
    n, b = coord.shape[0], offset.numel()
    off32 = offset.int().contiguous()
    if n == 0:
        z32 = torch.zeros(0, dtype=torch.int32, device=dev)
        return VoxelPartition(z32, torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
                              z32, torch.zeros(b, dtype=torch.int64, device=dev), 0)
    order32 = torch.empty(n, dtype=torch.int32, device=dev)
    cluster32 = torch.empty(n, dtype=torch.int32, device=dev)
    cluster = torch.empty(n, dtype=torch.int64, device=dev)
    idx_ptr_full = torch.empty(n + 1, dtype=torch.int32, device=dev)
    new_offset = torch.empty(b, dtype=torch.int64, device=dev)
    meta = torch.empty(8, dtype=torch.int32, device=dev)            # [n_vox, flags, passes needed, key bits, shifts]
    start32 = None if start is None else start.float().contiguous()
    if start32 is not None and start32.numel() != 3 * b:
        raise ValueError("voxel_partition: start must be (b, 3)")

    def run(passes):
        with _lib.on_device(dev):
            ws = _lib.workspace(lib.aopt_voxel_grid_workspace_bytes(n, b), dev)
            # bounding boxes -> compact (scene, z, y, x) keys -> own stable radix sort -> boundaries / voxel ids /
            # idx_ptr / per-scene offsets: one entry point, no library sort, ONE host synchronisation (the voxel count)
            _lib.check(
                lib.aopt_voxel_grid(n, b, _lib.ptr(coord), _lib.ptr(off32), _lib.ptr(start32), float(grid_size), passes,
                                    _lib.ptr(order32), _lib.ptr(cluster32), _lib.ptr(cluster), _lib.ptr(idx_ptr_full),
                                    _lib.ptr(new_offset), _lib.ptr(meta), _lib.ptr(ws), ws.numel(), _lib.stream()),
                "voxel_grid",
            )
        return meta.tolist()                                         # the one host sync (torch.unique has one too)

    # Passes enqueued: what the previous partition with this (grid size, scene count) needed — the key width is known
    # only on the device, but it barely moves from one batch to the next, and an unneeded pass still costs its three
    # launches.  Too few is detected (flag 2) and repaired by a second call.
    guess = _PASSES_NEEDED.get((float(grid_size), b), _VOXEL_PASSES)
    m = run(guess)
    if m[1] & 2 and not m[1] & 1:
        m = run(min(6, max(m[2], guess + 1)))
    if not m[1] & 1:
        _PASSES_NEEDED[(float(grid_size), b)] = max(1, min(6, m[2]))
    n_vox, flags = m[0], m[1]
    if flags & 1:
        raise ValueError("voxel_partition: a point lies below its scene's `start`, or the voxel grid needs more than "
                         "64 key bits (extent / grid_size too large)")
    idx_ptr = idx_ptr_full[: n_vox + 1]
    # the partition IS the CSR of `cluster` (rows = voxels, entries ascending) → free backward map
    cluster32._aopt_csr = {(n_vox, 0): NeighbourCSR(idx_ptr, order32, n_vox, 0, cluster32._version)}
    return VoxelPartition(order32, idx_ptr, cluster, cluster32, new_offset, n_vox)


class _PoolFn(Function):
    @staticmethod
    def forward(ctx, feat, coord, order, idx_ptr, cluster32, n_vox):
        lib = _lib.load()
        n, c = feat.shape
        dev = feat.device
        out_feat = torch.empty((n_vox, c), dtype=torch.float32, device=dev)
        argmax = torch.empty((n_vox, c), dtype=torch.int32, device=dev)
        out_coord = torch.empty((n_vox, 3), dtype=torch.float32, device=dev) if coord is not None else None
        if n_vox > 0:
            with _lib.on_device(dev):
                _lib.check(
                    lib.aopt_pool_forward(n_vox, c, _lib.ptr(feat), _lib.ptr(coord), _lib.ptr(order),
                                          _lib.ptr(idx_ptr), _lib.ptr(out_feat), _lib.ptr(argmax),
                                          _lib.ptr(out_coord), _lib.stream()),
                    "pool_forward",
                )
        ctx.save_for_backward(argmax, cluster32)
        ctx.shape = (n, c)
        if out_coord is None:
            ctx.mark_non_differentiable(argmax)
        else:
            ctx.mark_non_differentiable(out_coord, argmax)
        return out_feat, out_coord, argmax

    @staticmethod
    def backward(ctx, grad_feat, _gc, _ga):
        lib = _lib.load()
        argmax, cluster32 = ctx.saved_tensors
        n, c = ctx.shape
        grad_feat = grad_feat.contiguous().float()
        grad_in = torch.empty((n, c), dtype=torch.float32, device=grad_feat.device)
        if n > 0:
            with _lib.on_device(grad_feat.device):
                _lib.check(
                    lib.aopt_pool_backward(n, c, _lib.ptr(grad_feat), _lib.ptr(argmax), _lib.ptr(cluster32),
                                           _lib.ptr(grad_in), _lib.stream()),
                    "pool_backward",
                )
        return grad_in, None, None, None, None, None


def pool_coord(coord, part: VoxelPartition):
    """Per-voxel mean of the coordinates only (the `segment_csr(coord, mean)` of …v2m2_base.py:265), same
    sequential sum in voxel order as the fused pool kernel.  Used by prepare_pyramid, which needs the coarse
    coordinates before any feature exists."""
    dev = _lib.require_cuda(coord)
    lib = _lib.load()
    coord = coord.float().contiguous()
    out_coord = torch.empty((part.n_vox, 3), dtype=torch.float32, device=dev)
    if part.n_vox > 0:
        scratch = torch.empty((part.n_vox, 3), dtype=torch.float32, device=dev)      # max / arg-max of the coords:
        scratch_i = torch.empty((part.n_vox, 3), dtype=torch.int32, device=dev)      # by-products, not used
        with _lib.on_device(dev):
            _lib.check(
                lib.aopt_pool_forward(part.n_vox, 3, _lib.ptr(coord), _lib.ptr(coord), _lib.ptr(part.order),
                                      _lib.ptr(part.idx_ptr), _lib.ptr(scratch), _lib.ptr(scratch_i),
                                      _lib.ptr(out_coord), _lib.stream()),
                "pool_coord",
            )
    return out_coord


def prepare_pyramid(coord, offset, grid_sizes, knn=None, interp_k=None):
    """Voxel partitions and coarse coordinates of every GridPool stage, computed up front.

    The partition is the one step of the path that needs a host synchronisation (the voxel count sizes every
    tensor of the next level; the reference's torch.unique has the same sync).  Inside the model it sits between
    the stages, so three times per step the host stops, the device queue drains, and the small kernels of the
    coarse levels then run at the speed of the Python launch loop.  Coordinates do not depend on features
    (…v2m2_base.py:246-268 uses `coord` and `offset` only), so the whole pyramid can be built first; the feature
    path then runs without a host sync and the host stays ahead of the device.

    Results are cached on the coordinate tensors: grid_pool(coord_l, feat, offset_l, grid_sizes[l]) picks them up
    (same values as computing them inline — it is the same code).  Returns [(coord_l, offset_l int64)], l = 0..L;
    the level-l tensors are the ones grid_pool will return for stage l-1.

    knn = k (or one k per coarse level): also start the self neighbour searches of levels 1..L on the geometry
    side stream; interp_k: and the coarse -> fine searches of the interpolation up path (query.prefetch_knn)."""
    from .query import knn_query_sets, prefetch_knn, stash_knn

    _lib.require_cuda(coord, offset)
    levels = [(coord, offset)]
    for li, gs in enumerate(grid_sizes):
        c, o = levels[-1]
        c32 = c.float().contiguous()
        part = voxel_partition(c32, o, gs)
        pooled = pool_coord(c32, part)
        cache = getattr(c, "_aopt_pyramid", None)
        if cache is None:
            cache = {}
            c._aopt_pyramid = cache
        # level 0 belongs to the caller: the entry is valid for this offset tensor only.  Coarser coordinates are
        # created here together with their offsets, so any cast of those offsets (same length) is accepted.
        # (the entry keeps the offset tensor alive, so its data_ptr cannot be recycled by another tensor)
        cache[(float(gs), int(c._version), int(o.numel()))] = (part, pooled, o if li == 0 else None)
        levels.append((pooled, part.offset))
    if knn is not None or interp_k is not None:
        offs32 = [levels[0][1].int()] + [o.int() for _, o in levels[1:]]
        n_lv = len(levels)
        ks = None if knn is None else [int(knn[l - 1] if isinstance(knn, (list, tuple)) else knn) for l in range(1, n_lv)]
        if _lib.overlap(role="knn"):
            # experiment: one search per level, on the geometry side stream
            for l in range(1, n_lv):
                if ks is not None:
                    prefetch_knn(ks[l - 1], levels[l][0], offs32[l])
            if interp_k is not None:
                for l in reversed(range(1, n_lv)):       # the decoder walks up from the coarsest level
                    prefetch_knn(int(interp_k), levels[l][0], offs32[l], levels[l - 1][0], offs32[l - 1])
        else:
            # The coarse levels' searches are independent of each other once the coordinates exist: ONE batched
            # search per distinct k for the self lists of levels 1..L, ONE for the coarse -> fine searches of the
            # interpolation up path (query.knn_query_sets).  The results wait on the coordinate tensors for the
            # knn_query / interpolation calls of the feature path.
            if ks is not None and n_lv > 2:
                for k in sorted(set(ks)):
                    ls = [l for l in range(1, n_lv) if ks[l - 1] == k and levels[l][0].shape[0] > 0]
                    if len(ls) < 2:
                        continue
                    res = knn_query_sets(k, [(levels[l][0], offs32[l], None, None) for l in ls], root=True)
                    for l, (idx, dist) in zip(ls, res):
                        stash_knn(k, levels[l][0], None, idx, dist, True)
            if interp_k is not None and n_lv > 2:
                ls = [l for l in range(1, n_lv) if levels[l][0].shape[0] > 0 and levels[l - 1][0].shape[0] > 0]
                if len(ls) >= 2:
                    res = knn_query_sets(int(interp_k), [(levels[l][0], offs32[l], levels[l - 1][0].float().contiguous()
                                                           if levels[l - 1][0].dtype != torch.float32 else levels[l - 1][0],
                                                           offs32[l - 1]) for l in ls], root=False)
                    for l, (idx, dist2) in zip(ls, res):
                        stash_knn(int(interp_k), levels[l][0], levels[l - 1][0], idx, dist2, False)
    return levels


def _prepared(coord, offset, grid_size, start):
    if start is not None:
        return None
    cache = getattr(coord, "_aopt_pyramid", None)
    if not cache:
        return None
    hit = cache.get((float(grid_size), int(coord._version), int(offset.numel())))
    if hit is None:
        return None
    if hit[2] is not None and hit[2] is not offset and (
            hit[2].data_ptr() != offset.data_ptr() or hit[2].dtype != offset.dtype):
        return None
    return hit


def grid_pool(coord, feat, offset, grid_size, start=None, return_partition=False):
    """GridPool after its fc/norm/act: returns ([coord', feat', offset'], cluster) like
    GridPool.forward (…v2m2_base.py:244-269); feat' = per-voxel max (gradient to the arg-max row),
    coord' = per-voxel mean, offset' int64, cluster (n,) int64.
    Uses the partition and coarse coordinates of prepare_pyramid when they were built for this coord tensor."""
    assert coord.is_contiguous() and feat.is_contiguous()
    hit = _prepared(coord, offset, grid_size, start)
    if hit is not None:
        part, pooled, _ = hit
        out_feat, _, _ = _PoolFn.apply(feat.float(), None, part.order, part.idx_ptr, part.cluster32, part.n_vox)
        out_coord = pooled
    else:
        part = voxel_partition(coord.float(), offset, grid_size, start)
        out_feat, out_coord, _ = _PoolFn.apply(feat.float(), coord.float(), part.order, part.idx_ptr, part.cluster32,
                                               part.n_vox)
    if return_partition:
        return [out_coord, out_feat, part.offset], part.cluster, part
    return [out_coord, out_feat, part.offset], part.cluster


class _UnpoolMapFn(Function):
    @staticmethod
    def forward(ctx, feat, cluster32):
        lib = _lib.load()
        n_vox, c = feat.shape
        n = cluster32.numel()
        out = torch.empty((n, c), dtype=torch.float32, device=feat.device)
        if n > 0:
            with _lib.on_device(feat.device):
                _lib.check(
                    lib.aopt_grouping_forward(n, 1, c, _lib.ptr(feat), _lib.ptr(cluster32), _lib.ptr(out), c,
                                              _lib.stream()),
                    "unpool_map_forward",
                )
        ctx.cluster32, ctx.shape = cluster32, (n_vox, c)
        return out

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        n_vox, c = ctx.shape
        grad = grad.contiguous().float()
        csr = get_csr(ctx.cluster32, n_vox, 0)
        grad_in = torch.empty((n_vox, c), dtype=torch.float32, device=grad.device)
        with _lib.on_device(grad.device):
            _lib.check(
                lib.aopt_grouping_backward(n_vox, c, _lib.ptr(grad), c, _lib.ptr(csr.rowptr), _lib.ptr(csr.perm),
                                           1.0, _lib.ptr(grad_in), _lib.stream()),
                "unpool_map_backward",
            )
        return grad_in, None


def unpool_map(feat, cluster):
    """`feat[cluster]` (…v2m2_base.py:308-309) with an atomic-free segmented-sum backward.
    `cluster` may be the int64 tensor returned by grid_pool or a VoxelPartition."""
    if isinstance(cluster, VoxelPartition):
        c32 = cluster.cluster32
    else:
        c32 = getattr(cluster, "_aopt_c32", None)
        if c32 is None:
            c32 = cluster.int().contiguous()
            try:
                cluster._aopt_c32 = c32
            except Exception:  # pragma: no cover
                pass
    _lib.require_cuda(feat, c32)
    return _UnpoolMapFn.apply(feat.float().contiguous(), c32)
