"""knn_query — same signature and return as /root/reference/libs/pointops/functions/query.py:7-24,111.

ball_query / random_ball_query are outside the PTv2m2 hot path (no caller) and
raise NotImplementedError rather than silently falling back.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from .. import _lib

_METHODS = {"auto": _lib.KNN_AUTO, "tile": _lib.KNN_TILE, "grid": _lib.KNN_GRID}


def knn_query_raw(nsample, xyz, offset, new_xyz=None, new_offset=None, method=None, root=False):
    """Returns (idx int32 (m,k), dist2 float32 (m,k)) — squared distances, the kernel's own output.
    root=True: the kernel writes sqrt(dist2) instead (AOPT_KNN_SQRT_DIST), the reference's return (query.py:24)."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    if method is None:
        hit = _prefetched(int(nsample), xyz, new_xyz)
        if hit is not None:
            return (hit[0], torch.sqrt(hit[1])) if root else hit
    dev = _lib.require_cuda(xyz, new_xyz, offset, new_offset)
    assert xyz.is_contiguous() and new_xyz.is_contiguous()
    if xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise ValueError("knn_query: coordinates must be float32")
    if xyz.dim() != 2 or xyz.shape[1] != 3 or new_xyz.dim() != 2 or new_xyz.shape[1] != 3:
        raise ValueError("knn_query: coordinates must be (n, 3)")
    nsample = int(nsample)
    if not 1 <= nsample <= 128:
        raise ValueError("knn_query: nsample must be in [1, 128] (reference kernel limit)")
    same = new_offset is offset
    offset = offset.int().contiguous()                       # query.py:22
    new_offset = offset if same else new_offset.int().contiguous()
    if offset.numel() != new_offset.numel():
        raise ValueError("knn_query: offset and new_offset must have the same number of scenes")
    n, m, b = xyz.shape[0], new_xyz.shape[0], offset.numel()
    if method is None:
        method = os.environ.get("AOPT_KNN_METHOD", "auto")
    meth = _METHODS[method] if isinstance(method, str) else int(method)
    if root:
        meth |= _lib.KNN_SQRT_DIST
    lib = _lib.load()
    idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
    dist2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
    if m == 0:
        return idx, dist2
    with _lib.on_device(dev):
        ws = _lib.workspace(lib.aopt_knn_workspace_bytes(n, m, b, nsample, meth), dev)
        _lib.check(
            lib.aopt_knn_query(m, nsample, n, b, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(offset),
                               _lib.ptr(new_offset), _lib.ptr(idx), _lib.ptr(dist2), meth, _lib.ptr(ws),
                               ws.numel(), _lib.stream()),
            "knn_query",
        )
    return idx, dist2


# ---- searches issued ahead of time on the geometry side stream (pointops.prepare_pyramid(..., knn=...)) ----------
def _knn_key(nsample, xyz, new_xyz):
    if new_xyz is xyz:
        return ("self", nsample, xyz.shape[0], xyz._version)
    return ("cross", nsample, new_xyz.data_ptr(), new_xyz.shape[0], new_xyz._version, xyz._version)


def prefetch_knn(nsample, xyz, offset, new_xyz=None, new_offset=None):
    """Starts the search NOW on the geometry side stream; the next knn_query / knn_query_raw / interpolation call
    with the same tensors returns its result (after making the caller's stream wait for it).  The neighbour
    search is ALU-bound and moves almost no memory, the streaming kernels of the feature path leave the ALUs idle:
    run side by side they overlap (unlike the CSR walk / CSR build, see _lib.py).  No-op with overlap off."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    if not _lib.overlap(role="knn") or new_xyz.shape[0] == 0:
        return
    dev = _lib.require_cuda(xyz, new_xyz, offset, new_offset)
    main = torch.cuda.current_stream(dev)
    side = _lib.side_stream(dev, "knn")
    side.wait_stream(main)                                   # coordinates / offsets are produced on the caller's stream
    with torch.cuda.stream(side):
        idx, dist2 = knn_query_raw(nsample, xyz, offset, new_xyz, new_offset, method="auto")
        ev = torch.cuda.Event()
        ev.record(side)
    for t in (xyz, new_xyz, offset, new_offset):
        t.record_stream(side)                                # read by the side stream
    cache = getattr(xyz, "_aopt_knn", None)
    if cache is None:
        cache = {}
        xyz._aopt_knn = cache
    cache[_knn_key(int(nsample), xyz, new_xyz)] = (idx, dist2, ev)


def _prefetched(nsample, xyz, new_xyz):
    cache = getattr(xyz, "_aopt_knn", None)
    if not cache:
        return None
    hit = cache.pop(_knn_key(nsample, xyz, new_xyz), None)   # one consumer: the result is handed over, not kept
    if hit is None:
        return None
    idx, dist2, ev = hit
    cur = torch.cuda.current_stream(idx.device)
    cur.wait_event(ev)
    idx.record_stream(cur)                                   # allocated in the side stream's pool
    dist2.record_stream(cur)
    return idx, dist2


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, offset, new_xyz=None, new_offset=None):
        """
        input: xyz: (n, 3), new_xyz: (m, 3), offset: (b), new_offset: (b)
        output: idx: (m, nsample) -1 is placeholder, dist: (m, nsample)
        """
        idx, dist = knn_query_raw(nsample, xyz, offset, new_xyz, new_offset, root=True)   # sqrt of query.py:24 in-kernel
        ctx.mark_non_differentiable(idx)
        ctx.mark_non_differentiable(dist)
        return idx, dist

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None


knn_query = KNNQuery.apply


def ball_query(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.ball_query: outside the PTv2m2 hot path (not built)")


def random_ball_query(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.random_ball_query: outside the PTv2m2 hot path (not built)")
