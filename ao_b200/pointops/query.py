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


def knn_query_raw(nsample, xyz, offset, new_xyz=None, new_offset=None, method=None):
    """Returns (idx int32 (m,k), dist2 float32 (m,k)) — squared distances, the kernel's own output."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
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


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, offset, new_xyz=None, new_offset=None):
        """
        input: xyz: (n, 3), new_xyz: (m, 3), offset: (b), new_offset: (b)
        output: idx: (m, nsample) -1 is placeholder, dist: (m, nsample)
        """
        idx, dist2 = knn_query_raw(nsample, xyz, offset, new_xyz, new_offset)
        ctx.mark_non_differentiable(idx)
        dist = torch.sqrt(dist2)                             # query.py:24
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
