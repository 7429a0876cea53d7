"""Layout helpers and query+group conveniences — API of
/root/reference/libs/pointops/functions/utils.py:5-119 and pointcept/models/utils.py:11-28."""
from __future__ import annotations

import torch

from .. import _lib
from .grouping import grouping
from .query import knn_query


def offset2batch(offset):
    """(b,) cumulative end indices → (n,) int64 scene id per point.  One kernel, no host loop
    (the reference builds it with a Python loop over a CUDA tensor, utils.py:102-115)."""
    dev = _lib.require_cuda(offset)
    lib = _lib.load()
    off32 = offset.int().contiguous()
    b = off32.numel()
    n = int(off32[-1].item()) if b > 0 else 0
    batch = torch.empty(n, dtype=torch.int64, device=dev)
    if n > 0:
        with _lib.on_device(dev):
            _lib.check(lib.aopt_offset2batch(n, b, _lib.ptr(off32), _lib.ptr(batch), _lib.stream()), "offset2batch")
    return batch


def batch2offset(batch):
    return torch.cumsum(batch.bincount(), dim=0).int()      # utils.py:118-119


def knn_query_and_group(feat, xyz, offset=None, new_xyz=None, new_offset=None, idx=None, nsample=None,
                        with_xyz=False):
    if idx is None:
        assert nsample is not None
        idx, _ = knn_query(nsample, xyz, offset, new_xyz, new_offset)
    return grouping(idx, feat, xyz, new_xyz, with_xyz), idx


def ball_query_and_group(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.ball_query_and_group: outside the PTv2m2 hot path (not built)")


def query_and_group(nsample, xyz, new_xyz, feat, idx, offset, new_offset, dilation=0, with_feat=True,
                    with_xyz=True):
    """
    input: coords: (n, 3), new_xyz: (m, 3), feat: (n, c), idx: (m, nsample), offset: (b), new_offset: (b)
    output: new_feat: (m, nsample, c+3), grouped_idx: (m, nsample)
    """
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        num_samples_total = 1 + (nsample - 1) * (dilation + 1)
        idx_no_dilation, _ = knn_query(num_samples_total, xyz, offset, new_xyz, new_offset)
        idx = []
        batch_end = offset.tolist()
        batch_start = [0] + batch_end[:-1]
        new_batch_end = new_offset.tolist()
        new_batch_start = [0] + new_batch_end[:-1]
        for i in range(offset.shape[0]):
            if batch_end[i] - batch_start[i] < num_samples_total:
                soft_dilation = (batch_end[i] - batch_start[i] - 1) / (nsample - 1) - 1
            else:
                soft_dilation = dilation
            cols = [int((soft_dilation + 1) * j) for j in range(nsample)]
            idx.append(idx_no_dilation[new_batch_start[i]: new_batch_end[i], cols])
        idx = torch.cat(idx, dim=0).contiguous()
    if not with_feat:
        return idx
    # the reference gathers without -1 handling here; the kernel's zero-row rule is a superset
    grouped = grouping(idx, feat, xyz, new_xyz, with_xyz=with_xyz)
    return grouped, idx
