"""Scene sharding for multi-GPU runs.

Every operator of the path is confined to one batch segment (kNN scans only
[offset[b-1], offset[b]): /root/reference/libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:70-76;
voxel keys carry the scene id: …/point_transformer_v2m2_base.py:257-259), so a batch in the
offset-encoded layout splits by scene with no data-path collective.  The reference does the same
through DistributedSampler + per-rank collate (pointcept/engines/train.py:225-252,
engines/defaults.py:139); its only collective is DDP's fp32 gradient all-reduce (defaults.py:38).
"""
from __future__ import annotations

from typing import Tuple

import torch


def scene_range(n_scenes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of scenes owned by `rank` (remainder spread over the first ranks)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(coord: torch.Tensor, feat: torch.Tensor, offset: torch.Tensor, rank: int, world: int):
    """Rank-local sub-batch of a global offset-encoded batch.
    Returns (coord_r, feat_r, offset_r, point_base): offset_r is rebased to start at 0 and
    point_base is the global index of the shard's first point (global idx = local idx + point_base)."""
    b = offset.numel()
    lo, hi = scene_range(b, rank, world)
    off = offset.long()
    p0 = int(off[lo - 1]) if lo > 0 else 0
    p1 = int(off[hi - 1]) if hi > lo else p0
    offset_r = (off[lo:hi] - p0).to(offset.dtype)
    return coord[p0:p1].contiguous(), feat[p0:p1].contiguous(), offset_r, p0


def ddp_wrap(model: torch.nn.Module, device_index=None, bucket_cap_mb: float = 4.0):
    """DistributedDataParallel as the reference builds it (create_ddp_model, engines/defaults.py:30-43, called with
    broadcast_buffers=False, find_unused_parameters=False at engines/train.py:209-213).

    The S3DIS-cfg gradients are 14.9 MB (ScanNet cfg 43.2 MB): over NVLink 5 the all-reduce is latency-bound, so what
    matters is WHEN it is issued.  One flat bucket would fire only after the last gradient (the patch-embed weights,
    at the very end of the backward pass) and serialise behind it; small buckets (default 4 MB) let the decoder /
    deep-encoder gradients — ready while the expensive level-0 backward kernels are still ahead — go out under the
    backward pass, leaving one small tail bucket.  gradient_as_bucket_view avoids the grad -> bucket copies.

    device_index None: CPU model (gloo) — the host-side plumbing test."""
    from torch.nn.parallel import DistributedDataParallel

    if device_index is None:
        return DistributedDataParallel(model, broadcast_buffers=False, find_unused_parameters=False,
                                       bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)
    return DistributedDataParallel(model, device_ids=[device_index], output_device=device_index,
                                   broadcast_buffers=False, find_unused_parameters=False,
                                   bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)


def max_over_ranks(value: float, device) -> float:
    """Max of a per-rank timing (multi-GPU numbers are reported as the slowest rank)."""
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
