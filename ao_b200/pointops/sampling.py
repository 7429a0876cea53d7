"""farthest_point_sampling — same signature and return as
/root/reference/libs/pointops/functions/sampling.py:7-27 (caller: PTv1 TransitionDown,
/root/reference/pointcept/models/point_transformer/point_transformer_seg.py:101).

The kernel (csrc/fps.cu) gives one thread-block cluster to every scene and keeps the scene's points
and running distances in registers; results equal the reference kernel's, ties included."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib


class FarthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        """
        input: coords: (n, 3), offset: (b), new_offset: (b)
        output: idx: (m)
        """
        dev = _lib.require_cuda(xyz, offset, new_offset)
        assert xyz.is_contiguous()
        if xyz.dtype != torch.float32 or xyz.dim() != 2 or xyz.shape[1] != 3:
            raise ValueError("farthest_point_sampling: xyz must be float32 (n, 3)")
        if offset.numel() != new_offset.numel():
            raise ValueError("farthest_point_sampling: offset and new_offset must have the same number of scenes")
        n, b = xyz.shape[0], offset.numel()
        off, noff = offset.int().contiguous(), new_offset.int().contiguous()     # sampling.py:22
        # one host read for both: the largest scene (sampling.py:15-17, a python loop of .item() reads
        # there) and the output length new_offset[b-1] (sampling.py:18)
        if b == 0:
            return torch.zeros(0, dtype=torch.int32, device=dev)
        sizes = torch.diff(off, prepend=off.new_zeros(1))
        n_max, m = (int(v) for v in torch.stack([sizes.max(), noff[-1]]).tolist())
        idx = torch.zeros(m, dtype=torch.int32, device=dev)
        tmp = torch.full((n,), 1e10, dtype=torch.float32, device=dev)
        if m > 0:
            with _lib.on_device(dev):
                _lib.check(_lib.load().aopt_farthest_point_sampling(b, n_max, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff),
                                                                    _lib.ptr(tmp), _lib.ptr(idx), _lib.stream()),
                           "farthest_point_sampling")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None


farthest_point_sampling = FarthestPointSampling.apply
