"""interpolation / interpolation2 — API of /root/reference/libs/pointops/functions/interpolation.py:8-59.

Three-NN inverse-distance interpolation: kNN (k=3 by default, coarse → fine), weights, weighted
gather; backward is a CSR segmented sum instead of k index_put_(accumulate) / float atomics.
idx == -1 is NOT masked: like the reference (interpolation.py:21) it wraps to the last row.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import get_csr, prefetch_csr
from .query import knn_query_raw


def interpolation_weights(dist2: torch.Tensor) -> torch.Tensor:
    """(n,k) squared distances → normalised inverse-distance weights (interpolation.py:15-17)."""
    lib = _lib.load()
    n, k = dist2.shape
    w = torch.empty_like(dist2)
    if n > 0:
        with _lib.on_device(dist2.device):
            _lib.check(lib.aopt_interp_weights(n, k, _lib.ptr(dist2), _lib.ptr(w), _lib.stream()), "interp_weights")
    return w


class _InterpFn(Function):
    @staticmethod
    def forward(ctx, feat, idx, weight):
        lib = _lib.load()
        n, k = idx.shape
        m, c = feat.shape
        out = torch.empty((n, c), dtype=torch.float32, device=feat.device)
        if n > 0:
            with _lib.on_device(feat.device):
                _lib.check(
                    lib.aopt_interpolation_forward(n, c, k, m, _lib.ptr(feat), _lib.ptr(idx), _lib.ptr(weight),
                                                   _lib.ptr(out), _lib.stream()),
                    "interpolation_forward",
                )
        ctx.idx, ctx.weight, ctx.shape = idx, weight, (m, c, k)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        m, c, k = ctx.shape
        grad_out = grad_out.contiguous().float()
        csr = get_csr(ctx.idx, m, 1)  # negative indices wrap, as in the forward
        grad_in = torch.empty((m, c), dtype=torch.float32, device=grad_out.device)
        with _lib.on_device(grad_out.device):
            _lib.check(
                lib.aopt_interpolation_backward(m, c, k, _lib.ptr(grad_out), _lib.ptr(ctx.weight),
                                                _lib.ptr(csr.rowptr), _lib.ptr(csr.perm), _lib.ptr(grad_in),
                                                _lib.stream()),
                "interpolation_backward",
            )
        return grad_in, None, None


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """
    input: coords: (m, 3), new_xyz: (n, 3), feat: (m, c), offset: (b), new_offset: (b)
    output: (n, c)
    """
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    _lib.require_cuda(xyz, new_xyz, feat)
    idx, dist2 = knn_query_raw(k, xyz, offset, new_xyz, new_offset)
    weight = interpolation_weights(dist2)
    if feat.requires_grad and torch.is_grad_enabled():
        prefetch_csr(idx, feat.shape[0], 1)               # the backward's transposed graph, built under the forward
    return _InterpFn.apply(feat.float(), idx, weight)     # interpolation.py:19: fp32 accumulator


# interpolation2 has the same signature and math (the reference's CUDA-kernel variant).
interpolation2 = interpolation
