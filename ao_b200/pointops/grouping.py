"""grouping / grouping2 — API of /root/reference/libs/pointops/functions/grouping.py:7-60.

`grouping` reproduces the pure-torch semantics PTv2m2 relies on (idx == -1 selects a zero row;
relative coordinates are masked by sign(idx+1); output fp32) with one gather kernel writing straight
into the (m, nsample, 3+c) layout, and an atomic-free CSR backward instead of index_put_(accumulate).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import get_csr


def _gather(inp, idx, out, out_stride, col0=0):
    lib = _lib.load()
    m, k = idx.shape
    c = inp.shape[1]
    with _lib.on_device(inp.device):
        _lib.check(
            lib.aopt_grouping_forward(m, k, c, _lib.ptr(inp), _lib.ptr(idx),
                                      out.data_ptr() + 4 * col0, out_stride, _lib.stream()),
            "grouping_forward",
        )


def _scatter(grad_out, go_stride, col0, csr, n, c, scale=1.0):
    lib = _lib.load()
    grad_in = torch.empty((n, c), dtype=torch.float32, device=grad_out.device)
    with _lib.on_device(grad_out.device):
        _lib.check(
            lib.aopt_grouping_backward(n, c, grad_out.data_ptr() + 4 * col0, go_stride, _lib.ptr(csr.rowptr),
                                       _lib.ptr(csr.perm), float(scale), _lib.ptr(grad_in), _lib.stream()),
            "grouping_backward",
        )
    return grad_in


class _GroupingFn(Function):
    """feat (n,c) → (m,k,c) or, with xyz, (m,k,3+c) = [masked relative xyz | feat]."""

    @staticmethod
    def forward(ctx, idx, feat, xyz, new_xyz, with_xyz):
        lib = _lib.load()
        m, k = idx.shape
        n, c = feat.shape
        width = c + 3 if with_xyz else c
        out = torch.empty((m, k, width), dtype=torch.float32, device=feat.device)
        if m > 0:
            _gather(feat, idx, out, width, 3 if with_xyz else 0)
            if with_xyz:
                with _lib.on_device(feat.device):
                    _lib.check(
                        lib.aopt_group_xyz(m, k, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(idx),
                                           _lib.ptr(out), width, _lib.stream()),
                        "group_xyz",
                    )
        ctx.idx = idx
        ctx.shape = (n, c, with_xyz)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        n, c, with_xyz = ctx.shape
        idx = ctx.idx
        grad_out = grad_out.contiguous().float()
        width = c + 3 if with_xyz else c
        csr = get_csr(idx, n, 0)
        grad_feat = _scatter(grad_out, width, 3 if with_xyz else 0, csr, n, c)
        # coordinates carry no gradient on this path (reference coords never require grad)
        return None, grad_feat, None, None, None


def _as_idx(idx):
    if idx.dtype != torch.int32:
        idx = idx.int()
    return idx.contiguous()


def grouping(idx, feat, xyz, new_xyz=None, with_xyz=False):
    if new_xyz is None:
        new_xyz = xyz
    assert xyz.is_contiguous() and feat.is_contiguous()
    _lib.require_cuda(idx, feat, xyz, new_xyz)
    if with_xyz:
        assert new_xyz.is_contiguous()
    idx = _as_idx(idx)
    if idx.dim() != 2:
        raise ValueError("grouping: idx must be (m, nsample)")
    # grouping.py:41-42 concatenates an fp32 zero row → the result is fp32 also under autocast
    return _GroupingFn.apply(idx, feat.float(), xyz.float(), new_xyz.float(), bool(with_xyz))


class Grouping(Function):
    """grouping2: plain gather (m,k,c) with scatter-add backward (grouping.py:7-33)."""

    @staticmethod
    def forward(ctx, input, idx):
        assert input.is_contiguous() and idx.is_contiguous()
        _lib.require_cuda(input, idx)
        idx = _as_idx(idx)
        m, k = idx.shape
        n, c = input.shape
        out = torch.empty((m, k, c), dtype=torch.float32, device=input.device)
        if m > 0:
            _gather(input.float(), idx, out, c)
        ctx.n, ctx.c = n, c
        ctx.idx = idx
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_output = grad_output.contiguous().float()
        csr = get_csr(ctx.idx, ctx.n, 0)
        return _scatter(grad_output, ctx.c, 0, csr, ctx.n, ctx.c), None


grouping2 = Grouping.apply
