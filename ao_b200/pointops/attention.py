"""PTv2 GroupedVectorAttention operators (new names; the fused successors of the op chains in
/root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:109-128).

  group_xyz(idx, xyz, new_xyz)            pos = (xyz[idx] - new_xyz[:,None]) * sign(idx+1)      (:109,:111)
  gva_relation(key, query, idx)           key[idx] - query[:,None]                              (:109,:112)
  gva_aggregate(value, peb, logits, idx, groups)
                                          einsum((value[idx]+peb) , softmax_k(logits)*mask)      (:110,:119-128)

attention_relation_step / attention_fusion_step of the reference API have no caller in pointcept
 and are not built.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import get_csr
from .grouping import _as_idx, _scatter


import contextlib
import os

_SPLIT_RELATION_BWD = os.environ.get("AOPT_RELBWD", "fused") == "split"   # A/B switch (measurements only)


def group_xyz(idx, xyz, new_xyz=None):
    """(m,k,3) masked relative coordinates.  No gradient (coordinates are inputs)."""
    if new_xyz is None:
        new_xyz = xyz
    _lib.require_cuda(idx, xyz, new_xyz)
    idx = _as_idx(idx)
    lib = _lib.load()
    m, k = idx.shape
    out = torch.empty((m, k, 3), dtype=torch.float32, device=xyz.device)
    if m > 0:
        with _lib.on_device(xyz.device):
            _lib.check(
                lib.aopt_group_xyz(m, k, _lib.ptr(xyz.float().contiguous()), _lib.ptr(new_xyz.float().contiguous()),
                                   _lib.ptr(idx), _lib.ptr(out), 3, _lib.stream()),
                "group_xyz",
            )
    return out


class _RelationFn(Function):
    @staticmethod
    def forward(ctx, key, query, idx):
        lib = _lib.load()
        m, k = idx.shape
        n, c = key.shape
        out = torch.empty((m, k, c), dtype=torch.float32, device=key.device)
        if m > 0:
            with _lib.on_device(key.device):
                _lib.check(
                    lib.aopt_gather_sub_forward(m, k, c, _lib.ptr(key), _lib.ptr(query), _lib.ptr(idx),
                                                _lib.ptr(out), _lib.stream()),
                    "gather_sub_forward",
                )
        ctx.idx, ctx.shape = idx, (n, m, k, c)
        return out

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        n, m, k, c = ctx.shape
        grad = grad.contiguous().float()
        grad_key = grad_query = None
        if ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and m == n and m > 0 and not _SPLIT_RELATION_BWD:
            # queries == sources: one pass over grad for both gradients (csrc/grouping.cu relation_backward_kernel)
            csr = get_csr(ctx.idx, n, 0)
            grad_key = torch.empty((n, c), dtype=torch.float32, device=grad.device)
            grad_query = torch.empty((m, c), dtype=torch.float32, device=grad.device)
            with _lib.on_device(grad.device):
                _lib.check(
                    lib.aopt_relation_backward(n, k, c, _lib.ptr(grad), _lib.ptr(csr.rowptr), _lib.ptr(csr.perm),
                                               _lib.ptr(grad_key), _lib.ptr(grad_query), _lib.stream()),
                    "relation_backward",
                )
            return grad_key, grad_query, None
        if ctx.needs_input_grad[0]:
            grad_key = _scatter(grad, c, 0, get_csr(ctx.idx, n, 0), n, c)
        if ctx.needs_input_grad[1]:
            grad_query = torch.empty((m, c), dtype=torch.float32, device=grad.device)
            with _lib.on_device(grad.device):
                _lib.check(
                    lib.aopt_sum_over_k(m, k, c, _lib.ptr(grad), -1.0, _lib.ptr(grad_query), _lib.stream()),
                    "sum_over_k",
                )
        return grad_key, grad_query, None


def gva_relation(key, query, idx):
    """relation_qk before the positional bias: (m,k,c) = key[idx] - query[:,None]."""
    _lib.require_cuda(key, query, idx)
    idx = _as_idx(idx)
    if key.shape[1] != query.shape[1] or query.shape[0] != idx.shape[0]:
        raise ValueError("gva_relation: shape mismatch")
    return _RelationFn.apply(key.float().contiguous(), query.float().contiguous(), idx)


class _AggregateFn(Function):
    @staticmethod
    def forward(ctx, value, peb, logits, idx, groups):
        lib = _lib.load()
        n, k = idx.shape
        n_src, c = value.shape
        out = torch.empty((n, c), dtype=torch.float32, device=value.device)
        prob = torch.empty((n, k, groups), dtype=torch.float32, device=value.device)
        if n > 0:
            with _lib.on_device(value.device):
                _lib.check(
                    lib.aopt_gva_forward(n, k, c, groups, _lib.ptr(value), _lib.ptr(peb), _lib.ptr(logits),
                                         _lib.ptr(idx), _lib.ptr(out), _lib.ptr(prob), _lib.stream()),
                    "gva_forward",
                )
        ctx.save_for_backward(value, peb, prob)
        ctx.idx, ctx.groups, ctx.has_peb = idx, groups, peb is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        value, peb, prob = ctx.saved_tensors
        idx, g = ctx.idx, ctx.groups
        n, k = idx.shape
        n_src, c = value.shape
        grad_out = grad_out.contiguous().float()
        dev = grad_out.device
        need_peb = ctx.has_peb and ctx.needs_input_grad[1]
        grad_peb = torch.empty((n, k, c), dtype=torch.float32, device=dev) if need_peb else None
        grad_logits = torch.empty((n, k, g), dtype=torch.float32, device=dev)
        grad_value = None
        want_value = ctx.needs_input_grad[0]
        with _lib.on_device(dev):
            side = _lib.side_stream(dev, "walk") if (want_value and n > 0 and _lib.overlap(role="walk")) else None
            if want_value:
                csr = get_csr(idx, n_src, 0)
                grad_value = torch.empty((n_src, c), dtype=torch.float32, device=dev)
            if want_value and n_src == n and n > 0 and side is None:
                # self-attention (queries == sources): ONE kernel — the latency-bound CSR walk of grad_value runs in
                # the shadow of the HBM-bound per-query pass (csrc/gva.cu gva_backward_fused_ns_kernel)
                _lib.check(
                    lib.aopt_gva_backward(n, k, c, g, _lib.ptr(grad_out), _lib.ptr(value), _lib.ptr(peb), _lib.ptr(prob),
                                          _lib.ptr(idx), _lib.ptr(csr.rowptr), _lib.ptr(csr.perm), _lib.ptr(grad_peb),
                                          _lib.ptr(grad_logits), _lib.ptr(grad_value), _lib.stream()),
                    "gva_backward",
                )
                return grad_value, grad_peb, grad_logits, None, None
            # cross attention, or the side-stream experiment: grad_value (CSR walk) and grad_peb / grad_logits
            # (streaming) as two kernels.  Everything is allocated on the caller's stream; the side stream starts
            # after an event that follows the allocations and inputs, and the caller's stream joins it before returning.
            if side is not None:
                main = torch.cuda.current_stream(dev)
                side.wait_stream(main)
            if n > 0:
                _lib.check(
                    lib.aopt_gva_backward_query(n, k, c, g, _lib.ptr(grad_out), _lib.ptr(value), _lib.ptr(peb),
                                                _lib.ptr(prob), _lib.ptr(idx), _lib.ptr(grad_peb),
                                                _lib.ptr(grad_logits), _lib.stream()),
                    "gva_backward_query",
                )
            if want_value:
                with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                    _lib.check(
                        lib.aopt_gva_backward_value(n_src, k, c, g, _lib.ptr(grad_out), _lib.ptr(prob),
                                                    _lib.ptr(csr.rowptr), _lib.ptr(csr.perm), _lib.ptr(grad_value),
                                                    _lib.stream()),
                        "gva_backward_value",
                    )
                if side is not None:
                    main.wait_stream(side)
        return grad_value, grad_peb, grad_logits, None, None


def gva_aggregate(value, peb, logits, idx, groups):
    """value (n_src,c) un-gathered, peb (n,k,c) or None, logits (n,k,groups), idx (n,k) → (n,c)."""
    _lib.require_cuda(value, peb, logits, idx)
    idx = _as_idx(idx)
    n, k = idx.shape
    c = value.shape[1]
    if c % groups != 0:
        raise ValueError("gva_aggregate: channels must be divisible by groups")
    if logits.shape != (n, k, groups) or (peb is not None and peb.shape != (n, k, c)):
        raise ValueError("gva_aggregate: shape mismatch")
    return _AggregateFn.apply(value.float().contiguous(), None if peb is None else peb.float().contiguous(),
                              logits.float().contiguous(), idx, int(groups))


def attention_relation_step(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.attention_relation_step: no caller on the PTv2m2 path (not built)")


def attention_fusion_step(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.attention_fusion_step: no caller on the PTv2m2 path (not built)")
