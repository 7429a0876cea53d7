"""aggregation — PTv1 "share-planes" fused op, API of
/root/reference/libs/pointops/functions/aggregation.py:7-57 (kernels aggregation_cuda_kernel.cu:5-39).
Kept for API parity; PTv2m2 uses gva_aggregate (attention.py)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import get_csr
from .grouping import _as_idx


class Aggregation(Function):
    @staticmethod
    def forward(ctx, input, position, weight, idx):
        """
        input: input: (n, c), position: (n, nsample, c), weight : (n, nsample, c'), idx: (n, nsample)
        output: (n, c)
        """
        assert input.is_contiguous() and position.is_contiguous() and weight.is_contiguous()
        _lib.require_cuda(input, position, weight, idx)
        lib = _lib.load()
        idx = _as_idx(idx)
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        input, position, weight = input.float(), position.float(), weight.float()
        output = torch.empty((n, c), dtype=torch.float32, device=input.device)
        if n > 0:
            with _lib.on_device(input.device):
                _lib.check(
                    lib.aopt_aggregation_forward(n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                                                 _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(output), _lib.stream()),
                    "aggregation_forward",
                )
        ctx.save_for_backward(input, position, weight)
        ctx.idx = idx
        return output

    @staticmethod
    def backward(ctx, grad_output):
        """
        input: grad_out: (n, c)
        output: grad_input: (n, c), grad_position: (n, nsample, c), grad_weight : (n, nsample, c')
        """
        lib = _lib.load()
        input, position, weight = ctx.saved_tensors
        idx = ctx.idx
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        grad_output = grad_output.contiguous().float()
        dev = grad_output.device
        grad_input = torch.empty((input.shape[0], c), dtype=torch.float32, device=dev)
        grad_position = torch.empty((n, nsample, c), dtype=torch.float32, device=dev)
        grad_weight = torch.empty((n, nsample, w_c), dtype=torch.float32, device=dev)
        csr = get_csr(idx, input.shape[0], 0)
        if n > 0:
            with _lib.on_device(dev):
                _lib.check(
                    lib.aopt_aggregation_backward(n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                                                  _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(csr.rowptr),
                                                  _lib.ptr(csr.perm), _lib.ptr(grad_output), _lib.ptr(grad_input),
                                                  _lib.ptr(grad_position), _lib.ptr(grad_weight), _lib.stream()),
                    "aggregation_backward",
                )
        return grad_input, grad_position, grad_weight, None


aggregation = Aggregation.apply
