"""subtraction — API of /root/reference/libs/pointops/functions/subtraction.py:7-38
(kernels subtraction_cuda_kernel.cu:5-30): out[n,s,:] = input1[n,:] - input2[idx[n,s],:]."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from ._csr import get_csr
from .grouping import _as_idx, _scatter


class Subtraction(Function):
    @staticmethod
    def forward(ctx, input1, input2, idx):
        """
        input: input1: (n, c), input2: (n, c), idx: (n, nsample)
        output:  (n, nsample, c)
        """
        assert input1.is_contiguous() and input2.is_contiguous()
        _lib.require_cuda(input1, input2, idx)
        lib = _lib.load()
        idx = _as_idx(idx)
        n, c = input1.shape
        nsample = idx.shape[-1]
        output = torch.empty((n, nsample, c), dtype=torch.float32, device=input1.device)
        if n > 0:
            with _lib.on_device(input1.device):
                _lib.check(
                    lib.aopt_subtraction_forward(n, nsample, c, _lib.ptr(input1.float()), _lib.ptr(input2.float()),
                                                 _lib.ptr(idx), _lib.ptr(output), _lib.stream()),
                    "subtraction_forward",
                )
        ctx.idx, ctx.n2 = idx, input2.shape[0]
        return output

    @staticmethod
    def backward(ctx, grad_output):
        """
        input: grad_out: (n, nsample, c)
        output: grad_input1: (n, c), grad_input2: (n, c)
        """
        lib = _lib.load()
        idx = ctx.idx
        n, nsample, c = grad_output.shape
        grad_output = grad_output.contiguous().float()
        grad_input1 = torch.empty((n, c), dtype=torch.float32, device=grad_output.device)
        with _lib.on_device(grad_output.device):
            _lib.check(
                lib.aopt_sum_over_k(n, nsample, c, _lib.ptr(grad_output), 1.0, _lib.ptr(grad_input1), _lib.stream()),
                "sum_over_k",
            )
        grad_input2 = _scatter(grad_output, c, 0, get_csr(idx, ctx.n2, 0), ctx.n2, c, scale=-1.0)
        return grad_input1, grad_input2, None


subtraction = Subtraction.apply
