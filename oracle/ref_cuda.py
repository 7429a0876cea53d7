"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes access to oracle/_ref/libpointops_ref.so: the UNMODIFIED reference CUDA launchers
(/root/reference/libs/pointops/src/*/ *_cuda_kernel.cu), compiled by oracle/Makefile.ref from the
sources where they lie.  Used by the `-m gpu` tests to pin the C oracle and the new kernels against
the reference itself on the B200.  The launchers take raw device pointers and launch on the legacy
default stream (knn_query_cuda_kernel.cu:111), which is torch's default stream.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libpointops_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
        for name in ("knn_query_cuda_launcher", "grouping_forward_cuda_launcher", "grouping_backward_cuda_launcher",
                     "interpolation_forward_cuda_launcher", "interpolation_backward_cuda_launcher",
                     "aggregation_forward_cuda_launcher", "aggregation_backward_cuda_launcher",
                     "subtraction_forward_cuda_launcher", "subtraction_backward_cuda_launcher",
                     "farthest_point_sampling_cuda_launcher"):
            getattr(_lib, name).restype = None
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None):
    """Mirrors KNNQuery.forward (libs/pointops/functions/query.py:9-24) but returns dist2 (no sqrt)."""
    if new_xyz is None:
        new_xyz, new_offset = xyz, offset
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    off, noff = offset.int().contiguous(), new_offset.int().contiguous()
    torch.cuda.synchronize()
    lib().knn_query_cuda_launcher(ctypes.c_int(m), ctypes.c_int(nsample), _p(xyz), _p(new_xyz), _p(off), _p(noff),
                                  _p(idx), _p(dist2))
    torch.cuda.synchronize()
    return idx, dist2


def grouping_forward(inp, idx):
    m, k = idx.shape
    c = inp.shape[1]
    out = torch.zeros((m, k, c), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    lib().grouping_forward_cuda_launcher(ctypes.c_int(m), ctypes.c_int(k), ctypes.c_int(c), _p(inp), _p(idx), _p(out))
    torch.cuda.synchronize()
    return out


def grouping_backward(grad_out, idx, n):
    m, k, c = grad_out.shape
    gin = torch.zeros((n, c), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    lib().grouping_backward_cuda_launcher(ctypes.c_int(m), ctypes.c_int(k), ctypes.c_int(c), _p(grad_out), _p(idx), _p(gin))
    torch.cuda.synchronize()
    return gin


def interpolation_forward(inp, idx, weight):
    n, k = idx.shape
    c = inp.shape[1]
    out = torch.zeros((n, c), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    lib().interpolation_forward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(c), ctypes.c_int(k), _p(inp), _p(idx),
                                              _p(weight), _p(out))
    torch.cuda.synchronize()
    return out


def interpolation_backward(grad_out, idx, weight, m):
    n, c = grad_out.shape
    k = idx.shape[1]
    gin = torch.zeros((m, c), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    lib().interpolation_backward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(c), ctypes.c_int(k), _p(grad_out), _p(idx),
                                               _p(weight), _p(gin))
    torch.cuda.synchronize()
    return gin


def aggregation_forward(inp, position, weight, idx):
    n, k, c = position.shape
    w_c = weight.shape[-1]
    out = torch.zeros((n, c), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    lib().aggregation_forward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(c), ctypes.c_int(w_c),
                                            _p(inp), _p(position), _p(weight), _p(idx), _p(out))
    torch.cuda.synchronize()
    return out


def aggregation_backward(inp, position, weight, idx, grad_out):
    n, k, c = position.shape
    w_c = weight.shape[-1]
    gi = torch.zeros_like(inp)
    gp = torch.zeros_like(position)
    gw = torch.zeros_like(weight)
    torch.cuda.synchronize()
    lib().aggregation_backward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(c), ctypes.c_int(w_c),
                                             _p(inp), _p(position), _p(weight), _p(idx), _p(grad_out), _p(gi), _p(gp), _p(gw))
    torch.cuda.synchronize()
    return gi, gp, gw


def subtraction_forward(input1, input2, idx):
    n, c = input1.shape
    k = idx.shape[1]
    out = torch.zeros((n, k, c), dtype=torch.float32, device=input1.device)
    torch.cuda.synchronize()
    lib().subtraction_forward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(c), _p(input1), _p(input2),
                                            _p(idx), _p(out))
    torch.cuda.synchronize()
    return out


def subtraction_backward(idx, grad_out, n2):
    n, k, c = grad_out.shape
    g1 = torch.zeros((n, c), dtype=torch.float32, device=grad_out.device)
    g2 = torch.zeros((n2, c), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    lib().subtraction_backward_cuda_launcher(ctypes.c_int(n), ctypes.c_int(k), ctypes.c_int(c), _p(idx), _p(grad_out),
                                             _p(g1), _p(g2))
    torch.cuda.synchronize()
    return g1, g2


def farthest_point_sampling(xyz, offset, new_offset):
    """FarthestPointSampling.forward (libs/pointops/functions/sampling.py:9-24) through the reference launcher."""
    off, noff = offset.int().contiguous(), new_offset.int().contiguous()
    b = off.numel()
    sizes = torch.diff(off, prepend=off.new_zeros(1))
    n_max = int(sizes.max().item())
    idx = torch.zeros(int(noff[-1].item()), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    lib().farthest_point_sampling_cuda_launcher(ctypes.c_int(b), ctypes.c_int(n_max), _p(xyz), _p(off), _p(noff),
                                                _p(tmp), _p(idx))
    torch.cuda.synchronize()
    return idx, tmp
