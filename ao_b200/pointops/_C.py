"""`pointops._C` — the hot-path functions of the reference's native module, same names and argument
lists (/root/reference/libs/pointops/src/pointops_api.cpp:15-32 and src/*/*_cuda.cpp), implemented
by forwarding `data_ptr()` to libao_pointops.so.  With this module registered as `pointops._C`
(ao_b200.install_as_pointops(native_only=True)), the reference's own
`libs/pointops/functions/*.py` runs unmodified on the B200 kernels — see INTEGRATION.md §2.

Semantics kept from the reference launchers:
  * every function returns None and writes into caller-allocated tensors;
  * forward/backward functions whose reference kernels ACCUMULATE (`+=` / atomicAdd into buffers the
    Python side zero-fills: interpolation_cuda_kernel.cu:16,31, grouping_cuda_kernel.cu:24,
    aggregation_cuda_kernel.cu:18,35-37, subtraction_cuda_kernel.cu:27-28) accumulate here too;
  * no dtype/contiguity checks in the reference (knn_query_cuda.cpp:9-14); here wrong dtypes or CPU
    tensors raise ValueError instead of reading garbage.
Differences (documented contract): kNN tie order is (dist2, idx) lexicographic rather
than heap-state dependent; idx == -1 selects a zero row in grouping / wraps in interpolation instead
of reading out of bounds; backward passes use a CSR segmented sum (deterministic) instead of atomics.
Functions outside the PTv2m2 path raise NotImplementedError.
"""
from __future__ import annotations

import torch

from .. import _lib
from ._csr import get_csr


def _chk(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"pointops._C: {name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise ValueError(f"pointops._C: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"pointops._C: {name} must be contiguous")
    return t


def _f(t, name):
    return _chk(t, torch.float32, name)


def _i(t, name):
    return _chk(t, torch.int32, name)


def knn_query_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """knn_query/knn_query_cuda.cpp:7-16.  Writes idx (m,nsample) int32 and dist2 (m,nsample) fp32."""
    lib = _lib.load()
    _f(xyz, "xyz"), _f(new_xyz, "new_xyz"), _i(offset, "offset"), _i(new_offset, "new_offset")
    _i(idx, "idx"), _f(dist2, "dist2")
    n, b = xyz.shape[0], offset.numel()
    # the library writes idx through its raw pointer, which does not bump the tensor's version counter: a transposed
    # graph cached on a re-used idx buffer (pointops._csr) would be stale
    if getattr(idx, "_aopt_csr", None):
        idx._aopt_csr.clear()
    if m == 0:
        return
    with _lib.on_device(xyz.device):
        ws = _lib.workspace(lib.aopt_knn_workspace_bytes(n, m, b, nsample, _lib.KNN_AUTO), xyz.device)
        _lib.check(lib.aopt_knn_query(m, nsample, n, b, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(offset),
                                      _lib.ptr(new_offset), _lib.ptr(idx), _lib.ptr(dist2), _lib.KNN_AUTO,
                                      _lib.ptr(ws), ws.numel(), _lib.stream()), "knn_query_cuda")


def grouping_forward_cuda(m, nsample, c, input, idx, output):
    """grouping/grouping_cuda.cpp:7-13: output[m,s,:] = input[idx[m,s],:]."""
    lib = _lib.load()
    _f(input, "input"), _i(idx, "idx"), _f(output, "output")
    with _lib.on_device(input.device):
        _lib.check(lib.aopt_grouping_forward(m, nsample, c, _lib.ptr(input), _lib.ptr(idx), _lib.ptr(output), c,
                                             _lib.stream()), "grouping_forward_cuda")


def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
    """grouping/grouping_cuda.cpp:15-21: grad_input[idx[m,s],:] += grad_output[m,s,:]."""
    lib = _lib.load()
    _f(grad_output, "grad_output"), _i(idx, "idx"), _f(grad_input, "grad_input")
    n = grad_input.shape[0]
    csr = get_csr(idx, n, 0)
    tmp = torch.empty_like(grad_input)
    with _lib.on_device(grad_output.device):
        _lib.check(lib.aopt_grouping_backward(n, c, _lib.ptr(grad_output), c, _lib.ptr(csr.rowptr), _lib.ptr(csr.perm),
                                              1.0, _lib.ptr(tmp), _lib.stream()), "grouping_backward_cuda")
    grad_input.add_(tmp)


def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
    """interpolation/interpolation_cuda.cpp:7-14: output[n,:] += Σ_i input[idx[n,i],:]·weight[n,i]."""
    lib = _lib.load()
    _f(input, "input"), _i(idx, "idx"), _f(weight, "weight"), _f(output, "output")
    tmp = torch.empty_like(output)
    with _lib.on_device(input.device):
        _lib.check(lib.aopt_interpolation_forward(n, c, k, input.shape[0], _lib.ptr(input), _lib.ptr(idx),
                                                  _lib.ptr(weight), _lib.ptr(tmp), _lib.stream()),
                   "interpolation_forward_cuda")
    output.add_(tmp)


def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
    """interpolation/interpolation_cuda.cpp:16-23: grad_input[idx[n,i],:] += grad_output[n,:]·weight[n,i]."""
    lib = _lib.load()
    _f(grad_output, "grad_output"), _i(idx, "idx"), _f(weight, "weight"), _f(grad_input, "grad_input")
    m = grad_input.shape[0]
    csr = get_csr(idx, m, 1)
    tmp = torch.empty_like(grad_input)
    with _lib.on_device(grad_output.device):
        _lib.check(lib.aopt_interpolation_backward(m, c, k, _lib.ptr(grad_output), _lib.ptr(weight),
                                                   _lib.ptr(csr.rowptr), _lib.ptr(csr.perm), _lib.ptr(tmp),
                                                   _lib.stream()), "interpolation_backward_cuda")
    grad_input.add_(tmp)


def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
    """subtraction/subtraction_cuda.cpp:7-14: output[n,s,:] = input1[n,:] − input2[idx[n,s],:]."""
    lib = _lib.load()
    _f(input1, "input1"), _f(input2, "input2"), _i(idx, "idx"), _f(output, "output")
    with _lib.on_device(input1.device):
        _lib.check(lib.aopt_subtraction_forward(n, nsample, c, _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(idx),
                                                _lib.ptr(output), _lib.stream()), "subtraction_forward_cuda")


def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
    """subtraction/subtraction_cuda.cpp:16-23: grad_input1[n,:] += Σ_s g[n,s,:]; grad_input2[idx[n,s],:] −= g[n,s,:]."""
    lib = _lib.load()
    _i(idx, "idx"), _f(grad_output, "grad_output"), _f(grad_input1, "grad_input1"), _f(grad_input2, "grad_input2")
    n2 = grad_input2.shape[0]
    csr = get_csr(idx, n2, 0)
    t1, t2 = torch.empty_like(grad_input1), torch.empty_like(grad_input2)
    with _lib.on_device(grad_output.device):
        _lib.check(lib.aopt_sum_over_k(n, nsample, c, _lib.ptr(grad_output), 1.0, _lib.ptr(t1), _lib.stream()),
                   "subtraction_backward_cuda")
        _lib.check(lib.aopt_grouping_backward(n2, c, _lib.ptr(grad_output), c, _lib.ptr(csr.rowptr), _lib.ptr(csr.perm),
                                              -1.0, _lib.ptr(t2), _lib.stream()), "subtraction_backward_cuda")
    grad_input1.add_(t1)
    grad_input2.add_(t2)


def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
    """aggregation/aggregation_cuda.cpp:7-15 (PTv1 share-planes layout): output[n,ch] += Σ_s (input[idx]+pos)·w[ch % w_c]."""
    lib = _lib.load()
    _f(input, "input"), _f(position, "position"), _f(weight, "weight"), _i(idx, "idx"), _f(output, "output")
    tmp = torch.empty_like(output)
    with _lib.on_device(input.device):
        _lib.check(lib.aopt_aggregation_forward(n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                                                _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(tmp), _lib.stream()),
                   "aggregation_forward_cuda")
    output.add_(tmp)


def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input,
                              grad_position, grad_weight):
    """aggregation/aggregation_cuda.cpp:17-28."""
    lib = _lib.load()
    _f(input, "input"), _f(position, "position"), _f(weight, "weight"), _i(idx, "idx")
    _f(grad_output, "grad_output"), _f(grad_input, "grad_input"), _f(grad_position, "grad_position")
    _f(grad_weight, "grad_weight")
    csr = get_csr(idx, input.shape[0], 0)
    gi, gp, gw = torch.empty_like(grad_input), torch.empty_like(grad_position), torch.empty_like(grad_weight)
    with _lib.on_device(input.device):
        _lib.check(lib.aopt_aggregation_backward(n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                                                 _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(csr.rowptr),
                                                 _lib.ptr(csr.perm), _lib.ptr(grad_output), _lib.ptr(gi), _lib.ptr(gp),
                                                 _lib.ptr(gw), _lib.stream()), "aggregation_backward_cuda")
    grad_input.add_(gi)
    grad_position.add_(gp)
    grad_weight.add_(gw)


def farthest_point_sampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    """sampling/sampling_cuda.cpp: b scenes, n = the largest scene, tmp (total points) pre-filled with 1e10,
    idx (new_offset[-1]) int32 output."""
    lib = _lib.load()
    _f(xyz, "xyz"), _i(offset, "offset"), _i(new_offset, "new_offset"), _f(tmp, "tmp"), _i(idx, "idx")
    if idx.numel() == 0:
        return
    with _lib.on_device(xyz.device):
        _lib.check(lib.aopt_farthest_point_sampling(int(b), int(n), _lib.ptr(xyz), _lib.ptr(offset), _lib.ptr(new_offset),
                                                    _lib.ptr(tmp), _lib.ptr(idx), _lib.stream()),
                   "farthest_point_sampling_cuda")


def _not_built(name):
    def f(*args, **kwargs):
        raise NotImplementedError(f"pointops._C.{name}: outside the PTv2m2 hot path (SURVEY.md §8), not built")
    f.__name__ = name
    return f


ball_query_cuda = _not_built("ball_query_cuda")
random_ball_query_cuda = _not_built("random_ball_query_cuda")
attention_relation_step_forward_cuda = _not_built("attention_relation_step_forward_cuda")
attention_relation_step_backward_cuda = _not_built("attention_relation_step_backward_cuda")
attention_fusion_step_forward_cuda = _not_built("attention_fusion_step_forward_cuda")
attention_fusion_step_backward_cuda = _not_built("attention_fusion_step_backward_cuda")
