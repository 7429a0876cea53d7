"""Fused positional-bias MLP of GroupedVectorAttention (new operator).

    peb = linear_p_bias(pos) = Linear(C,C)(ReLU(PointBatchNorm(C)(Linear(3,C)(pos))))
    /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:88-93,116-118

`pos` is the (N,k,3) tensor of masked relative neighbour coordinates (pointops.group_xyz).  The hidden
(N,k,C) tensors are never materialised: training-mode BatchNorm statistics follow in closed form from
the mean / covariance of pos (`pos_moments`, computed once per neighbour list), the first layer is
recomputed on the fly, and the C x C layer runs on the tensor cores in bf16 with fp32 accumulation —
the precision of the autocast path it replaces.  Kernels: ao_b200/csrc/pe_mlp.cu.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib


def pe_mlp_supported(channels: int) -> bool:
    return bool(_lib.load().aopt_pe_mlp_supported(int(channels)))


def pos_moments(pos: torch.Tensor) -> torch.Tensor:
    """(9,) float64: Σp (3) and Σ xx, xy, xz, yy, yz, zz over all rows of pos (..., 3)."""
    dev = _lib.require_cuda(pos)
    lib = _lib.load()
    pos = pos.float().contiguous()
    rows = pos.numel() // 3
    out = torch.empty(9, dtype=torch.float64, device=dev)
    with _lib.on_device(dev):
        ws = _lib.workspace(lib.aopt_pos_moments_workspace_bytes(), dev)
        _lib.check(lib.aopt_pos_moments(rows, _lib.ptr(pos), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream()),
                   "pos_moments")
    return out


class _PeMlpFn(Function):
    @staticmethod
    def forward(ctx, pos, moments, w1, b1, gamma, beta, w2, b2, running_mean, running_var, eps, use_batch, aux_w):
        lib = _lib.load()
        dev = pos.device
        c = w2.shape[0]
        rows = pos.numel() // 3
        ga = 0 if aux_w is None else aux_w.shape[0]
        out = torch.empty(pos.shape[:-1] + (c,), dtype=torch.float32, device=dev)
        aux_out = None if aux_w is None else torch.empty(pos.shape[:-1] + (ga,), dtype=torch.float32, device=dev)
        state = torch.empty(lib.aopt_pe_mlp_state_bytes(c), dtype=torch.uint8, device=dev)
        stats = torch.empty(3 * c, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(
                lib.aopt_pe_mlp_forward(rows, c, _lib.ptr(pos), _lib.ptr(moments), _lib.ptr(w1), _lib.ptr(b1),
                                        _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(running_mean), _lib.ptr(running_var),
                                        float(eps), int(use_batch), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(out),
                                        _lib.ptr(aux_w), ga, _lib.ptr(aux_out), _lib.ptr(state), state.numel(),
                                        _lib.stream()),
                "pe_mlp_forward",
            )
            _lib.check(lib.aopt_pe_mlp_stats(c, _lib.ptr(state), _lib.ptr(stats), _lib.stream()), "pe_mlp_stats")
        ctx.save_for_backward(pos, moments, w1, gamma, state)
        ctx.meta = (rows, c, int(use_batch), ga, tuple(out.shape))
        ctx.mark_non_differentiable(stats)
        if aux_out is None:
            aux_out = out.new_empty(0)
            ctx.mark_non_differentiable(aux_out)
        return out, aux_out, stats

    @staticmethod
    def backward(ctx, grad_out, grad_aux, _grad_stats):
        lib = _lib.load()
        pos, moments, w1, gamma, state = ctx.saved_tensors
        rows, c, use_batch, ga, out_shape = ctx.meta
        dev = pos.device
        grad_out = torch.zeros(out_shape, dtype=torch.float32, device=dev) if grad_out is None else grad_out.contiguous().float()
        if ga > 0:
            grad_aux = (torch.zeros(out_shape[:-1] + (ga,), dtype=torch.float32, device=dev) if grad_aux is None
                        else grad_aux.contiguous().float())
        else:
            grad_aux = None
        gaw = torch.empty((ga, c), dtype=torch.float32, device=dev) if ga > 0 else None
        gw1 = torch.empty((c, 3), dtype=torch.float32, device=dev)
        gb1, gg, gbeta, gb2 = (torch.empty(c, dtype=torch.float32, device=dev) for _ in range(4))
        gw2 = torch.empty((c, c), dtype=torch.float32, device=dev)
        if rows == 0:
            for t in (gw1, gb1, gg, gbeta, gb2, gw2) + ((gaw,) if gaw is not None else ()):
                t.zero_()
        else:
            with _lib.on_device(dev):
                ws = _lib.workspace(lib.aopt_pe_mlp_backward_workspace_bytes(rows, c), dev)
                _lib.check(
                    lib.aopt_pe_mlp_backward(rows, c, _lib.ptr(pos), _lib.ptr(moments), _lib.ptr(w1), _lib.ptr(gamma),
                                             use_batch, _lib.ptr(grad_out), _lib.ptr(state), _lib.ptr(gw1), _lib.ptr(gb1),
                                             _lib.ptr(gg), _lib.ptr(gbeta), _lib.ptr(gw2), _lib.ptr(gb2), ga,
                                             _lib.ptr(grad_aux), _lib.ptr(gaw), _lib.ptr(ws), ws.numel(), _lib.stream()),
                    "pe_mlp_backward",
                )
        return None, None, gw1, gb1, gg, gbeta, gw2, gb2, None, None, None, None, gaw


def pe_bias_mlp(pos: torch.Tensor, mlp: torch.nn.Sequential, moments: torch.Tensor = None, aux_weight: torch.Tensor = None):
    """`mlp` is the reference's linear_p_bias Sequential: [Linear(3,C), PointBatchNorm(C), ReLU, Linear(C,C)]
    (the BatchNorm1d may be wrapped in a module with a `.norm` attribute).  Returns peb (..., C) fp32 and, in
    training mode, updates the BatchNorm running statistics exactly like nn.BatchNorm1d does.

    aux_weight (ga <= 16, C): additionally returns aux = h @ aux_weightᵀ (..., ga), where h is the hidden
    activation (the input of the last Linear) — i.e. any Linear L applied to peb is available as
    aux = pe_bias_mlp(..., aux_weight=L.weight @ mlp[3].weight) + L.weight @ mlp[3].bias without reading peb."""
    lin1, bn, lin2 = mlp[0], mlp[1], mlp[3]
    bn = getattr(bn, "norm", bn)
    c = lin2.out_features
    _lib.require_cuda(pos, lin1.weight, lin2.weight)
    if not pe_mlp_supported(c) or lin1.in_features != 3 or lin2.in_features != c:
        raise ValueError(f"pe_bias_mlp: unsupported width {c} (aopt_pe_mlp_supported)")
    pos = pos.float().contiguous()
    use_batch = bn.training or not bn.track_running_stats
    if use_batch and moments is None:
        moments = pos_moments(pos)
    f = lambda t: t.detach().float().contiguous() if not t.requires_grad else t.float().contiguous()
    b2 = lin2.bias if lin2.bias is not None else torch.zeros(c, device=pos.device)
    b1 = lin1.bias if lin1.bias is not None else torch.zeros(c, device=pos.device)
    if aux_weight is not None:
        if aux_weight.dim() != 2 or aux_weight.shape[1] != c or not 1 <= aux_weight.shape[0] <= 16:
            raise ValueError("pe_bias_mlp: aux_weight must be (ga <= 16, C)")
        aux_weight = aux_weight.float().contiguous()
    out, aux, stats = _PeMlpFn.apply(pos, moments, f(lin1.weight), f(b1), f(bn.weight), f(bn.bias), f(lin2.weight), f(b2),
                                     None if use_batch else bn.running_mean.float(),
                                     None if use_batch else bn.running_var.float(), bn.eps, use_batch, aux_weight)
    if bn.training and bn.track_running_stats:
        with torch.no_grad():
            rows = pos.numel() // 3
            mean, var = stats[:c], stats[c:2 * c]
            bn.num_batches_tracked += 1
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
            bn.running_var.mul_(1 - m).add_((var * (rows / max(rows - 1, 1))).to(bn.running_var.dtype), alpha=m)
    return out if aux_weight is None else (out, aux)
