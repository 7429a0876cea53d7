"""BatchNorm-shaped element-wise work around the dense layers of a PTv2 block (new operators).

    bn_act(x, bn, relu, residual, row_scale)   [ReLU]([residual +] [row_scale ·] BatchNorm_train(x))  on (rows, C)
        PointBatchNorm (+ nn.ReLU, DropPath, residual add):
        /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:25-45,187-197
    we_tail(rel, upe, cst, bn, lin)            Linear(G,G)(ReLU(BatchNorm_train(rel + upe + cst)))    on (N, k, G)
        weight_encoding[1:] of GroupedVectorAttention (:94-99,120)

Kernels: ao_b200/csrc/dense.cu (fp32 statistics, fp64 combination, no atomics).  These replace ATen's BatchNorm /
ReLU / cast / skinny-GEMM kernels in TRAINING mode; in evaluation mode (running statistics) the torch modules run.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from .. import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


def fused_dense_enabled() -> bool:
    """AOPT_FUSED_DENSE=0 routes every BatchNorm site back through torch.nn (A/B measurements, tests)."""
    return os.environ.get("AOPT_FUSED_DENSE", "1") != "0"


def bn_act_supported(channels: int) -> bool:
    return bool(_lib.load().aopt_bn_act_supported(int(channels)))


def we_tail_supported(groups: int) -> bool:
    return bool(_lib.load().aopt_we_tail_supported(int(groups)))


_WS = {}


def _dense_ws(width: int, dev: torch.device):
    """Scratch of the dense kernels (partial rows + fp64 sums), allocated once per (device, stream, width): the
    calls are stream-ordered, so consecutive operators on one stream can share it; ~100 BatchNorm sites per step would
    otherwise each pay a size query and an allocator round trip on a host-bound path."""
    key = (dev.index, _lib.stream(), width)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = _lib.workspace(_lib.load().aopt_dense_workspace_bytes(width), dev)
    return ws


class _BnActFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, row_scale, running_mean, running_var, momentum, eps, relu, out_dtype,
                pre_bias, tracked):
        lib = _lib.load()
        dev = x.device
        rows, c = x.shape
        out = torch.empty((rows, c), dtype=out_dtype, device=dev)
        stats = torch.empty(2 * c, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            _lib.check(
                lib.aopt_bn_act_forward(rows, c, x.data_ptr(), c, _DT[x.dtype], gamma.data_ptr(), beta.data_ptr(), eps,
                                        _lib.ptr(residual), _lib.ptr(row_scale), int(relu), out.data_ptr(), _DT[out_dtype],
                                        stats.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var), momentum,
                                        _lib.ptr(pre_bias), _lib.ptr(tracked), ws.data_ptr(), ws.numel(), _lib.stream()),
                "bn_act_forward")
        ctx.save_for_backward(x, out if relu else None, gamma, stats, row_scale)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        ctx.bias_like = pre_bias
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        x, out, gamma, stats, row_scale = ctx.saved_tensors
        dev = x.device
        rows, c = x.shape
        odt = out.dtype if out is not None else grad_out.dtype
        if grad_out.dtype != odt or not grad_out.is_contiguous():
            grad_out = grad_out.to(odt).contiguous()
        gx = torch.empty_like(x)
        want_res = ctx.has_res and ctx.needs_input_grad[3]
        # without a ReLU the residual's gradient IS grad_out
        gres = torch.empty_like(grad_out) if (want_res and ctx.relu) else None
        gg = torch.empty(c, dtype=torch.float32, device=dev)
        gb = torch.empty(c, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            _lib.check(
                lib.aopt_bn_act_backward(rows, c, grad_out.data_ptr(), _lib.ptr(out), _DT[odt], x.data_ptr(), c, _DT[x.dtype],
                                         stats.data_ptr(), gamma.data_ptr(), _lib.ptr(row_scale), gx.data_ptr(), c,
                                         _lib.ptr(gres), gg.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream()),
                "bn_act_backward")
        if want_res and not ctx.relu:
            gres = grad_out
        # a bias in front of a training-mode BatchNorm has an identically zero gradient
        gpb = torch.zeros_like(ctx.bias_like) if (ctx.bias_like is not None and ctx.needs_input_grad[11]) else None
        return gx, gg, gb, gres, None, None, None, None, None, None, None, gpb, None


def _torch_bn_act(x, bn, relu, residual, row_scale, out_dtype):
    y = bn(x)
    if row_scale is not None:
        y = y * row_scale.to(y.dtype).unsqueeze(-1)
    if residual is not None:
        y = residual + y
    if relu:
        y = torch.relu(y)
    return y if out_dtype is None else y.to(out_dtype)


_WIDTH_OK = {}


def bn_fusable(bn: torch.nn.Module, c: int, rows: int, is_cuda: bool, dtype: torch.dtype, out_dtype: torch.dtype = None) -> bool:
    """True when bn_act would run the fused kernels for a (rows, c) input of `dtype` (training-mode statistics, CUDA,
    fp32 / bf16, supported width)."""
    bn = getattr(bn, "norm", bn)
    ok = _WIDTH_OK.get(c)
    if ok is None:
        ok = _WIDTH_OK[c] = bn_act_supported(c)
    return (ok and fused_dense_enabled() and (bn.training or not bn.track_running_stats) and is_cuda and dtype in _DT
            and rows >= 2 and bn.weight is not None and bn.weight.dtype == torch.float32
            and (out_dtype is None or out_dtype in _DT))


def bn_act_usable(x: torch.Tensor, bn: torch.nn.Module, out_dtype: torch.dtype = None) -> bool:
    c = x.shape[-1]
    return bn_fusable(bn, c, x.numel() // max(c, 1), x.is_cuda, x.dtype, out_dtype)


def _momentum(bn, track):
    """(momentum, counter the kernel increments).  With a fixed momentum the num_batches_tracked increment rides in the
    apply kernel (one launch and ~8 us of host time less per BatchNorm site); the cumulative average (momentum=None)
    needs the count on the host, as in nn.BatchNorm1d."""
    if not track:
        return 0.0, None
    nbt = bn.num_batches_tracked
    if bn.momentum is not None and nbt is not None and nbt.is_cuda and nbt.dtype == torch.int64:
        return float(bn.momentum), nbt
    if nbt is not None:
        nbt.add_(1)
    return (float(bn.momentum) if bn.momentum is not None else 1.0 / float(nbt)), None


def bn_act(x: torch.Tensor, bn: torch.nn.Module, relu: bool = False, residual: torch.Tensor = None,
           row_scale: torch.Tensor = None, out_dtype: torch.dtype = None, pre_bias: torch.Tensor = None) -> torch.Tensor:
    """`bn` is a BatchNorm1d (or a module with a `.norm` BatchNorm1d: the reference's PointBatchNorm).  x is (N, C) or
    (N, L, C); statistics over all leading dimensions, like PointBatchNorm.  Returns
        [relu]( [residual +] [row_scale[:, None] *] bn(x) )       in out_dtype (default: x.dtype, or the residual's),
    and updates the running statistics like nn.BatchNorm1d.  row_scale (N,) fp32 carries DropPath (mask / keep_prob).
    pre_bias (C,): a bias the caller did NOT add to x because training-mode normalisation removes it (the bias of the
    Linear in front): it only shifts the running mean, and receives a zero gradient.  Only with bn_act_usable()."""
    bn = getattr(bn, "norm", bn)
    c = x.shape[-1]
    rows = x.numel() // max(c, 1)
    if not bn_act_usable(x, bn, out_dtype):
        if pre_bias is not None:
            x = x + pre_bias.to(x.dtype)
        shape = x.shape
        y = _torch_bn_act(x.reshape(rows, c), bn, relu, None if residual is None else residual.reshape(rows, c), row_scale,
                          out_dtype)
        return y.view(shape)
    shape = x.shape
    x2 = x.reshape(rows, c)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if out_dtype is None:
        out_dtype = x.dtype if residual is None else torch.promote_types(x.dtype, residual.dtype)
    if residual is not None:
        residual = residual.reshape(rows, c)
        if residual.dtype != out_dtype or not residual.is_contiguous():
            residual = residual.to(out_dtype).contiguous()
    if row_scale is not None:
        row_scale = row_scale.reshape(rows).float().contiguous()
    track = bn.training and bn.track_running_stats
    momentum, tracked = _momentum(bn, track)
    if pre_bias is not None and (pre_bias.dtype != torch.float32 or not pre_bias.is_contiguous()):
        pre_bias = pre_bias.float().contiguous()
    out = _BnActFn.apply(x2, bn.weight, bn.bias, residual, row_scale, bn.running_mean if track else None,
                         bn.running_var if track else None, momentum, float(bn.eps), bool(relu), out_dtype, pre_bias, tracked)
    return out.view(shape)


class _WeTailFn(Function):
    """rel is either the materialised (N, k, G) tensor (kp = qp = idx = None) or None (gather mode: rel = kp[idx] - qp)."""

    @staticmethod
    def forward(ctx, rel, kp, qp, idx, upe, cst, gamma, beta, w2, b2, running_mean, running_var, momentum, eps, tracked):
        lib = _lib.load()
        src = rel if rel is not None else kp
        dev = src.device
        g = src.shape[-1]
        shape = tuple(rel.shape) if rel is not None else (idx.shape[0], idx.shape[1], g)
        nsample = shape[-2] if len(shape) >= 2 else 1
        rows = 1
        for d in shape[:-1]:
            rows *= d
        logits = torch.empty(shape, dtype=torch.float32, device=dev)
        stats = torch.empty(2 * g, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            ws = _dense_ws(3 * g + g * g, dev)
            _lib.check(
                lib.aopt_we_tail_forward(rows, g, _lib.ptr(rel), _lib.ptr(kp), _lib.ptr(qp), _lib.ptr(idx), nsample, _lib.ptr(upe),
                                         _lib.ptr(cst), gamma.data_ptr(), beta.data_ptr(), eps, w2.data_ptr(), _lib.ptr(b2),
                                         logits.data_ptr(), stats.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var),
                                         momentum, _lib.ptr(tracked), ws.data_ptr(), ws.numel(), _lib.stream()),
                "we_tail_forward")
        ctx.save_for_backward(rel, kp, qp, upe, cst, gamma, beta, w2, stats)
        ctx.idx = idx
        ctx.meta = (rows, g, nsample, shape)
        ctx.has_b2 = b2 is not None
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        lib = _lib.load()
        rel, kp, qp, upe, cst, gamma, beta, w2, stats = ctx.saved_tensors
        idx = ctx.idx
        rows, g, nsample, shape = ctx.meta
        dev = grad_logits.device
        grad_logits = grad_logits.float().contiguous()
        gu = torch.empty(shape, dtype=torch.float32, device=dev)
        gg, gb, gb2 = (torch.empty(g, dtype=torch.float32, device=dev) for _ in range(3))
        gw2 = torch.empty((g, g), dtype=torch.float32, device=dev)
        gkp = gqp = None
        with _lib.on_device(dev):
            ws = _dense_ws(3 * g + g * g, dev)
            _lib.check(
                lib.aopt_we_tail_backward(rows, g, _lib.ptr(rel), _lib.ptr(kp), _lib.ptr(qp), _lib.ptr(idx), nsample, _lib.ptr(upe),
                                          _lib.ptr(cst), grad_logits.data_ptr(), stats.data_ptr(), gamma.data_ptr(),
                                          beta.data_ptr(), w2.data_ptr(), gu.data_ptr(), gg.data_ptr(), gb.data_ptr(),
                                          gb2.data_ptr(), gw2.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream()),
                "we_tail_backward")
            if rel is None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                # gather mode: grad_u scattered back to the key projection (CSR walk) and summed over k for the query
                # projection in one pass (aopt_relation_backward, the G-wide backward of gva_relation)
                from ._csr import get_csr

                n = kp.shape[0]
                csr = get_csr(idx, n, 0)
                gkp = torch.empty((n, g), dtype=torch.float32, device=dev)
                gqp = torch.empty((n, g), dtype=torch.float32, device=dev)
                _lib.check(lib.aopt_relation_backward(n, nsample, g, gu.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(),
                                                      gkp.data_ptr(), gqp.data_ptr(), _lib.stream()), "relation_backward")
        # cst sits in front of a training-mode BatchNorm: its gradient (the column sums of grad_u) is identically zero
        gcst = torch.zeros_like(cst) if (cst is not None and ctx.needs_input_grad[5]) else None
        return (gu if rel is not None else None, gkp, gqp, None, gu if upe is not None else None, gcst, gg, gb, gw2,
                gb2 if ctx.has_b2 else None, None, None, None, None, None)


def we_tail_usable(rel: torch.Tensor, bn: torch.nn.Module, rows: int = None) -> bool:
    """rel: the (…, G) first addend — or, for gather mode, the (N, G) key projection with rows = N·k."""
    bn = getattr(bn, "norm", bn)
    if rows is None:
        rows = rel.numel() // rel.shape[-1]
    return (fused_dense_enabled() and rel.is_cuda and rel.dtype == torch.float32 and (bn.training or not bn.track_running_stats)
            and bn.weight is not None and rows >= 2 and we_tail_supported(rel.shape[-1]))


def we_tail(rel: torch.Tensor, upe: torch.Tensor, cst: torch.Tensor, bn: torch.nn.Module, lin: torch.nn.Linear,
            gather=None) -> torch.Tensor:
    """logits = lin(ReLU(bn(rel + upe + cst))) with `bn` / `lin` = weight_encoding[1] / weight_encoding[3]; rel, upe
    (..., G) fp32, cst (G) or None.  Training-mode statistics only (check with we_tail_usable); G in {6, 12}.
    gather = (kp, qp, idx) with rel = None: rel = kp[idx] - qp[:, None] ((N, G) key / query projections, idx (N, k), the
    G-wide gva_relation) is formed inside the kernels and never stored; gradients flow to kp and qp."""
    bn = getattr(bn, "norm", bn)
    kp = qp = idx = None
    if gather is not None:
        if rel is not None:
            raise ValueError("we_tail: pass either rel or gather=(kp, qp, idx)")
        kp, qp, idx = gather
        _lib.require_cuda(kp, qp, idx, lin.weight)
        if idx.dtype != torch.int32 or not idx.is_contiguous():
            idx = idx.int().contiguous()
        kp, qp = kp.float().contiguous(), qp.float().contiguous()
        if kp.shape != qp.shape or idx.shape[0] != qp.shape[0] or idx.dim() != 2:
            raise ValueError("we_tail: gather mode needs kp, qp (N, G) and idx (N, k) of one point set")
        first, rows = kp, idx.numel()
        shape = (idx.shape[0], idx.shape[1], kp.shape[1])
    else:
        _lib.require_cuda(rel, lin.weight)
        rel = rel.contiguous()
        first, rows = rel, None
        shape = rel.shape
    if not we_tail_usable(first, bn, rows):
        raise ValueError("we_tail: unsupported input (see we_tail_usable)")
    g = first.shape[-1]
    if upe is not None:
        upe = upe.float().contiguous()
        if tuple(upe.shape) != tuple(shape):
            raise ValueError("we_tail: upe must have rel's shape")
    if cst is not None:
        cst = cst.float().contiguous()
        if cst.numel() != g:
            raise ValueError("we_tail: cst must have G entries")
    track = bn.training and bn.track_running_stats
    momentum, tracked = _momentum(bn, track)
    f = lambda t: t if t.dtype == torch.float32 and t.is_contiguous() else t.float().contiguous()
    return _WeTailFn.apply(rel, kp, qp, idx, upe, cst, f(bn.weight), f(bn.bias), f(lin.weight),
                           None if lin.bias is None else f(lin.bias), bn.running_mean if track else None,
                           bn.running_var if track else None, momentum, float(bn.eps), tracked)


# ---- autocast Linear with cached low-precision weights ------------------------------------------------------------------
# torch.autocast casts every fp32 weight to bf16 through autograd: a copy kernel, a ToCopyBackward node and a second copy
# kernel per weight and step, plus ~40 us of dispatcher / autograd host time per nn.Linear call — on a path whose GEMMs
# take 10-30 us (profiles/r02o_model_step_torch_profile.txt: `aten::_to_copy` 537 calls, 6.6 ms of host time).  `linear`
# is the same computation (bf16 operands, fp32 accumulation in cuBLAS, the weight gradient returned in fp32) with the cast
# weight cached until the optimizer changes the parameter (`_version`) and no autograd node for the casts.
_MM_OUT_DTYPE = None


def _shadow(p: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """Low-precision copy of parameter p, kept on the parameter until an in-place update changes its version."""
    hit = getattr(p, "_aopt_lo", None)
    if hit is not None and hit[0] == p._version and hit[1] == dt:
        return hit[2]
    lo = p.detach().to(dt)
    p._aopt_lo = (p._version, dt, lo)
    return lo


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with low-precision operands and an fp32 result (cuBLAS writes the fp32 accumulator; no cast kernel)."""
    global _MM_OUT_DTYPE
    if _MM_OUT_DTYPE is None:
        try:
            torch.mm(a, b, out_dtype=torch.float32)
            _MM_OUT_DTYPE = True
        except (TypeError, RuntimeError):
            _MM_OUT_DTYPE = False
    return torch.mm(a, b, out_dtype=torch.float32) if _MM_OUT_DTYPE else torch.mm(a, b).float()


def col_sum(x: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a dense (rows, c) fp32 / bf16 matrix (a bias gradient); c % 4 == 0, c <= 1024."""
    rows, c = x.shape
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = _dense_ws(2 * c, x.device)
    _lib.check(_lib.load().aopt_col_sum(rows, c, x.data_ptr(), c, _DT[x.dtype], out.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _lib.stream()), "col_sum")
    return out


def _bias_grad(g2: torch.Tensor) -> torch.Tensor:
    c = g2.shape[1]
    if g2.dtype in _DT and g2.is_contiguous() and g2.shape[0] >= 1024 and _WIDTH_OK.get(c, bn_act_supported(c)):
        return col_sum(g2)
    return g2.sum(0, dtype=torch.float32)


_SKINNY_OK = {}


def _weight_grad(g2: torch.Tensor, xb: torch.Tensor) -> torch.Tensor:
    """g2ᵀ @ xb in fp32.  A Linear with 6 / 12 outputs over >= 16k rows goes through aopt_skinny_wgrad."""
    rows, g = g2.shape
    c = xb.shape[1]
    if g <= 16 and rows >= 16384 and g2.dtype in _DT and xb.dtype in _DT and g2.is_contiguous() and xb.is_contiguous():
        if _skinny_ok(g, c):
            out = torch.empty((g, c), dtype=torch.float32, device=xb.device)
            ws = _dense_ws(g * c, xb.device)
            _lib.check(_lib.load().aopt_skinny_wgrad(rows, g, c, g2.data_ptr(), _DT[g2.dtype], xb.data_ptr(), c, _DT[xb.dtype],
                                                     out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream()), "skinny_wgrad")
            return out
    return _wgrad_mm(g2.t(), xb) if xb.dtype != torch.float32 else torch.mm(g2.float().t(), xb)


def _skinny_ok(g: int, c: int) -> bool:
    ok = _SKINNY_OK.get((g, c))
    if ok is None:
        ok = _SKINNY_OK[(g, c)] = bool(_lib.load().aopt_skinny_wgrad_supported(g, c))
    return ok


def _wgrad_mm(gt: torch.Tensor, xb: torch.Tensor) -> torch.Tensor:
    """gt (out, rows) @ xb (rows, in) with an fp32 result (cuBLAS writes the accumulator; the output is tiny)."""
    return _mm_f32(gt, xb)


class _SkinnyLinearFn(Function):
    """x (rows, c) -> x wᵀ (+ b) (rows, g) fp32, w (g, c) fp32 with a handful of outputs: all three products in own kernels."""

    @staticmethod
    def forward(ctx, x, w, b):
        rows, c = x.shape
        g = w.shape[0]
        out = torch.empty((rows, g), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().aopt_skinny_linear(rows, g, c, x.data_ptr(), c, _DT[x.dtype], w.data_ptr(), _lib.ptr(b),
                                                  out.data_ptr(), _lib.stream()), "skinny_linear")
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        rows, c = x.shape
        g = w.shape[0]
        if gy.dtype != torch.float32 or not gy.is_contiguous():
            gy = gy.float().contiguous()
        gx = gw = gb = None
        lib = _lib.load()
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            _lib.check(lib.aopt_skinny_dgrad(rows, g, c, gy.data_ptr(), w.data_ptr(), gx.data_ptr(), c, _DT[x.dtype],
                                             _lib.stream()), "skinny_dgrad")
        if ctx.needs_input_grad[1]:
            gw = torch.empty((g, c), dtype=torch.float32, device=x.device)
            ws = _dense_ws(g * c, x.device)
            _lib.check(lib.aopt_skinny_wgrad(rows, g, c, gy.data_ptr(), 0, x.data_ptr(), c, _DT[x.dtype], gw.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _lib.stream()), "skinny_wgrad")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb


def _small_k_usable(x2: torch.Tensor, w: torch.Tensor) -> bool:
    """A Linear with a handful of INPUT channels over many rows (the patch-embedding projection, 6 -> 48)."""
    return (x2.dtype == torch.float32 and x2.is_contiguous() and x2.shape[0] >= 16384 and w.shape[1] <= 20
            and w.dtype == torch.float32 and _skinny_ok(w.shape[1], w.shape[0]))


def _small_k_forward(x2: torch.Tensor, w: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """y (rows, out) in dt = x2 (rows, k) fp32 · wᵀ through aopt_skinny_dgrad (its "gradient" operand is x2)."""
    rows, k = x2.shape
    cout = w.shape[0]
    y = torch.empty((rows, cout), dtype=dt, device=x2.device)
    wt = w.detach().t().contiguous()
    _lib.check(_lib.load().aopt_skinny_dgrad(rows, k, cout, x2.data_ptr(), wt.data_ptr(), y.data_ptr(), cout, _DT[dt],
                                             _lib.stream()), "skinny_dgrad")
    return y


def _small_k_wgrad(gy: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """(out, k) fp32 = gyᵀ · x2 through aopt_skinny_wgrad with the operands exchanged."""
    rows, k = x2.shape
    cout = gy.shape[1]
    gwt = torch.empty((k, cout), dtype=torch.float32, device=x2.device)
    ws = _dense_ws(k * cout, x2.device)
    _lib.check(_lib.load().aopt_skinny_wgrad(rows, k, cout, x2.data_ptr(), 0, gy.data_ptr(), cout, _DT[gy.dtype], gwt.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _lib.stream()), "skinny_wgrad")
    return gwt.t().contiguous()


class _LinearFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, dt, out_f32):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        xb = x2 if x2.dtype == dt else x2.to(dt)
        wb = _shadow(w, dt)
        if out_f32:
            y = _mm_f32(xb, wb.t())
            if b is not None:
                y.add_(b)
        else:
            y = torch.mm(xb, wb.t()) if b is None else torch.addmm(_shadow(b, dt), xb, wb.t())
        ctx.save_for_backward(xb, wb)
        ctx.x_dtype = x.dtype
        ctx.has_bias = b is not None
        return y.view(shape[:-1] + (w.shape[0],))

    @staticmethod
    def backward(ctx, gy):
        xb, wb = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if g2.dtype != xb.dtype:
            g2 = g2.to(xb.dtype)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(g2, wb).view(gy.shape[:-1] + (wb.shape[1],))
            if gx.dtype != ctx.x_dtype:
                gx = gx.to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            gw = _weight_grad(g2, xb)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _bias_grad(g2)
        return gx, gw, gb, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None, out_f32: bool = False) -> torch.Tensor:
    """torch.nn.functional.linear; under CUDA autocast the low-precision copy of the fp32 parameters is cached across
    calls (see above).  Same arithmetic as the autocast Linear it replaces.  out_f32: return fp32 (under autocast the
    GEMM writes its fp32 accumulator instead of a rounded copy that the caller would cast back)."""
    if (out_f32 and x.is_cuda and x.dim() == 2 and weight.shape[0] <= 20 and x.shape[0] >= 16384
            and fused_dense_enabled() and x.dtype in _DT and weight.dtype == torch.float32 and x.is_contiguous()
            and weight.is_contiguous() and (bias is None or bias.dtype == torch.float32)
            and _skinny_ok(weight.shape[0], weight.shape[1])):
        return _SkinnyLinearFn.apply(x, weight, bias)    # a handful of outputs over many rows: own kernels, any mode
    if (x.is_cuda and torch.is_autocast_enabled() and fused_dense_enabled() and weight.dtype == torch.float32
            and x.dtype in (torch.float32, torch.bfloat16, torch.float16) and x.numel() > 0):
        return _LinearFn.apply(x, weight, bias, torch.get_autocast_dtype("cuda"), bool(out_f32))
    y = torch.nn.functional.linear(x, weight, bias)
    return y.float() if out_f32 else y


# ---- q | k | v of GroupedVectorAttention as ONE GEMM -----------------------------------------------------------------
# linear_q / linear_k (Linear -> PointBatchNorm -> ReLU) and linear_v (Linear) read the same input (…v2m2_base.py:104-108):
# one (N, C) x (C, 3C) product instead of three, the two BatchNorm stages read their column block of the result in
# place (row stride 3C) and, in the backward pass, write their input gradients into the column blocks of ONE (N, 3C)
# buffer, so dX and dW are one product each as well: 3 cuBLAS calls per block instead of 9, one autograd node instead of 5.
def _cat_weights(ws, dt):
    """[Wq; Wk; Wv] in dtype dt, cached on Wq until one of the three parameters changes."""
    wq = ws[0]
    ver = tuple(w._version for w in ws)
    hit = getattr(wq, "_aopt_cat", None)
    if hit is not None and hit[0] == ver and hit[1] == dt and hit[3] is ws[1] and hit[4] is ws[2]:
        return hit[2]
    cat = torch.cat([w.detach() for w in ws], 0).to(dt)
    wq._aopt_cat = (ver, dt, cat, ws[1], ws[2])
    return cat


class _QkvFn(Function):
    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, gq, betaq, gk, betak, rmq, rvq, rmk, rvk, momq, momk, epsq, epsk,
                trq, trk, dt, qk_dtype):
        lib = _lib.load()
        dev = x.device
        rows, c = x.shape
        xb = x if x.dtype == dt else x.to(dt)
        wcat = _cat_weights((wq, wk, wv), dt)
        y = torch.mm(xb, wcat.t())                                           # (rows, 3c): q_pre | k_pre | v_pre
        q = torch.empty((rows, c), dtype=qk_dtype, device=dev)
        k = torch.empty((rows, c), dtype=qk_dtype, device=dev)
        stats = torch.empty(4 * c, dtype=torch.float32, device=dev)
        esz = y.element_size()
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            for j, (out, g, b, rm, rv, mom, eps, bias, tr) in enumerate(((q, gq, betaq, rmq, rvq, momq, epsq, bq, trq),
                                                                         (k, gk, betak, rmk, rvk, momk, epsk, bk, trk))):
                _lib.check(
                    lib.aopt_bn_act_forward(rows, c, y.data_ptr() + j * c * esz, 3 * c, _DT[dt], g.data_ptr(), b.data_ptr(),
                                            eps, 0, 0, 1, out.data_ptr(), _DT[qk_dtype], stats.data_ptr() + j * 8 * c,
                                            _lib.ptr(rm), _lib.ptr(rv), mom, _lib.ptr(bias), _lib.ptr(tr), ws.data_ptr(),
                                            ws.numel(), _lib.stream()),
                    "bn_act_forward")
        v = torch.empty((rows, c), dtype=torch.float32, device=dev)          # dense fp32 copy of the v block (+ bias)
        _lib.check(lib.aopt_copy_cols(rows, c, y.data_ptr() + 2 * c * esz, 3 * c, _DT[dt], _lib.ptr(bv), v.data_ptr(), c, 0,
                                      _lib.stream()), "copy_cols")
        ctx.save_for_backward(xb, wcat, y, q, k, gq, gk, stats)
        ctx.x_dtype = x.dtype
        ctx.biases = (bq, bk, bv)
        return q, k, v

    @staticmethod
    def backward(ctx, gq_out, gk_out, gv):
        lib = _lib.load()
        xb, wcat, y, q, k, gq, gk, stats = ctx.saved_tensors
        dev = xb.device
        rows, c = xb.shape
        dt = y.dtype
        gy = torch.empty_like(y)
        ggq, gbq, ggk, gbk = (torch.empty(c, dtype=torch.float32, device=dev) for _ in range(4))
        esz = y.element_size()
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            for j, (go, out, g, gg, gb) in enumerate(((gq_out, q, gq, ggq, gbq), (gk_out, k, gk, ggk, gbk))):
                if go is None:
                    gy[:, j * c:(j + 1) * c].zero_()
                    gg.zero_()
                    gb.zero_()
                    continue
                if go.dtype != out.dtype or not go.is_contiguous():
                    go = go.to(out.dtype).contiguous()
                _lib.check(
                    lib.aopt_bn_act_backward(rows, c, go.data_ptr(), out.data_ptr(), _DT[out.dtype], y.data_ptr() + j * c * esz,
                                             3 * c, _DT[dt], stats.data_ptr() + j * 8 * c, g.data_ptr(), 0,
                                             gy.data_ptr() + j * c * esz, 3 * c, 0, gg.data_ptr(), gb.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _lib.stream()),
                    "bn_act_backward")
        if gv is None:
            gy[:, 2 * c:].zero_()
        else:
            if gv.dtype != torch.float32 or not gv.is_contiguous():
                gv = gv.float().contiguous()
            _lib.check(lib.aopt_copy_cols(rows, c, gv.data_ptr(), c, 0, 0, gy.data_ptr() + 2 * c * esz, 3 * c, _DT[dt],
                                          _lib.stream()), "copy_cols")
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(gy, wcat)
            if gx.dtype != ctx.x_dtype:
                gx = gx.to(ctx.x_dtype)
        gw = _wgrad_mm(gy.t(), xb) if dt != torch.float32 else torch.mm(gy.t(), xb)
        bq, bk, bv = ctx.biases
        zq = torch.zeros_like(bq) if bq is not None else None          # in front of a training-mode BatchNorm
        zk = torch.zeros_like(bk) if bk is not None else None
        gbv = _bias_grad(gv) if (bv is not None and gv is not None) else (None if bv is None else torch.zeros_like(bv))
        return (gx, gw[:c], zq, gw[c:2 * c], zk, gw[2 * c:], gbv, ggq, gbq, ggk, gbk) + (None,) * 12


def qkv_usable(x: torch.Tensor, seq_q, seq_k, lin_v) -> bool:
    """linear_q / linear_k are [Linear(C,C), PointBatchNorm(C), ReLU], linear_v is Linear(C,C), all on x (N, C), and the
    BatchNorm stages run on batch statistics."""
    try:
        lq, bq, lk, bk = seq_q[0], getattr(seq_q[1], "norm", seq_q[1]), seq_k[0], getattr(seq_k[1], "norm", seq_k[1])
    except (TypeError, IndexError):
        return False
    c = x.shape[-1]
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
    return (x.dim() == 2 and len(seq_q) == 3 and len(seq_k) == 3 and isinstance(seq_q[2], torch.nn.ReLU)
            and isinstance(seq_k[2], torch.nn.ReLU) and isinstance(bq, torch.nn.BatchNorm1d) and isinstance(bk, torch.nn.BatchNorm1d)
            and all(isinstance(l, torch.nn.Linear) and l.weight.shape == (c, c) and l.weight.dtype == torch.float32
                    for l in (lq, lk, lin_v))
            and bn_fusable(bq, c, x.shape[0], x.is_cuda, dt) and bn_fusable(bk, c, x.shape[0], x.is_cuda, dt))


def qkv_bn(x: torch.Tensor, seq_q, seq_k, lin_v, qk_dtype: torch.dtype = None):
    """(ReLU(BN_q(x Wqᵀ + bq)), ReLU(BN_k(x Wkᵀ + bk)), x Wvᵀ + bv) with one GEMM (see above).  q, k in qk_dtype
    (default: the GEMM's dtype — the autocast dtype, or x's), v in fp32 (it feeds gva_aggregate).  Check qkv_usable()."""
    lq, bnq, lk, bnk = seq_q[0], getattr(seq_q[1], "norm", seq_q[1]), seq_k[0], getattr(seq_k[1], "norm", seq_k[1])
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
    if qk_dtype is None:
        qk_dtype = dt
    if not x.is_contiguous():
        x = x.contiguous()
    trackq = bnq.training and bnq.track_running_stats
    trackk = bnk.training and bnk.track_running_stats
    momq, trq = _momentum(bnq, trackq)
    momk, trk = _momentum(bnk, trackk)
    return _QkvFn.apply(x, lq.weight, lq.bias, lk.weight, lk.bias, lin_v.weight, lin_v.bias, bnq.weight, bnq.bias, bnk.weight,
                        bnk.bias, bnq.running_mean if trackq else None, bnq.running_var if trackq else None,
                        bnk.running_mean if trackk else None, bnk.running_var if trackk else None, momq, momk,
                        float(bnq.eps), float(bnk.eps), trq, trk, dt, qk_dtype)


# ---- Linear -> BatchNorm [-> ReLU] as ONE autograd node ----------------------------------------------------------------
# The step is host-bound: every autograd Function costs ~20 us of Python / engine time per direction.  A Linear whose
# output goes straight into a training-mode BatchNorm (fc1 + norm1, fc3 + norm3 + residual, GridPool.fc + norm, the
# Linear -> PointBatchNorm -> ReLU triples) is one node here: the cuBLAS product, then the bn_act kernels on its result;
# backward = bn_act backward into the GEMM-output gradient, then the two products.
class _LinearBnFn(Function):
    @staticmethod
    def forward(ctx, x, w, lin_bias, gamma, beta, residual, row_scale, rm, rv, mom, eps, relu, out_dtype, tracked, dt):
        lib = _lib.load()
        dev = x.device
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        small_k = _small_k_usable(x2, w) and not ctx.needs_input_grad[0]
        if small_k:
            xb, wb = x2, w                                                   # fp32 rows of <= 20 inputs: own kernels
            y = _small_k_forward(x2, w, dt)
        else:
            xb = x2 if x2.dtype == dt else x2.to(dt)
            wb = w if dt == torch.float32 else _shadow(w, dt)
            y = torch.mm(xb, wb.t())
        ctx.small_k = small_k
        rows, c = y.shape
        out = torch.empty((rows, c), dtype=out_dtype, device=dev)
        stats = torch.empty(2 * c, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            _lib.check(
                lib.aopt_bn_act_forward(rows, c, y.data_ptr(), c, _DT[dt], gamma.data_ptr(), beta.data_ptr(), eps,
                                        _lib.ptr(residual), _lib.ptr(row_scale), int(relu), out.data_ptr(), _DT[out_dtype],
                                        stats.data_ptr(), _lib.ptr(rm), _lib.ptr(rv), mom, _lib.ptr(lin_bias), _lib.ptr(tracked),
                                        ws.data_ptr(), ws.numel(), _lib.stream()),
                "bn_act_forward")
        ctx.save_for_backward(xb, wb, y, out if relu else None, gamma, stats, row_scale)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        ctx.lin_bias = lin_bias
        ctx.x_meta = (x.dtype, shape)
        return out.view(shape[:-1] + (c,))

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        xb, wb, y, out, gamma, stats, row_scale = ctx.saved_tensors
        dev = xb.device
        rows, c = y.shape
        odt = out.dtype if out is not None else grad_out.dtype
        grad_out = grad_out.reshape(rows, c)
        if grad_out.dtype != odt or not grad_out.is_contiguous():
            grad_out = grad_out.to(odt).contiguous()
        gy = torch.empty_like(y)
        want_res = ctx.has_res and ctx.needs_input_grad[5]
        gres = torch.empty_like(grad_out) if (want_res and ctx.relu) else None
        gg = torch.empty(c, dtype=torch.float32, device=dev)
        gb = torch.empty(c, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            ws = _dense_ws(2 * c, dev)
            _lib.check(
                lib.aopt_bn_act_backward(rows, c, grad_out.data_ptr(), _lib.ptr(out), _DT[odt], y.data_ptr(), c, _DT[y.dtype],
                                         stats.data_ptr(), gamma.data_ptr(), _lib.ptr(row_scale), gy.data_ptr(), c,
                                         _lib.ptr(gres), gg.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream()),
                "bn_act_backward")
        if want_res and not ctx.relu:
            gres = grad_out
        x_dtype, shape = ctx.x_meta
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(gy, wb)
            if gx.dtype != x_dtype:
                gx = gx.to(x_dtype)
            gx = gx.view(shape)
        if ctx.needs_input_grad[1]:
            if ctx.small_k:
                gw = _small_k_wgrad(gy, xb)
            else:
                gw = _weight_grad(gy, xb) if y.dtype != torch.float32 else torch.mm(gy.t(), xb)
        gbias = torch.zeros_like(ctx.lin_bias) if (ctx.lin_bias is not None and ctx.needs_input_grad[2]) else None
        if gres is not None and ctx.has_res:
            gres = gres.view(shape[:-1] + (c,))
        return (gx, gw, gbias, gg, gb, gres) + (None,) * 9


def linear_bn_act(x: torch.Tensor, lin: torch.nn.Linear, bn: torch.nn.Module, relu: bool = False,
                  residual: torch.Tensor = None, row_scale: torch.Tensor = None, out_dtype: torch.dtype = None) -> torch.Tensor:
    """[relu]([residual +] [row_scale *] bn(lin(x))) as one autograd node when bn runs on batch statistics (see above);
    otherwise bn_act(linear(x))."""
    bn = getattr(bn, "norm", bn)
    c = lin.out_features
    rows = x.numel() // max(x.shape[-1], 1)
    dt = torch.get_autocast_dtype("cuda") if (x.is_cuda and torch.is_autocast_enabled()) else x.dtype
    if not (bn_fusable(bn, c, rows, x.is_cuda, dt, out_dtype) and lin.weight.dtype == torch.float32 and dt in _DT):
        return bn_act(linear(x, lin.weight, lin.bias), bn, relu=relu, residual=residual, row_scale=row_scale, out_dtype=out_dtype)
    if out_dtype is None:
        out_dtype = dt if residual is None else torch.promote_types(dt, residual.dtype)
    if residual is not None:
        residual = residual.reshape(rows, c)
        if residual.dtype != out_dtype or not residual.is_contiguous():
            residual = residual.to(out_dtype).contiguous()
    if row_scale is not None:
        row_scale = row_scale.reshape(rows).float().contiguous()
    track = bn.training and bn.track_running_stats
    momentum, tracked = _momentum(bn, track)
    return _LinearBnFn.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, residual, row_scale,
                             bn.running_mean if track else None, bn.running_var if track else None, momentum, float(bn.eps),
                             bool(relu), out_dtype, tracked, dt)
