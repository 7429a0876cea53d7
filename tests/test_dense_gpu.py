"""GPU parity of the BatchNorm-shaped operators (ao_b200/csrc/dense.cu, through the C ABI via ao_b200.pointops.bn_act /
we_tail) against the torch modules they replace in a PTv2 block — PointBatchNorm + ReLU + DropPath + residual
(/root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:25-45,187-197) and
weight_encoding[1:] (:94-99,120) — evaluated in fp32 on the same inputs.
Tolerances: fp32 in / out: rtol 2e-5, atol 2e-5 (the sums are formed in a different order than ATen's);
bf16 outputs: one bf16 rounding of the fp32 result (rtol 2^-7)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ref_bn_act(x, bn, relu, residual, row_scale):
    y = bn(x.float())
    if row_scale is not None:
        y = y * row_scale[:, None]
    if residual is not None:
        y = residual.float() + y
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("c,rows", [(48, 50139), (96, 1000), (192, 12534), (384, 2868), (24, 3001), (8, 7)])
@pytest.mark.parametrize("xdt,odt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("relu,res,drop", [(True, False, False), (False, False, False), (True, True, True), (False, True, False)])
def test_bn_act_matches_torch(c, rows, xdt, odt, relu, res, drop):
    from ao_b200 import pointops

    torch.manual_seed(c + rows)
    x = (torch.randn(rows, c, device=DEV) * 1.7 + torch.linspace(-3, 5, c, device=DEV)).to(xdt).requires_grad_(True)
    residual = torch.randn(rows, c, device=DEV).to(odt).requires_grad_(True) if res else None
    row_scale = (torch.rand(rows, device=DEV) < 0.7).float() / 0.7 if drop else None
    bn_a, bn_b = nn.BatchNorm1d(c).to(DEV), nn.BatchNorm1d(c).to(DEV)
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-0.5, 0.5)
    bn_b.load_state_dict(bn_a.state_dict())
    out = pointops.bn_act(x, bn_a, relu=relu, residual=residual, row_scale=row_scale, out_dtype=odt)
    assert out.dtype == odt and out.shape == x.shape
    xr = x.detach().clone().requires_grad_(True)
    rr = residual.detach().clone().requires_grad_(True) if res else None
    ref = _ref_bn_act(xr, bn_b, relu, rr, row_scale)
    tol = dict(rtol=2e-5, atol=2e-5) if odt == torch.float32 else dict(rtol=2 ** -7, atol=2 ** -7)
    torch.testing.assert_close(out.float(), ref, **tol)
    torch.testing.assert_close(bn_a.running_mean, bn_b.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn_a.running_var, bn_b.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn_a.num_batches_tracked) == 1
    # backward: the mask of the reference is taken from the fused output so that a value rounded to / away from 0 in
    # bf16 does not count as a mismatch
    g = torch.randn(rows, c, device=DEV).to(odt)
    out.backward(g)
    if relu:
        gm = g.float() * (out.detach().float() > 0)
        pre = _ref_bn_act(xr, bn_b, False, rr, row_scale)
        pre.backward(gm)
    else:
        ref.backward(g.float())
    gtol = dict(rtol=2e-4, atol=2e-4) if xdt == torch.float32 else dict(rtol=2 ** -6, atol=2 ** -6)
    scale = max(1.0, float(xr.grad.abs().max()))
    torch.testing.assert_close(x.grad.float() / scale, xr.grad.float() / scale, **gtol)
    ptol = dict(rtol=5e-4, atol=5e-4 * max(1.0, rows ** 0.5))
    torch.testing.assert_close(bn_a.weight.grad, bn_b.weight.grad, **ptol)
    torch.testing.assert_close(bn_a.bias.grad, bn_b.bias.grad, **ptol)
    if res:
        rtol_ = dict(rtol=1e-6, atol=1e-6) if odt == torch.float32 else dict(rtol=2 ** -7, atol=2 ** -7)
        torch.testing.assert_close(residual.grad.float(), rr.grad.float(), **rtol_)


def test_bn_act_three_dims_eval_mode_and_repeatability():
    from ao_b200 import pointops

    torch.manual_seed(3)
    x = torch.randn(4000, 16, 48, device=DEV)
    bn = nn.BatchNorm1d(48).to(DEV)
    a = pointops.bn_act(x, bn, relu=True)
    bn2 = nn.BatchNorm1d(48).to(DEV)
    b = pointops.bn_act(x, bn2, relu=True)
    assert torch.equal(a, b)                                     # no atomics: bitwise repeatable
    ref = torch.relu(nn.BatchNorm1d(48).to(DEV)(x.reshape(-1, 48))).view_as(x)
    torch.testing.assert_close(a, ref, rtol=2e-5, atol=2e-5)
    bn.eval()                                                    # running statistics: the torch module runs
    e = pointops.bn_act(x, bn, relu=True)
    torch.testing.assert_close(e, torch.relu(bn(x.reshape(-1, 48))).view_as(x))
    with pytest.raises(ValueError):
        pointops.we_tail(x[..., :5].contiguous(), None, None, bn, nn.Linear(5, 5).to(DEV))


@pytest.mark.parametrize("g,n", [(6, 50139), (12, 12534), (6, 3), (12, 1000)])
@pytest.mark.parametrize("with_upe", [True, False])
def test_we_tail_matches_torch(g, n, with_upe):
    from ao_b200 import pointops

    torch.manual_seed(g + n)
    k = 16
    rel = (torch.randn(n, k, g, device=DEV) * 2.0 + 0.3).requires_grad_(True)
    upe = torch.randn(n, k, g, device=DEV).requires_grad_(True) if with_upe else None
    cst = torch.randn(g, device=DEV).requires_grad_(True) if with_upe else None
    bn_a, lin_a = nn.BatchNorm1d(g).to(DEV), nn.Linear(g, g).to(DEV)
    bn_b, lin_b = nn.BatchNorm1d(g).to(DEV), nn.Linear(g, g).to(DEV)
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-0.5, 0.5)
    bn_b.load_state_dict(bn_a.state_dict())
    lin_b.load_state_dict(lin_a.state_dict())
    assert pointops.we_tail_usable(rel, bn_a)
    out = pointops.we_tail(rel, upe, cst, bn_a, lin_a)
    rel_r = rel.detach().clone().requires_grad_(True)
    upe_r = upe.detach().clone().requires_grad_(True) if with_upe else None
    cst_r = cst.detach().clone().requires_grad_(True) if with_upe else None
    u = rel_r if not with_upe else rel_r + upe_r + cst_r
    ref = lin_b(torch.relu(bn_b(u.reshape(-1, g)))).view(n, k, g)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(bn_a.running_mean, bn_b.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn_a.running_var, bn_b.running_var, rtol=1e-5, atol=1e-6)
    gl = torch.randn(n, k, g, device=DEV)
    out.backward(gl)
    ref.backward(gl)
    torch.testing.assert_close(rel.grad, rel_r.grad, rtol=2e-4, atol=2e-5)
    rows = n * k
    ptol = dict(rtol=5e-4, atol=5e-4 * max(1.0, rows ** 0.5))
    for a, b in ((bn_a.weight, bn_b.weight), (bn_a.bias, bn_b.bias), (lin_a.weight, lin_b.weight), (lin_a.bias, lin_b.bias)):
        torch.testing.assert_close(a.grad, b.grad, **ptol)
    if with_upe:
        torch.testing.assert_close(upe.grad, upe_r.grad, rtol=2e-4, atol=2e-5)
        assert float(cst.grad.abs().max()) == 0.0                 # in front of a training-mode BatchNorm
        assert float(cst_r.grad.abs().max()) < 1e-2 * max(1.0, rows ** 0.5)
    # repeatable bit for bit
    bn_c = nn.BatchNorm1d(g).to(DEV)
    bn_c.load_state_dict(bn_b.state_dict())
    bn_c.running_mean.zero_(); bn_c.running_var.fill_(1.0)
    out2 = pointops.we_tail(rel.detach(), None if upe is None else upe.detach(), None if cst is None else cst.detach(), bn_c, lin_a)
    assert torch.equal(out2, out.detach())


def test_block_with_fused_dense_equals_torch_modules(monkeypatch):
    """One PTv2 Block + GridPool stage, fp32, training mode: the bn_act / we_tail routing against the same modules
    with AOPT_FUSED_DENSE=0 (plain torch.nn BatchNorm / ReLU)."""
    from ao_b200 import ptv2, scenes

    coord_np, feat_np, off_np = scenes.small_batch(seed=5, sizes=(3000, 2500))
    coord, offset = torch.from_numpy(coord_np).to(DEV), torch.from_numpy(off_np).to(DEV).int()
    torch.manual_seed(0)
    feat = torch.randn(coord.shape[0], 48, device=DEV)

    def run(flag):
        monkeypatch.setenv("AOPT_FUSED_DENSE", flag)
        monkeypatch.setenv("AOPT_FUSED_PE", "1")                      # the relation-free schedule (uses we_tail)
        torch.manual_seed(1)
        seq = ptv2.BlockSequence(depth=2, embed_channels=48, groups=6, neighbours=16, drop_path_rate=0.0).to(DEV).train()
        pool = ptv2.GridPool(48, 96, 0.2).to(DEV).train()
        f = feat.clone().requires_grad_(True)
        pts = seq([coord, f, offset])
        (c2, f2, o2), _ = pool(pts)
        loss = (f2 * torch.linspace(-1, 1, 96, device=DEV)).sum() + pts[1].square().mean()
        loss.backward()
        grads = {n: p.grad.clone() for n, p in list(seq.named_parameters()) + list(pool.named_parameters())}
        return pts[1].detach(), f2.detach(), f.grad.clone(), grads

    a, b = run("1"), run("0")
    torch.testing.assert_close(a[0], b[0], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(a[1], b[1], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(a[2], b[2], rtol=2e-2, atol=2e-3)
    gmax = max(float(g.abs().max()) for g in b[3].values())
    for n in a[3]:
        # gradients that are zero up to rounding (a bias in front of a BatchNorm) are compared on the scale of the others
        ga, gb = a[3][n], b[3][n]
        scale = max(1e-2 * gmax, float(gb.abs().max()))
        assert float((ga - gb).abs().max()) / scale < 3e-2, n


def test_linear_bias_in_front_of_batchnorm_is_folded(monkeypatch):
    """ptv2.run_seq leaves the bias of a Linear -> PointBatchNorm -> ReLU triple out of the GEMM (training mode): same output,
    same running statistics (the running mean sees the bias), zero bias gradient."""
    from ao_b200 import ptv2

    torch.manual_seed(7)
    seq = nn.Sequential(nn.Linear(48, 96), ptv2.PointBatchNorm(96), nn.ReLU(inplace=True)).to(DEV).train()
    with torch.no_grad():
        seq[0].bias.uniform_(-2, 2)
    ref = nn.Sequential(nn.Linear(48, 96), ptv2.PointBatchNorm(96), nn.ReLU(inplace=True)).to(DEV).train()
    ref.load_state_dict(seq.state_dict())
    x = torch.randn(5000, 48, device=DEV)
    out = ptv2.run_seq(seq, x)
    want = ref(x)
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(seq[1].norm.running_mean, ref[1].norm.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(seq[1].norm.running_var, ref[1].norm.running_var, rtol=1e-4, atol=1e-5)
    out.square().sum().backward()
    want.square().sum().backward()
    assert float(seq[0].bias.grad.abs().max()) == 0.0
    assert float(ref[0].bias.grad.abs().max()) < 1e-1                 # rounding noise around the exact zero
    torch.testing.assert_close(seq[0].weight.grad, ref[0].weight.grad, rtol=1e-3, atol=1e-2)
    seq.eval()                                                        # evaluation: the bias is back in the GEMM
    ref.eval()
    torch.testing.assert_close(ptv2.run_seq(seq, x), ref(x), rtol=1e-5, atol=1e-5)


def test_cached_weight_linear_equals_autocast_linear():
    """pointops.linear under bf16 autocast = nn.Linear under bf16 autocast (same operand rounding, fp32 accumulation);
    the cached low-precision weight follows in-place parameter updates."""
    from ao_b200 import pointops

    torch.manual_seed(11)
    lin = nn.Linear(96, 48).to(DEV)
    ref = nn.Linear(96, 48).to(DEV)
    ref.load_state_dict(lin.state_dict())
    for shape, f32 in (((7000, 96), False), ((500, 16, 96), True)):
        x = torch.randn(*shape, device=DEV, requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = pointops.linear(x, lin.weight, lin.bias, out_f32=f32)
            yr = ref(xr)
        assert y.dtype == (torch.float32 if f32 else torch.bfloat16) and y.shape == yr.shape
        torch.testing.assert_close(y.float(), yr.float(), rtol=2 ** -7, atol=2 ** -7)
        g = torch.randn_like(yr)
        y.backward(g.to(y.dtype))
        yr.backward(g)
        torch.testing.assert_close(x.grad, xr.grad, rtol=2 ** -6, atol=2 ** -6)
        torch.testing.assert_close(lin.weight.grad, ref.weight.grad, rtol=2e-2, atol=2e-2 * float(ref.weight.grad.abs().max()))
        torch.testing.assert_close(lin.bias.grad, ref.bias.grad, rtol=2e-2, atol=2e-2 * float(ref.bias.grad.abs().max()))
        assert lin.weight.grad.dtype == torch.float32
        lin.zero_grad(); ref.zero_grad()
    with torch.no_grad():                                        # an optimizer step: in-place update bumps the version
        lin.weight.mul_(2.0)
        ref.weight.mul_(2.0)
    x = torch.randn(100, 96, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        torch.testing.assert_close(pointops.linear(x, lin.weight, lin.bias).float(), ref(x).float(), rtol=2 ** -7, atol=2 ** -7)
    # no autocast: plain fp32 F.linear
    torch.testing.assert_close(pointops.linear(x, lin.weight, lin.bias), ref(x))


@pytest.mark.parametrize("autocast", [False, True])
def test_qkv_one_gemm_equals_three_layers(autocast):
    """pointops.qkv_bn (one (N,C)x(C,3C) product, BatchNorm stages on column blocks) against linear_q / linear_k /
    linear_v evaluated one by one with torch modules."""
    from ao_b200 import pointops, ptv2

    torch.manual_seed(21)
    c, n = 96, 6000
    gva = ptv2.GroupedVectorAttention(c, 12).to(DEV).train()
    ref = ptv2.GroupedVectorAttention(c, 12).to(DEV).train()
    ref.load_state_dict(gva.state_dict())
    x = torch.randn(n, c, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    assert pointops.qkv_usable(x, gva.linear_q, gva.linear_k, gva.linear_v)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        q, k, v = pointops.qkv_bn(x, gva.linear_q, gva.linear_k, gva.linear_v)
        qr, kr, vr = ref.linear_q(xr), ref.linear_k(xr), ref.linear_v(xr)
    assert v.dtype == torch.float32 and v.is_contiguous() and q.dtype == qr.dtype
    tol = dict(rtol=2 ** -6, atol=2 ** -6) if autocast else dict(rtol=1e-4, atol=1e-4)
    for a, b in ((q, qr), (k, kr), (v, vr)):
        torch.testing.assert_close(a.float(), b.float(), **tol)
    for name in ("running_mean", "running_var"):
        for s in ("linear_q", "linear_k"):
            torch.testing.assert_close(getattr(getattr(gva, s)[1].norm, name), getattr(getattr(ref, s)[1].norm, name),
                                       rtol=1e-2 if autocast else 1e-4, atol=1e-2 if autocast else 1e-5)
    wts = [torch.randn(n, c, device=DEV) for _ in range(3)]
    (q.float() * wts[0] + k.float() * wts[1] + v * wts[2]).sum().backward()
    (qr.float() * wts[0] + kr.float() * wts[1] + vr.float() * wts[2]).sum().backward()
    if autocast:
        # bf16: a pre-activation that rounds to the other side of zero flips a ReLU mask in one of the two
        # implementations — isolated elements differ, so the input gradient is compared in the L2 norm
        assert float((x.grad - xr.grad).norm() / xr.grad.norm()) < 2e-2
    else:
        torch.testing.assert_close(x.grad, xr.grad, rtol=1e-3, atol=1e-3 * float(xr.grad.abs().max()))
    for s in ("linear_q", "linear_k"):
        wa, wb = getattr(gva, s)[0].weight.grad, getattr(ref, s)[0].weight.grad
        torch.testing.assert_close(wa, wb, rtol=5e-2 if autocast else 1e-3, atol=(5e-2 if autocast else 1e-3) * float(wb.abs().max()))
        ga, gb = getattr(gva, s)[1].norm.weight.grad, getattr(ref, s)[1].norm.weight.grad
        torch.testing.assert_close(ga, gb, rtol=5e-2 if autocast else 1e-3, atol=(5e-2 if autocast else 1e-3) * float(gb.abs().max()))
        assert float(getattr(gva, s)[0].bias.grad.abs().max()) == 0.0
    torch.testing.assert_close(gva.linear_v.weight.grad, ref.linear_v.weight.grad, rtol=5e-2 if autocast else 1e-3,
                               atol=(5e-2 if autocast else 1e-3) * float(ref.linear_v.weight.grad.abs().max()))
    torch.testing.assert_close(gva.linear_v.bias.grad, ref.linear_v.bias.grad, rtol=2e-2, atol=2e-2 * float(ref.linear_v.bias.grad.abs().max()))


@pytest.mark.parametrize("g,c,rows", [(6, 48, 320000), (12, 96, 50139), (6, 96, 20000)])
@pytest.mark.parametrize("gdt,xdt", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)])
def test_skinny_wgrad_col_sum_copy_cols(g, c, rows, gdt, xdt):
    """The three small helpers of the Linear backward passes against torch in fp64 / exact copies."""
    from ao_b200 import _lib
    from ao_b200.pointops import dense

    torch.manual_seed(g * c)
    grad = torch.randn(rows, g, device=DEV).to(gdt)
    x = (torch.randn(rows, c, device=DEV) + 0.5).to(xdt)
    got = dense._weight_grad(grad, x)
    want = (grad.double().t() @ x.double()).float()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))
    assert torch.equal(got, dense._weight_grad(grad, x))                       # fixed summation order
    s = dense.col_sum(x)
    torch.testing.assert_close(s, x.double().sum(0).float(), rtol=1e-4, atol=1e-4 * float(x.double().sum(0).abs().max()))
    # column block of a (rows, 3c) matrix -> dense fp32 (+ bias) and back
    wide = torch.randn(rows, 3 * c, device=DEV).to(xdt)
    bias = torch.randn(c, device=DEV)
    dst = torch.empty(rows, c, device=DEV)
    lib = _lib.load()
    dt = {torch.float32: 0, torch.bfloat16: 1}
    _lib.check(lib.aopt_copy_cols(rows, c, wide.data_ptr() + 2 * c * wide.element_size(), 3 * c, dt[xdt], bias.data_ptr(),
                                  dst.data_ptr(), c, 0, _lib.stream()), "copy_cols")
    torch.testing.assert_close(dst, wide[:, 2 * c:].float() + bias, rtol=0, atol=0)
    back = torch.zeros_like(wide)
    _lib.check(lib.aopt_copy_cols(rows, c, dst.data_ptr(), c, 0, 0, back.data_ptr() + 2 * c * back.element_size(), 3 * c, dt[xdt],
                                  _lib.stream()), "copy_cols")
    torch.testing.assert_close(back[:, 2 * c:], dst.to(xdt), rtol=0, atol=0)
    assert float(back[:, :2 * c].abs().max()) == 0.0


@pytest.mark.parametrize("g,c,rows", [(6, 48, 320000), (12, 96, 50139), (13, 48, 100000), (20, 48, 30000)])
@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float32])
def test_skinny_linear_forward_backward(g, c, rows, xdt):
    """pointops.linear(x, w, out_f32=True) with 6 / 12 outputs over many rows (own forward / dgrad / wgrad kernels)
    against torch in fp64."""
    from ao_b200 import pointops

    torch.manual_seed(g + c)
    x = torch.randn(rows, c, device=DEV).to(xdt).requires_grad_(True)
    w = (torch.randn(g, c, device=DEV) * 0.2).requires_grad_(True)
    b = torch.randn(g, device=DEV, requires_grad=True) if g in (13, 20) else None      # the segmentation head has a bias
    y = pointops.linear(x, w, b, out_f32=True)
    assert y.dtype == torch.float32 and y.shape == (rows, g)
    want = x.detach().double() @ w.detach().double().t()
    if b is not None:
        want = want + b.detach().double()
    torch.testing.assert_close(y.double(), want, rtol=1e-5, atol=1e-5)
    gy = torch.randn(rows, g, device=DEV)
    y.backward(gy)
    wx = gy.double() @ w.detach().double()
    tol = dict(rtol=1e-5, atol=1e-5) if xdt == torch.float32 else dict(rtol=2 ** -7, atol=2 ** -7)
    torch.testing.assert_close(x.grad.double(), wx, **tol)
    ww = gy.double().t() @ x.detach().double()
    torch.testing.assert_close(w.grad.double(), ww, rtol=1e-4, atol=1e-4 * float(ww.abs().max()))
    if b is not None:
        torch.testing.assert_close(b.grad.double(), gy.double().sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cin", [6, 9, 4, 3])
@pytest.mark.parametrize("autocast", [False, True])
def test_patch_projection_small_input_width(cin, autocast):
    """Linear(in_channels, 48, bias=False) -> PointBatchNorm -> ReLU on 10^5 rows of raw features (GVAPatchEmbed.proj,
    …v2m2_base.py:363-364): the product runs in the small-K kernels; against the torch modules."""
    from ao_b200 import ptv2

    torch.manual_seed(cin)
    seq = nn.Sequential(nn.Linear(cin, 48, bias=False), ptv2.PointBatchNorm(48), nn.ReLU(inplace=True)).to(DEV).train()
    ref = nn.Sequential(nn.Linear(cin, 48, bias=False), ptv2.PointBatchNorm(48), nn.ReLU(inplace=True)).to(DEV).train()
    ref.load_state_dict(seq.state_dict())
    x = torch.randn(100000, cin, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = ptv2.run_seq(seq, x)
        want = ref(x)
    tol = dict(rtol=3e-2, atol=3e-2) if autocast else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out.float(), want.float(), **tol)
    wts = torch.randn_like(want, dtype=torch.float32)
    (out.float() * wts).sum().backward()
    (want.float() * wts).sum().backward()
    ga, gb = seq[0].weight.grad, ref[0].weight.grad
    # under autocast the torch reference returns this gradient rounded to bf16 (3 significant digits)
    assert float((ga - gb).norm() / gb.norm()) < (6e-2 if autocast else 1e-3)


@pytest.mark.parametrize("g", [6, 12])
def test_we_tail_gather_mode_equals_materialised_relation(g):
    """we_tail(gather=(kp, qp, idx)) forms rel = kp[idx] - qp inside the kernels: same bits as gva_relation followed by
    we_tail on the stored tensor, forward and backward, -1 padding included."""
    from ao_b200 import pointops, scenes

    coord_np, _, off_np = scenes.small_batch(seed=g, sizes=(4000, 9, 3000))       # a scene with 9 < k points: -1 padding
    coord, offset = torch.from_numpy(coord_np).to(DEV), torch.from_numpy(off_np).to(DEV).int()
    idx, _ = pointops.knn_query(16, coord, offset)
    assert int((idx < 0).sum()) > 0
    n = coord.shape[0]
    torch.manual_seed(g)
    bn_a, lin_a = nn.BatchNorm1d(g).to(DEV), nn.Linear(g, g).to(DEV)
    bn_b, lin_b = nn.BatchNorm1d(g).to(DEV), nn.Linear(g, g).to(DEV)
    bn_b.load_state_dict(bn_a.state_dict()); lin_b.load_state_dict(lin_a.state_dict())
    kp, qp = torch.randn(n, g, device=DEV, requires_grad=True), torch.randn(n, g, device=DEV, requires_grad=True)
    upe = torch.randn(n, 16, g, device=DEV, requires_grad=True)
    cst = torch.randn(g, device=DEV)
    kp2, qp2, upe2 = (t.detach().clone().requires_grad_(True) for t in (kp, qp, upe))
    a = pointops.we_tail(None, upe, cst, bn_a, lin_a, gather=(kp, qp, idx))
    b = pointops.we_tail(pointops.gva_relation(kp2, qp2, idx), upe2, cst, bn_b, lin_b)
    assert torch.equal(a, b)
    gl = torch.randn_like(a)
    a.backward(gl); b.backward(gl)
    assert torch.equal(kp.grad, kp2.grad) and torch.equal(qp.grad, qp2.grad) and torch.equal(upe.grad, upe2.grad)
    assert torch.equal(lin_a.weight.grad, lin_b.weight.grad) and torch.equal(bn_a.weight.grad, bn_b.weight.grad)
    assert torch.equal(bn_a.running_var, bn_b.running_var)
