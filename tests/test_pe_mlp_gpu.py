"""Fused positional-bias MLP (pointops.pe_bias_mlp, csrc/pe_mlp.cu) against the torch modules it replaces
(linear_p_bias of GroupedVectorAttention: …v2m2_base.py:88-93) run in fp32.

Tolerance: the C x C layer uses bf16 operands with fp32 accumulation (the precision of the autocast path
it replaces), so values are compared at 2e-2 of the tensor's scale; everything upstream of that product
(BatchNorm statistics from the closed form, first layer, ReLU mask) is fp32/fp64 and is checked tightly
through the statistics and the running-stat update."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def make_mlp(c, seed):
    from ao_b200 import ptv2

    torch.manual_seed(seed)
    mlp = nn.Sequential(nn.Linear(3, c), ptv2.PointBatchNorm(c), nn.ReLU(inplace=True), nn.Linear(c, c)).cuda()
    with torch.no_grad():
        mlp[1].norm.weight.uniform_(0.5, 1.5)
        mlp[1].norm.bias.uniform_(-0.3, 0.3)
    return mlp


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("c", [48, 96])
@pytest.mark.parametrize("n,k", [(1501, 16), (7, 16), (4000, 8)])
def test_pe_mlp_training_matches_torch_modules(c, n, k):
    from ao_b200 import pointops

    torch.manual_seed(c + n)
    pos = (0.2 * torch.randn(n, k, 3, device="cuda"))
    pos[::7, -3:] = 0.0                                            # masked slots are exact zeros
    ref = make_mlp(c, 1).train()
    fused = copy.deepcopy(ref).train()
    y_ref = ref(pos)
    y = pointops.pe_bias_mlp(pos, fused)
    assert y.shape == (n, k, c) and y.dtype == torch.float32
    assert rel_err(y, y_ref) < 2e-2
    # running statistics: closed-form batch statistics == what BatchNorm1d measured (fp32 tight)
    bn_r, bn_f = ref[1].norm, fused[1].norm
    assert torch.allclose(bn_f.running_mean, bn_r.running_mean, rtol=1e-4, atol=1e-6)
    assert torch.allclose(bn_f.running_var, bn_r.running_var, rtol=1e-4, atol=1e-7)
    assert int(bn_f.num_batches_tracked) == 1
    g = torch.randn_like(y_ref)
    y_ref.backward(g)
    y.backward(g)
    for (name, p_r), (_, p_f) in zip(ref.named_parameters(), fused.named_parameters()):
        if name == "0.bias":                                       # analytically zero (bias before BatchNorm)
            assert p_f.grad.abs().max().item() == 0.0
            assert p_r.grad.abs().max().item() < 1e-2 * max(1.0, g.abs().sum().item() ** 0.5)
            continue
        assert rel_err(p_f.grad, p_r.grad) < 2e-2, name
    # deterministic: per-CTA partial sums in a fixed order, no atomics
    fused2 = copy.deepcopy(ref).train()
    fused2.zero_grad()
    y2 = pointops.pe_bias_mlp(pos, fused2)
    y2.backward(g)
    assert torch.equal(y2, y)
    for p_a, p_b in zip(fused.parameters(), fused2.parameters()):
        assert torch.equal(p_a.grad, p_b.grad)


@pytest.mark.parametrize("c", [48, 96])
def test_pe_mlp_eval_mode_uses_running_statistics(c):
    from ao_b200 import pointops

    torch.manual_seed(3)
    pos = 0.3 * torch.randn(900, 16, 3, device="cuda")
    ref = make_mlp(c, 2)
    with torch.no_grad():
        ref[1].norm.running_mean.uniform_(-0.2, 0.2)
        ref[1].norm.running_var.uniform_(0.5, 2.0)
    ref.eval()
    fused = copy.deepcopy(ref).eval()
    y_ref = ref(pos)
    y = pointops.pe_bias_mlp(pos, fused)
    assert rel_err(y, y_ref) < 2e-2
    assert torch.equal(fused[1].norm.running_mean, ref[1].norm.running_mean)      # untouched in eval
    g = torch.randn_like(y_ref)
    y_ref.backward(g)
    y.backward(g)
    for (name, p_r), (_, p_f) in zip(ref.named_parameters(), fused.named_parameters()):
        assert rel_err(p_f.grad, p_r.grad) < 2e-2, name


@pytest.mark.parametrize("c,ga", [(48, 6), (96, 12), (48, 16), (96, 1)])
def test_pe_mlp_aux_head_is_a_linear_on_peb(c, ga):
    """aux = h @ Wf^T with Wf = L.weight @ W2  ==  L(peb) - L.weight @ b2 - L.bias, forward and backward."""
    from ao_b200 import pointops

    torch.manual_seed(7 + ga)
    pos = 0.2 * torch.randn(1203, 16, 3, device="cuda")
    ref = make_mlp(c, 4).train()
    lin_ref = nn.Linear(c, ga).cuda()
    fused, lin_f = copy.deepcopy(ref).train(), copy.deepcopy(lin_ref)
    peb_ref = ref(pos)
    u_ref = lin_ref(peb_ref)
    wf = lin_f.weight @ fused[3].weight
    peb, aux = pointops.pe_bias_mlp(pos, fused, aux_weight=wf)
    u = aux + (lin_f.weight @ fused[3].bias + lin_f.bias)
    assert aux.shape == (1203, 16, ga)
    assert rel_err(peb, peb_ref) < 2e-2 and rel_err(u, u_ref) < 2e-2
    g_p, g_u = torch.randn_like(peb_ref), torch.randn_like(u_ref)
    (peb_ref * g_p).sum().add((u_ref * g_u).sum()).backward()
    (peb * g_p).sum().add((u * g_u).sum()).backward()
    for (name, p_r), (_, p_f) in zip(list(ref.named_parameters()) + list(lin_ref.named_parameters()),
                                     list(fused.named_parameters()) + list(lin_f.named_parameters())):
        if name == "0.bias":
            continue
        assert rel_err(p_f.grad, p_r.grad) < 2e-2, name
    # only one of the two outputs used downstream
    fused.zero_grad()
    peb2, aux2 = pointops.pe_bias_mlp(pos, fused, aux_weight=wf.detach())
    (aux2 * g_u).sum().backward()
    assert fused[3].weight.grad is not None and torch.isfinite(fused[3].weight.grad).all()


def test_pos_moments_against_float64():
    from ao_b200 import pointops

    torch.manual_seed(0)
    pos = torch.randn(50001, 16, 3, device="cuda") * 0.1 + 0.02
    m = pointops.pos_moments(pos).cpu().numpy()
    p = pos.double().view(-1, 3).cpu().numpy()
    ref = np.concatenate([p.sum(0), [(p[:, 0] * p[:, 0]).sum(), (p[:, 0] * p[:, 1]).sum(), (p[:, 0] * p[:, 2]).sum(),
                                     (p[:, 1] * p[:, 1]).sum(), (p[:, 1] * p[:, 2]).sum(), (p[:, 2] * p[:, 2]).sum()]])
    assert np.allclose(m, ref, rtol=1e-10, atol=1e-9)


def test_unsupported_width_is_an_error():
    from ao_b200 import pointops

    assert pointops.pe_mlp_supported(48) and pointops.pe_mlp_supported(96)
    assert not pointops.pe_mlp_supported(192) and not pointops.pe_mlp_supported(50)
    with pytest.raises(ValueError):
        pointops.pe_bias_mlp(torch.zeros(4, 16, 3, device="cuda"), make_mlp(192, 0))


def test_model_with_fused_pe_tracks_the_unfused_model():
    """Whole PTv2m2 in fp32 with and without the fused positional MLP (forced on for levels with C in
    {48, 96}): the only difference between the two runs is the bf16 rounding inside the fused C x C layer."""
    import os

    from ao_b200 import ptv2, scenes

    coord, feat, offset = scenes.s3dis_batch(2, n_points=5000)
    data = dict(coord=torch.from_numpy(coord).cuda(), feat=torch.from_numpy(feat).cuda(), offset=torch.from_numpy(offset).cuda())
    torch.manual_seed(0)
    model = ptv2.PointTransformerV2(**dict(ptv2.S3DIS_CFG, drop_path_rate=0.0)).cuda().train()
    target = torch.randint(0, 13, (coord.shape[0],), device="cuda")
    out = {}
    for flag in ("0", "1"):
        os.environ["AOPT_FUSED_PE"] = flag
        try:
            m = copy.deepcopy(model)
            logits = m(data)
            loss = torch.nn.functional.cross_entropy(logits.float(), target)
            loss.backward()
            out[flag] = (logits.float().detach(), loss.item(),
                         m.patch_embed.blocks.blocks[0].attn.linear_p_bias[3].weight.grad.detach().clone(),
                         m.patch_embed.blocks.blocks[0].attn.linear_p_bias[0].weight.grad.detach().clone())
        finally:
            os.environ.pop("AOPT_FUSED_PE", None)
    assert abs(out["0"][1] - out["1"][1]) < 2e-2 * max(1.0, abs(out["0"][1]))
    assert rel_err(out["1"][0], out["0"][0]) < 0.1
    for i in (2, 3):
        cos = torch.nn.functional.cosine_similarity(out["0"][i].flatten(), out["1"][i].flatten(), dim=0).item()
        assert cos > 0.9, (i, cos)
