"""Host logic of ao_b200.pointops.dense / ptv2.run_seq without a GPU: on CPU tensors (or with AOPT_FUSED_DENSE=0, or in
evaluation mode) every entry point must fall through to the torch modules it mirrors and give their results exactly
(/root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:25-45,86-99,187-197).  The CUDA
kernels themselves are covered by tests/test_dense_gpu.py."""
import pytest
import torch
import torch.nn as nn


def test_bn_act_falls_back_to_the_torch_module_on_cpu():
    from ao_b200.pointops import dense

    torch.manual_seed(0)
    x = torch.randn(50, 7, 48)
    bn_a, bn_b = nn.BatchNorm1d(48), nn.BatchNorm1d(48)
    res = torch.randn(50, 7, 48)
    scale = (torch.rand(350) < 0.7).float() / 0.7
    assert not dense.bn_act_usable(x, bn_a)
    got = dense.bn_act(x, bn_a, relu=True, residual=res, row_scale=scale)
    want = torch.relu(res + (bn_b(x.reshape(-1, 48)) * scale[:, None]).view_as(x))
    assert torch.equal(got, want)
    assert torch.equal(bn_a.running_mean, bn_b.running_mean) and int(bn_a.num_batches_tracked) == 1
    # a bias the caller left out of x comes back in on the fallback path
    bias = torch.randn(48)
    bn_c, bn_d = nn.BatchNorm1d(48), nn.BatchNorm1d(48)
    assert torch.equal(dense.bn_act(x, bn_c, pre_bias=bias), bn_d((x + bias).reshape(-1, 48)).view_as(x))


def test_run_seq_equals_sequential_on_cpu_train_and_eval():
    from ao_b200 import ptv2

    torch.manual_seed(1)
    seq = nn.Sequential(nn.Linear(6, 48), ptv2.PointBatchNorm(48), nn.ReLU(inplace=True), nn.Linear(48, 13))
    ref = nn.Sequential(nn.Linear(6, 48), ptv2.PointBatchNorm(48), nn.ReLU(inplace=True), nn.Linear(48, 13))
    ref.load_state_dict(seq.state_dict())
    x = torch.randn(200, 6)
    for mode in (True, False):
        seq.train(mode); ref.train(mode)
        torch.testing.assert_close(ptv2.run_seq(seq, x, torch.float32), ref(x))
        h = ref[:3](x)                                          # (in training mode every call updates the running statistics:
        torch.testing.assert_close(ptv2.run_seq(seq, x, stop=3), h)   # both copies see the same number of batches)
        torch.testing.assert_close(ptv2.run_seq(seq, h, start=3), ref[3](h))
    assert torch.equal(seq[1].norm.running_mean, ref[1].norm.running_mean)


def test_linear_and_qkv_guards_on_cpu():
    from ao_b200 import ptv2
    from ao_b200.pointops import dense

    torch.manual_seed(2)
    lin = nn.Linear(48, 6)
    x = torch.randn(100, 48)
    assert torch.equal(dense.linear(x, lin.weight, lin.bias), nn.functional.linear(x, lin.weight, lin.bias))
    assert dense.linear(x, lin.weight, out_f32=True).dtype == torch.float32
    gva = ptv2.GroupedVectorAttention(48, 6)
    assert not dense.qkv_usable(x, gva.linear_q, gva.linear_k, gva.linear_v)          # CPU tensors: torch modules
    assert not dense.we_tail_usable(torch.randn(10, 16, 6), gva.weight_encoding[1])
    with pytest.raises(ValueError):
        dense.we_tail(torch.randn(10, 16, 6), None, None, gva.weight_encoding[1], gva.weight_encoding[3])


def test_fused_dense_switch_and_thresholds(monkeypatch):
    from ao_b200 import ptv2
    from ao_b200.pointops import dense

    monkeypatch.setenv("AOPT_FUSED_DENSE", "0")
    assert not dense.fused_dense_enabled()
    monkeypatch.delenv("AOPT_FUSED_DENSE")
    assert dense.fused_dense_enabled()
    assert ptv2.relation_free_min_elems() == 64e6
    monkeypatch.setenv("AOPT_RELFREE_ALL", "1")
    assert ptv2.relation_free_min_elems() == 0.0
    monkeypatch.delenv("AOPT_RELFREE_ALL")
    monkeypatch.setenv("AOPT_RELFREE_MIN_ELEMS", "1e6")
    assert ptv2.relation_free_min_elems() == 1e6
    assert not ptv2.we_gather_enabled()


def test_vote_accumulate_rejects_what_the_kernel_cannot_take():
    """pointops.vote_accumulate (tester fragment vote, pointcept/engines/test.py:106-113) has no CPU path: host tensors and
    malformed arguments are refused before anything is launched."""
    from ao_b200.pointops import vote_accumulate

    pred, logits, index = torch.zeros(10, 13), torch.randn(4, 13), torch.arange(4)
    with pytest.raises(ValueError, match="CUDA"):
        vote_accumulate(pred, logits, index)
    with pytest.raises(ValueError, match="classes"):
        vote_accumulate(pred, torch.randn(4, 12), index)
    with pytest.raises(ValueError, match="one entry per logits row"):
        vote_accumulate(pred, logits, torch.arange(3))
    with pytest.raises(ValueError):
        vote_accumulate(pred.double(), logits, index)
