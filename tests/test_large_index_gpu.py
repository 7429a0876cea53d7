"""Tensors past 2^31 elements (the reference kernels index with 32-bit ints,
grouping_cuda_kernel.cu:7-12, and overflow at e.g. 640k points x 16 x 256 channels).  Every (N,k,C) kernel here is
run at N*k*C = 2.17e9 elements and checked on sampled rows (including rows past the 2^31-th element) against
torch indexing."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N, K, C, G = 530_000, 16, 256, 32          # N*K*C = 2_170_880_000 > 2**31


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 1e9


@pytest.fixture(scope="module")
def big():
    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of free device memory")
    assert N * K * C > 2 ** 31
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, N, (N, K), device="cuda", generator=g, dtype=torch.int32)
    idx[::1000, -1] = -1                                              # some padded slots
    rows = torch.cat([torch.arange(0, 64, device="cuda"), torch.randint(0, N, (256,), device="cuda", generator=g),
                      torch.arange(N - 64, N, device="cuda")])       # the last rows lie past element 2^31
    return dict(idx=idx, rows=rows, gen=g)


def test_gather_sub_and_relation_backward_past_2_31_elements(big):
    from ao_b200 import pointops

    idx, rows, g = big["idx"], big["rows"], big["gen"]
    key = torch.randn(N, C, device="cuda", generator=g, requires_grad=True)
    query = torch.randn(N, C, device="cuda", generator=g, requires_grad=True)
    rel = pointops.gva_relation(key, query, idx)
    assert rel.numel() > 2 ** 31
    j = idx[rows].long()
    ref = torch.where((j >= 0)[..., None], key.detach()[j.clamp(min=0)], torch.zeros((), device="cuda")) - query.detach()[rows][:, None]
    assert torch.equal(rel.detach()[rows], ref)
    grad = torch.randn(N, K, C, device="cuda", generator=g)
    gk, gq = torch.autograd.grad(rel, [key, query], grad)
    del rel
    assert torch.allclose(gq[rows], -grad[rows].sum(1), rtol=1e-5, atol=1e-4)
    # grad_key of a few sources: sum of the gradient rows that gathered them
    for src in (0, int(idx[N - 1, 0]), N - 1):
        hit = (idx == src).nonzero()
        want = grad[hit[:, 0], hit[:, 1]].double().sum(0)
        assert torch.allclose(gk[src].double(), want, rtol=1e-5, atol=1e-4)


def test_gva_aggregate_past_2_31_elements(big):
    from ao_b200 import pointops

    idx, rows, g = big["idx"], big["rows"], big["gen"]
    value = torch.randn(N, C, device="cuda", generator=g, requires_grad=True)
    peb = torch.randn(N, K, C, device="cuda", generator=g, requires_grad=True)
    logits = torch.randn(N, K, G, device="cuda", generator=g, requires_grad=True)
    out = pointops.gva_aggregate(value, peb, logits, idx, G)
    j = idx[rows].long()
    v = torch.where((j >= 0)[..., None], value.detach()[j.clamp(min=0)], torch.zeros((), device="cuda")) + peb.detach()[rows]
    w = torch.softmax(logits.detach()[rows], dim=1) * (j >= 0)[..., None]
    ref = torch.einsum("nsgi,nsg->ngi", v.view(-1, K, G, C // G), w).reshape(-1, C)
    assert torch.allclose(out.detach()[rows], ref, rtol=1e-5, atol=2e-5)
    go = torch.randn(N, C, device="cuda", generator=g)
    gv, gp, gl = torch.autograd.grad(out, [value, peb, logits], go)
    assert gp.numel() > 2 ** 31
    ref_gp = go[rows][:, None, :] * w.repeat_interleave(C // G, dim=2)
    assert torch.allclose(gp[rows], ref_gp, rtol=1e-5, atol=2e-5)
    assert torch.isfinite(gv[rows]).all() and torch.isfinite(gl[rows]).all()
