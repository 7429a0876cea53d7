"""Parity at BASELINE.json's FULL sizes (configs[1]: 4 rooms x 80k points, k=16, C=48, G=6) through
size-independent properties — the oracle cannot finish these sizes in seconds, so each check is an
identity the operator must satisfy exactly or to fp32 tolerance:
  kNN        self-neighbour first with dist 0, rows ascending, ids inside the query's scene, TILE == GRID
             on a query slab, oracle bit-equality on sampled rows
  gather     adjoint identity  <gather(x), y> == <x, scatter(y)>  and conservation  sum(scatter(y)) == sum(y)
  GVA        linearity in value/peb, rows of prob sum to 1, masked slots inert; backward adjoint vs autograd
             of the torch restatement on a row subset
  GridPool   max is idempotent under point permutation inside a scene, argmax rows carry the max,
             cluster ↔ idx_ptr consistent, checksum of per-voxel counts == N
  interp     weights sum to 1, constant features are reproduced, adjoint identity
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K, C, G = 16, 48, 6


@pytest.fixture(scope="module")
def batch():
    from ao_b200 import pointops, scenes

    coord, feat, offset = scenes.s3dis_batch(4, 80000)
    xyz = torch.from_numpy(coord).cuda()
    off = torch.from_numpy(offset).cuda()
    idx, d2 = pointops.knn_query_raw(K, xyz, off)
    return dict(coord=coord, offset=offset, xyz=xyz, off=off, idx=idx, d2=d2)


def test_knn_full_size_properties(batch, oracle):
    from ao_b200 import pointops

    xyz, off, idx, d2 = batch["xyz"], batch["off"], batch["idx"], batch["d2"]
    n = xyz.shape[0]
    assert idx.shape == (n, K) and idx.dtype == torch.int32
    assert torch.equal(idx[:, 0].long(), torch.arange(n, device="cuda"))           # tie-free scene: self first
    assert bool((d2[:, 0] == 0).all()) and bool((d2[:, 1:] >= d2[:, :-1]).all())  # ascending
    lo = torch.repeat_interleave(torch.cat([off.new_zeros(1), off[:-1]]).long(), torch.diff(off, prepend=off.new_zeros(1)).long())
    hi = torch.repeat_interleave(off.long(), torch.diff(off, prepend=off.new_zeros(1)).long())
    assert bool(((idx.long() >= lo[:, None]) & (idx.long() < hi[:, None])).all())   # neighbours stay in the scene
    # exhaustive TILE scan == GRID on a slab of queries spanning a scene boundary
    sl = slice(79000, 81000)
    q = xyz[sl].contiguous()
    qoff = torch.tensor([1000, 2000, 2000, 2000], dtype=torch.int32, device="cuda")
    ti, td = pointops.knn_query_raw(K, xyz, off, q, qoff, method="tile")
    assert torch.equal(ti, idx[sl]) and torch.equal(td.view(torch.int32), d2[sl].view(torch.int32))
    # C oracle on sampled rows (bit-exact)
    for b, e in ((0, 64), (159990, 160010), (319936, 320000)):
        ri, rd = oracle.knn_query(K, batch["coord"], batch["offset"], rule="lex", rows=(b, e))
        assert np.array_equal(idx[b:e].cpu().numpy(), ri)
        assert np.array_equal(d2[b:e].cpu().numpy().view(np.uint32), rd.view(np.uint32))


def test_gather_scatter_adjoint_full_size(batch):
    from ao_b200 import pointops

    idx = batch["idx"]
    n = idx.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    key = torch.randn(n, C, device="cuda", generator=g, requires_grad=True)
    query = torch.randn(n, C, device="cuda", generator=g, requires_grad=True)
    y = torch.randn(n, K, C, device="cuda", generator=g)
    rel = pointops.gva_relation(key, query, idx)
    # spot-check the forward on random rows
    rows = torch.randint(0, n, (4096,), device="cuda")
    assert torch.equal(rel[rows], key.detach()[idx[rows].long()] - query.detach()[rows][:, None])
    gk, gq = torch.autograd.grad(rel, [key, query], y)
    # adjoint identity in float64 accumulations
    lhs = (rel.detach().double() * y.double()).sum()
    rhs = (key.detach().double() * gk.double()).sum() + (query.detach().double() * gq.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * max(1.0, abs(lhs.item())) + 1e-2
    # conservation: every (query, slot) gradient lands on exactly one source row
    assert torch.allclose(gk.double().sum(0), y.double().sum((0, 1)), rtol=1e-6, atol=1e-2)
    assert torch.allclose(gq.double().sum(0), -y.double().sum((0, 1)), rtol=1e-6, atol=1e-2)
    # deterministic (no atomics)
    gk2, _ = torch.autograd.grad(pointops.gva_relation(key, query, idx), [key, query], y)
    assert torch.equal(gk, gk2)


def test_gva_aggregate_full_size(batch, oracle):
    from ao_b200 import pointops

    idx = batch["idx"].clone()
    n = idx.shape[0]
    idx[::1000, 11:] = -1                                          # padded slots
    g = torch.Generator(device="cuda").manual_seed(2)
    value = torch.randn(n, C, device="cuda", generator=g, requires_grad=True)
    peb = torch.randn(n, K, C, device="cuda", generator=g, requires_grad=True)
    logits = 2 * torch.randn(n, K, G, device="cuda", generator=g)
    logits.requires_grad_(True)
    out = pointops.gva_aggregate(value, peb, logits, idx, G)
    # linearity in (value, peb)
    out2 = pointops.gva_aggregate(2 * value.detach(), 2 * peb.detach(), logits.detach(), idx, G)
    assert torch.allclose(out2, 2 * out.detach(), rtol=1e-6, atol=1e-6)
    # constant value and zero peb: out = value * (sum of unmasked probabilities) ≤ value
    ones = torch.ones(n, C, device="cuda")
    o1 = pointops.gva_aggregate(ones, None, logits.detach(), idx, G)
    full = (idx >= 0).all(1)
    assert torch.allclose(o1[full], torch.ones_like(o1[full]), rtol=1e-5, atol=1e-5)
    assert bool((o1[~full] < 1).all())
    # row subset against the torch restatement, forward and backward
    rows = torch.cat([torch.arange(0, n, 997, device="cuda"), torch.arange(0, n, 1000, device="cuda")]).unique()
    v_c, p_c, l_c = value.detach().clone().requires_grad_(True), peb.detach()[rows].clone().requires_grad_(True), \
        logits.detach()[rows].clone().requires_grad_(True)
    ref = oracle.gva_aggregate(v_c, p_c, l_c, idx[rows], G)      # torch ops on the GPU tensors
    assert torch.allclose(out[rows], ref, rtol=1e-5, atol=2e-5)
    go = torch.randn(n, C, device="cuda", generator=g)
    gv, gp, gl = torch.autograd.grad(out, [value, peb, logits], go)
    rp, rl = torch.autograd.grad(ref, [p_c, l_c], go[rows])
    assert torch.allclose(gp[rows], rp, rtol=1e-5, atol=2e-5)
    assert torch.allclose(gl[rows], rl, rtol=1e-4, atol=5e-5)
    # grad_value: adjoint identity over the whole batch (out is linear in value)
    lhs = (pointops.gva_aggregate(value.detach(), None, logits.detach(), idx, G).double() * go.double()).sum()
    rhs = (value.detach().double() * gv.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * abs(lhs.item()) + 1e-2


def test_grid_pool_and_interpolation_full_size(batch):
    from ao_b200 import pointops

    xyz, off = batch["xyz"], batch["off"]
    n = xyz.shape[0]
    g = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.relu(torch.randn(n, 96, device="cuda", generator=g)).requires_grad_(True)
    (nc, nf, noff), cluster, part = pointops.grid_pool(xyz, feat, off, 0.1, return_partition=True)
    nv = nc.shape[0]
    counts = torch.diff(part.idx_ptr.long())
    assert int(counts.sum()) == n and int(counts.min()) >= 1 and int(noff[-1]) == nv
    assert torch.equal(torch.bincount(cluster, minlength=nv), counts)          # cluster ↔ idx_ptr
    assert bool((cluster[part.order.long()][1:] >= cluster[part.order.long()][:-1]).all())
    # max / mean against torch scatter ops
    ref_max = torch.full((nv, 96), -1.0, device="cuda").scatter_reduce(0, cluster[:, None].expand(-1, 96), feat.detach(), "amax")
    assert torch.equal(nf.detach(), ref_max)
    ref_mean = torch.zeros(nv, 3, device="cuda", dtype=torch.float64).index_add_(0, cluster, xyz.double()) / counts[:, None]
    assert torch.allclose(nc.double(), ref_mean, rtol=0, atol=1e-5)
    # voxels are ordered scene-major and every voxel's points share its scene
    scene_of_pt = torch.bucketize(torch.arange(n, device="cuda"), off.long(), right=True)
    scene_of_vox = torch.bucketize(torch.arange(nv, device="cuda"), noff.long(), right=True)
    assert torch.equal(scene_of_vox[cluster], scene_of_pt)
    # backward: gradient goes to exactly one point per (voxel, channel) and is conserved
    go = torch.randn(nv, 96, device="cuda", generator=g)
    (gf,) = torch.autograd.grad(nf, [feat], go)
    assert torch.allclose(gf.double().sum(0), go.double().sum(0), rtol=1e-6, atol=1e-3)
    assert int((gf != 0).sum()) <= nv * 96
    # interpolation coarse → fine
    src = torch.randn(nv, C, device="cuda", generator=g, requires_grad=True)
    up = pointops.interpolation(nc, xyz, src, noff.int(), off)
    const = pointops.interpolation(nc, xyz, torch.ones(nv, 8, device="cuda"), noff.int(), off)
    assert torch.allclose(const, torch.ones_like(const), rtol=1e-5, atol=1e-5)   # weights sum to 1
    y = torch.randn(n, C, device="cuda", generator=g)
    (gs,) = torch.autograd.grad(up, [src], y)
    lhs, rhs = (up.detach().double() * y.double()).sum(), (src.detach().double() * gs.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * abs(lhs.item()) + 1e-2


@pytest.mark.parametrize("k", [16, 3, 8])
def test_group_xyz_matches_reference_grouping(oracle, k):
    from ao_b200 import pointops, scenes

    coord, feat, offset = scenes.small_batch(31, sizes=(600, 2, 900))
    xyz, off = torch.from_numpy(coord).cuda(), torch.from_numpy(offset).cuda()
    idx, _ = pointops.knn_query_raw(k, xyz, off)                  # scene of 2 points → -1 padding
    pos = pointops.group_xyz(idx, xyz)
    ref = oracle.grouping(idx.cpu(), torch.from_numpy(feat), torch.from_numpy(coord), with_xyz=True)[:, :, :3]
    assert torch.equal(pos.cpu(), ref)                              # == treats -0.0 and 0.0 alike
