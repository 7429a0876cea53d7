"""GPU parity of the gather / scatter / aggregation / pooling / interpolation kernels (through the
C ABI via ao_b200.pointops) against the torch restatements in oracle/torch_ref.py, the reference CUDA
launchers (oracle/_ref) and the golden vectors made by the reference Python (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import to_cuda

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")

# fp32 tolerances (stated per the north star): forward values are sums of <= k (<= 32) fp32 products,
# backward sums run over the in-degree of a point (tens of terms, different order than torch).
RTOL, ATOL = 1e-5, 1e-5


def gold(name):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def rand_idx(rng, m, k, n, pad_rows=5):
    idx = rng.integers(0, n, (m, k)).astype(np.int32)
    for r in rng.integers(0, m, pad_rows):
        idx[r, rng.integers(0, k):] = -1
    return idx


# ------------------------------------------------------------------------------------------ CSR
@pytest.fixture(params=["sort", "count"])
def csr_impl(request):
    """Both CSR builders (csrc/csr.cu): stable radix sort of the entries, and count / fill / rank."""
    from ao_b200 import _lib

    _lib.set_tuning("csr_impl", 1 if request.param == "sort" else 2)
    yield request.param
    _lib.set_tuning("csr_impl", 0)


@pytest.mark.parametrize("n,m,k", [(1000, 1000, 16), (50, 4000, 3), (5000, 77, 1), (4097 * 3, 9000, 8), (2047, 2049, 2),
                                   (2048, 1, 1), (70000, 40000, 16), (5_000_000, 3000, 4)])
def test_csr_build(n, m, k, csr_impl):
    """n = 2047 / 2048: one and two 11-bit digits; 70000 x 16 entries: several tiles per CTA of the sort;
    n = 5M: three digits."""
    from ao_b200 import pointops

    rng = np.random.default_rng(n + k)
    idx = rand_idx(rng, m, k, n, pad_rows=min(5, m))
    (d_idx,) = to_cuda(idx)
    for mode in (0, 1):
        csr = pointops.build_csr(d_idx, n, mode)
        rowptr, perm = csr.rowptr.cpu().numpy(), csr.perm.cpu().numpy()
        flat = idx.reshape(-1).astype(np.int64)
        if mode == 1:
            flat = np.where(flat < 0, flat + n, flat)
        keep = flat >= 0
        order = np.argsort(flat[keep], kind="stable")
        expect_perm = np.flatnonzero(keep)[order]
        counts = np.bincount(flat[keep], minlength=n)
        assert np.array_equal(rowptr, np.r_[0, np.cumsum(counts)])
        assert np.array_equal(perm[: rowptr[-1]], expect_perm)      # ascending positions inside each row


def test_csr_hub_row(csr_impl):
    """All queries reference the same source: one row holds every entry (rank kernel worst case)."""
    from ao_b200 import pointops

    idx = np.zeros((3000, 3), np.int32)
    idx[:, 1] = 7
    (d_idx,) = to_cuda(idx)
    csr = pointops.build_csr(d_idx, 10, 0)
    rowptr, perm = csr.rowptr.cpu().numpy(), csr.perm.cpu().numpy()
    assert rowptr.tolist() == [0, 6000] + [6000] * 6 + [9000] * 3
    assert np.array_equal(perm[:6000], np.sort(np.r_[np.arange(0, 9000, 3), np.arange(2, 9000, 3)]))


# ------------------------------------------------------------------------------------------ grouping
def test_grouping_golden_from_reference_python():
    from ao_b200 import pointops

    g = gold("grouping")
    feat = g["feat"].cuda().requires_grad_(True)
    idx, xyz, new_xyz = g["idx"].cuda(), g["xyz"].cuda(), g["new_xyz"].cuda()
    out = pointops.grouping(idx, feat, xyz, new_xyz, with_xyz=True)
    assert torch.equal(out.cpu(), g["out_xyz"])                         # bit-exact incl. zero rows (== treats -0.0 as 0.0)
    (gf,) = torch.autograd.grad(out, feat, g["grad_out"].cuda())
    assert torch.allclose(gf.cpu(), g["grad_feat"], rtol=RTOL, atol=ATOL)
    assert torch.equal(pointops.grouping(idx, feat, xyz, new_xyz).cpu(), g["out_plain"])


@pytest.mark.parametrize("c", [48, 96, 7])
@pytest.mark.parametrize("with_xyz", [False, True])
def test_grouping_vs_torch_restatement(oracle, c, with_xyz):
    from ao_b200 import pointops

    rng = np.random.default_rng(c)
    n, m, k = 3000, 2500, 16
    idx = rand_idx(rng, m, k, n)
    feat = rng.standard_normal((n, c)).astype(np.float32)
    xyz = rng.standard_normal((n, 3)).astype(np.float32)
    new_xyz = rng.standard_normal((m, 3)).astype(np.float32)
    go = rng.standard_normal((m, k, c + (3 if with_xyz else 0))).astype(np.float32)
    d_idx, d_feat, d_xyz, d_new, d_go = to_cuda(idx, feat, xyz, new_xyz, go)
    d_feat.requires_grad_(True)
    out = pointops.grouping(d_idx, d_feat, d_xyz, d_new, with_xyz=with_xyz)
    (gf,) = torch.autograd.grad(out, d_feat, d_go)
    rf = torch.from_numpy(feat).requires_grad_(True)
    ro = oracle.grouping(torch.from_numpy(idx), rf, torch.from_numpy(xyz), torch.from_numpy(new_xyz), with_xyz=with_xyz)
    (rg,) = torch.autograd.grad(ro, rf, torch.from_numpy(go))
    assert torch.equal(out.cpu(), ro)
    assert torch.allclose(gf.cpu(), rg, rtol=RTOL, atol=ATOL)
    # determinism: the CSR backward has a fixed summation order
    (gf2,) = torch.autograd.grad(pointops.grouping(d_idx, d_feat, d_xyz, d_new, with_xyz=with_xyz), d_feat, d_go)
    assert torch.equal(gf, gf2)


def test_grouping2_vs_reference_cuda():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    from ao_b200 import pointops

    rng = np.random.default_rng(0)
    n, m, k, c = 4000, 4000, 16, 48
    idx = rng.integers(0, n, (m, k)).astype(np.int32)
    d_idx, d_in, d_go = to_cuda(idx, rng.standard_normal((n, c)).astype(np.float32),
                                rng.standard_normal((m, k, c)).astype(np.float32))
    d_in.requires_grad_(True)
    out = pointops.grouping2(d_in, d_idx)
    assert torch.equal(out, ref_cuda.grouping_forward(d_in.detach(), d_idx))
    (gi,) = torch.autograd.grad(out, d_in, d_go)
    assert torch.allclose(gi, ref_cuda.grouping_backward(d_go, d_idx, n), rtol=RTOL, atol=ATOL)   # atomics: order differs


# ------------------------------------------------------------------------------------------ GVA
@pytest.mark.parametrize("c,g,k", [(48, 6, 16), (96, 12, 16), (384, 48, 16), (64, 4, 16), (30, 5, 16),
                                   (32, 8, 16), (48, 6, 3), (48, 6, 18), (96, 12, 7), (64, 4, 32), (20, 4, 5)])
@pytest.mark.parametrize("use_peb", [True, False])
def test_gva_relation_and_aggregate(oracle, c, g, k, use_peb):
    """I = c/g in {4, 8, 16} takes the 128-bit chunk kernels (1, 2, 4 lanes per group), other widths the
    scalar kernels; k not a multiple of 4 exercises the remainder loops; n*c/4 is not a multiple of 32
    (partial last warp in the backward-query kernel)."""
    from ao_b200 import pointops

    rng = np.random.default_rng(c + g + k)
    n = 1501
    idx = rand_idx(rng, n, k, n, pad_rows=20)
    arrs = dict(key=rng.standard_normal((n, c)), query=rng.standard_normal((n, c)), value=rng.standard_normal((n, c)),
                peb=rng.standard_normal((n, k, c)), logits=3 * rng.standard_normal((n, k, g)),
                g_rel=rng.standard_normal((n, k, c)), g_out=rng.standard_normal((n, c)))
    arrs = {a: v.astype(np.float32) for a, v in arrs.items()}
    cpu = {a: torch.from_numpy(v) for a, v in arrs.items()}
    dev = {a: v.cuda() for a, v in cpu.items()}
    for d in (cpu, dev):
        for a in ("key", "query", "value", "peb", "logits"):
            d[a].requires_grad_(True)
    t_idx = torch.from_numpy(idx)
    d_idx = t_idx.cuda()

    rel = pointops.gva_relation(dev["key"], dev["query"], d_idx)
    r_rel = oracle.gva_relation(cpu["key"], cpu["query"], t_idx)
    assert torch.equal(rel.cpu(), r_rel)
    gk, gq = torch.autograd.grad(rel, [dev["key"], dev["query"]], dev["g_rel"])
    rk, rq = torch.autograd.grad(r_rel, [cpu["key"], cpu["query"]], cpu["g_rel"])
    assert torch.allclose(gk.cpu(), rk, rtol=RTOL, atol=ATOL) and torch.allclose(gq.cpu(), rq, rtol=RTOL, atol=ATOL)

    peb_d, peb_c = (dev["peb"], cpu["peb"]) if use_peb else (None, None)
    out = pointops.gva_aggregate(dev["value"], peb_d, dev["logits"], d_idx, g)
    r_out = oracle.gva_aggregate(cpu["value"], peb_c if use_peb else torch.zeros_like(cpu["peb"]), cpu["logits"], t_idx, g)
    assert torch.allclose(out.cpu(), r_out, rtol=1e-5, atol=2e-5)
    ins_d = [dev["value"], dev["logits"]] + ([dev["peb"]] if use_peb else [])
    ins_c = [cpu["value"], cpu["logits"]] + ([cpu["peb"]] if use_peb else [])
    gd = torch.autograd.grad(out, ins_d, dev["g_out"])
    gc = torch.autograd.grad(r_out, ins_c, cpu["g_out"])
    for a, b_ in zip(gd, gc):
        assert torch.allclose(a.cpu(), b_, rtol=1e-4, atol=5e-5)
    gd2 = torch.autograd.grad(pointops.gva_aggregate(dev["value"], peb_d, dev["logits"], d_idx, g), ins_d, dev["g_out"])
    assert all(torch.equal(a, b_) for a, b_ in zip(gd, gd2))         # deterministic


def test_gva_masked_slots_get_no_weight(oracle):
    """idx == -1: softmax runs over all k logits, then the padded slots are zeroed WITHOUT
    renormalisation (…v2m2_base.py:122-125)."""
    from ao_b200 import pointops

    n, k, c, g = 4, 4, 8, 1
    idx = torch.tensor([[0, 1, -1, -1]] * n, dtype=torch.int32).cuda()
    value = torch.ones(n, c).cuda()
    logits = torch.zeros(n, k, g).cuda()
    out = pointops.gva_aggregate(value, None, logits, idx, g)
    assert torch.allclose(out, torch.full((n, c), 0.5).cuda())     # two of four equal weights survive


# ------------------------------------------------------------------------------------------ GridPool
@pytest.mark.parametrize("c", [96, 10])
def test_grid_pool_vs_restatement(oracle, c):
    from ao_b200 import pointops, scenes

    coord, _, offset = scenes.s3dis_batch(2, n_points=6000)
    rng = np.random.default_rng(c)
    feat = np.maximum(rng.standard_normal((coord.shape[0], c)), 0).astype(np.float32)     # post-ReLU: ties at 0
    d_coord, d_feat, d_off = to_cuda(coord, feat, offset)
    d_feat.requires_grad_(True)
    (nc, nf, noff), cluster, part = pointops.grid_pool(d_coord, d_feat, d_off, 0.1, return_partition=True)
    rc, rf, roff, rcluster, rarg = oracle.grid_pool(torch.from_numpy(coord), torch.from_numpy(feat),
                                                    torch.from_numpy(offset), 0.1)
    assert torch.equal(cluster.cpu(), rcluster)                      # same voxels, same (scene,z,y,x) numbering
    assert torch.equal(noff.cpu(), roff) and noff.dtype == torch.int64
    assert torch.equal(nf.cpu(), rf)                                 # max is order-free → bit-exact
    assert torch.allclose(nc.cpu(), rc, rtol=0, atol=1e-6)           # same sequential order → expect equal
    go = torch.from_numpy(rng.standard_normal(tuple(nf.shape)).astype(np.float32))
    (gf,) = torch.autograd.grad(nf, d_feat, go.cuda())
    expect = torch.zeros_like(torch.from_numpy(feat))
    cols = torch.arange(c).expand_as(rarg)
    expect[rarg, cols] = go                                          # gradient to the first maximal point
    assert torch.equal(gf.cpu(), expect)
    # the reference path through its own ops: segment_csr(feat[sorted], idx_ptr) == direct reduction
    assert part.n_vox == rc.shape[0]


def test_grid_pool_hand_computed_fixture():
    """Known-answer test written out by hand (tests/golden/gridpool_hand.json; …v2m2_base.py:249-268): points exactly on
    cell faces, negative coordinates, single-point voxels, duplicated points, two scenes — and the same batch with an
    EMPTY scene in the middle."""
    import json

    from ao_b200 import pointops

    fx = json.load(open(os.path.join(GOLD, "gridpool_hand.json")))
    coord = torch.tensor(fx["coord"], dtype=torch.float32).cuda()
    feat = torch.tensor(fx["feat"], dtype=torch.float32).cuda().requires_grad_(True)
    counts = np.diff(np.array(fx["idx_ptr"]))
    mean = (np.array(fx["coord_sum"], np.float32) / counts[:, None].astype(np.float32)).astype(np.float32)
    for offset, new_offset in ((fx["offset"], fx["new_offset"]),
                               (fx["with_empty_scene"]["offset"], fx["with_empty_scene"]["new_offset"])):
        off = torch.tensor(offset, dtype=torch.int32).cuda()
        (nc, nf, noff), cluster, part = pointops.grid_pool(coord, feat, off, fx["grid_size"], return_partition=True)
        assert cluster.tolist() == fx["cluster"] and cluster.dtype == torch.int64
        assert part.idx_ptr.tolist() == fx["idx_ptr"]
        assert [sorted(part.order[a:b].tolist()) for a, b in zip(fx["idx_ptr"], fx["idx_ptr"][1:])] == fx["voxel_points"]
        assert part.order.tolist() == [p for mem in fx["voxel_points"] for p in mem]    # stable: ascending id in a voxel
        assert noff.tolist() == new_offset
        assert np.array_equal(nc.cpu().numpy(), mean)
        assert nf.tolist() == fx["feat_max"]
        go = torch.arange(1, nf.numel() + 1, dtype=torch.float32, device="cuda").view_as(nf)
        (gf,) = torch.autograd.grad(nf, feat, go)
        expect = torch.zeros_like(gf)
        for v, row in enumerate(fx["argmax"]):
            for c, p in enumerate(row):
                expect[p, c] = go[v, c]
        assert torch.equal(gf, expect)


@pytest.mark.parametrize("case", ["many_scenes", "wide_keys", "one_voxel"])
def test_voxel_partition_key_layouts(oracle, case):
    """The compact-key radix sort of aopt_voxel_grid: 700 scenes (10 scene bits; the old fixed layout stopped at
    512), extents that need more than 33 key bits (second call with six passes), every point in one voxel."""
    from ao_b200 import pointops

    rng = np.random.default_rng(11)
    if case == "many_scenes":
        sizes = rng.integers(1, 9, 700)
        coord = rng.uniform(-1, 1, (int(sizes.sum()), 3)).astype(np.float32)
        grid = 0.25
    elif case == "wide_keys":
        sizes = np.array([1500, 900])
        coord = rng.uniform(-4000, 4000, (2400, 3)).astype(np.float32)      # 8000 / 0.25 = 2^15 cells per axis: 46 bits
        grid = 0.25
    else:
        sizes = np.array([300])
        coord = rng.uniform(0, 0.01, (300, 3)).astype(np.float32)
        grid = 1.0
    offset = np.cumsum(sizes).astype(np.int32)
    d_coord, d_off = to_cuda(coord, offset)
    part = pointops.voxel_partition(d_coord, d_off, grid)
    feat = torch.from_numpy(coord.copy())
    rc, rf, roff, rcluster, rarg = oracle.grid_pool(torch.from_numpy(coord), feat, torch.from_numpy(offset), grid)
    assert torch.equal(part.cluster.cpu(), rcluster)
    assert torch.equal(part.offset.cpu(), roff)
    assert part.n_vox == rc.shape[0]
    order = part.order.cpu().numpy()
    ptr = part.idx_ptr.cpu().numpy()
    assert ptr[0] == 0 and ptr[-1] == coord.shape[0]
    cl = rcluster.numpy()
    assert np.array_equal(cl[order], np.repeat(np.arange(part.n_vox), np.diff(ptr)))
    assert all(np.all(np.diff(order[a:b]) > 0) for a, b in zip(ptr[:-1], ptr[1:]))     # ascending id inside a voxel


def test_voxel_partition_rejects_points_below_start():
    from ao_b200 import pointops

    coord = torch.rand(100, 3, device="cuda")
    off = torch.tensor([100], dtype=torch.int32, device="cuda")
    start = torch.full((1, 3), 0.5, device="cuda")
    with pytest.raises(ValueError):
        pointops.voxel_partition(coord, off, 0.1, start)
    part = pointops.voxel_partition(coord, off, 0.1, torch.full((1, 3), -0.25, device="cuda"))   # start below the minimum
    cells = torch.floor((coord + 0.25) / 0.1).long()
    key = (cells[:, 2] * 100 + cells[:, 1]) * 100 + cells[:, 0]
    assert torch.equal(part.cluster, torch.unique(key, return_inverse=True)[1])


def test_unpool_map(oracle):
    from ao_b200 import pointops, scenes

    coord, _, offset = scenes.s3dis_batch(1, n_points=5000)
    d_coord, d_off = to_cuda(coord, offset)
    feat = torch.randn(5000, 32, device="cuda")
    (nc, nf, noff), cluster = pointops.grid_pool(d_coord, feat, d_off, 0.2)
    coarse = torch.randn(nc.shape[0], 64, device="cuda", requires_grad=True)
    up = pointops.unpool_map(coarse, cluster)
    assert torch.equal(up, coarse[cluster])
    go = torch.randn_like(up)
    (g1,) = torch.autograd.grad(up, coarse, go)
    (g2,) = torch.autograd.grad(coarse[cluster], coarse, go)
    assert torch.allclose(g1, g2, rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------------------------------ interpolation
def test_interpolation_golden_from_reference_python():
    from ao_b200 import pointops

    g = gold("interpolation")
    feat = g["feat"].cuda().requires_grad_(True)
    out = pointops.interpolation(g["xyz"].cuda(), g["new_xyz"].cuda(), feat, g["offset"].cuda(), g["new_offset"].cuda())
    assert torch.allclose(out.cpu(), g["out"], rtol=1e-6, atol=1e-6)
    (gf,) = torch.autograd.grad(out, feat, g["grad_out"].cuda())
    assert torch.allclose(gf.cpu(), g["grad_feat"], rtol=RTOL, atol=ATOL)
    idx, d2 = pointops.knn_query_raw(3, g["xyz"].cuda(), g["offset"].cuda(), g["new_xyz"].cuda(), g["new_offset"].cuda())
    assert torch.equal(idx.cpu(), g["knn_idx"]) and torch.equal(d2.cpu(), g["knn_dist2"])


@pytest.mark.parametrize("c", [48, 192, 5])
def test_interpolation_vs_restatement_and_reference_cuda(oracle, c):
    from ao_b200 import pointops, scenes
    from oracle import ref_cuda

    fine, _, foff = scenes.s3dis_batch(2, n_points=8000)
    d_fine, d_foff = to_cuda(fine, foff)
    (cc, _, coff), _ = pointops.grid_pool(d_fine, torch.zeros(fine.shape[0], 4, device="cuda"), d_foff, 0.1)
    coarse, coff_np = cc.cpu().numpy(), coff.cpu().numpy().astype(np.int32)
    rng = np.random.default_rng(c)
    feat = rng.standard_normal((coarse.shape[0], c)).astype(np.float32)
    go = rng.standard_normal((fine.shape[0], c)).astype(np.float32)
    d_feat, d_go = to_cuda(feat, go)
    d_feat.requires_grad_(True)
    out = pointops.interpolation(cc, d_fine, d_feat, coff, d_foff)
    (gf,) = torch.autograd.grad(out, d_feat, d_go)
    rf = torch.from_numpy(feat).requires_grad_(True)
    ro = oracle.interpolation(torch.from_numpy(coarse), torch.from_numpy(fine), rf, torch.from_numpy(coff_np),
                              torch.from_numpy(foff))
    (rg,) = torch.autograd.grad(ro, rf, torch.from_numpy(go))
    assert torch.allclose(out.cpu(), ro, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gf.cpu(), rg, rtol=1e-4, atol=1e-5)
    if ref_cuda.available():                                          # interpolation2's kernels on the same idx/weights
        idx, d2 = pointops.knn_query_raw(3, cc, coff, d_fine, d_foff)
        w = pointops.interpolation_weights(d2)
        assert torch.allclose(out, ref_cuda.interpolation_forward(d_feat.detach(), idx, w), rtol=1e-5, atol=1e-6)
        assert torch.allclose(gf, ref_cuda.interpolation_backward(d_go, idx, w, coarse.shape[0]), rtol=1e-4, atol=1e-5)


def test_interpolation_weights_formula(oracle):
    from ao_b200 import pointops

    d2 = torch.rand(1000, 3).cuda() * 4
    d2[0] = 0.0
    d2[1] = torch.tensor([0.0, 1e10, 1e10])
    w = pointops.interpolation_weights(d2)
    ref = oracle.interpolation_weights(torch.sqrt(d2.cpu()))
    assert torch.allclose(w.cpu(), ref, rtol=1e-6, atol=1e-9)


# ------------------------------------------------------------------------------------------ layout helpers
def test_offset_helpers():
    from ao_b200 import pointops

    g = gold("offsets")
    off = g["offset"].cuda()
    batch = pointops.offset2batch(off)
    assert torch.equal(batch.cpu(), g["batch"]) and batch.dtype == torch.int64
    assert torch.equal(pointops.batch2offset(batch).cpu(), g["back"])


# ------------------------------------------------------------------------------------------ PTv1-layout API parity
def test_aggregation_and_subtraction_vs_reference_cuda(oracle):
    from ao_b200 import pointops
    from oracle import ref_cuda

    rng = np.random.default_rng(8)
    n, k, c, w_c = 2000, 16, 32, 8
    idx = rng.integers(0, n, (n, k)).astype(np.int32)
    inp, pos, wgt, go = (rng.standard_normal(s).astype(np.float32) for s in ((n, c), (n, k, c), (n, k, w_c), (n, c)))
    d_idx, d_in, d_pos, d_w, d_go = to_cuda(idx, inp, pos, wgt, go)
    for t in (d_in, d_pos, d_w):
        t.requires_grad_(True)
    out = pointops.aggregation(d_in, d_pos, d_w, d_idx)
    gi, gp, gw = torch.autograd.grad(out, [d_in, d_pos, d_w], d_go)
    ci, cp, cw = (torch.from_numpy(a).requires_grad_(True) for a in (inp, pos, wgt))
    r = oracle.aggregation(ci, cp, cw, torch.from_numpy(idx))
    ri, rp, rw = torch.autograd.grad(r, [ci, cp, cw], torch.from_numpy(go))
    assert torch.allclose(out.cpu(), r, rtol=1e-5, atol=1e-5)
    for a, b_ in ((gi, ri), (gp, rp), (gw, rw)):
        assert torch.allclose(a.cpu(), b_, rtol=1e-4, atol=1e-4)
    if ref_cuda.available():
        assert torch.allclose(out, ref_cuda.aggregation_forward(d_in.detach(), d_pos.detach(), d_w.detach(), d_idx), rtol=1e-5, atol=1e-5)
        qi, qp, qw = ref_cuda.aggregation_backward(d_in.detach(), d_pos.detach(), d_w.detach(), d_idx, d_go)
        for a, b_ in ((gi, qi), (gp, qp), (gw, qw)):
            assert torch.allclose(a, b_, rtol=1e-4, atol=1e-4)
    a1, a2 = torch.randn(n, c, device="cuda", requires_grad=True), torch.randn(n, c, device="cuda", requires_grad=True)
    sub = pointops.subtraction(a1, a2, d_idx)
    assert torch.equal(sub.cpu(), oracle.subtraction(a1.detach().cpu(), a2.detach().cpu(), torch.from_numpy(idx)))
    gsub = torch.randn_like(sub)
    g1, g2 = torch.autograd.grad(sub, [a1, a2], gsub)
    if ref_cuda.available():
        assert torch.equal(sub, ref_cuda.subtraction_forward(a1.detach(), a2.detach(), d_idx))
        q1, q2 = ref_cuda.subtraction_backward(d_idx, gsub, n)
        assert torch.allclose(g1, q1, rtol=1e-4, atol=1e-4) and torch.allclose(g2, q2, rtol=1e-4, atol=1e-4)


def test_non_default_stream_and_device_guard():
    """Kernels launch on the caller's current stream (the reference always uses the legacy stream)."""
    from ao_b200 import pointops, scenes

    coord, _, offset = scenes.small_batch(seed=9)
    xyz, off = to_cuda(coord, offset)
    ref, _ = pointops.knn_query(8, xyz, off)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a, _ = pointops.knn_query(8, xyz, off)
        feat = torch.randn(xyz.shape[0], 16, device="cuda")
        gq = pointops.grouping(a, feat, xyz, with_xyz=True)
    s.synchronize()
    assert torch.equal(a, ref) and gq.shape == (xyz.shape[0], 8, 19)


@pytest.mark.parametrize("c", [48, 7])
def test_relation_backward_fused_equals_the_two_pass_kernels(c):
    """aopt_relation_backward (one pass over the (n,k,c) gradient) gives bit-identical grad_key / grad_query to
    aopt_grouping_backward + aopt_sum_over_k (same summation orders), -1 padded rows included."""
    from ao_b200 import _lib, pointops, scenes

    coord, _, off = scenes.small_batch(77, sizes=(900, 5, 1400))
    xyz, o = to_cuda(coord, off)
    n, k = xyz.shape[0], 16
    idx, _ = pointops.knn_query(k, xyz, o)
    assert int((idx < 0).sum()) > 0                              # the 5-point scene pads with -1
    g = torch.Generator(device="cuda").manual_seed(c)
    grad = torch.randn(n, k, c, device="cuda", generator=g)
    csr = pointops.get_csr(idx, n)
    lib = _lib.load()
    gk1, gq1, gk2, gq2 = (torch.empty(n, c, device="cuda") for _ in range(4))
    _lib.check(lib.aopt_grouping_backward(n, c, grad.data_ptr(), c, csr.rowptr.data_ptr(), csr.perm.data_ptr(), 1.0, gk1.data_ptr(), _lib.stream()), "a")
    _lib.check(lib.aopt_sum_over_k(n, k, c, grad.data_ptr(), -1.0, gq1.data_ptr(), _lib.stream()), "b")
    _lib.check(lib.aopt_relation_backward(n, k, c, grad.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(), gk2.data_ptr(), gq2.data_ptr(), _lib.stream()), "c")
    assert torch.equal(gk1, gk2) and torch.equal(gq1, gq2)
    ref = torch.zeros(n + 1, c, device="cuda", dtype=torch.float64).index_add_(0, torch.where(idx < 0, n, idx).reshape(-1).long(), grad.reshape(-1, c).double())[:n]
    assert torch.allclose(gk2.double(), ref, rtol=1e-5, atol=1e-4)
    # through autograd (pointops.gva_relation uses the fused call when queries == sources)
    key, query = (torch.randn(n, c, device="cuda", generator=g, requires_grad=True) for _ in range(2))
    a, b = torch.autograd.grad(pointops.gva_relation(key, query, idx), [key, query], grad)
    assert torch.equal(a, gk2) and torch.equal(b, gq2)


@pytest.mark.parametrize("c,g,k,use_peb", [(48, 6, 16, True), (96, 12, 16, True), (384, 48, 16, False), (64, 4, 8, True),
                                           (64, 16, 32, True), (48, 6, 12, True)])
def test_gva_backward_fused_equals_the_two_kernels(c, g, k, use_peb):
    """aopt_gva_backward (query pass + CSR walk in one kernel, self-attention) gives bit-identical grad_peb /
    grad_logits / grad_value to aopt_gva_backward_query + aopt_gva_backward_value: real neighbour lists with -1
    padded rows, hub rows far longer than one batch of the walk, a partial last warp; k = 12 takes the fallback."""
    from ao_b200 import _lib, pointops, scenes

    coord, _, off = scenes.small_batch(31, sizes=(1100, 5, 1700, 3))
    coord[:40] = coord[0] + 1e-3 * np.arange(40, dtype=np.float32)[:, None]   # a tight clump: its members are hubs
    xyz, o = to_cuda(coord, off)
    n = xyz.shape[0]
    idx, _ = pointops.knn_query(k, xyz, o)
    assert int((idx < 0).sum()) > 0
    gen = torch.Generator(device="cuda").manual_seed(c + k)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    value, peb, logits, grad_out = rnd(n, c), (rnd(n, k, c) if use_peb else None), rnd(n, k, g), rnd(n, c)
    prob = torch.softmax(logits, dim=1).contiguous()
    csr = pointops.get_csr(idx, n)
    assert int(torch.diff(csr.rowptr).max()) > 2 * 8
    lib = _lib.load()
    P = _lib.ptr
    gp1, gp2 = ((torch.empty(n, k, c, device="cuda"), torch.empty(n, k, c, device="cuda")) if use_peb else (None, None))
    gl1, gl2, gv1, gv2 = (torch.empty(n, k, g, device="cuda"), torch.empty(n, k, g, device="cuda"),
                          torch.empty(n, c, device="cuda"), torch.empty(n, c, device="cuda"))
    _lib.check(lib.aopt_gva_backward_query(n, k, c, g, P(grad_out), P(value), P(peb), P(prob), P(idx), P(gp1), P(gl1), _lib.stream()), "q")
    _lib.check(lib.aopt_gva_backward_value(n, k, c, g, P(grad_out), P(prob), P(csr.rowptr), P(csr.perm), P(gv1), _lib.stream()), "v")
    _lib.check(lib.aopt_gva_backward(n, k, c, g, P(grad_out), P(value), P(peb), P(prob), P(idx), P(csr.rowptr), P(csr.perm),
                                     P(gp2), P(gl2), P(gv2), _lib.stream()), "fused")
    assert torch.equal(gl1, gl2) and torch.equal(gv1, gv2)
    if use_peb:
        assert torch.equal(gp1, gp2)
    # grad_value against an fp64 scatter
    w = prob.double() * (idx >= 0).unsqueeze(-1)
    contrib = (w.unsqueeze(-1) * grad_out.double().view(n, 1, g, c // g)).reshape(n * k, c)
    ref = torch.zeros(n + 1, c, device="cuda", dtype=torch.float64).index_add_(0, torch.where(idx < 0, n, idx).reshape(-1).long(), contrib)[:n]
    assert torch.allclose(gv2.double(), ref, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------ per-scene bounding box
@pytest.mark.parametrize("sizes", [
    [5000, 1, 0, 3000, 2047, 2049, 7],            # empty scene, scenes ending right around the 2048-point chunks
    [1] * 300,                                    # hundreds of one-point scenes inside one chunk
    [3, 0, 0, 4100, 0, 2],                        # consecutive empty scenes
    [10000],                                      # B = 1
])
def test_segment_min3_matches_numpy_on_adversarial_layouts(sizes):
    """aopt_segment_min3 (csrc/bbox.cu): chunks that straddle scene boundaries take the masked warp-reduction path;
    the result must be the exact per-scene minimum (…v2m2_base.py:249-253 `segment_csr(coord, ptr, reduce="min")`)."""
    from ao_b200 import _lib

    rng = np.random.default_rng(len(sizes))
    n = int(sum(sizes))
    coord = (rng.standard_normal((n + 5, 3)) * 10).astype(np.float32)     # 5 points past the last offset: no scene
    off = np.cumsum(sizes).astype(np.int32)
    c, o = to_cuda(coord, off)
    start = torch.empty((len(sizes), 3), dtype=torch.float32, device="cuda")
    lib = _lib.load()
    _lib.check(lib.aopt_segment_min3(n + 5, len(sizes), _lib.ptr(c), _lib.ptr(o), _lib.ptr(start), _lib.stream()),
               "segment_min3")
    got = start.cpu().numpy()
    s = 0
    for b, e in enumerate(off):
        if e > s:
            assert np.array_equal(got[b], coord[s:e].min(axis=0)), b
        else:
            assert np.array_equal(got[b], np.zeros(3, np.float32)), b       # segment_csr's fill value for an empty segment
        s = e
