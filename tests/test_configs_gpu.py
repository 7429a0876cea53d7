"""BASELINE.json configs[3] and configs[4] at their full sizes (parity cases, not bench lines):

  configs[3]  ScanNet-shaped room (150k points, 0.02 m voxels): the four GridPool stages of
              configs/scannet/semseg-pt-v2m2-0-base.py (grid 0.06 / 0.15 / 0.375 / 0.9375), both up paths
              (`map` gather = UnpoolWithSkip backend "map", …v2m2_base.py:308-309, and `interp`, :311-313), kNN k = 16 / 32.
  configs[4]  SemanticKITTI-shaped scan (120k points, sparse): kNN k in {8, 16, 32} + the fused GVA pair over
              C in {48, 96, 192, 384}.

Full-size checks are size-independent properties plus oracle comparisons on sampled rows (the C oracle's
brute-force scan of a 150k-point scene is restricted to a query sample so the test stays in seconds)."""
import numpy as np
import pytest
import torch

from helpers import to_cuda

pytestmark = pytest.mark.gpu
SCANNET_GRIDS = (0.06, 0.15, 0.375, 0.9375)     # configs/scannet/semseg-pt-v2m2-0-base.py:28
SCANNET_CH = (96, 192, 384, 512)
KITTI_C_G = ((48, 6), (96, 12), (192, 24), (384, 48))


def _knn_sample_check(oracle, k, coord, off, idx, d2, rows):
    """Bit-exact idx / dist2 on the sampled query rows against the C oracle's exhaustive scan."""
    q = np.ascontiguousarray(coord[rows])
    scene = np.searchsorted(off, rows, side="right")
    order = np.argsort(scene, kind="stable")
    rows, q, scene = rows[order], q[order], scene[order]
    q_off = np.cumsum(np.bincount(scene, minlength=len(off))).astype(np.int32)
    ri, rd = oracle.knn_query(k, coord, off, q, q_off, rule="lex")
    assert np.array_equal(d2[rows].view(np.uint32), rd.view(np.uint32))
    same = (idx[rows] == ri).all(1)
    if not same.all():                                          # only exact-distance ties may permute
        from helpers import tie_rows

        assert tie_rows(rd)[~same].all()
    return float(same.mean())


@pytest.fixture(scope="module")
def scannet_room():
    from ao_b200 import scenes

    coord, feat, off = scenes.scannet_batch(1, 150000)
    assert coord.shape == (150000, 3) and feat.shape[1] == 9
    return coord, feat, off


@pytest.mark.parametrize("k", [16, 32])
def test_scannet_room_knn(oracle, scannet_room, k):
    from ao_b200 import pointops

    coord, _, off = scannet_room
    xyz, o = to_cuda(coord, off)
    idx, d2 = pointops.knn_query_raw(k, xyz, o)
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    assert (idx[:, 0] == np.arange(coord.shape[0])).all() and (d2[:, 0] == 0).all()   # self first (tie-free data)
    assert (np.diff(d2, axis=1) >= 0).all() and idx.min() >= 0 and idx.max() < coord.shape[0]
    rows = np.random.default_rng(k).choice(coord.shape[0], 256, replace=False)
    assert _knn_sample_check(oracle, k, coord, off, idx, d2, np.sort(rows)) > 0.99
    ti, td = pointops.knn_query_raw(k, xyz, o, method="tile")
    assert np.array_equal(ti.cpu().numpy(), idx) and np.array_equal(td.cpu().numpy().view(np.uint32), d2.view(np.uint32))


def test_scannet_pool_pyramid_and_both_up_paths(oracle, scannet_room):
    """GridPool down path over the four ScanNet stages, then `map` and `interp` up paths, forward + backward."""
    from ao_b200 import pointops

    coord, _, off = scannet_room
    xyz, o = to_cuda(coord, off)
    g = torch.Generator(device="cuda").manual_seed(7)
    levels = [(xyz, o.int())]
    clusters, feats = [], []
    for gs, c in zip(SCANNET_GRIDS, SCANNET_CH):
        cx, co = levels[-1]
        f = torch.relu(torch.randn(cx.shape[0], c, device="cuda", generator=g)).requires_grad_(True)
        (nc, nf, noff), cluster, part = pointops.grid_pool(cx, f, co, gs, return_partition=True)
        nv, n = nc.shape[0], cx.shape[0]
        counts = torch.diff(part.idx_ptr.long())
        assert int(counts.sum()) == n and int(counts.min()) >= 1 and int(noff[-1]) == nv
        # same voxel <=> same cell of the reference's voxel_grid (start = scene minimum, …v2m2_base.py:249-259)
        start = cx.min(0).values
        cell = torch.floor((cx - start) / gs).long()
        key = (cell[:, 2] * 4096 + cell[:, 1]) * 4096 + cell[:, 0]
        uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
        assert uniq.numel() == nv and torch.equal(inv, cluster)
        ref_max = torch.full((nv, c), -1.0, device="cuda").scatter_reduce(0, cluster[:, None].expand(-1, c), f.detach(), "amax")
        assert torch.equal(nf.detach(), ref_max)
        ref_mean = torch.zeros(nv, 3, device="cuda", dtype=torch.float64).index_add_(0, cluster, cx.double()) / counts[:, None]
        assert torch.allclose(nc.double(), ref_mean, rtol=0, atol=1e-5)
        go = torch.randn(nv, c, device="cuda", generator=g)
        (gf,) = torch.autograd.grad(nf, [f], go)
        assert torch.allclose(gf.double().sum(0), go.double().sum(0), rtol=1e-6, atol=1e-3)
        levels.append((nc.contiguous(), noff.int()))
        clusters.append(cluster)
        feats.append(c)
    sizes = [lv[0].shape[0] for lv in levels]
    assert sizes[0] == 150000 and all(a > b for a, b in zip(sizes, sizes[1:]))
    # up paths, coarsest -> finest
    for li in range(len(SCANNET_GRIDS) - 1, -1, -1):
        (fx, fo), (cx, co) = levels[li], levels[li + 1]
        c = 48
        src = torch.randn(cx.shape[0], c, device="cuda", generator=g, requires_grad=True)
        # map backend: feat[cluster]  (…v2m2_base.py:308-309), backward = segment sum over the voxel partition
        up = pointops.unpool_map(src, clusters[li])
        assert torch.equal(up.detach(), src.detach()[clusters[li]])
        y = torch.randn(fx.shape[0], c, device="cuda", generator=g)
        (gs_map,) = torch.autograd.grad(up, [src], y)
        ref = torch.zeros(cx.shape[0], c, device="cuda", dtype=torch.float64).index_add_(0, clusters[li], y.double())
        assert torch.allclose(gs_map.double(), ref, rtol=1e-5, atol=1e-4)
        # interp backend (:311-313)
        up2 = pointops.interpolation(cx, fx, src, co, fo)
        const = pointops.interpolation(cx, fx, torch.ones(cx.shape[0], 4, device="cuda"), co, fo)
        assert torch.allclose(const, torch.ones_like(const), rtol=1e-5, atol=1e-5)
        (gs_i,) = torch.autograd.grad(up2, [src], y)
        lhs, rhs = (up2.detach().double() * y.double()).sum(), (src.detach().double() * gs_i.double()).sum()
        assert abs(lhs.item() - rhs.item()) <= 1e-6 * abs(lhs.item()) + 1e-2
        if cx.shape[0] <= 3000:                                   # small levels: full oracle comparison
            ref_up = oracle.interpolation(cx.cpu(), fx.cpu(), src.detach().cpu(), co.cpu(), fo.cpu())
            assert torch.allclose(up2.detach().cpu(), ref_up, rtol=1e-5, atol=2e-5)


@pytest.fixture(scope="module")
def kitti_scan():
    from ao_b200 import scenes

    coord, feat, off = scenes.kitti_batch(1, 120000)
    assert coord.shape[0] == 120000 and feat.shape[1] in (1, 4)
    return coord, off


@pytest.mark.parametrize("k", [8, 16, 32])
def test_kitti_scan_knn(oracle, kitti_scan, k):
    from ao_b200 import pointops

    coord, off = kitti_scan
    xyz, o = to_cuda(coord, off)
    idx, d2 = pointops.knn_query_raw(k, xyz, o)
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    assert (np.diff(d2, axis=1) >= 0).all() and idx.min() >= 0
    rows = np.sort(np.random.default_rng(100 + k).choice(coord.shape[0], 256, replace=False))
    assert _knn_sample_check(oracle, k, coord, off, idx, d2, rows) > 0.99


@pytest.mark.parametrize("c,g", KITTI_C_G)
def test_kitti_scan_gva_sweep(oracle, kitti_scan, c, g):
    """Fused GVA pair (gather_sub + softmax-aggregate) over the channel sweep, k = 16: sampled rows against
    the torch restatement of …v2m2_base.py:109-128, adjointness of forward/backward on the whole scan."""
    from ao_b200 import pointops

    coord, off = kitti_scan
    xyz, o = to_cuda(coord, off)
    n, k = coord.shape[0], 16
    idx, _ = pointops.knn_query(k, xyz, o)
    gen = torch.Generator(device="cuda").manual_seed(c)
    key, query, value = (torch.randn(n, c, device="cuda", generator=gen, requires_grad=True) for _ in range(3))
    peb = torch.randn(n, k, c, device="cuda", generator=gen, requires_grad=True)
    logits = torch.randn(n, k, g, device="cuda", generator=gen, requires_grad=True)
    rel = pointops.gva_relation(key, query, idx)
    out = pointops.gva_aggregate(value, peb, logits, idx, g)
    rows = torch.from_numpy(np.sort(np.random.default_rng(c).choice(n, 512, replace=False))).cuda()
    ref_rel = oracle.gva_relation(key.detach(), query.detach()[rows], idx[rows])
    assert torch.equal(rel.detach()[rows], ref_rel)
    ref_out = oracle.gva_aggregate(value.detach(), peb.detach()[rows], logits.detach()[rows], idx[rows], g)
    assert torch.allclose(out.detach()[rows], ref_out, rtol=1e-5, atol=2e-5)
    y = torch.randn(n, c, device="cuda", generator=gen)
    gv, gp, gl = torch.autograd.grad(out, [value, peb, logits], y)
    # <out, y> = <value, gv> + <peb, gp> for fixed softmax weights (out is linear in value and peb)
    lhs = (out.detach().double() * y.double()).sum()
    rhs = (value.detach().double() * gv.double()).sum() + (peb.detach().double() * gp.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * abs(lhs.item()) + 5e-2
    assert abs(float(gl.double().sum())) <= 1e-2 * max(1.0, float(gl.abs().double().sum()) ** 0.5)   # softmax grads sum to 0 over k
    yr = torch.randn(n, k, c, device="cuda", generator=gen)
    gk, gq = torch.autograd.grad(rel, [key, query], yr)
    assert torch.allclose(gq, -yr.sum(1), rtol=1e-5, atol=1e-4)
    lhs = (rel.detach().double() * yr.double()).sum()
    rhs = (key.detach().double() * gk.double()).sum() + (query.detach().double() * gq.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-6 * abs(lhs.item()) + 5e-2
