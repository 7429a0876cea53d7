"""GPU parity: aopt_knn_query (TILE and GRID methods, through the C ABI via ao_b200.pointops)
against the C oracle and against the reference's own CUDA kernel (oracle/_ref)."""
import numpy as np
import pytest
import torch

from helpers import assert_knn_equal, tie_rows, to_cuda

pytestmark = pytest.mark.gpu


def run(method, k, coord, offset, q=None, qoff=None):
    from ao_b200 import pointops

    xyz, off = to_cuda(coord, offset)
    if q is None:
        idx, d2 = pointops.knn_query_raw(k, xyz, off, method=method)
    else:
        nq, noff = to_cuda(q, qoff)
        idx, d2 = pointops.knn_query_raw(k, xyz, off, nq, noff, method=method)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), d2.cpu().numpy()


@pytest.mark.parametrize("method", ["tile", "grid"])
@pytest.mark.parametrize("k", [1, 3, 8, 16, 32])
def test_small_adversarial_batch_bit_exact(oracle, method, k):
    """Unequal scenes, one with 5 < k points (→ -1 / 1e10 padding), tie-free coordinates."""
    from ao_b200 import scenes

    coord, _, offset = scenes.small_batch(seed=k, sizes=(700, 5, 1300, 257))
    idx, d2 = run(method, k, coord, offset)
    ri, rd = oracle.knn_query(k, coord, offset, rule="lex")
    assert_knn_equal(idx, d2, ri, rd)
    hi, hd = oracle.knn_query(k, coord, offset, rule="heap")      # the reference algorithm
    assert_knn_equal(idx, d2, hi, hd)


@pytest.mark.parametrize("k", [8, 16, 32])
def test_heap_container_equals_register_list(oracle, k):
    """The GRID query kernel's opt-in top-k container (tuning "knn_topk" = 1: binary heap in shared memory + heap sort)
    returns the bits of the default sorted register list — ties (duplicated points), short scenes and padding included."""
    from ao_b200 import _lib, scenes

    coord, _, offset = scenes.small_batch(seed=40 + k, sizes=(3000, 5, 4100, 257))
    coord[100:140] = coord[60:100]                   # duplicates: equal distances, ranked by index
    try:
        _lib.set_tuning("knn_topk", 1)
        hi, hd = run("grid", k, coord, offset)
        _lib.set_tuning("knn_topk", 2)
        li, ld = run("grid", k, coord, offset)
        _lib.set_tuning("knn_pend", 1)               # per-lane pending list instead of inserting in place
        pi, pd = run("grid", k, coord, offset)
        _lib.set_tuning("knn_topk", 1)               # ... and both together
        qi, qd = run("grid", k, coord, offset)
        assert np.array_equal(qi, li) and np.array_equal(qd, ld)
    finally:
        _lib.set_tuning("knn_topk", 0)
        _lib.set_tuning("knn_pend", 0)
    assert np.array_equal(hi, li) and np.array_equal(hd, ld)
    assert np.array_equal(pi, li) and np.array_equal(pd, ld)
    ri, rd = oracle.knn_query(k, coord, offset, rule="lex")
    assert np.array_equal(hi, ri) and np.array_equal(hd, rd)


@pytest.mark.parametrize("k", [1, 3, 8, 16, 32])
def test_single_site_query_kernel_equals_the_default(oracle, k):
    """Tuning "knn_site" = 1 selects knn_grid1_kernel (one scan and one insert site, tail masked instead of a tail loop,
    the exhaustive fallback folded into the ring loop): same bits as the default kernel and the oracle, self and cross
    searches, duplicates, scenes shorter than k."""
    from ao_b200 import _lib, scenes

    coord, _, offset = scenes.small_batch(seed=70 + k, sizes=(3000, 5, 4100, 257))
    coord[100:140] = coord[60:100]
    rng = np.random.default_rng(k)
    query = (coord[rng.integers(0, coord.shape[0], 900)] + rng.normal(0, 0.05, (900, 3))).astype(np.float32)
    q_off = np.array([300, 300, 650, 900], np.int32)             # scene 1 has no queries
    try:
        _lib.set_tuning("knn_site", 1)
        si, sd = run("grid", k, coord, offset)
        sci, scd = run("grid", k, coord, offset, query, q_off)
    finally:
        _lib.set_tuning("knn_site", 0)
    di, dd = run("grid", k, coord, offset)
    dci, dcd = run("grid", k, coord, offset, query, q_off)
    assert np.array_equal(si, di) and np.array_equal(sd, dd)
    assert np.array_equal(sci, dci) and np.array_equal(scd, dcd)
    ri, rd = oracle.knn_query(k, coord, offset, rule="lex")
    assert np.array_equal(si, ri) and np.array_equal(sd, rd)


@pytest.mark.parametrize("method", ["tile", "grid", "auto"])
def test_no_candidates_gives_padding(method):
    """Queries against an EMPTY candidate set (n = 0): every row is padding (idx -1, dist2 1e10), whichever method is
    asked for — the grid has nothing to build and hands over to the scan kernel."""
    from ao_b200 import pointops

    q = torch.rand(3000, 3, device="cuda")
    xyz = torch.zeros(0, 3, device="cuda")
    off, qoff = (torch.tensor([v], dtype=torch.int32, device="cuda") for v in (0, 3000))
    idx, d2 = pointops.knn_query_raw(4, xyz, off, q, qoff, method=method)
    assert bool((idx == -1).all()) and bool((d2 == 1e10).all())


@pytest.mark.parametrize("method", ["tile", "grid", "auto"])
def test_batched_sets_equal_one_search_per_set(method):
    """pointops.knn_query_sets (aopt_knn_query_multi: several point sets concatenated scene by scene, indices rebased per
    set) returns, bit for bit, what one search per set returns — self lists (k = 16) and coarse -> fine cross lists
    (k = 3), scenes shorter than k included."""
    from ao_b200 import pointops, scenes

    sets = []
    for seed, sizes in ((1, (3000, 7, 2500)), (2, (900, 400, 5)), (3, (60, 33, 12))):
        coord, _, off = scenes.small_batch(seed, sizes=sizes)
        sets.append(to_cuda(coord, off))
    res = pointops.knn_query_sets(16, [(c, o, None, None) for c, o in sets], root=True, method=method)
    for (c, o), (idx, dist) in zip(sets, res):
        ri, rd2 = pointops.knn_query_raw(16, c, o, method=method)
        assert torch.equal(idx, ri) and torch.equal(dist, torch.sqrt(rd2))
        assert int(idx.max()) < c.shape[0]
    cross = [(sets[l][0], sets[l][1], sets[l - 1][0], sets[l - 1][1]) for l in (1, 2)]
    res = pointops.knn_query_sets(3, cross, root=False, method=method)
    for (c, o, q, qo), (idx, d2) in zip(cross, res):
        ri, rd2 = pointops.knn_query_raw(3, c, o, q, qo, method=method)
        assert torch.equal(idx, ri) and torch.equal(d2, rd2)


@pytest.mark.parametrize("method", ["tile", "grid"])
def test_bigk_and_many_scenes(oracle, method):
    rng = np.random.default_rng(1)
    sizes = rng.integers(1, 90, 64)                                 # B = 64, some scenes < k
    coord = rng.uniform(-1, 1, (int(sizes.sum()), 3)).astype(np.float32)
    offset = np.cumsum(sizes).astype(np.int32)
    for k in ([16, 100] if method == "tile" else [16]):
        idx, d2 = run(method, k, coord, offset)
        ri, rd = oracle.knn_query(k, coord, offset, rule="lex")
        assert_knn_equal(idx, d2, ri, rd)


@pytest.mark.parametrize("method", ["tile", "grid"])
@pytest.mark.parametrize("k", [1, 3])
def test_cross_set_queries(oracle, method, k):
    """interpolation (k=3, coarse→fine) and evaluator (k=1) call shapes: m != n, queries outside the
    candidates' bounding box."""
    rng = np.random.default_rng(2)
    cs, fs = [900, 2, 400], [5000, 30, 2100]
    coarse = rng.uniform(-1, 1, (sum(cs), 3)).astype(np.float32)
    fine = rng.uniform(-1.5, 1.5, (sum(fs), 3)).astype(np.float32)
    off, noff = np.cumsum(cs).astype(np.int32), np.cumsum(fs).astype(np.int32)
    idx, d2 = run(method, k, coarse, off, fine, noff)
    ri, rd = oracle.knn_query(k, coarse, off, fine, noff, rule="lex")
    assert_knn_equal(idx, d2, ri, rd)


@pytest.mark.parametrize("method", ["tile", "grid"])
def test_duplicate_points_follow_the_lexicographic_contract(oracle, method):
    from ao_b200 import scenes

    coord, _, offset = scenes.small_batch(seed=3, sizes=(600, 900), dup=200)   # exact duplicates → ties
    idx, d2 = run(method, 16, coord, offset)
    ri, rd = oracle.knn_query(16, coord, offset, rule="lex")
    assert tie_rows(rd).sum() > 100
    assert_knn_equal(idx, d2, ri, rd)                                # ties broken by lower index
    hi, hd = oracle.knn_query(16, coord, offset, rule="heap")
    changed = assert_knn_equal(idx, d2, hi, hd, allow_tie_perm=True)  # reference differs only inside tie groups
    print("tie rows where the reference heap order differs:", changed)


@pytest.mark.parametrize("k", [16, 3])
def test_s3dis_room_grid_equals_tile_and_oracle_sample(oracle, k):
    """Full-size room (80k points): GRID == TILE everywhere (bitwise), and both equal the oracle on a
    sample of rows (the oracle is O(n^2) on one core)."""
    from ao_b200 import scenes

    coord, _, offset = scenes.s3dis_batch(2)
    it, dt = run("tile", k, coord, offset)
    ig, dg = run("grid", k, coord, offset)
    assert np.array_equal(dt.view(np.uint32), dg.view(np.uint32))
    assert np.array_equal(it, ig)
    for b, e in ((0, 256), (79990, 80250), (159744, 160000)):
        ri, rd = oracle.knn_query(k, coord, offset, rule="lex", rows=(b, e))
        assert_knn_equal(ig[b:e], dg[b:e], ri, rd)
    # self query: nearest neighbour of every point is itself at distance 0
    assert (ig[:, 0] == np.arange(ig.shape[0])).all() and (dg[:, 0] == 0).all()
    assert (np.diff(dg, axis=1) >= 0).all()


def test_outdoor_scan_grid_equals_tile(oracle):
    """Strongly non-uniform density (LiDAR): shells up to the cap and the full-scan fallback."""
    from ao_b200 import scenes

    coord, _, offset = scenes.kitti_batch(1, n_points=60000)
    for k in (8, 16, 32):
        it, dt = run("tile", k, coord, offset)
        ig, dg = run("grid", k, coord, offset)
        assert np.array_equal(dt.view(np.uint32), dg.view(np.uint32))
        assert np.array_equal(it, ig)


def test_against_reference_cuda_kernel(oracle):
    """The reference's own kernel (oracle/_ref, built from /root/reference sources) on the same
    device: pins both the C oracle and the new kernels.  Tie-free input → everything bit-equal."""
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpointops_ref.so not built")
    from ao_b200 import pointops, scenes

    coord, _, offset = scenes.s3dis_batch(1, n_points=20000)
    xyz, off = to_cuda(coord, offset)
    for k in (16, 3, 1):
        ri, rd = ref_cuda.knn_query(k, xyz, off)
        ri, rd = ri.cpu().numpy(), rd.cpu().numpy()
        assert tie_rows(rd).sum() == 0
        for method in ("tile", "grid"):
            idx, d2 = pointops.knn_query_raw(k, xyz, off, method=method)
            assert_knn_equal(idx.cpu().numpy(), d2.cpu().numpy(), ri, rd)
        oi, od = oracle.knn_query(k, coord, offset, rule="heap", )
        assert_knn_equal(oi, od, ri, rd)          # C oracle == reference CUDA kernel
    # cross-set, small adversarial batch incl. a scene shorter than k
    c2, _, o2 = scenes.small_batch(seed=11)
    q2 = np.random.default_rng(4).uniform(-2, 2, (1000, 3)).astype(np.float32)
    qo = np.array([300, 310, 800, 1000], np.int32)
    x2, of2, qq, qof = to_cuda(c2, o2, q2, qo)
    ri, rd = ref_cuda.knn_query(8, x2, of2, qq, qof)
    idx, d2 = pointops.knn_query_raw(8, x2, of2, qq, qof, method="tile")
    assert_knn_equal(idx.cpu().numpy(), d2.cpu().numpy(), ri.cpu().numpy(), rd.cpu().numpy())


def test_python_api_matches_reference_signature():
    from ao_b200 import pointops, scenes

    coord, _, offset = scenes.small_batch(seed=5)
    xyz, off = to_cuda(coord, offset)
    idx, dist = pointops.knn_query(16, xyz, off)               # positional call as in …v2m2_base.py:223
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32 and idx.shape == (xyz.shape[0], 16)
    idx64, _ = pointops.knn_query(16, xyz, off.long())          # int64 offsets (models/utils.py:28)
    assert torch.equal(idx, idx64)
    # the sqrt of query.py:24 is applied inside the kernel (AOPT_KNN_SQRT_DIST): same bits as torch.sqrt(dist2)
    for method in ("tile", "grid"):
        i2, d2 = pointops.knn_query_raw(16, xyz, off, method=method)
        i1, d1 = pointops.knn_query_raw(16, xyz, off, method=method, root=True)
        assert torch.equal(i1, i2) and torch.equal(d1, torch.sqrt(d2))
    big = torch.from_numpy(np.random.default_rng(0).random((3000, 3)).astype(np.float32)).cuda()
    boff = torch.tensor([3000], dtype=torch.int32, device="cuda")
    i2, d2 = pointops.knn_query_raw(40, big, boff)             # nsample > 32: the local-memory kernel
    i1, d1 = pointops.knn_query_raw(40, big, boff, root=True)
    assert torch.equal(i1, i2) and torch.equal(d1, torch.sqrt(d2))
    # padding convention of query.py:24: sqrt(1e10) = 1e5
    assert torch.all(dist[700:705, 5:] == 1e5) and torch.all(idx[700:705, 5:] == -1)
    with pytest.raises(ValueError):
        pointops.knn_query(129, xyz, off)
