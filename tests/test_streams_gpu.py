"""Side-stream overlap (ao_b200._lib.overlap): the CSR prefetch and the overlapped GVA backward must give
bit-identical results to the single-stream path, repeatedly and under allocator reuse, and the schedule's
gradients must not depend on the switch."""
import numpy as np
import pytest
import torch

from helpers import to_cuda

pytestmark = pytest.mark.gpu


def _case(seed, n=20000, k=16, c=48, g=6):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, n, (n, k)).astype(np.int32)
    idx[rng.integers(0, n, 50), k // 2:] = -1
    value = rng.standard_normal((n, c)).astype(np.float32)
    peb = rng.standard_normal((n, k, c)).astype(np.float32)
    logits = rng.standard_normal((n, k, g)).astype(np.float32)
    gout = rng.standard_normal((n, c)).astype(np.float32)
    return to_cuda(idx, value, peb, logits, gout)


def _aggregate_grads(tensors, g, prefetch):
    from ao_b200 import pointops

    idx, value, peb, logits, gout = tensors
    idx = idx.clone()                      # fresh tensor object: no cached CSR
    v, p, l = (t.clone().requires_grad_(True) for t in (value, peb, logits))
    if prefetch:
        pointops.prefetch_csr(idx, v.shape[0], 0)
    out = pointops.gva_aggregate(v, p, l, idx, g)
    gv, gp, gl = torch.autograd.grad(out, [v, p, l], gout)
    return out, gv, gp, gl


def test_overlapped_gva_backward_is_bit_identical():
    from ao_b200 import _lib

    t = _case(1)
    was = _lib.overlap_mode()
    try:
        _lib.overlap(False)
        ref = _aggregate_grads(t, 6, prefetch=False)
        _lib.overlap(True)
        for rep in range(6):               # repeated: allocator blocks get recycled across the two streams
            got = _aggregate_grads(t, 6, prefetch=(rep % 2 == 0))
            for a, b in zip(ref, got):
                assert torch.equal(a, b)
            junk = [torch.randn(1 << 20, device="cuda") for _ in range(4)]   # churn the caching allocator
            del junk
        torch.cuda.synchronize()
    finally:
        _lib.overlap_mode(was)


def test_prefetched_csr_equals_lazy_build():
    from ao_b200 import _lib, pointops

    (idx,) = to_cuda(np.random.default_rng(3).integers(-1, 5000, (7000, 8)).astype(np.int32))
    was = _lib.overlap_mode()
    try:
        _lib.overlap(True)
        a = idx.clone()
        pointops.prefetch_csr(a, 5000, 0)
        ca = pointops.get_csr(a, 5000, 0)
        _lib.overlap(False)
        b = idx.clone()
        pointops.prefetch_csr(b, 5000, 0)          # no-op when overlap is off
        assert not getattr(b, "_aopt_csr", None)
        cb = pointops.get_csr(b, 5000, 0)
        e = int(ca.rowptr[-1])                     # entries past rowptr[-1] (dropped -1 slots) are never written
        assert torch.equal(ca.rowptr, cb.rowptr) and torch.equal(ca.perm[:e], cb.perm[:e])
    finally:
        _lib.overlap_mode(was)


def test_schedule_step_does_not_depend_on_overlap():
    from ao_b200 import _lib, scenes
    from ao_b200.schedule import PointOpsSchedule, ScheduleConfig

    coord, _, off = scenes.s3dis_batch(2, n_points=6000)
    c, o = to_cuda(coord, off)
    was = _lib.overlap_mode()
    try:
        outs = []
        for on in (False, True, True):
            _lib.overlap(on)
            sched = PointOpsSchedule(ScheduleConfig.s3dis(), device="cuda", seed=0)
            outs.append(sched.step(c, o).clone())
            outs.append(sched.step(c, o).clone())   # second step: reused level tensors, recycled scratch
        torch.cuda.synchronize()
        for x in outs[1:]:
            assert torch.equal(outs[0], x)
    finally:
        _lib.overlap_mode(was)


def test_prepared_pyramid_equals_inline_grid_pool():
    """pointops.prepare_pyramid hoists the voxel partitions (the host syncs) in front of the feature path; grid_pool
    must return exactly what it returns without it, level after level, gradients included."""
    from ao_b200 import pointops, scenes

    coord, _, off = scenes.s3dis_batch(3, n_points=5000)
    grids = (0.1, 0.2, 0.4)
    rng = np.random.default_rng(0)

    def run(prepare):
        c, o = to_cuda(coord, off)
        if prepare:
            levels = pointops.prepare_pyramid(c, o, grids)
            assert len(levels) == len(grids) + 1
        outs = []
        for li, gs in enumerate(grids):
            feat = torch.from_numpy(np.abs(np.random.default_rng(li).standard_normal((c.shape[0], 8))).astype(np.float32)).cuda()
            feat.requires_grad_(True)
            (nc, nf, no), cluster = pointops.grid_pool(c, feat, o, gs)
            (g,) = torch.autograd.grad(nf, [feat], torch.ones_like(nf))
            outs += [nc.clone(), nf.detach().clone(), no.clone(), cluster.clone(), g.clone()]
            if prepare:
                assert nc is levels[li + 1][0]
            c, o = nc, no.int()          # an int32 cast of the offsets, as the model / schedule pass them on
        return outs

    a, b = run(False), run(True)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)


def test_prefetched_knn_equals_direct_search():
    from ao_b200 import _lib, pointops, scenes

    coord, _, off = scenes.s3dis_batch(2, n_points=7000)
    c, o = to_cuda(coord, off)
    fine, foff = c, o
    (coarse, _, coff), _ = pointops.grid_pool(c, c.clone(), o, 0.2)
    coff = coff.int()
    was = _lib.overlap_mode()
    try:
        _lib.overlap(False)
        ref_self = pointops.knn_query_raw(16, coarse, coff)
        ref_cross = pointops.knn_query_raw(3, coarse, coff, fine, foff)
        _lib.overlap(True)
        for _ in range(3):
            pointops.prefetch_knn(16, coarse, coff)
            pointops.prefetch_knn(3, coarse, coff, fine, foff)
            assert len(coarse._aopt_knn) == 2
            got_cross = pointops.knn_query_raw(3, coarse, coff, fine, foff)
            got_self = pointops.knn_query_raw(16, coarse, coff)
            assert len(coarse._aopt_knn) == 0                 # handed over, not kept
            for a, b in zip(ref_self + ref_cross, got_self + got_cross):
                assert torch.equal(a, b)
        torch.cuda.synchronize()
    finally:
        _lib.overlap_mode(was)
