"""Other callers of the same pointops API: each test is written the way the reference call
site uses the package, on CUDA tensors, and checked against the oracle.

  evaluator        pointcept/engines/hooks/evaluator.py:125-131     k = 1 cross-set kNN, pred[idx]
  tester vote      pointcept/engines/test.py:106-113                softmax + per-fragment `pred[index] += prob`
  MSC matching     pointcept/models/masked_scene_contrast/masked_scene_contrast_v1m1_base.py:144-153
  PTv1 layer       pointcept/models/point_transformer/point_transformer_seg.py:47-60   knn_query_and_group x2
  PTv1 down        …/point_transformer_seg.py:94-111    farthest_point_sampling + knn_query_and_group
  PTv1 up          …/point_transformer_seg.py:165       interpolation (coarse -> fine)
  CAC backbone     pointcept/models/context_aware_classifier/context_aware_classifier_v1m1_base.py:200-202
                   (PT-v2m2 built with num_classes = 0 returns the decoder features)
"""
import numpy as np
import pytest
import torch

from helpers import to_cuda

pytestmark = pytest.mark.gpu


def _two_scenes(seed, sizes=(900, 400)):
    from ao_b200 import scenes

    coord, feat, off = scenes.small_batch(seed, sizes=sizes)
    return coord, feat, off


def test_evaluator_maps_predictions_back_to_the_original_points(oracle):
    """hooks/evaluator.py:125-131: idx, _ = knn_query(1, coord, offset, origin_coord, origin_offset);
    pred = pred[idx.flatten().long()]."""
    from ao_b200 import pointops

    coord, _, off = _two_scenes(41)
    rng = np.random.default_rng(0)
    # "origin" = the un-voxelised cloud: 3 jittered copies of every kept point
    origin, o_off, s = [], [], 0
    for e in off:
        pts = np.repeat(coord[s:e], 3, axis=0) + rng.normal(0, 0.01, (3 * (e - s), 3)).astype(np.float32)
        origin.append(pts.astype(np.float32)); o_off.append(len(pts)); s = e
    origin, o_off = np.concatenate(origin), np.cumsum(o_off).astype(np.int32)
    c, o, oc, oo = to_cuda(coord, off, origin, o_off)
    idx, dist = pointops.knn_query(1, c.float(), o.int(), oc.float(), oo.int())
    ri, rd2 = oracle.knn_query(1, coord, off, origin, o_off, rule="lex")
    assert np.array_equal(idx.cpu().numpy(), ri)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))          # query.py:24 returns sqrt(dist2)
    pred = torch.arange(c.shape[0], device="cuda") % 13
    mapped = pred[idx.flatten().long()]
    assert mapped.shape[0] == origin.shape[0]
    assert np.array_equal(mapped.cpu().numpy(), (np.arange(coord.shape[0]) % 13)[ri[:, 0]])


def test_msc_contrastive_pair_matching(oracle):
    """masked_scene_contrast_v1m1_base.py:144-153: knn_query(max_k, view2, view2_offset, view1, view1_offset),
    then (row, index) pairs with distance < max_radius."""
    from ao_b200 import pointops

    coord, _, off = _two_scenes(42, sizes=(700, 300))
    rng = np.random.default_rng(1)
    v2, v2o, s = [], [], 0
    for e in off:                                                     # view 2 = a jittered subset of view 1
        keep = rng.random(e - s) < 0.7
        v2.append((coord[s:e][keep] + rng.normal(0, 0.02, (int(keep.sum()), 3))).astype(np.float32))
        v2o.append(int(keep.sum())); s = e
    v2, v2o = np.concatenate(v2), np.cumsum(v2o).astype(np.int32)
    max_k, max_radius = 8, 0.1
    c1, o1, c2, o2 = to_cuda(coord, off, v2, v2o)
    index, distance = pointops.knn_query(max_k, c2.float(), o2.int(), c1.float(), o1.int())
    ri, rd2 = oracle.knn_query(max_k, v2, v2o, coord, off, rule="lex")
    assert np.array_equal(index.cpu().numpy(), ri)
    assert np.array_equal(distance.cpu().numpy(), np.sqrt(rd2))       # query.py:24: sqrt of the kernel's dist2
    pairs = torch.cat([torch.arange(index.shape[0], device="cuda", dtype=torch.long).view(-1, 1, 1).expand(-1, max_k, 1),
                       index.view(-1, max_k, 1)], dim=-1)[distance.squeeze(-1) < max_radius]
    ref_rows, ref_cols = np.nonzero(np.sqrt(rd2) < max_radius)
    assert np.array_equal(pairs[:, 0].cpu().numpy(), ref_rows)
    assert np.array_equal(pairs[:, 1].cpu().numpy(), ri[ref_rows, ref_cols])
    # matched points lie in the same scene (offset-encoded batch layout is honoured)
    scene1 = np.searchsorted(off, pairs[:, 0].cpu().numpy(), side="right")
    scene2 = np.searchsorted(v2o, pairs[:, 1].cpu().numpy(), side="right")
    assert np.array_equal(scene1, scene2)


def test_ptv1_layer_and_transitions(oracle):
    """PointTransformerLayer (…seg.py:47-60), TransitionDown (:94-111) and TransitionUp (:165) call patterns."""
    from ao_b200 import pointops

    coord, feat, off = _two_scenes(43, sizes=(1200, 500))
    p, x, o = to_cuda(coord, np.ascontiguousarray(np.tile(feat, (1, 4))[:, :8]), off)
    x = x.clone().requires_grad_(True)
    nsample, stride = 8, 4
    # layer: grouped keys with relative xyz in front, grouped values re-using idx
    x_k, idx = pointops.knn_query_and_group(x, p, o, new_xyz=p, new_offset=o, nsample=nsample, with_xyz=True)
    x_v, _ = pointops.knn_query_and_group(x, p, o, new_xyz=p, new_offset=o, idx=idx, nsample=nsample, with_xyz=False)
    ri, _ = oracle.knn_query(nsample, coord, off, rule="lex")
    assert np.array_equal(idx.cpu().numpy(), ri)
    ref_k = oracle.grouping(torch.from_numpy(ri), x.detach().cpu(), torch.from_numpy(coord), torch.from_numpy(coord), with_xyz=True)
    assert x_k.shape == (p.shape[0], nsample, 3 + 8) and torch.equal(x_k.detach().cpu(), ref_k)
    assert torch.equal(x_v.detach().cpu(), ref_k[:, :, 3:])
    # down: FPS to n // stride points per scene, then group the old features around the new points
    n_o, count = [int(o[0].item()) // stride], int(o[0].item()) // stride
    for i in range(1, o.shape[0]):
        count += (int(o[i].item()) - int(o[i - 1].item())) // stride
        n_o.append(count)
    n_o = torch.tensor(n_o, dtype=torch.int32, device="cuda")
    fidx = pointops.farthest_point_sampling(p, o, n_o)
    assert np.array_equal(fidx.cpu().numpy(), oracle.farthest_point_sampling(coord, off, n_o.cpu().numpy()))
    n_p = p[fidx.long(), :]
    xg, gidx = pointops.knn_query_and_group(x, p, offset=o, new_xyz=n_p, new_offset=n_o, nsample=nsample, with_xyz=True)
    ri2, _ = oracle.knn_query(nsample, coord, off, n_p.cpu().numpy(), n_o.cpu().numpy(), rule="lex")
    assert np.array_equal(gidx.cpu().numpy(), ri2)
    ref_g = oracle.grouping(torch.from_numpy(ri2), x.detach().cpu(), torch.from_numpy(coord), n_p.cpu(), with_xyz=True)
    assert torch.equal(xg.detach().cpu(), ref_g)
    pooled = xg[:, :, 3:].max(1)[0]
    # up: interpolate the coarse features back onto the fine points; gradient reaches x through both ops
    up = pointops.interpolation(n_p, p, pooled, n_o, o)
    ref_up = oracle.interpolation(n_p.cpu(), torch.from_numpy(coord), pooled.detach().cpu(), n_o.cpu(), torch.from_numpy(off))
    assert torch.allclose(up.detach().cpu(), ref_up, rtol=1e-5, atol=2e-5)
    (gx,) = torch.autograd.grad(up.sum() + x_v.sum(), [x])
    assert torch.isfinite(gx).all() and float(gx.abs().sum()) > 0


def test_cac_feature_backbone_num_classes_zero():
    """context_aware_classifier_v1m1_base.py:200-202 runs PT-v2m2 with num_classes = 0 as a feature backbone:
    the output is the last decoder's features (N, dec_channels[0])."""
    from ao_b200 import ptv2, scenes

    coord, feat, off = scenes.s3dis_batch(2, n_points=3000)
    cfg = dict(ptv2.S3DIS_CFG, num_classes=0, drop_path_rate=0.0)
    model = ptv2.PointTransformerV2(**cfg).cuda().train()
    c, f, o = to_cuda(coord, feat, off)
    out = model(dict(coord=c, feat=f, offset=o))
    assert out.shape == (coord.shape[0], cfg["dec_channels"][0])
    out.square().mean().backward()
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("classes", [13, 20, 19, 1, 40])
def test_tester_fragment_vote(oracle, classes):
    """engines/test.py:106-113: pred_part = softmax(logits); for every fragment of the batch
    pred[idx_part[bs:be], :] += pred_part[bs:be].  Fragments come from the test-mode GridSample
    (transform.py:834-837): fragment i takes point (i mod count) of every voxel, so indices are distinct inside a
    fragment and the same point is voted for by several fragments.  Tolerance: fp32, rtol 1e-6 / atol 1e-7 against
    the torch ops on the same GPU (same expf, the class sum in a different order)."""
    from ao_b200 import pointops

    rng = np.random.default_rng(100 + classes)
    n_pred, n_vox, n_frag = 5000, 1200, 5
    voxel_of = rng.integers(0, n_vox, n_pred)
    order = np.argsort(voxel_of, kind="stable")
    count = np.bincount(voxel_of, minlength=n_vox)
    keep = count > 0
    starts = np.cumsum(np.insert(count, 0, 0)[:-1])
    frags = [order[(starts + i % np.maximum(count, 1))[keep]] for i in range(n_frag)]
    index = np.concatenate(frags).astype(np.int64)
    offset = np.cumsum([len(f) for f in frags])
    for f in frags:
        assert len(np.unique(f)) == len(f)
    logits = (rng.normal(0, 4, (len(index), classes))).astype(np.float32)
    lg, ix = to_cuda(logits, index)
    pred = torch.zeros(n_pred, classes, device="cuda")
    ref = torch.zeros(n_pred, classes, device="cuda")
    for _ in range(2):                                   # two batches into the same accumulator (:102-113)
        out = pointops.vote_accumulate(pred, lg, ix, torch.from_numpy(offset))
        oracle.vote_accumulate(ref, lg, ix, offset.tolist())
    assert out is pred
    assert torch.allclose(pred, ref, rtol=1e-6, atol=1e-7)
    votes = np.bincount(index, minlength=n_pred) * 2
    assert torch.allclose(pred.sum(1).cpu(), torch.from_numpy(votes).float(), atol=1e-4)   # every vote sums to 1
    assert torch.equal(pred.argmax(1)[votes > 0], ref.argmax(1)[votes > 0])
    # one fragment (offset=None), int32 index, bf16 logits (upcast like .float()), negative (wrapping) index
    one = torch.zeros(n_pred, classes, device="cuda")
    pointops.vote_accumulate(one, lg[: offset[0]].bfloat16(), ix[: offset[0]].int())
    ref1 = torch.zeros(n_pred, classes, device="cuda")
    oracle.vote_accumulate(ref1, lg[: offset[0]].bfloat16().float(), ix[: offset[0]], [int(offset[0])])
    assert torch.allclose(one, ref1, rtol=1e-6, atol=1e-7)
    neg = torch.zeros(4, classes, device="cuda")
    pointops.vote_accumulate(neg, lg[:1], torch.tensor([-1], device="cuda"))
    assert float(neg[3].sum()) == pytest.approx(1.0, abs=1e-6) and float(neg[:3].abs().sum()) == 0.0
    with pytest.raises(IndexError):
        pointops.vote_accumulate(neg, lg[:1], torch.tensor([4], device="cuda"), check_index=True)
    assert float(neg[:3].abs().sum()) == 0.0
    # empty fragment list / empty batch
    pointops.vote_accumulate(neg, lg[:0], ix[:0])
    with pytest.raises(ValueError):
        pointops.vote_accumulate(neg, lg[:2], ix[:1])
