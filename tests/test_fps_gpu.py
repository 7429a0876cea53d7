"""Farthest point sampling (the PTv1 caller) — csrc/fps.cu against oracle/fps_oracle.c and
against the unmodified reference launcher (sampling_cuda_kernel.cu), bit for bit, ties included."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import to_cuda

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    spec = importlib.util.spec_from_file_location("make_fps_golden_gpu", os.path.join(GOLD, "make_fps_golden_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", ["small_s4", "small_dup_s2", "small_all", "room_s4", "room_s16"])
@pytest.mark.parametrize("cluster", ["auto", "1", "2", "4", "8", "16"])
def test_fps_equals_oracle_for_every_cluster_size(oracle, monkeypatch, case, cluster):
    from ao_b200 import pointops

    if cluster == "auto":
        monkeypatch.delenv("AOPT_FPS_CLUSTER", raising=False)
    else:
        monkeypatch.setenv("AOPT_FPS_CLUSTER", cluster)
    xyz, off, noff = _cases().inputs(case)
    ref = oracle.farthest_point_sampling(xyz, off, noff)
    got = pointops.farthest_point_sampling(*to_cuda(xyz, off, noff)).cpu().numpy()
    assert got.dtype == np.int32 and got.shape == ref.shape
    assert np.array_equal(got, ref), f"first mismatch at {np.flatnonzero(got != ref)[:5]}"


@pytest.mark.parametrize("case", ["small_s4", "small_dup_s2", "small_all", "room_s4"])
def test_fps_equals_reference_cuda_kernel(oracle, case):
    """idx AND the final running distances `tmp` equal the unmodified reference launcher's."""
    from ao_b200 import _lib
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpointops_ref.so not built")
    xyz, off, noff = to_cuda(*_cases().inputs(case))
    ridx, rtmp = ref_cuda.farthest_point_sampling(xyz, off, noff)
    from ao_b200.pointops import _C

    b = off.numel()
    n_max = int(torch.diff(off, prepend=off.new_zeros(1)).max().item())
    idx = torch.zeros_like(ridx)
    tmp = torch.full_like(rtmp, 1e10)
    _C.farthest_point_sampling_cuda(b, n_max, xyz, off, noff, tmp, idx)
    torch.cuda.synchronize()
    assert torch.equal(idx, ridx)
    assert torch.equal(tmp.view(torch.int32), rtmp.view(torch.int32))
    assert np.array_equal(oracle.farthest_point_sampling(xyz.cpu().numpy(), off.cpu().numpy(), noff.cpu().numpy()), ridx.cpu().numpy())


def test_fps_full_size_rooms_match_reference_kernel():
    """BASELINE.json-sized rooms: 2 x 80k points -> 20k samples each (PTv1 stride 4), register-resident path,
    bit-equal to the reference kernel; and the defining property: every sample maximises the distance to
    the samples before it."""
    from ao_b200 import pointops, scenes
    from oracle import ref_cuda

    coord, _, off = scenes.s3dis_batch(2, 80000)
    noff = (off // 4).astype(np.int32)
    xyz, o, no = to_cuda(coord, off, noff)
    got = pointops.farthest_point_sampling(xyz, o, no)
    assert got.shape[0] == int(noff[-1])
    if ref_cuda.available():
        ridx, _ = ref_cuda.farthest_point_sampling(xyz, o, no)
        assert torch.equal(got, ridx)
    g = got.long()
    assert int(g[0]) == 0 and int(g[int(noff[0])]) == int(off[0])
    assert torch.unique(g).numel() == g.numel()              # continuous data: no point is taken twice
    assert bool(((g[: int(noff[0])] >= 0) & (g[: int(noff[0])] < int(off[0]))).all())
    # greedy property on a prefix of scene 0 (fp64 distances, tie-free data)
    pts = xyz[: int(off[0])].double()
    sel = g[:40]
    dmin = torch.full((pts.shape[0],), float("inf"), dtype=torch.float64, device=xyz.device)
    for j in range(39):
        dmin = torch.minimum(dmin, ((pts - pts[sel[j]]) ** 2).sum(1))
        assert abs(float(dmin[sel[j + 1]]) - float(dmin.max())) <= 1e-6 * float(dmin.max())


def test_fps_scene_larger_than_the_register_budget():
    """n_max > 16 x 512 x 20 points: the same kernel with the points left in global memory."""
    from ao_b200 import pointops
    from oracle import ref_cuda

    rng = np.random.default_rng(5)
    n = 16 * 512 * 24 + 777
    coord = (rng.random((n + 900, 3)) * np.array([30, 20, 3])).astype(np.float32)
    off = np.array([n, n + 900], np.int32)
    noff = np.array([48, 48 + 30], np.int32)
    xyz, o, no = to_cuda(coord, off, noff)
    got = pointops.farthest_point_sampling(xyz, o, no)
    if ref_cuda.available():
        ridx, _ = ref_cuda.farthest_point_sampling(xyz, o, no)
        assert torch.equal(got, ridx)
    assert int(got[0]) == 0 and int(got[48]) == n


def test_fps_python_api_signature_and_edge_cases():
    from ao_b200 import pointops

    xyz = torch.rand(50, 3, device="cuda")
    off = torch.tensor([20, 50], device="cuda")                       # int64 offsets are accepted (sampling.py:22)
    idx = pointops.farthest_point_sampling(xyz, off, torch.tensor([1, 2], device="cuda"))
    assert idx.dtype == torch.int32 and idx.tolist() == [0, 20]       # m = 1 per scene: the first point
    with pytest.raises(ValueError):
        pointops.farthest_point_sampling(xyz.cpu(), off.cpu(), off.cpu())
    with pytest.raises(ValueError):
        pointops.farthest_point_sampling(xyz.double(), off, off)
