"""`pointops._C` shim (ao_b200/pointops/_C.py): the reference's native-module surface
(/root/reference/libs/pointops/src/pointops_api.cpp:15-32) on top of libao_pointops.so."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from helpers import assert_knn_equal, to_cuda

REF_FUNCTIONS = "/root/reference/libs/pointops/functions"
HOT = ["knn_query_cuda", "grouping_forward_cuda", "grouping_backward_cuda", "interpolation_forward_cuda",
       "interpolation_backward_cuda", "subtraction_forward_cuda", "subtraction_backward_cuda",
       "aggregation_forward_cuda", "aggregation_backward_cuda", "farthest_point_sampling_cuda"]
COLD = ["ball_query_cuda", "random_ball_query_cuda",
        "attention_relation_step_forward_cuda", "attention_relation_step_backward_cuda",
        "attention_fusion_step_forward_cuda", "attention_fusion_step_backward_cuda"]


def test_shim_exports_the_reference_native_surface():
    from ao_b200.pointops import _C

    for name in HOT + COLD:
        assert callable(getattr(_C, name)), name
    for name in COLD:
        with pytest.raises(NotImplementedError):
            getattr(_C, name)()
    with pytest.raises(ValueError):          # CPU tensors are rejected, never silently computed
        _C.grouping_forward_cuda(1, 1, 4, torch.zeros(1, 4), torch.zeros(1, 1, dtype=torch.int32), torch.zeros(1, 1, 4))


@pytest.mark.skipif(not os.path.isdir(REF_FUNCTIONS), reason="reference checkout not present (GPU box)")
def test_reference_python_imports_on_top_of_the_shim(tmp_path):
    """The reference's own functions/*.py, packaged as `pointops` the way its setup.py does
    (libs/pointops/setup.py:23-24), imports with `pointops._C` provided by this repo."""
    import ao_b200

    pkg = tmp_path / "pointops"
    os.symlink(REF_FUNCTIONS, pkg)
    saved = {k: v for k, v in sys.modules.items() if k == "pointops" or k.startswith("pointops.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(tmp_path))
    try:
        ao_b200.install_native_only()
        ref = importlib.import_module("pointops")
        assert ref.__file__.startswith(str(tmp_path))
        from ao_b200.pointops import _C

        assert sys.modules["pointops._C"] is _C
        assert sys.modules["pointops.query"].knn_query_cuda is _C.knn_query_cuda
        assert sys.modules["pointops.grouping"].grouping_forward_cuda is _C.grouping_forward_cuda
        assert sys.modules["pointops.interpolation"].interpolation_backward_cuda is _C.interpolation_backward_cuda
        # the pure-torch entry points of the reference still work on CPU tensors
        idx = torch.tensor([[0, 1], [1, -1]], dtype=torch.int32)
        out = ref.grouping(idx, torch.arange(6.).view(2, 3), torch.zeros(2, 3), with_xyz=False)
        assert out.shape == (2, 2, 3) and torch.equal(out[1, 1], torch.zeros(3))
    finally:
        sys.path.remove(str(tmp_path))
        for k in [k for k in sys.modules if k == "pointops" or k.startswith("pointops.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_shim_against_reference_launchers(oracle):
    """Same tensors through pointops._C (this repo) and through the unmodified reference launchers."""
    from ao_b200 import scenes
    from ao_b200.pointops import _C
    from oracle import ref_cuda

    coord, feat, offset = scenes.small_batch(11, sizes=(900, 40, 1500))
    xyz, off = to_cuda(coord, offset)
    m, k, c = xyz.shape[0], 8, 20
    idx = torch.zeros((m, k), dtype=torch.int32, device="cuda")
    d2 = torch.zeros((m, k), dtype=torch.float32, device="cuda")
    _C.knn_query_cuda(m, k, xyz, xyz, off, off, idx, d2)
    ri, rd = oracle.knn_query(k, coord, offset, rule="lex")
    assert_knn_equal(idx.cpu().numpy(), d2.cpu().numpy(), ri, rd)
    torch.manual_seed(3)
    inp = torch.randn(m, c, device="cuda")
    out = torch.zeros(m, k, c, device="cuda")
    _C.grouping_forward_cuda(m, k, c, inp, idx, out)
    go = torch.randn(m, k, c, device="cuda")
    gi = torch.zeros(m, c, device="cuda")
    _C.grouping_backward_cuda(m, k, c, go, idx, gi)
    w = torch.rand(m, k, device="cuda")
    io = torch.zeros(m, c, device="cuda")
    _C.interpolation_forward_cuda(m, c, k, inp, idx, w, io)
    ig = torch.zeros(m, c, device="cuda")
    _C.interpolation_backward_cuda(m, c, k, go[:, 0].contiguous(), idx, w, ig)
    sub = torch.zeros(m, k, c, device="cuda")
    inp2 = torch.randn(m, c, device="cuda")
    _C.subtraction_forward_cuda(m, k, c, inp, inp2, idx, sub)
    g1, g2 = torch.zeros(m, c, device="cuda"), torch.zeros(m, c, device="cuda")
    _C.subtraction_backward_cuda(m, k, c, idx, go, g1, g2)
    # torch restatements
    assert torch.equal(out, inp[idx.long()])
    ref_gi = torch.zeros(m, c, device="cuda").index_add_(0, idx.view(-1).long(), go.view(-1, c))
    assert torch.allclose(gi, ref_gi, rtol=1e-5, atol=1e-5)
    assert torch.allclose(io, (inp[idx.long()] * w.unsqueeze(-1)).sum(1), rtol=1e-5, atol=1e-5)
    assert torch.equal(sub, inp.unsqueeze(1) - inp2[idx.long()])
    assert torch.allclose(g1, go.sum(1), rtol=1e-5, atol=1e-5)
    # accumulate semantics of the reference launchers: a second call adds
    _C.grouping_backward_cuda(m, k, c, go, idx, gi)
    assert torch.allclose(gi, 2 * ref_gi, rtol=1e-5, atol=2e-5)
    if ref_cuda.available():
        assert torch.equal(out, ref_cuda.grouping_forward(inp, idx))
        assert torch.allclose(io, ref_cuda.interpolation_forward(inp, idx, w), rtol=1e-5, atol=1e-5)
        assert torch.allclose(ig, ref_cuda.interpolation_backward(go[:, 0].contiguous(), idx, w, m), rtol=1e-4, atol=1e-4)
        assert torch.equal(sub, ref_cuda.subtraction_forward(inp, inp2, idx))
        q1, q2 = ref_cuda.subtraction_backward(idx, go, m)
        assert torch.allclose(g1, q1, rtol=1e-5, atol=1e-5) and torch.allclose(g2, q2, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_reused_index_buffer_does_not_keep_a_stale_transposed_graph():
    """A caller that refills ONE preallocated idx buffer (the reference's pattern: query.py:19 allocates, the launcher
    writes through the raw pointer) must get gradients of the NEW neighbour lists: the library's writes do not bump
    the tensor version, so `_C.knn_query_cuda` drops the transposed graph cached on the buffer."""
    from ao_b200 import scenes
    from ao_b200.pointops import _C

    c, k = 12, 8
    idx = torch.zeros((1500, k), dtype=torch.int32, device="cuda")
    d2 = torch.zeros((1500, k), dtype=torch.float32, device="cuda")
    for seed in (21, 22):
        coord, _, offset = scenes.small_batch(seed, sizes=(700, 800))
        xyz, off = to_cuda(coord, offset)
        m = xyz.shape[0]
        _C.knn_query_cuda(m, k, xyz, xyz, off, off, idx, d2)
        go = torch.randn(m, k, c, device="cuda")
        gi = torch.zeros(m, c, device="cuda")
        _C.grouping_backward_cuda(m, k, c, go, idx, gi)      # builds (and caches) the CSR of the current idx
        ref = torch.zeros(m, c, device="cuda").index_add_(0, idx.view(-1).long(), go.view(-1, c))
        assert torch.allclose(gi, ref, rtol=1e-5, atol=1e-5), f"stale CSR after refill (seed {seed})"
