#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN PYTHON (imported from /root/reference,
CPU tensors) on seeded inputs.  Run in the build container only (the GPU box has no /root/reference);
the fixtures it writes are committed.

What is real reference code here, and what is substituted:
  * libs/pointops/functions/{grouping,interpolation,utils}.py — imported unmodified.
  * pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py — imported unmodified
    (classes PointBatchNorm, GroupedVectorAttention, Block, UnpoolWithSkip).
  * `pointops._C` (CUDA-only) is stubbed; `knn_query` is replaced by the C oracle
    (oracle/knn_oracle.c, lex rule) — the kNN itself is pinned on the GPU against oracle/_ref.
  * torch.cuda.FloatTensor → torch.FloatTensor so `interpolation` runs on CPU.
  * torch_geometric / torch_scatter / timm are absent: stub modules satisfy the imports; no fixture
    exercises them (GridPool parity is "unpinned", see oracle/torch_ref.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from oracle import torch_ref  # noqa: E402


def _load(name, path, package=False):
    spec = importlib.util.spec_from_file_location(
        name, path, submodule_search_locations=[os.path.dirname(path)] if package else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    # ---- pointops (python) with a stubbed _C ----
    c = types.ModuleType("pointops._C")
    for n in ("knn_query_cuda random_ball_query_cuda ball_query_cuda farthest_point_sampling_cuda "
              "grouping_forward_cuda grouping_backward_cuda interpolation_forward_cuda interpolation_backward_cuda "
              "subtraction_forward_cuda subtraction_backward_cuda aggregation_forward_cuda aggregation_backward_cuda "
              "attention_relation_step_forward_cuda attention_relation_step_backward_cuda "
              "attention_fusion_step_forward_cuda attention_fusion_step_backward_cuda").split():
        setattr(c, n, None)
    sys.modules["pointops._C"] = c
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.IntTensor = torch.IntTensor
    pointops = _load("pointops", os.path.join(REF, "libs/pointops/functions/__init__.py"), package=True)

    def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None):
        idx, d2 = torch_ref.knn_query(nsample, xyz, offset, new_xyz, new_offset, rule="lex")
        return torch.from_numpy(idx), torch.sqrt(torch.from_numpy(d2))

    pointops.knn_query = knn_query
    sys.modules["pointops.interpolation"].knn_query = knn_query
    sys.modules["pointops.utils"].knn_query = knn_query

    # ---- model file with stubbed third-party imports ----
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("torch_geometric"); stub("torch_geometric.nn"); stub("torch_geometric.nn.pool", voxel_grid=None)
    stub("torch_scatter", segment_csr=None)
    stub("timm"); stub("timm.models"); stub("timm.models.layers", DropPath=torch.nn.Identity)

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    stub("pointcept"); stub("pointcept.models")
    stub("pointcept.models.builder", MODELS=_Reg())
    _load("pointcept.models.utils", os.path.join(REF, "pointcept/models/utils.py"))
    model = _load("ref_ptv2m2", os.path.join(REF, "pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py"))
    return pointops, model


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path)/1024:.1f} KiB")


def main():
    pointops, model = import_reference()
    rng = np.random.default_rng(20241017)

    # ---------------- offset helpers ----------------
    offset = torch.tensor([5, 5 + 1, 6 + 9, 15 + 3], dtype=torch.int32)
    batch = pointops.offset2batch(offset)
    save("offsets", offset=offset, batch=batch, back=pointops.batch2offset(batch))

    # ---------------- grouping (pure torch), with -1 padding ----------------
    n, m, k, c = 40, 23, 5, 8
    xyz = torch.from_numpy(rng.standard_normal((n, 3)).astype(np.float32))
    new_xyz = torch.from_numpy(rng.standard_normal((m, 3)).astype(np.float32))
    feat = torch.from_numpy(rng.standard_normal((n, c)).astype(np.float32)).requires_grad_(True)
    idx = torch.from_numpy(rng.integers(0, n, (m, k)).astype(np.int32))
    idx[3, 2:] = -1
    idx[7, :] = -1
    idx[20, 4] = -1
    out_xyz = pointops.grouping(idx, feat, xyz, new_xyz, with_xyz=True)
    g = torch.from_numpy(rng.standard_normal(tuple(out_xyz.shape)).astype(np.float32))
    (grad_feat,) = torch.autograd.grad(out_xyz, feat, g)
    out_plain = pointops.grouping(idx, feat, xyz, new_xyz, with_xyz=False)
    save("grouping", xyz=xyz, new_xyz=new_xyz, feat=feat, idx=idx, out_xyz=out_xyz, out_plain=out_plain,
         grad_out=g, grad_feat=grad_feat)

    # ---------------- interpolation (reference python, oracle kNN), one scene shorter than k ----------------
    sizes_c, sizes_f = [30, 2, 17], [70, 9, 40]
    cx = torch.from_numpy(rng.uniform(-1, 1, (sum(sizes_c), 3)).astype(np.float32))
    fx = torch.from_numpy(rng.uniform(-1, 1, (sum(sizes_f), 3)).astype(np.float32))
    cf = torch.from_numpy(rng.standard_normal((sum(sizes_c), 12)).astype(np.float32)).requires_grad_(True)
    off_c = torch.tensor(np.cumsum(sizes_c), dtype=torch.int32)
    off_f = torch.tensor(np.cumsum(sizes_f), dtype=torch.int32)
    out = pointops.interpolation(cx, fx, cf, off_c, off_f)
    g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32))
    (grad_cf,) = torch.autograd.grad(out, cf, g)
    kidx, kd2 = torch_ref.knn_query(3, cx, off_c, fx, off_f)
    save("interpolation", xyz=cx, new_xyz=fx, feat=cf, offset=off_c, new_offset=off_f, out=out, grad_out=g,
         grad_feat=grad_cf, knn_idx=kidx, knn_dist2=kd2)

    # ---------------- GroupedVectorAttention + Block (reference modules, train-mode BN) ----------------
    torch.manual_seed(4242)
    sizes = [120, 9, 90]          # the middle scene has fewer points than k → -1 padded neighbours
    npts, k, C, G = sum(sizes), 16, 48, 6
    coord = torch.from_numpy(rng.uniform(-1, 1, (npts, 3)).astype(np.float32))
    off = torch.tensor(np.cumsum(sizes), dtype=torch.int32)
    ref_idx, _ = pointops.knn_query(k, coord, off)
    x = torch.from_numpy(rng.standard_normal((npts, C)).astype(np.float32)).requires_grad_(True)
    gva = model.GroupedVectorAttention(embed_channels=C, groups=G, attn_drop_rate=0.0, qkv_bias=True,
                                       pe_multiplier=False, pe_bias=True)
    gva.train()
    y = gva(x, coord, ref_idx)
    gy = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    params = list(gva.parameters())
    grads = torch.autograd.grad(y, [x] + params, gy)
    arrays = dict(coord=coord, offset=off, idx=ref_idx, x=x, y=y, grad_y=gy, grad_x=grads[0])
    for (name, p), gparam in zip(gva.named_parameters(), grads[1:]):
        arrays["param." + name] = p
        arrays["grad." + name] = gparam
    save("gva_module", **arrays)

    torch.manual_seed(7)
    blk = model.Block(embed_channels=C, groups=G, qkv_bias=True, pe_multiplier=False, pe_bias=True,
                      attn_drop_rate=0.0, drop_path_rate=0.0)
    blk.train()
    x2 = torch.from_numpy(rng.standard_normal((npts, C)).astype(np.float32)).requires_grad_(True)
    _, y2, _ = blk([coord, x2, off], ref_idx)
    gy2 = torch.from_numpy(rng.standard_normal(tuple(y2.shape)).astype(np.float32))
    params = list(blk.parameters())
    grads = torch.autograd.grad(y2, [x2] + params, gy2)
    arrays = dict(coord=coord, offset=off, idx=ref_idx, x=x2, y=y2, grad_y=gy2, grad_x=grads[0])
    for (name, p), gparam in zip(blk.named_parameters(), grads[1:]):
        arrays["param." + name] = p
        arrays["grad." + name] = gparam
    save("block_module", **arrays)

    # ---------------- UnpoolWithSkip, interp backend (reference module) ----------------
    torch.manual_seed(11)
    up = model.UnpoolWithSkip(in_channels=24, skip_channels=12, out_channels=12, backend="interp")
    up.train()
    feat_c = torch.from_numpy(rng.standard_normal((sum(sizes_c), 24)).astype(np.float32)).requires_grad_(True)
    feat_s = torch.from_numpy(rng.standard_normal((sum(sizes_f), 12)).astype(np.float32)).requires_grad_(True)
    _, yu, _ = up([cx, feat_c, off_c], [fx, feat_s, off_f])
    gyu = torch.from_numpy(rng.standard_normal(tuple(yu.shape)).astype(np.float32))
    params = list(up.parameters())
    grads = torch.autograd.grad(yu, [feat_c, feat_s] + params, gyu)
    arrays = dict(coord=cx, offset=off_c, skip_coord=fx, skip_offset=off_f, feat=feat_c, skip_feat=feat_s, y=yu,
                  grad_y=gyu, grad_feat=grads[0], grad_skip_feat=grads[1])
    for (name, p), gparam in zip(up.named_parameters(), grads[2:]):
        arrays["param." + name] = p
        arrays["grad." + name] = gparam
    save("unpool_interp_module", **arrays)


if __name__ == "__main__":
    main()
