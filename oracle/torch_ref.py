"""ORACLE — TEST INFRASTRUCTURE ONLY.

Pure-torch (CPU or CUDA tensor) restatement of the reference op chains that sit on the
PTv2m2 point-operator path.  Never imported by the product package ``ao_b200``; allowed users
are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference legs.

Each function cites the reference lines it follows (paths relative to /root/reference):

  knn_query            libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104 (via C oracle)
  grouping             libs/pointops/functions/grouping.py:36-60
  interpolation        libs/pointops/functions/interpolation.py:8-22
  gva_tail             pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:112-128
  grid_pool            .../point_transformer_v2m2_base.py:244-269
  segment_csr / voxel_grid   third-party (torch_scatter / torch_geometric→torch_cluster.grid_cluster),
                       NOT vendored and NOT pinned by the reference (no requirements file) →
                       "parity unpinned" for those two; restated from their published semantics
                       and anchored on the call sites above.
  vote_accumulate      pointcept/engines/test.py:106-113 (tester: softmax + per-fragment indexed add)
  aggregation (PTv1)   libs/pointops/src/aggregation/aggregation_cuda_kernel.cu:5-39
  subtraction          libs/pointops/src/subtraction/subtraction_cuda_kernel.cu:5-30

Pinning status: grouping/interpolation/offset helpers are checked against the imported reference
Python (tests/golden/make_golden.py, run in the build container, fixtures committed);
knn/grouping2/interpolation2/aggregation/subtraction are checked against the compiled reference
CUDA launchers (oracle/_ref) on the GPU box.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _oracle_lib() -> ctypes.CDLL:
    """Loads oracle/_build/liboracle.so (built by oracle/Makefile from knn_oracle.c)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        i, p = ctypes.c_int, ctypes.c_void_p
        for name in ("knn_oracle_heap", "knn_oracle_lex"):
            fn = getattr(lib, name)
            fn.argtypes = [i, i, p, p, p, p, p, p]
            fn.restype = i
        lib.knn_oracle_lex_range.argtypes = [i, i, i, p, p, p, p, p, p]
        lib.knn_oracle_lex_range.restype = i
        lib.fps_oracle.argtypes = [i, i, p, p, p, p, p]
        lib.fps_oracle.restype = i
        _LIB = lib
    return _LIB


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def _npi(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.int32)


def farthest_point_sampling(xyz, offset, new_offset) -> np.ndarray:
    """FarthestPointSampling.forward (libs/pointops/functions/sampling.py:9-24) on the CPU through
    oracle/fps_oracle.c (thread-by-thread emulation of sampling_cuda_kernel.cu).  Returns idx (m,) int32."""
    x, off, noff = _np32(xyz), _npi(offset), _npi(new_offset)
    b = off.shape[0]
    if b == 0:
        return np.zeros(0, np.int32)
    sizes = np.diff(np.concatenate([[0], off]))
    n_max = int(sizes.max())                                  # sampling.py:15-17
    idx = np.zeros(int(noff[-1]), np.int32)                   # sampling.py:18
    tmp = np.full(x.shape[0], 1e10, np.float32)               # sampling.py:19
    rc = _oracle_lib().fps_oracle(b, n_max, x.ctypes.data, off.ctypes.data, noff.ctypes.data, tmp.ctypes.data,
                                  idx.ctypes.data)
    if rc != 0:
        raise RuntimeError("fps_oracle failed")
    return idx


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None, rule: str = "lex",
              rows: Optional[Tuple[int, int]] = None):
    """C-oracle kNN.  Returns (idx int32 (m,k), dist2 float32 (m,k)) as numpy — note dist2,
    i.e. BEFORE the sqrt of libs/pointops/functions/query.py:24, so that bit-exactness can be
    asserted on the kernel's own output.

    rule = "heap": the reference's heap (tie order heap-dependent);
    rule = "lex":  ascending (d2, idx), the deterministic contract.
    rows = (begin, end): only those query rows (lex rule).
    """
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    x, q = _np32(xyz), _np32(new_xyz)
    off, noff = _npi(offset), _npi(new_offset)
    lib = _oracle_lib()
    if rows is None:
        m = q.shape[0]
        idx = np.empty((m, nsample), np.int32)
        d2 = np.empty((m, nsample), np.float32)
        fn = lib.knn_oracle_heap if rule == "heap" else lib.knn_oracle_lex
        rc = fn(m, nsample, x.ctypes.data, q.ctypes.data, off.ctypes.data, noff.ctypes.data,
                idx.ctypes.data, d2.ctypes.data)
    else:
        assert rule == "lex"
        b, e = rows
        idx = np.empty((e - b, nsample), np.int32)
        d2 = np.empty((e - b, nsample), np.float32)
        rc = lib.knn_oracle_lex_range(b, e, nsample, x.ctypes.data, q.ctypes.data, off.ctypes.data,
                                      noff.ctypes.data, idx.ctypes.data, d2.ctypes.data)
    if rc != 0:
        raise ValueError("nsample out of range [1,128]")
    return idx, d2


def knn_query_numpy(nsample, xyz, offset, new_xyz=None, new_offset=None):
    """Independent numpy restatement of the lex rule for SMALL inputs (cross-checks the C oracle).
    float32 arithmetic with an explicit fma emulated in float64 (exact for these operand sizes:
    a product of two float32 is exact in float64, and the sum is rounded once to float32 — the
    double-rounding cases of float64→float32 are excluded by the caller through tie-free inputs
    and verified equal to the C oracle in tests)."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    x, q = _np32(xyz), _np32(new_xyz)
    off, noff = _npi(offset), _npi(new_offset)
    m = q.shape[0]
    idx = np.full((m, nsample), -1, np.int32)
    d2o = np.full((m, nsample), np.float32(1e10), np.float32)
    for b in range(len(off)):
        s, e = (0 if b == 0 else off[b - 1]), off[b]
        qs, qe = (0 if b == 0 else noff[b - 1]), noff[b]
        if qe <= qs or e <= s:
            continue
        d = q[qs:qe, None, :] - x[None, s:e, :]           # float32 differences (FADD)
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        t = (dy * dy).astype(np.float32)                   # FMUL
        t = (dx.astype(np.float64) * dx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        d2 = (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        cand = np.arange(s, e, dtype=np.int64)
        order = np.lexsort((np.broadcast_to(cand, d2.shape), d2), axis=1)[:, :nsample]
        kk = order.shape[1]
        idx[qs:qe, :kk] = cand[order].astype(np.int32)
        d2o[qs:qe, :kk] = np.take_along_axis(d2, order, axis=1)
    return idx, d2o


# ----------------------------------------------------------------------------------------------
# offset-encoded batch layout: libs/pointops/functions/utils.py:102-119,
# pointcept/models/utils.py:11-28
# ----------------------------------------------------------------------------------------------
def offset2batch(offset: torch.Tensor) -> torch.Tensor:
    off = offset.detach().cpu().long()
    counts = torch.diff(off, prepend=off.new_zeros(1))
    return torch.repeat_interleave(torch.arange(len(off)), counts).long().to(offset.device)


def batch2offset(batch: torch.Tensor, long: bool = False) -> torch.Tensor:
    out = torch.cumsum(batch.bincount(), dim=0)
    return out.long() if long else out.int()


# ----------------------------------------------------------------------------------------------
# grouping: libs/pointops/functions/grouping.py:36-60
# ----------------------------------------------------------------------------------------------
def grouping(idx, feat, xyz, new_xyz=None, with_xyz=False):
    if new_xyz is None:
        new_xyz = xyz
    m, k, c = idx.shape[0], idx.shape[1], feat.shape[1]
    # :41-42 a zero row is appended so that idx == -1 selects zeros
    xyz_p = torch.cat([xyz, xyz.new_zeros(1, 3)], dim=0)
    feat_p = torch.cat([feat, feat.new_zeros(1, c)], dim=0)
    flat = idx.reshape(-1).long()
    g_feat = feat_p[flat].view(m, k, c)                      # :43-45
    if not with_xyz:
        return g_feat
    mask = torch.sign(idx + 1)                               # :49
    g_xyz = xyz_p[flat].view(m, k, 3) - new_xyz.unsqueeze(1)  # :50-54
    g_xyz = g_xyz * mask.unsqueeze(-1).to(g_xyz.dtype)       # :55-57 (einsum "n s c, n s -> n s c")
    return torch.cat((g_xyz, g_feat), dim=-1)                # :58


# grouping2 (CUDA Function): libs/pointops/src/grouping/grouping_cuda_kernel.cu:5-25 — gather with
# no -1 handling, backward = scatter-add.  idx must be >= 0.
def grouping2(inp, idx):
    return inp[idx.reshape(-1).long()].view(idx.shape[0], idx.shape[1], inp.shape[1])


# ----------------------------------------------------------------------------------------------
# interpolation: libs/pointops/functions/interpolation.py:8-22
# ----------------------------------------------------------------------------------------------
def interpolation_weights(dist: torch.Tensor) -> torch.Tensor:
    dist_recip = 1.0 / (dist + 1e-8)                         # :15
    norm = torch.sum(dist_recip, dim=1, keepdim=True)        # :16
    return dist_recip / norm                                 # :17


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3, idx_dist2=None):
    """idx_dist2: optional precomputed (idx, dist2) numpy pair (from knn_query) to skip the search."""
    if idx_dist2 is None:
        idx_dist2 = knn_query(k, xyz, offset, new_xyz, new_offset)
    idx = torch.from_numpy(idx_dist2[0]).to(feat.device)
    dist = torch.sqrt(torch.from_numpy(idx_dist2[1])).to(feat.device)   # query.py:24
    weight = interpolation_weights(dist)
    out = feat.new_zeros(new_xyz.shape[0], feat.shape[1], dtype=torch.float32)   # :19
    for i in range(k):                                       # :20-21 (python negative index wraps)
        out = out + feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return out


# interpolation2 (CUDA Function): libs/pointops/src/interpolation/interpolation_cuda_kernel.cu:5-33
def interpolation2_apply(inp, idx, weight):
    out = inp.new_zeros(idx.shape[0], inp.shape[1])
    for i in range(idx.shape[1]):
        out = out + inp[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return out


# ----------------------------------------------------------------------------------------------
# tester fragment vote: pointcept/engines/test.py:106-113
# ----------------------------------------------------------------------------------------------
def vote_accumulate(pred, logits, index, offset):
    """The tester's loop, verbatim in structure: softmax over the classes, then one indexed `+=` per fragment."""
    import torch.nn.functional as F

    pred_part = F.softmax(logits, -1)                        # :107
    bs = 0
    for be in offset:                                        # :110-113
        pred[index[bs:be], :] += pred_part[bs:be]
        bs = be
    return pred


# ----------------------------------------------------------------------------------------------
# GroupedVectorAttention tail: point_transformer_v2m2_base.py:112-128
# ----------------------------------------------------------------------------------------------
def gva_relation(key, query, idx):
    """relation_qk before the positional bias: key[idx] - query[:,None]  (:109,:112)."""
    c = key.shape[1]
    key_p = torch.cat([key, key.new_zeros(1, c)], dim=0)
    g = key_p[idx.reshape(-1).long()].view(idx.shape[0], idx.shape[1], c)
    return g - query.unsqueeze(1)


def gva_aggregate(value, peb, logits, idx, groups):
    """value: (N,C) un-gathered; peb: (N,k,C); logits: (N,k,G); idx: (N,k).
    Returns (N,C) = einsum("n s g i, n s g -> n g i") of (value[idx]+peb) with
    softmax_k(logits) * sign(idx+1)   (:110,:119,:122-128; attn_drop is identity at rate 0)."""
    n, k = idx.shape
    c = value.shape[1]
    val_p = torch.cat([value, value.new_zeros(1, c)], dim=0)
    v = val_p[idx.reshape(-1).long()].view(n, k, c) + peb          # :110,:119
    w = torch.softmax(logits, dim=1)                              # :122
    mask = torch.sign(idx + 1)                                    # :124
    w = w * mask.unsqueeze(-1).to(w.dtype)                        # :125
    v = v.view(n, k, groups, c // groups)                         # :126
    out = torch.einsum("nsgi,nsg->ngi", v, w)                      # :127
    return out.reshape(n, c)                                      # :128


# ----------------------------------------------------------------------------------------------
# third-party restatements (unpinned): torch_scatter.segment_csr, torch_cluster.grid_cluster
# ----------------------------------------------------------------------------------------------
def segment_csr(src: torch.Tensor, indptr: torch.Tensor, reduce: str):
    """Sequential per-segment reduce in storage order, as torch_scatter's segment_csr kernels do:
    sum starts at 0; mean = sum / max(count,1); min/max start at the dtype's extreme and update on
    strict </>; returns (out, arg) for min/max (arg = position in src of the selected element,
    src.shape[0] for empty segments)."""
    ptr = indptr.detach().cpu().long().tolist()
    nseg = len(ptr) - 1
    out = src.new_zeros((nseg,) + tuple(src.shape[1:]))
    arg = torch.full((nseg,) + tuple(src.shape[1:]), src.shape[0], dtype=torch.long, device=src.device)
    for s in range(nseg):
        a, b = ptr[s], ptr[s + 1]
        if b <= a:
            continue
        seg = src[a:b]
        if reduce in ("sum", "mean"):
            acc = torch.zeros_like(seg[0])
            for r in range(b - a):          # sequential order matters in float32
                acc = acc + seg[r]
            out[s] = acc / float(b - a) if reduce == "mean" else acc
        elif reduce == "max":
            v, i = seg.max(dim=0)           # torch.max returns the FIRST maximal index on CPU
            out[s], arg[s] = v, i + a
        elif reduce == "min":
            v, i = seg.min(dim=0)
            out[s], arg[s] = v, i + a
        else:
            raise ValueError(reduce)
    return (out, arg) if reduce in ("min", "max") else out


def voxel_grid_keys(pos: torch.Tensor, size: float, batch: torch.Tensor) -> torch.Tensor:
    """torch_geometric.nn.pool.voxel_grid(pos, size, batch, start=0) → torch_cluster.grid_cluster:
    4-D position (x,y,z,batch) in float32, voxel size (s,s,s,1), start 0, end = per-dim max;
    key = Σ_d trunc((p_d - start_d)/size_d) · stride_d, stride = exclusive cumprod of
    trunc((end_d - start_d)/size_d) + 1."""
    p4 = torch.cat([pos, batch.to(pos.dtype).unsqueeze(-1)], dim=1)
    sz = torch.tensor([size, size, size, 1.0], dtype=pos.dtype, device=pos.device)
    end = p4.max(dim=0).values
    nvox = (end / sz).long() + 1
    stride = torch.cumprod(nvox, 0)
    stride = torch.cat([stride.new_ones(1), stride[:-1]])
    coords = (p4 / sz).long()
    return (coords * stride).sum(dim=1)


def grid_pool(coord, feat, offset, grid_size):
    """GridPool.forward after the fc/norm/act (…v2m2_base.py:246-268).  `feat` is the already
    projected feature.  Returns (coord', feat', offset', cluster, argmax) with intra-voxel order =
    ascending original index (the reference's torch.sort(cluster) at :263 is unstable, so the order
    inside a voxel is unspecified there; a stable order is one valid instance)."""
    batch = offset2batch(offset)                                                   # :246
    ptr = torch.cat([batch.new_zeros(1), torch.cumsum(batch.bincount(), dim=0)])   # :251
    start, _ = segment_csr(coord, ptr, "min")                                      # :249-253
    keys = voxel_grid_keys(coord - start[batch], grid_size, batch)                 # :257-259
    unique, cluster, counts = torch.unique(keys, sorted=True, return_inverse=True,
                                           return_counts=True)                     # :260-262
    order = torch.sort(cluster, stable=True).indices                               # :263
    idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, dim=0)])        # :264
    new_coord = segment_csr(coord[order], idx_ptr, "mean")                         # :265
    new_feat, arg = segment_csr(feat[order], idx_ptr, "max")                       # :266
    new_batch = batch[idx_ptr[:-1]]                                                # :267
    new_offset = batch2offset(new_batch, long=True)                                # :268
    argmax = order[arg.clamp(max=order.numel() - 1)]                               # original point ids
    return new_coord, new_feat, new_offset, cluster, argmax


# ----------------------------------------------------------------------------------------------
# PTv1-layout fused ops (API parity): aggregation_cuda_kernel.cu:5-39, subtraction_cuda_kernel.cu:5-30
# ----------------------------------------------------------------------------------------------
def aggregation(inp, position, weight, idx):
    n, k, c = position.shape
    w_c = weight.shape[-1]
    g = inp[idx.reshape(-1).long()].view(n, k, c) + position
    w = weight.repeat(1, 1, c // w_c)          # channel ch uses weight[..., ch % w_c]  (:11)
    return (g * w).sum(dim=1)


def subtraction(input1, input2, idx):
    n, k = idx.shape
    c = input1.shape[1]
    return input1.unsqueeze(1) - input2[idx.reshape(-1).long()].view(n, k, c)
