"""knn_query — same signature and return as /root/reference/libs/pointops/functions/query.py:7-24,111.

ball_query / random_ball_query are outside the PTv2m2 hot path (no caller) and
raise NotImplementedError rather than silently falling back.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from .. import _lib

_METHODS = {"auto": _lib.KNN_AUTO, "tile": _lib.KNN_TILE, "grid": _lib.KNN_GRID}


def knn_query_raw(nsample, xyz, offset, new_xyz=None, new_offset=None, method=None, root=False):
    """Returns (idx int32 (m,k), dist2 float32 (m,k)) — squared distances, the kernel's own output.
    root=True: the kernel writes sqrt(dist2) instead (AOPT_KNN_SQRT_DIST), the reference's return (query.py:24)."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    if method is None:
        hit = _prefetched(int(nsample), xyz, new_xyz)
        if hit is not None:
            idx_h, d_h, is_root = hit
            if is_root == bool(root):
                return idx_h, d_h
            return (idx_h, torch.sqrt(d_h)) if root else (idx_h, d_h * d_h)
    dev = _lib.require_cuda(xyz, new_xyz, offset, new_offset)
    assert xyz.is_contiguous() and new_xyz.is_contiguous()
    if xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise ValueError("knn_query: coordinates must be float32")
    if xyz.dim() != 2 or xyz.shape[1] != 3 or new_xyz.dim() != 2 or new_xyz.shape[1] != 3:
        raise ValueError("knn_query: coordinates must be (n, 3)")
    nsample = int(nsample)
    if not 1 <= nsample <= 128:
        raise ValueError("knn_query: nsample must be in [1, 128] (reference kernel limit)")
    same = new_offset is offset
    offset = offset.int().contiguous()                       # query.py:22
    new_offset = offset if same else new_offset.int().contiguous()
    if offset.numel() != new_offset.numel():
        raise ValueError("knn_query: offset and new_offset must have the same number of scenes")
    n, m, b = xyz.shape[0], new_xyz.shape[0], offset.numel()
    if method is None:
        method = os.environ.get("AOPT_KNN_METHOD", "auto")
    meth = _METHODS[method] if isinstance(method, str) else int(method)
    if root:
        meth |= _lib.KNN_SQRT_DIST
    lib = _lib.load()
    idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
    dist2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
    if m == 0:
        return idx, dist2
    with _lib.on_device(dev):
        ws = _lib.workspace(lib.aopt_knn_workspace_bytes(n, m, b, nsample, meth), dev)
        _lib.check(
            lib.aopt_knn_query(m, nsample, n, b, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(offset),
                               _lib.ptr(new_offset), _lib.ptr(idx), _lib.ptr(dist2), meth, _lib.ptr(ws),
                               ws.numel(), _lib.stream()),
            "knn_query",
        )
    return idx, dist2


# ---- searches issued ahead of time on the geometry side stream (pointops.prepare_pyramid(..., knn=...)) ----------
def _knn_key(nsample, xyz, new_xyz):
    if new_xyz is xyz:
        return ("self", nsample, xyz.shape[0], xyz._version)
    return ("cross", nsample, new_xyz.data_ptr(), new_xyz.shape[0], new_xyz._version, xyz._version)


def prefetch_knn(nsample, xyz, offset, new_xyz=None, new_offset=None):
    """Starts the search NOW on the geometry side stream; the next knn_query / knn_query_raw / interpolation call
    with the same tensors returns its result (after making the caller's stream wait for it).  The neighbour
    search is ALU-bound and moves almost no memory, the streaming kernels of the feature path leave the ALUs idle:
    run side by side they overlap (unlike the CSR walk / CSR build, see _lib.py).  No-op with overlap off."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    if not _lib.overlap(role="knn") or new_xyz.shape[0] == 0:
        return
    dev = _lib.require_cuda(xyz, new_xyz, offset, new_offset)
    main = torch.cuda.current_stream(dev)
    side = _lib.side_stream(dev, "knn")
    side.wait_stream(main)                                   # coordinates / offsets are produced on the caller's stream
    with torch.cuda.stream(side):
        idx, dist2 = knn_query_raw(nsample, xyz, offset, new_xyz, new_offset, method="auto")
        ev = torch.cuda.Event()
        ev.record(side)
    for t in (xyz, new_xyz, offset, new_offset):
        t.record_stream(side)                                # read by the side stream
    cache = getattr(xyz, "_aopt_knn", None)
    if cache is None:
        cache = {}
        xyz._aopt_knn = cache
    cache[_knn_key(int(nsample), xyz, new_xyz)] = (idx, dist2, ev, False)


def _prefetched(nsample, xyz, new_xyz):
    cache = getattr(xyz, "_aopt_knn", None)
    if not cache:
        return None
    hit = cache.pop(_knn_key(nsample, xyz, new_xyz), None)   # one consumer: the result is handed over, not kept
    if hit is None:
        return None
    idx, dist, ev, is_root = hit
    if ev is not None:                                       # searched on the side stream
        cur = torch.cuda.current_stream(idx.device)
        cur.wait_event(ev)
        idx.record_stream(cur)                               # allocated in the side stream's pool
        dist.record_stream(cur)
    return idx, dist, is_root


def knn_query_sets(nsample, sets, root=False, method="auto"):
    """ONE search for several independent point sets (aopt_knn_query_multi): sets = [(xyz, offset, new_xyz, new_offset)],
    new_xyz / new_offset None for a self query (all sets self, or all cross).  The sets are concatenated scene by scene —
    a search never leaves its scene, so nothing changes for any query — and the indices come back relative to each set's
    own xyz.  Returns [(idx, dist2 or dist)] per set: row ranges of one (M, nsample) result.

    Used for the pyramid levels 1..L of a PTv2m2 forward (pooling.prepare_pyramid): one grid build + one query launch
    instead of one search (7 launches) per level; the coarse levels are launch-latency work."""
    nsample = int(nsample)
    if not 1 <= nsample <= 128:
        raise ValueError("knn_query: nsample must be in [1, 128] (reference kernel limit)")
    self_q = all(s[2] is None for s in sets)
    if not self_q and any(s[2] is None for s in sets):
        raise ValueError("knn_query_sets: self and cross queries cannot be mixed")
    tensors = [t for s in sets for t in s if t is not None]
    dev = _lib.require_cuda(*tensors)
    n_l = [s[0].shape[0] for s in sets]
    m_l = n_l if self_q else [s[2].shape[0] for s in sets]
    b_l = [s[1].numel() for s in sets]
    for s, bb in zip(sets, b_l):
        if s[0].dtype != torch.float32 or (s[2] is not None and (s[2].dtype != torch.float32 or s[3].numel() != bb)):
            raise ValueError("knn_query_sets: float32 coordinates and matching scene counts expected")
    base, qbase = [0], [0]
    for n, m in zip(n_l, m_l):
        base.append(base[-1] + n)
        qbase.append(qbase[-1] + m)
    xyz = torch.cat([s[0] for s in sets]).contiguous()
    offset = torch.cat([s[1].int() + base[i] for i, s in enumerate(sets)]).contiguous()
    if self_q:
        new_xyz, new_offset = xyz, offset
    else:
        new_xyz = torch.cat([s[2] for s in sets]).contiguous()
        new_offset = torch.cat([s[3].int() + qbase[i] for i, s in enumerate(sets)]).contiguous()
    index_base = torch.tensor([base[i] for i, bb in enumerate(b_l) for _ in range(bb)], dtype=torch.int32).to(dev, non_blocking=True)
    n, m, b = base[-1], qbase[-1], sum(b_l)
    meth = _METHODS[method] if isinstance(method, str) else int(method)
    if root:
        meth |= _lib.KNN_SQRT_DIST
    lib = _lib.load()
    idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
    dist = torch.empty((m, nsample), dtype=torch.float32, device=dev)
    if m > 0:
        with _lib.on_device(dev):
            ws = _lib.workspace(lib.aopt_knn_workspace_bytes(n, m, b, nsample, meth), dev)
            _lib.check(
                lib.aopt_knn_query_multi(m, nsample, n, b, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(offset),
                                         _lib.ptr(new_offset), _lib.ptr(index_base), _lib.ptr(idx), _lib.ptr(dist), meth,
                                         _lib.ptr(ws), ws.numel(), _lib.stream()),
                "knn_query_multi",
            )
    return [(idx[qbase[i]:qbase[i + 1]], dist[qbase[i]:qbase[i + 1]]) for i in range(len(sets))]


def stash_knn(nsample, xyz, new_xyz, idx, dist, is_root):
    """Hands a search result to the next knn_query / knn_query_raw / interpolation call on the same tensors."""
    cache = getattr(xyz, "_aopt_knn", None)
    if cache is None:
        cache = {}
        xyz._aopt_knn = cache
    cache[_knn_key(int(nsample), xyz, new_xyz if new_xyz is not None else xyz)] = (idx, dist, None, bool(is_root))


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, offset, new_xyz=None, new_offset=None):
        """
        input: xyz: (n, 3), new_xyz: (m, 3), offset: (b), new_offset: (b)
        output: idx: (m, nsample) -1 is placeholder, dist: (m, nsample)
        """
        idx, dist = knn_query_raw(nsample, xyz, offset, new_xyz, new_offset, root=True)   # sqrt of query.py:24 in-kernel
        ctx.mark_non_differentiable(idx)
        ctx.mark_non_differentiable(dist)
        return idx, dist

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None


knn_query = KNNQuery.apply


def ball_query(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.ball_query: outside the PTv2m2 hot path (not built)")


def random_ball_query(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.random_ball_query: outside the PTv2m2 hot path (not built)")
