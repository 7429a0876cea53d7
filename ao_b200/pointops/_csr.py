"""Transposed neighbour graph (CSR), built once per index tensor and cached on it.

One kNN result feeds every block of a BlockSequence
(/root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:223-225), so the
CSR that makes the backward passes atomic-free is built lazily on first backward use and memoised
on the idx tensor object itself.
"""
from __future__ import annotations

import torch

from .. import _lib


class NeighbourCSR:
    __slots__ = ("rowptr", "perm", "n_src", "mode", "version", "ready")

    def __init__(self, rowptr, perm, n_src, mode, version, ready=None):
        self.rowptr, self.perm, self.n_src, self.mode, self.version = rowptr, perm, n_src, mode, version
        self.ready = ready   # CUDA event of a side-stream build (prefetch_csr); consumers wait on it


def build_csr(idx: torch.Tensor, n_src: int, negative_mode: int = 0) -> NeighbourCSR:
    """idx: int32 CUDA tensor of any shape with values in [-1, n_src)."""
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    lib = _lib.load()
    n_entries = idx.numel()
    rowptr = torch.empty(n_src + 1, dtype=torch.int32, device=idx.device)
    perm = torch.empty(max(n_entries, 1), dtype=torch.int32, device=idx.device)
    with _lib.on_device(idx.device):
        ws = _lib.workspace(lib.aopt_csr_workspace_bytes(n_src, n_entries), idx.device)
        _lib.check(
            lib.aopt_csr_build(n_src, n_entries, _lib.ptr(idx), negative_mode, _lib.ptr(rowptr),
                               _lib.ptr(perm), _lib.ptr(ws), ws.numel(), _lib.stream()),
            "csr_build",
        )
    return NeighbourCSR(rowptr, perm, n_src, negative_mode, idx._version)


def _cache_of(idx):
    cache = getattr(idx, "_aopt_csr", None)
    if cache is None:
        cache = {}
        try:
            idx._aopt_csr = cache
        except Exception:  # pragma: no cover - tensors that refuse attributes
            return None
    return cache


def get_csr(idx: torch.Tensor, n_src: int, negative_mode: int = 0) -> NeighbourCSR:
    cache = _cache_of(idx)
    if cache is None:
        return build_csr(idx, n_src, negative_mode)
    key = (n_src, negative_mode)
    hit = cache.get(key)
    if hit is None or hit.version != idx._version:
        hit = build_csr(idx, n_src, negative_mode)
        cache[key] = hit
    elif hit.ready is not None:
        torch.cuda.current_stream(idx.device).wait_event(hit.ready)   # built on the side stream
    return hit


def prefetch_csr(idx: torch.Tensor, n_src: int, negative_mode: int = 0) -> None:
    """Starts the CSR build of `idx` NOW on a side stream, so that it overlaps the forward pass instead of
    sitting at the head of the backward pass (count / scan / fill / rank: small latency-bound kernels that use
    3 % of the HBM pipe).  The first backward that needs it waits on the build's event.  With overlap off this
    is a no-op (the CSR is then built lazily by get_csr, as before)."""
    if not _lib.overlap(role="geom") or idx.numel() == 0:
        return
    assert idx.dtype == torch.int32 and idx.is_contiguous() and idx.is_cuda
    cache = _cache_of(idx)
    if cache is None:
        return
    key = (n_src, negative_mode)
    hit = cache.get(key)
    if hit is not None and hit.version == idx._version:
        return
    dev = idx.device
    main = torch.cuda.current_stream(dev)
    side = _lib.side_stream(dev, "geom")
    side.wait_stream(main)                       # idx is produced on the caller's stream
    with torch.cuda.stream(side):
        csr = build_csr(idx, n_src, negative_mode)
        csr.ready = torch.cuda.Event()
        csr.ready.record(side)
    # memory handed across streams: idx is read by the side stream, rowptr / perm (side-stream pool) are read
    # by the caller's stream — the caching allocator must not recycle them under a kernel still in flight
    idx.record_stream(side)
    csr.rowptr.record_stream(main)
    csr.perm.record_stream(main)
    cache[key] = csr
