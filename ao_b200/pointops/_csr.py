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
    __slots__ = ("rowptr", "perm", "n_src", "mode", "version")

    def __init__(self, rowptr, perm, n_src, mode, version):
        self.rowptr, self.perm, self.n_src, self.mode, self.version = rowptr, perm, n_src, mode, version


def build_csr(idx: torch.Tensor, n_src: int, negative_mode: int = 0) -> NeighbourCSR:
    """idx: int32 CUDA tensor of any shape with values in [-1, n_src)."""
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    lib = _lib.load()
    n_entries = idx.numel()
    rowptr = torch.empty(n_src + 1, dtype=torch.int32, device=idx.device)
    perm = torch.empty(max(n_entries, 1), dtype=torch.int32, device=idx.device)
    with torch.cuda.device(idx.device):
        ws = _lib.workspace(lib.aopt_csr_workspace_bytes(n_src, n_entries), idx.device)
        _lib.check(
            lib.aopt_csr_build(n_src, n_entries, _lib.ptr(idx), negative_mode, _lib.ptr(rowptr),
                               _lib.ptr(perm), _lib.ptr(ws), ws.numel(), _lib.stream()),
            "csr_build",
        )
    return NeighbourCSR(rowptr, perm, n_src, negative_mode, idx._version)


def get_csr(idx: torch.Tensor, n_src: int, negative_mode: int = 0) -> NeighbourCSR:
    cache = getattr(idx, "_aopt_csr", None)
    if cache is None:
        cache = {}
        try:
            idx._aopt_csr = cache
        except Exception:  # pragma: no cover - tensors that refuse attributes
            return build_csr(idx, n_src, negative_mode)
    key = (n_src, negative_mode)
    hit = cache.get(key)
    if hit is None or hit.version != idx._version:
        hit = build_csr(idx, n_src, negative_mode)
        cache[key] = hit
    return hit
