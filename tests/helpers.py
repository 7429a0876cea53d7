"""Shared helpers for the parity tests."""
import numpy as np
import torch


def tie_rows(d2: np.ndarray) -> np.ndarray:
    """Rows of a (m,k) ascending dist2 matrix that contain an exact tie among real entries."""
    real = d2 < 1e9
    eq = (d2[:, 1:] == d2[:, :-1]) & real[:, 1:] & real[:, :-1]
    return eq.any(axis=1)


def assert_knn_equal(idx, d2, idx_ref, d2_ref, allow_tie_perm=False):
    """Bit-exact idx and dist2.  With allow_tie_perm, rows holding exact distance ties may differ by a
    permutation inside each tie group (reference heap order) and in which of several equal
    candidates at the k-th distance was kept; dist2 must still be identical."""
    idx, d2, idx_ref, d2_ref = map(np.asarray, (idx, d2, idx_ref, d2_ref))
    assert idx.shape == idx_ref.shape and d2.shape == d2_ref.shape
    assert np.array_equal(d2.view(np.uint32), d2_ref.view(np.uint32)), "dist2 differs bitwise"
    if not allow_tie_perm:
        assert np.array_equal(idx, idx_ref), f"idx differs on {np.flatnonzero((idx != idx_ref).any(1))[:10]}"
        return 0
    bad = (idx != idx_ref).any(axis=1)
    ties = tie_rows(d2_ref)
    assert not (bad & ~ties).any(), "idx differs on a tie-free row"
    return int(bad.sum())


def to_cuda(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]
