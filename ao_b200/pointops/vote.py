"""vote_accumulate — the tester's fragment vote (/root/reference/pointcept/engines/test.py:106-113).

    pred_part = F.softmax(pred_part, -1)
    bs = 0
    for be in input_dict["offset"]:
        pred[idx_part[bs:be], :] += pred_part[bs:be]
        bs = be

as one kernel per fragment (softmax fused into the indexed read-modify-write; the probabilities are never
materialised).  Fragments are applied in order on the caller's stream, so a point that several fragments vote
for receives its votes in the reference's order.  Inside one fragment the indices must be distinct — they are, a
fragment holds one point per voxel (pointcept/datasets/transform.py:834-837); with duplicates torch's
`pred[idx] += x` keeps one arbitrary writer and so does this kernel.
"""
from __future__ import annotations

import torch

from .. import _lib


def vote_accumulate(pred: torch.Tensor, logits: torch.Tensor, index: torch.Tensor, offset=None,
                    check_index: bool = False) -> torch.Tensor:
    """pred (n_pred, classes) fp32, updated in place and returned; logits (rows, classes); index (rows) integer;
    offset = cumulative end rows of the fragments in `logits` (None: one fragment).  `check_index=True` raises
    IndexError for an index outside [-n_pred, n_pred) like torch's indexing does (one host sync)."""
    if pred.dim() != 2 or logits.dim() != 2 or pred.shape[1] != logits.shape[1]:
        raise ValueError("vote_accumulate: pred (n_pred, classes) and logits (rows, classes) expected")
    if index.dim() != 1 or index.shape[0] != logits.shape[0]:
        raise ValueError("vote_accumulate: index must hold one entry per logits row")
    if not pred.is_cuda or pred.dtype != torch.float32 or not pred.is_contiguous():
        raise ValueError("vote_accumulate: pred must be a contiguous fp32 CUDA tensor")
    if logits.device != pred.device or index.device != pred.device:
        raise ValueError("vote_accumulate: all tensors must be on pred's device")
    lib = _lib.load()
    logits = logits.detach().float().contiguous()
    index = index.long().contiguous()
    rows, c = logits.shape
    if offset is None:
        ends = [rows]
    else:
        ends = [int(e) for e in (offset.tolist() if torch.is_tensor(offset) else offset)]
    bad = torch.zeros(1, dtype=torch.int32, device=pred.device) if check_index else None
    with _lib.on_device(pred.device):
        bs = 0
        for be in ends:
            if be < bs or be > rows:
                raise ValueError("vote_accumulate: offset must be non-decreasing and within the rows")
            if be > bs:
                _lib.check(lib.aopt_vote_accumulate(be - bs, c, pred.shape[0], logits.data_ptr() + bs * c * 4,
                                                    index.data_ptr() + bs * 8, _lib.ptr(pred),
                                                    _lib.ptr(bad) if bad is not None else None, _lib.stream()),
                           "vote_accumulate")
            bs = be
    if bad is not None and int(bad.item()) != 0:
        raise IndexError("vote_accumulate: index out of range for pred")
    return pred
