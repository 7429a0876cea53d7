"""ORACLE — TEST INFRASTRUCTURE ONLY (the `gpu_reference` leg of bench.py).

The reference's OWN op schedule for the PTv2m2 point-operator path, run on the same B200 with the
reference's own implementations, so that there is a like-for-like GPU number next to the new kernels
(BASELINE.md §4):

  kNN              the UNMODIFIED reference kernel (oracle/_ref/libpointops_ref.so, built from
                   /root/reference/libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104) — one launch
                   per BlockSequence (7 self searches per forward; the reference does not share neighbour
                   lists between encoder and decoder, …v2m2_base.py:223) + one k=3 search per interpolation
  grouping         libs/pointops/functions/grouping.py:36-60 (pure torch: cat / index / sub / cat), twice per
                   block (…v2m2_base.py:109-110)
  GVA tail         …v2m2_base.py:112,118-128 (sub, add, softmax, mask, einsum)
  GridPool         …v2m2_base.py:246-268 with torch.unique / torch.sort and torch.segment_reduce standing in for
                   the un-vendored torch_scatter.segment_csr
  interpolation    libs/pointops/functions/interpolation.py:8-22 ("interp" backend) or feat[cluster] ("map")
  backward         autograd of all of the above (index_put_(accumulate=True) scatters)

All torch ops run on CUDA tensors (torch 2.11 ATen kernels).  The dense per-point MLPs are stood in for by
resident random tensors exactly as in ao_b200.schedule.  Never imported by the product package.
"""
from __future__ import annotations

import torch

from . import ref_cuda, torch_ref


def _offset2batch(offset):
    off = offset.long()
    counts = torch.diff(off, prepend=off.new_zeros(1))
    return torch.repeat_interleave(torch.arange(off.numel(), device=offset.device), counts)


def grid_pool_batch(coord, feat, offset, grid_size):
    """…v2m2_base.py:246-268 on a whole offset-encoded batch (vectorised torch ops)."""
    batch = _offset2batch(offset)
    lengths = torch.diff(offset.long(), prepend=offset.new_zeros(1).long())
    start = torch.segment_reduce(coord, "min", lengths=lengths, axis=0)
    keys = torch_ref.voxel_grid_keys(coord - start[batch], grid_size, batch)
    _, cluster, counts = torch.unique(keys, sorted=True, return_inverse=True, return_counts=True)
    order = torch.sort(cluster).indices
    idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, dim=0)])
    new_coord = torch.segment_reduce(coord[order], "mean", lengths=counts, axis=0)
    new_feat = torch.segment_reduce(feat[order], "max", lengths=counts, axis=0)
    new_batch = batch[idx_ptr[:-1]]
    new_offset = torch.cumsum(new_batch.bincount(), dim=0)
    return new_coord, new_feat, new_offset, cluster


class GpuReferenceSchedule:
    """cfg: an ao_b200.schedule.ScheduleConfig-like object (attributes channels, groups, enc_depths, dec_depths,
    grid_sizes, patch_depth, unpool, interp_k and k_patch() / k_enc(i) / k_dec(i))."""

    def __init__(self, cfg, device, seed=0):
        self.cfg, self.device = cfg, torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.cache = {}

    def _rand(self, name, level, *shape, grad=False):
        key = (name, level) + tuple(shape)
        t = self.cache.get(key)
        if t is None:
            t = torch.randn(*shape, device=self.device, generator=self.gen)
            self.cache[key] = t
        return t.detach().requires_grad_(grad)

    def _block(self, li, coord, idx, k, tape):
        cfg = self.cfg
        n, c, g = coord.shape[0], cfg.channels[li], cfg.groups[li]
        key, value, query = (self._rand(nm, li, n, c, grad=True) for nm in ("key", "value", "query"))
        peb = self._rand("peb", li, n, k, c, grad=True)
        logits = self._rand("logits", li, n, k, g, grad=True)
        key_g = torch_ref.grouping(idx, key, coord, coord, with_xyz=True)               # :109
        value_g = torch_ref.grouping(idx, value, coord, coord, with_xyz=False)          # :110
        pos, key_g = key_g[:, :, 0:3], key_g[:, :, 3:]                                   # :111
        rel = key_g - query.unsqueeze(1) + peb                                           # :112,:118
        value_g = value_g + peb                                                          # :119
        w = torch.softmax(logits, dim=1)                                                 # :122
        w = w * torch.sign(idx + 1).unsqueeze(-1)                                        # :124-125
        out = torch.einsum("nsgi,nsg->ngi", value_g.view(n, k, g, c // g), w).reshape(n, c)
        tape.append(([rel, out], [key, value, query, peb, logits],
                     [self._rand("g_rel", li, n, k, c), self._rand("g_out", li, n, c)]))

    def step(self, coord, offset):
        cfg = self.cfg
        n_stage = len(cfg.grid_sizes)
        tape, levels, clusters = [], [(coord, offset)], []
        idx, _ = ref_cuda.knn_query(cfg.k_patch(), coord, offset)
        for _ in range(cfg.patch_depth):
            self._block(0, coord, idx, cfg.k_patch(), tape)
        for i in range(n_stage):
            c, o = levels[-1]
            feat = torch.relu(self._rand("pool_in", i, c.shape[0], cfg.channels[i + 1])).requires_grad_(True)
            nc, nf, no, cluster = grid_pool_batch(c, feat, o, cfg.grid_sizes[i])
            tape.append(([nf], [feat], [self._rand("g_pool", i, nf.shape[0], nf.shape[1])]))
            levels.append((nc.contiguous(), no))
            clusters.append(cluster)
            idx, _ = ref_cuda.knn_query(cfg.k_enc(i), levels[-1][0], no)                 # :223
            for _ in range(cfg.enc_depths[i]):
                self._block(i + 1, levels[-1][0], idx, cfg.k_enc(i), tape)
        for i in reversed(range(n_stage)):
            (cc, co), (fc, fo) = levels[i + 1], levels[i]
            c = cfg.channels[i]
            src = self._rand("up_in", i, cc.shape[0], c, grad=True)
            if cfg.unpool == "interp":
                idx3, d2 = ref_cuda.knn_query(cfg.interp_k, cc, co, fc, fo)             # interpolation.py:14
                w = torch_ref.interpolation_weights(torch.sqrt(d2))
                up = torch.zeros(fc.shape[0], c, device=self.device)
                for j in range(cfg.interp_k):
                    up = up + src[idx3[:, j].long(), :] * w[:, j].unsqueeze(-1)         # :20-21
            else:
                up = src[clusters[i]]                                                    # …v2m2_base.py:309
            tape.append(([up], [src], [self._rand("g_up", i, fc.shape[0], c)]))
            idx, _ = ref_cuda.knn_query(cfg.k_dec(i), fc, fo)                            # :223 (not shared)
            for _ in range(cfg.dec_depths[i]):
                self._block(i, fc, idx, cfg.k_dec(i), tape)
        outs = [o for rec in tape for o in rec[0]]
        leaves = [x for rec in tape for x in rec[1]]
        ups = [g for rec in tape for g in rec[2]]
        grads = torch.autograd.grad(outs, leaves, ups)
        return grads[0]
