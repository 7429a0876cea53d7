"""ORACLE — TEST INFRASTRUCTURE ONLY (CPU baseline leg of bench.py).

The reference's pure-torch CPU path for the PTv2m2 point-operator schedule, timed by bench.py as
`cpu_baseline` and as the `--impl reference` arm.  Never imported by the product package.

The reference has no CPU kNN (its only kNN is the CUDA kernel
libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104); BASELINE.json's north_star defines the
CPU path as "cdist+topk, gather, scatter".  Everything else is the reference's own torch code
restated (oracle/torch_ref.py cites the lines): `grouping` (libs/pointops/functions/grouping.py:36-60),
`interpolation` (interpolation.py:8-22), the GroupedVectorAttention tail
(pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:109-128), GridPool (:244-269,
with torch.segment_reduce standing in for torch_scatter.segment_csr), autograd for every backward.

Op schedule = the reference forward: one kNN per BlockSequence (7 per forward — the
reference does not share neighbour lists between encoder and decoder), two `grouping` calls per
block, 3 GridPools, 3 interpolations (each with its own k=3 search), then the backward of all of it.

Bounded sample.  A whole 80k-point room costs minutes on a CPU (kNN is O(n^2)), so a step processes
a FRACTION f of the query rows of every level — the first ceil(f*n_l) rows — against the FULL level
as candidates / gather sources / scatter targets.  Every operator is row-parallel over queries, so
cost is linear in f and throughput = f * n_0 points / time is an unbiased estimate of the full room.
"""
from __future__ import annotations

import math
import time
from typing import List

import numpy as np
import torch

from . import torch_ref

class Cfg:
    """Op-schedule description with the accessor names of ao_b200.schedule.ScheduleConfig (bench.py passes that
    dataclass in; this class keeps the oracle importable on its own).  Default = the S3DIS config
    (configs/s3dis/semseg-pt-v2m2-0-base.py:10-36)."""

    def __init__(self, k=16, patch_k=None, patch_depth=2, channels=(48, 96, 192, 384), groups=(6, 12, 24, 48),
                 enc_depths=(2, 6, 2), dec_depths=(1, 1, 1), grid_sizes=(0.1, 0.2, 0.4), interp_k=3, unpool="interp"):
        self.k, self.patch_k, self.patch_depth, self.channels, self.groups = k, patch_k, patch_depth, channels, groups
        self.enc_depths, self.dec_depths, self.grid_sizes, self.interp_k, self.unpool = (
            enc_depths, dec_depths, grid_sizes, interp_k, unpool)

    def k_patch(self):
        return self.patch_k or self.k

    def k_enc(self, i):
        return self.k

    def k_dec(self, i):
        return self.k


S3DIS = Cfg()
FIXED_FRACTION = 0.04    # query-row fraction of a reference-arm step (fixed: a per-run calibration made the arm's
#                          value move by +-40 % between runs of the same code)


def knn_cdist_topk(k, xyz, new_xyz, chunk=2048):
    """Exact kNN of one scene: non-mm cdist + topk, chunked over queries."""
    idx_out, dist_out = [], []
    kk = min(k, xyz.shape[0])
    for s in range(0, new_xyz.shape[0], chunk):
        d = torch.cdist(new_xyz[s:s + chunk], xyz, compute_mode="donot_use_mm_for_euclid_dist")
        dist, idx = d.topk(kk, dim=1, largest=False, sorted=True)
        if kk < k:
            idx = torch.cat([idx, idx.new_full((idx.shape[0], k - kk), -1)], 1)
            dist = torch.cat([dist, dist.new_full((dist.shape[0], k - kk), 1e5)], 1)
        idx_out.append(idx.int())
        dist_out.append(dist)
    return torch.cat(idx_out), torch.cat(dist_out)


def grid_pool_fast(coord, feat, grid_size):
    """GridPool tail for ONE scene with vectorised torch ops (…v2m2_base.py:249-268)."""
    start = coord.min(dim=0).values
    batch = torch.zeros(coord.shape[0], dtype=torch.long)
    keys = torch_ref.voxel_grid_keys(coord - start, grid_size, batch)
    _, cluster, counts = torch.unique(keys, sorted=True, return_inverse=True, return_counts=True)
    order = torch.sort(cluster, stable=True).indices
    new_coord = torch.segment_reduce(coord[order], "mean", lengths=counts, axis=0)
    new_feat = torch.segment_reduce(feat[order], "max", lengths=counts, axis=0)
    return new_coord, new_feat, cluster


class CpuRoom:
    """One room: level coordinates (full pyramid, built once, untimed) + synthetic stand-ins for the
    dense layers' outputs, like ao_b200.schedule.PointOpsSchedule."""

    def __init__(self, coord: np.ndarray, cfg=S3DIS, seed: int = 0):
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        self.coords: List[torch.Tensor] = [torch.from_numpy(np.ascontiguousarray(coord))]
        self.clusters = []
        for gs in cfg.grid_sizes:
            c = self.coords[-1]
            nc, _, cl = grid_pool_fast(c, c, gs)
            self.coords.append(nc.contiguous())
            self.clusters.append(cl)
        self.sizes = [c.shape[0] for c in self.coords]
        self.gen = g

    def _rand(self, *shape, grad=False):
        return torch.randn(*shape, generator=self.gen).requires_grad_(grad)

    def _block(self, li, rows, idx):
        cfg = self.cfg
        n, c, g, k = self.sizes[li], cfg.channels[li], cfg.groups[li], idx.shape[1]
        coord = self.coords[li]
        key, value = self._rand(n, c, grad=True), self._rand(n, c, grad=True)
        query, peb = self._rand(rows, c, grad=True), self._rand(rows, k, c, grad=True)
        logits = self._rand(rows, k, g, grad=True)
        g_rel, g_out = self._rand(rows, k, c), self._rand(rows, c)
        t0 = time.perf_counter()
        key_g = torch_ref.grouping(idx, key, coord, coord[:rows], with_xyz=True)       # :109
        value_g = torch_ref.grouping(idx, value, coord, coord[:rows], with_xyz=False)  # :110
        pos, key_g = key_g[:, :, 0:3], key_g[:, :, 3:]                                  # :111
        rel = key_g - query.unsqueeze(1) + peb                                          # :112,:118
        value_g = value_g + peb                                                         # :119
        w = torch.softmax(logits, dim=1)                                                # :122
        w = w * torch.sign(idx + 1).unsqueeze(-1)                                       # :124-125
        out = torch.einsum("nsgi,nsg->ngi", value_g.view(rows, k, g, c // g), w).reshape(rows, c)
        torch.autograd.grad([rel, out], [key, value, query, peb, logits], [g_rel, g_out])
        return time.perf_counter() - t0

    def step(self, f: float) -> dict:
        """One fwd+bwd of the reference op schedule on the first ceil(f*n_l) rows of every level.
        Returns seconds per operator family (only operator time is counted, not the synthetic
        tensor generation)."""
        cfg = self.cfg
        rows = [max(1, min(n, math.ceil(f * n))) for n in self.sizes]
        t = dict(knn=0.0, block=0.0, pool=0.0, interp=0.0)
        n_stage = len(cfg.grid_sizes)

        def knn(li, kk, q):
            t0 = time.perf_counter()
            idx, dist = knn_cdist_topk(kk, self.coords[li], q)
            t["knn"] += time.perf_counter() - t0
            return idx, dist

        idx, _ = knn(0, cfg.k_patch(), self.coords[0][:rows[0]])
        for _ in range(cfg.patch_depth):
            t["block"] += self._block(0, rows[0], idx)
        for i in range(n_stage):
            c_next = cfg.channels[i + 1]
            feat = torch.relu(self._rand(rows[i], c_next)).requires_grad_(True)
            t0 = time.perf_counter()
            _, nf, _ = grid_pool_fast(self.coords[i][:rows[i]], feat, cfg.grid_sizes[i])
            torch.autograd.grad(nf, feat, torch.ones_like(nf))
            t["pool"] += time.perf_counter() - t0
            idx, _ = knn(i + 1, cfg.k_enc(i), self.coords[i + 1][:rows[i + 1]])
            for _ in range(cfg.enc_depths[i]):
                t["block"] += self._block(i + 1, rows[i + 1], idx)
        for i in reversed(range(n_stage)):
            c = cfg.channels[i]
            src = self._rand(self.sizes[i + 1], c, grad=True)
            q = self.coords[i][:rows[i]]
            if cfg.unpool == "interp":
                idx3, dist3 = knn(i + 1, cfg.interp_k, q)                               # interpolation.py:14
                t0 = time.perf_counter()
                w = torch_ref.interpolation_weights(dist3)
                up = torch.zeros(rows[i], c)
                for j in range(cfg.interp_k):
                    up = up + src[idx3[:, j].long(), :] * w[:, j].unsqueeze(-1)        # :20-21
            else:
                t0 = time.perf_counter()
                up = src[self.clusters[i][:rows[i]]]                                    # …v2m2_base.py:309 ("map")
            torch.autograd.grad(up, src, torch.ones_like(up))
            t["interp"] += time.perf_counter() - t0
            idx, _ = knn(i, cfg.k_dec(i), q)                                             # …v2m2_base.py:223 (not shared)
            for _ in range(cfg.dec_depths[i]):
                t["block"] += self._block(i, rows[i], idx)
        t["total"] = sum(t.values())
        t["points"] = f * self.sizes[0] if rows[0] < self.sizes[0] else float(self.sizes[0])
        t["rows"] = rows
        return t


def calibrate_fraction(room: CpuRoom, target_s: float, f0: float = 0.004) -> float:
    """Fraction whose step takes about target_s seconds (cost is linear in f)."""
    r = room.step(f0)
    f = f0 * target_s / max(r["total"], 1e-6)
    return float(min(1.0, max(f0 / 4, f)))
