"""The point-operator schedule of one PTv2m2 forward+backward — the benchmark
workload ("PTv2 pointops fwd+bwd").

For a batch in the offset layout it issues exactly the point-operator calls the backbone makes
(…/point_transformer_v2m2_base.py:556-576), in order, with the shapes of the configured model, and
then their backward passes:

    level 0:   kNN(k)            patch_embed_depth blocks
    stage i:   GridPool L_i→L_{i+1} (feature width C_{i+1}), kNN, enc_depths[i] blocks
    decoder i: interpolation L_{i+1}→L_i (kNN k=3, width C_i), dec_depths[i] blocks
               (neighbour lists of level i are the encoder's — same coordinates, same k)
    block:     gva_relation (key[idx]-q), gva_aggregate;   backward: CSR build (once per neighbour
               list), segmented scatter for key/value, per-query sums, softmax backward
    per neighbour list (4 per step): group_xyz — the relative coordinates (N,k,3) that feed every
               block's positional-bias MLP depend only on (idx, coord)

The dense per-point MLPs between the point operators are NOT part of this schedule (they are
cuBLAS GEMMs / BatchNorm in ao_b200.ptv2); their outputs are stood in for by resident synthetic
tensors of the right shape (q/k/v (N,C), peb (N,k,C), logits (N,k,G), upstream gradients).
Everything data-dependent — neighbour search, voxel partition, CSR — is recomputed every step.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import pointops


@dataclass
class ScheduleConfig:
    name: str = "s3dis"
    k: int = 16                                # neighbours of every BlockSequence unless patch_k / enc_k / dec_k say otherwise
    patch_depth: int = 2
    channels: tuple = (48, 96, 192, 384)       # level widths (patch embed, enc stages)
    groups: tuple = (6, 12, 24, 48)
    enc_depths: tuple = (2, 6, 2)
    dec_depths: tuple = (1, 1, 1)
    grid_sizes: tuple = (0.1, 0.2, 0.4)
    unpool: str = "interp"                     # "interp" (pointops.interpolation) or "map" (feat[cluster])
    interp_k: int = 3
    patch_k: Optional[int] = None              # patch_embed_neighbours
    enc_k: Optional[tuple] = None              # enc_neighbours (levels 1..L)
    dec_k: Optional[tuple] = None              # dec_neighbours (levels 0..L-1)
    variant: str = "materialised"              # "materialised": gva_relation at width C (the (N,k,C) relation tensor exists,
    #                                            what the model runs in fp32) | "fused": what ptv2 runs under bf16 autocast
    #                                            for C in {48, 96} — G-wide relation + fused positional MLP (pe_mlp)

    def k_patch(self):
        return self.patch_k or self.k

    def k_enc(self, i):
        return self.enc_k[i] if self.enc_k else self.k

    def k_dec(self, i):
        return self.dec_k[i] if self.dec_k else self.k

    def level_ks(self):
        """Neighbour counts used on every level (a level with two different k runs ONE search with the larger k:
        the smaller list is its sorted prefix)."""
        n_stage = len(self.grid_sizes)
        ks = [set() for _ in range(n_stage + 1)]
        if self.patch_depth:
            ks[0].add(self.k_patch())
        for i in range(n_stage):
            if self.enc_depths[i]:
                ks[i + 1].add(self.k_enc(i))
            if self.dec_depths[i]:
                ks[i].add(self.k_dec(i))
        return [sorted(s) for s in ks]

    @staticmethod
    def s3dis(**kw):
        """configs/s3dis/semseg-pt-v2m2-0-base.py:10-36"""
        return ScheduleConfig(**kw)

    @staticmethod
    def scannet(k=16, unpool="map", **kw):
        """configs/scannet/semseg-pt-v2m2-0-base.py:10-36: patch k=8, 4 pooling stages, `map` unpooling.
        k = enc/dec neighbours (16 in the config; BASELINE.json configs[3] also asks for 32)."""
        return ScheduleConfig(name="scannet", k=k, patch_k=8, patch_depth=1, channels=(48, 96, 192, 384, 512),
                              groups=(6, 12, 24, 48, 64), enc_depths=(2, 2, 6, 2), dec_depths=(1, 1, 1, 1),
                              grid_sizes=(0.06, 0.15, 0.375, 0.9375), unpool=unpool, **kw)

    @staticmethod
    def kitti(k=16, unpool="map", **kw):
        """configs/semantic_kitti/semseg-pt-v2m2-0-base.py:10-36 (same backbone as ScanNet, outdoor grid sizes)."""
        return ScheduleConfig(name="kitti", k=k, patch_k=8, patch_depth=1, channels=(48, 96, 192, 384, 512),
                              groups=(6, 12, 24, 48, 64), enc_depths=(2, 2, 6, 2), dec_depths=(1, 1, 1, 1),
                              grid_sizes=(0.15, 0.375, 0.9375, 2.34375), unpool=unpool, **kw)


class Profiler:
    """CUDA-event timing per kernel family on the launching stream (bench.py roofline)."""

    def __init__(self):
        self.records: Dict[str, list] = {}
        self.enabled = False

    @contextlib.contextmanager
    def span(self, name: str, nbytes: float):
        if not self.enabled:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.records.setdefault(name, []).append((s, e, nbytes))

    def summary(self):
        out = {}
        for name, recs in self.records.items():
            ms = sum(s.elapsed_time(e) for s, e, _ in recs)
            nbytes = sum(b for _, _, b in recs)
            out[name] = dict(ms=ms, calls=len(recs), bytes=nbytes, gbs=(nbytes / (ms * 1e-3) / 1e9) if ms > 0 else 0.0)
        return out


@dataclass
class Level:
    n: int = 0
    c: int = 0
    g: int = 0
    coord: Optional[torch.Tensor] = None
    offset: Optional[torch.Tensor] = None
    tensors: dict = field(default_factory=dict)
    lists: dict = field(default_factory=dict)      # k -> resident (n,k,*) stand-ins of that neighbour list


class PointOpsSchedule:
    """Holds the resident synthetic activations; `step(coord, offset)` runs one fwd+bwd."""

    def __init__(self, cfg: ScheduleConfig, device="cuda", seed: int = 0):
        self.cfg = cfg
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.levels: List[Level] = []
        self.prof = Profiler()
        self.last_sizes = None
        self.between = None      # optional callable run between the forward and the backward pass (bench.py: the
        #                          gradient all-reduce stand-in is issued there, on its own stream)

    # ---- synthetic stand-ins for the dense layers' outputs (allocated once per level size) -------------
    def _rand(self, *shape, grad=False):
        t = torch.randn(*shape, device=self.device, generator=self.gen, dtype=torch.float32)
        return t.requires_grad_(grad)

    def _fused_level(self, li: int) -> bool:
        c, g = self.cfg.channels[li], self.cfg.groups[li]
        return self.cfg.variant == "fused" and g <= 16 and pointops.pe_mlp_supported(c)

    def _level_tensors(self, li: int, n: int):
        cfg = self.cfg
        while len(self.levels) <= li:
            self.levels.append(Level())
        lv = self.levels[li]
        if lv.n == n and lv.tensors:
            return lv
        c, g = cfg.channels[li], cfg.groups[li]
        lv.n, lv.c, lv.g = n, c, g
        t = dict(key=self._rand(n, c, grad=True), query=self._rand(n, c, grad=True), value=self._rand(n, c, grad=True),
                 g_out=self._rand(n, c))
        if self._fused_level(li):
            # what GroupedVectorAttention._forward_fused feeds the operators: key / query projected to G columns by
            # weight_encoding[0] (cuBLAS, outside this schedule), linear_p_bias itself, and W_e·W_2 as the auxiliary head
            t["kp"], t["qp"] = self._rand(n, g, grad=True), self._rand(n, g, grad=True)
            mlp = torch.nn.Sequential(torch.nn.Linear(3, c), torch.nn.BatchNorm1d(c), torch.nn.ReLU(inplace=True),
                                      torch.nn.Linear(c, c)).to(self.device).train()
            t["mlp"] = mlp
            t["aux_w"] = (self._rand(g, c) / c ** 0.5).requires_grad_(True)
        if li + 1 < len(cfg.channels):
            t["pool_in"] = torch.relu(self._rand(n, cfg.channels[li + 1])).requires_grad_(True)   # post-ReLU (…:247)
            t["g_interp"] = self._rand(n, c)                                                       # grad of unpool output
        lv.tensors = t
        lv.lists = {}
        return lv

    def _list_tensors(self, lv: Level, k: int):
        """(n,k,*) stand-ins of one neighbour list: positional bias, attention logits, upstream relation gradient."""
        t = lv.lists.get(k)
        if t is None:
            fused = "mlp" in lv.tensors
            t = dict(logits=self._rand(lv.n, k, lv.g, grad=True))
            if fused:
                t["g_u"] = self._rand(lv.n, k, lv.g)
            else:
                t["peb"] = self._rand(lv.n, k, lv.c, grad=True)
                t["g_rel"] = self._rand(lv.n, k, lv.c)
            lv.lists[k] = t
        return t

    def _coarse_tensors(self, li: int, n: int):
        """Tensors living on level li that feed level li-1: grad of the pooled feature, unpool input."""
        lv = self.levels[li]
        t = lv.tensors
        c_here, c_fine = self.cfg.channels[li], self.cfg.channels[li - 1]
        if t.get("_coarse_n") != n:
            t["g_pool"] = self._rand(n, c_here)
            t["interp_in"] = self._rand(n, c_fine, grad=True)
            t["_coarse_n"] = n
        return t

    # ---- algorithmic bytes (SURVEY.md §8d formulas) ----------------------------------------------------
    @staticmethod
    def bytes_relation_fwd(n, k, c): return 4.0 * n * k + 8.0 * n * c + 4.0 * n * k * c
    @staticmethod
    def bytes_aggregate_fwd(n, k, c, g): return 4.0 * n * c + 4.0 * n * k * c + 8.0 * n * k * g + 4.0 * n * k + 4.0 * n * c
    @staticmethod
    def bytes_group_xyz(n, k): return 12.0 * n + 12.0 * n + 4.0 * n * k + 12.0 * n * k

    # ---- one block ---------------------------------------------------------------------------------------
    def _block_forward(self, lv: Level, nl, tape):
        """Forward of one block's point operators on the neighbour list nl = (k, idx, pos, moments).  Every block
        gets its own autograd leaves (views of the level's resident tensors — no copy), so the single backward pass at
        the end of the step produces one gradient per block instead of accumulating into shared leaves (the
        accumulation would add (N,k,C) element-wise adds that the model does not have)."""
        k, idx, pos, mom = nl
        t, lt = lv.tensors, self._list_tensors(lv, k)
        leaf = lambda x: x.detach().requires_grad_(True)
        value, logits = leaf(t["value"]), leaf(lt["logits"])
        if "mlp" in t:                          # ptv2.GroupedVectorAttention._forward_fused
            kp, qp, aux_w = leaf(t["kp"]), leaf(t["qp"]), leaf(t["aux_w"])
            mlp = t["mlp"]
            peb, upe = pointops.pe_bias_mlp(pos, mlp, mom, aux_weight=aux_w)
            with self.prof.span("gva_relation_fwd", self.bytes_relation_fwd(lv.n, k, lv.g)):
                rel = pointops.gva_relation(kp, qp, idx)                      # (N,k,G)
            with self.prof.span("gva_aggregate_fwd", self.bytes_aggregate_fwd(lv.n, k, lv.c, lv.g)):
                out = pointops.gva_aggregate(value, peb, logits, idx, lv.g)
            # u = rel + upe + const (ptv2.py): its gradient reaches both terms unchanged
            tape.append(([rel, upe, out], [kp, qp, value, logits, aux_w] + list(mlp.parameters()),
                         [lt["g_u"], lt["g_u"], t["g_out"]]))
            return
        key, query, peb = leaf(t["key"]), leaf(t["query"]), leaf(lt["peb"])
        with self.prof.span("gva_relation_fwd", self.bytes_relation_fwd(lv.n, k, lv.c)):
            rel = pointops.gva_relation(key, query, idx)
        with self.prof.span("gva_aggregate_fwd", self.bytes_aggregate_fwd(lv.n, k, lv.c, lv.g)):
            out = pointops.gva_aggregate(value, peb, logits, idx, lv.g)
        tape.append(([rel, out], [key, query, value, peb, logits], [lt["g_rel"], t["g_out"]]))

    def _neighbour_lists(self, li: int, lv: Level):
        """One search per level with the largest k the level needs; smaller lists are sorted prefixes of it
        ((dist2, idx)-lexicographic order, ties included), e.g. the ScanNet cfg's patch-embed k=8 next to the last
        decoder's k=16 on the level-0 coordinates."""
        ks = self.cfg.level_ks()[li]
        out = {}
        if not ks:
            return out
        kmax = ks[-1]
        with self.prof.span("knn", 0.0):
            idx_max, _ = pointops.knn_query(kmax, lv.coord, lv.offset)
        for k in ks:
            idx = idx_max if k == kmax else idx_max[:, :k].contiguous()
            pointops.prefetch_csr(idx, lv.coord.shape[0], 0)   # as ptv2.BlockSequence does when training
            pos = pointops.group_xyz(idx, lv.coord)            # (N,k,3), shared by every block on this neighbour list
            mom = pointops.pos_moments(pos) if "mlp" in lv.tensors else None
            out[k] = (k, idx, pos, mom)
        return out

    # ---- one training-step worth of point operators -------------------------------------------------
    def step(self, coord: torch.Tensor, offset: torch.Tensor):
        cfg = self.cfg
        n_stage = len(cfg.grid_sizes)
        lv0 = self._level_tensors(0, coord.shape[0])
        lv0.coord, lv0.offset = coord, offset
        tape = []     # (outputs, leaves, upstream gradients) in forward order
        lists = []    # per level: {k: (k, idx, pos, moments)}
        parts = []
        # ---------------- forward ----------------
        lists.append(self._neighbour_lists(0, lv0))
        # the coordinate pyramid (voxel partitions + coarse coordinates: the step's only host syncs) is built while
        # the level-0 search is still running on the device; grid_pool below finds it cached on the coord tensors
        pointops.prepare_pyramid(coord, offset, cfg.grid_sizes, knn=[cfg.level_ks()[i + 1][-1] for i in range(n_stage)],
                                 interp_k=cfg.interp_k if cfg.unpool == "interp" else None)
        for _ in range(cfg.patch_depth):
            self._block_forward(lv0, lists[0][cfg.k_patch()], tape)
        for i in range(n_stage):
            fine = self.levels[i]
            c_next = cfg.channels[i + 1]
            nb = 4.0 * fine.n * c_next + 16.0 * fine.n
            pool_in = fine.tensors["pool_in"]
            with self.prof.span("grid_pool_fwd", nb):
                (nc, nf, noff), cluster, part = pointops.grid_pool(fine.coord, pool_in, fine.offset,
                                                                   cfg.grid_sizes[i], return_partition=True)
            lv = self._level_tensors(i + 1, nc.shape[0])
            lv.coord, lv.offset = nc, noff.int()
            self._coarse_tensors(i + 1, nc.shape[0])
            parts.append(part)
            tape.append(([nf], [pool_in], [lv.tensors["g_pool"]]))
            lists.append(self._neighbour_lists(i + 1, lv))
            for _ in range(cfg.enc_depths[i]):
                self._block_forward(lv, lists[i + 1][cfg.k_enc(i)], tape)
        for i in reversed(range(n_stage)):
            coarse, fine = self.levels[i + 1], self.levels[i]
            src = coarse.tensors["interp_in"]
            nb = 24.0 * fine.n + 4.0 * coarse.n * fine.c + 4.0 * fine.n * fine.c
            with self.prof.span("unpool_fwd", nb):
                if cfg.unpool == "interp":
                    up = pointops.interpolation(coarse.coord, fine.coord, src, coarse.offset, fine.offset, k=cfg.interp_k)
                else:
                    up = pointops.unpool_map(src, parts[i])
            tape.append(([up], [src], [fine.tensors["g_interp"]]))
            for _ in range(cfg.dec_depths[i]):
                self._block_forward(fine, lists[i][cfg.k_dec(i)], tape)   # encoder's neighbour list reused
        if self.between is not None:
            self.between()
        # ---------------- backward: ONE pass over the recorded graph, like loss.backward() in training ------
        # (the engine runs the nodes in reverse creation order: decoder blocks, unpool, encoder blocks, pool, ...)
        outs = [o for rec in tape for o in rec[0]]
        leaves = [x for rec in tape for x in rec[1]]
        ups = [g for rec in tape for g in rec[2]]
        grads = torch.autograd.grad(outs, leaves, ups)
        self.last_sizes = [l.n for l in self.levels[: n_stage + 1]]
        return grads[0]      # gradient of the first patch-embed block's first leaf: the last one the pass produces

    # ---- totals for reporting -------------------------------------------------------------------------
    def blocks_per_level(self):
        cfg = self.cfg
        per = [cfg.patch_depth] + list(cfg.enc_depths)
        for i, d in enumerate(cfg.dec_depths):
            per[i] += d
        return per
