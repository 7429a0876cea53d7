"""PointTransformer V2 (mode 2) backbone routed through the ao_b200 point operators.

Mirror of /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py: same
module tree, constructor arguments and parameter names, so a reference `state_dict` loads with
`strict=True` and the golden fixtures produced by the reference modules (tests/golden/) can be
replayed.  What changes is the op schedule underneath:

  reference (per block, :103-129)                         here
  -------------------------------------------------       ---------------------------------------
  grouping(idx, key, coord, with_xyz=True)  (N,k,3+C)     pointops.group_xyz            (N,k,3), once per
                                                           neighbour list instead of once per block
  key[...,3:] - query.unsqueeze(1)                        pointops.gva_relation         (N,k,C)
  grouping(idx, value); value + peb; softmax; mask;       pointops.gva_aggregate        (N,C)
  einsum                                                   (value never gathered to (N,k,C))
  PointBatchNorm 3-D: transpose+contiguous x2 (:36-41)    BatchNorm1d on the (N·k, C) view
  one kNN per BlockSequence (:223), 7 per forward          4: decoder stages reuse the encoder's
                                                           neighbour lists (same coords, same k)
  GridPool: offset2batch loop, voxel_grid, unique, sort,   pointops.grid_pool
  2 permuted copies, 2 segment_csr (:244-269)
  interpolation: 3 gather+mul+add passes (:311)            pointops.interpolation (fused fwd, CSR bwd)

  PointBatchNorm + ReLU (+ DropPath + residual) as 3-6   pointops.bn_act: statistics + apply, two passes, fp32 /
  ATen kernels per site, dtype casts in between            bf16 in and out (training mode)
  weight_encoding[1:] on (N,k,G): BN, ReLU, cast,          pointops.we_tail (with the additions that form its input)
  (N·k,G)x(G,G) GEMM, cast

Dense per-point Linear layers stay torch.nn.Linear (cuBLAS tensor cores; bf16 under autocast).
"""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn

from . import pointops


def fused_pe_enabled() -> bool:
    """The fused positional-bias MLP (pointops.pe_bias_mlp) computes its C x C layer in bf16 on the tensor
    cores; it replaces the torch layers only under autocast (where they run in bf16 too), so fp32 runs keep
    fp32 parity with the reference modules.  AOPT_FUSED_PE=0 disables it, =1 forces it."""
    import os

    flag = os.environ.get("AOPT_FUSED_PE", "auto")
    if flag == "0":
        return False
    if flag == "1":
        return True
    return torch.is_autocast_enabled()


def we_gather_enabled() -> bool:
    """AOPT_WE_GATHER=1: pointops.we_tail forms rel = kp[idx] - qp inside its kernels (gather mode) instead of reading the
    (N,k,G) tensor gva_relation stored.  Off: the four passes of the tail (two forward, two backward) each repeat the
    24-byte row gathers and a division per row, which costs more than the one store it saves — we_tail backward 172 ->
    305 us at level 0, training step 32.6 -> 33.5 ms (profiles/r03h_model_step.txt)."""
    import os

    return os.environ.get("AOPT_WE_GATHER", "0") == "1"


def relation_free_min_elems() -> float:
    """Smallest N·k·C for which the relation-free schedule is also used at the widths the fused positional-MLP kernel
    does not cover (C = 192, 384: hidden activation through cuBLAS + bn_act, one more GEMM for `upe`).  It trades the
    (N,k,C) relation tensor and its passes for two more operators per block, which pays once the level is large enough
    for the step to be device-bound there: S3DIS cfg, 4 rooms (level 2: 38 M elements): 38.3 ms with it at every level vs
    36.6 ms without; 8 rooms (77 M): 49.6 vs 51.3 ms (profiles/r02v_*, r02y_*).  AOPT_RELFREE_MIN_ELEMS overrides
    (0 = always, AOPT_RELFREE_ALL=1 is the same)."""
    import os

    if os.environ.get("AOPT_RELFREE_ALL", "") == "1":
        return 0.0
    return float(os.environ.get("AOPT_RELFREE_MIN_ELEMS", 64e6))


class DropPath(nn.Module):
    """Stochastic depth per row (timm.models.layers.DropPath semantics for a (N,C) input)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


class PointBatchNorm(nn.Module):
    """Batch Normalization for point features [N, C] or [N, L, C] (…v2m2_base.py:25-45)."""

    def __init__(self, embed_channels):
        super().__init__()
        self.norm = nn.BatchNorm1d(embed_channels)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if input.dim() == 3:
            # same per-channel statistics as BatchNorm1d over (N, C, L), without the two transposes
            n, l, c = input.shape
            return self.norm(input.reshape(n * l, c)).view(n, l, c)
        elif input.dim() == 2:
            return self.norm(input)
        raise NotImplementedError


def _autocast_dtype():
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else None


def run_seq(seq, x, out_dtype=None, start=0, stop=None):
    """nn.Sequential forward with every [PointBatchNorm, ReLU] pair (or lone PointBatchNorm) routed through
    pointops.bn_act — the Linear -> PointBatchNorm -> ReLU triples of the reference (…v2m2_base.py:86-93,240-242,
    288-295,363-364,566-571).  out_dtype: element type wanted from a trailing BatchNorm stage (saves a cast kernel);
    start / stop: run the sub-sequence seq[start:stop]."""
    mods = list(seq)[start:stop]
    i = 0
    pre_bias = None
    while i < len(mods):
        m = mods[i]
        if isinstance(m, PointBatchNorm):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            last = i + (2 if relu else 1) >= len(mods)
            x = pointops.bn_act(x, m.norm, relu=relu, out_dtype=out_dtype if last else None, pre_bias=pre_bias)
            pre_bias = None
            i += 2 if relu else 1
        elif isinstance(m, nn.Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], PointBatchNorm):
            # Linear -> PointBatchNorm [-> ReLU]: one autograd node; the Linear's bias cancels under batch statistics and
            # is left out of the GEMM (no bias-gradient reduction over the rows; the running mean still sees it)
            relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
            last = i + (3 if relu else 2) >= len(mods)
            x = pointops.linear_bn_act(x, m, mods[i + 1].norm, relu=relu, out_dtype=out_dtype if last else None)
            i += 3 if relu else 2
        elif isinstance(m, nn.Linear):
            x = pointops.linear(x, m.weight, m.bias, out_f32=(out_dtype == torch.float32 and i + 1 == len(mods)))
            i += 1
        else:
            x = m(x)
            i += 1
    return x


class GroupedVectorAttention(nn.Module):
    def __init__(self, embed_channels, groups, attn_drop_rate=0.0, qkv_bias=True, pe_multiplier=False,
                 pe_bias=True):
        super().__init__()
        self.embed_channels = embed_channels
        self.groups = groups
        assert embed_channels % groups == 0
        self.attn_drop_rate = attn_drop_rate
        self.qkv_bias = qkv_bias
        self.pe_multiplier = pe_multiplier
        self.pe_bias = pe_bias
        self.linear_q = nn.Sequential(nn.Linear(embed_channels, embed_channels, bias=qkv_bias),
                                      PointBatchNorm(embed_channels), nn.ReLU(inplace=True))
        self.linear_k = nn.Sequential(nn.Linear(embed_channels, embed_channels, bias=qkv_bias),
                                      PointBatchNorm(embed_channels), nn.ReLU(inplace=True))
        self.linear_v = nn.Linear(embed_channels, embed_channels, bias=qkv_bias)
        if self.pe_multiplier:
            self.linear_p_multiplier = nn.Sequential(nn.Linear(3, embed_channels), PointBatchNorm(embed_channels),
                                                     nn.ReLU(inplace=True), nn.Linear(embed_channels, embed_channels))
        if self.pe_bias:
            self.linear_p_bias = nn.Sequential(nn.Linear(3, embed_channels), PointBatchNorm(embed_channels),
                                               nn.ReLU(inplace=True), nn.Linear(embed_channels, embed_channels))
        self.weight_encoding = nn.Sequential(nn.Linear(embed_channels, groups), PointBatchNorm(groups),
                                             nn.ReLU(inplace=True), nn.Linear(groups, groups))
        self.softmax = nn.Softmax(dim=1)
        self.attn_drop = nn.Dropout(attn_drop_rate)

    def forward(self, feat, coord, reference_index, pos=None, pos_moments=None):
        # relation-free schedule (_forward_fused): every width under autocast; the positional MLP itself runs in the
        # tcgen05 kernel where it is supported (C in {48, 96}, G <= 16) and through cuBLAS + bn_act elsewhere
        fused = (self.pe_bias and not self.pe_multiplier and fused_pe_enabled()
                 and not (self.attn_drop_rate > 0.0 and self.training)
                 and ((pointops.pe_mlp_supported(self.embed_channels) and self.groups <= 16)
                      or reference_index.numel() * self.embed_channels >= relation_free_min_elems()))
        # the point operators compute in fp32: q / k feed a GEMM in the relation-free schedule (any dtype) and
        # gva_relation otherwise (fp32); value always feeds gva_aggregate
        qk_dtype = None if fused else torch.float32
        if pointops.qkv_usable(feat, self.linear_q, self.linear_k, self.linear_v):
            query, key, value = pointops.qkv_bn(feat, self.linear_q, self.linear_k, self.linear_v, qk_dtype)   # one GEMM
        else:
            query, key = run_seq(self.linear_q, feat, qk_dtype), run_seq(self.linear_k, feat, qk_dtype)
            value = pointops.linear(feat, self.linear_v.weight, self.linear_v.bias, out_f32=True)
        if pos is None:                                                       # (N,k,3): depends only on (idx, coord)
            pos = pointops.group_xyz(reference_index, coord)                  # :109,:111
        if fused:
            return self._forward_fused(query, key, value, pos, pos_moments, reference_index)
        relation_qk = pointops.gva_relation(key, query, reference_index)      # :109,:112
        peb = None
        if self.pe_multiplier:
            relation_qk = relation_qk * self.linear_p_multiplier(pos)
        if self.pe_bias:
            if fused_pe_enabled() and pointops.pe_mlp_supported(self.embed_channels):
                # bf16 tensor-core fused MLP: used where the torch path would run in bf16 anyway (autocast)
                peb = pointops.pe_bias_mlp(pos, self.linear_p_bias, pos_moments)
            else:
                peb = run_seq(self.linear_p_bias, pos, torch.float32).float()
            relation_qk = relation_qk + peb
        weight = run_seq(self.weight_encoding, relation_qk, torch.float32)    # (N,k,G) logits
        if self.attn_drop_rate > 0.0 and self.training:
            # dropout sits between softmax and mask (:122-125): un-fused tail for this rare setting
            value_g = pointops.grouping(reference_index, value, coord, with_xyz=False)
            if peb is not None:
                value_g = value_g + peb
            weight = self.attn_drop(self.softmax(weight.float()))
            mask = torch.sign(reference_index + 1)
            weight = weight * mask.unsqueeze(-1)
            n, k, c = value_g.shape
            out = torch.einsum("nsgi,nsg->ngi", value_g.view(n, k, self.groups, c // self.groups), weight)
            return out.reshape(n, c)
        return pointops.gva_aggregate(value, peb, weight, reference_index, self.groups)   # :110,:119-128


    def _forward_fused(self, query, key, value, pos, pos_moments, reference_index):
        """Same mathematics as forward() with the (N,k,C) relation tensor eliminated.  The first layer of
        weight_encoding is linear, so (:112,:118,:120)
            Linear_e(key[idx] - q + peb) = (key We^T)[idx] - (q We^T) + (We W2) h + We b2 + b_e ,
        where h is the hidden activation of linear_p_bias: two (N,G) GEMMs, a G-wide gather, and one extra
        column tile in the fused positional MLP kernel — instead of materialising relation_qk, adding peb,
        casting and a (N·k,C)x(C,G) GEMM (and their backward passes)."""
        import torch.nn.functional as F

        lin_e, lin2 = self.weight_encoding[0], self.linear_p_bias[3]
        # key / query projected by weight_encoding[0]: (N,C) x (C,G) in the autocast dtype, like the Linear it replaces
        # (in fp32 these two skinny GEMMs ran as SIMT sgemm kernels, 174 us each at level 0:
        # profiles/r02e_model_step_torch_profile.txt)
        kp = pointops.linear(key, lin_e.weight, out_f32=True)                 # (N, G)
        qp = pointops.linear(query, lin_e.weight, out_f32=True)
        with torch.autocast("cuda", enabled=False):
            we = lin_e.weight.float()                                         # (G, C)
            wf = we @ lin2.weight.float()                                     # (G, C) acting on h
        if pointops.pe_mlp_supported(self.embed_channels) and self.groups <= 16:
            with torch.autocast("cuda", enabled=False):
                peb, upe = pointops.pe_bias_mlp(pos, self.linear_p_bias, pos_moments, aux_weight=wf)
        else:
            # widths without the fused kernel (C = 192, 384): hidden activation h = ReLU(BN(Linear(3,C)(pos))) through
            # cuBLAS + bn_act, then the two products that read it
            h = run_seq(self.linear_p_bias, pos, stop=3)                      # (N, k, C)
            peb = pointops.linear(h, lin2.weight, lin2.bias, out_f32=True)    # (N, k, C)
            upe = pointops.linear(h, wf, out_f32=True)                        # (N, k, G)
        with torch.autocast("cuda", enabled=False):
            const = lin_e.bias.float() if lin_e.bias is not None else None
            if lin2.bias is not None:
                cb = F.linear(lin2.bias.float(), we)
                const = cb if const is None else const + cb
            if (we_gather_enabled() and kp.shape[0] == reference_index.shape[0]
                    and pointops.we_tail_usable(kp, self.weight_encoding[1], reference_index.numel())):
                # opt-in: the G-wide relation gathered inside the tail kernels instead of stored once (measured slower)
                weight = pointops.we_tail(None, upe, const, self.weight_encoding[1], self.weight_encoding[3],
                                          gather=(kp, qp, reference_index))
                return pointops.gva_aggregate(value, peb, weight, reference_index, self.groups)
            rel = pointops.gva_relation(kp, qp, reference_index)              # (N, k, G)
            if pointops.we_tail_usable(rel, self.weight_encoding[1]):
                # u = rel + upe + const = weight_encoding[0](relation_qk); BN(G), ReLU, Linear(G,G) in one operator
                weight = pointops.we_tail(rel, upe, const, self.weight_encoding[1], self.weight_encoding[3])
                return pointops.gva_aggregate(value, peb, weight, reference_index, self.groups)
            u = rel + upe
            if const is not None:
                u = u + const
        weight = run_seq(self.weight_encoding, u, torch.float32, start=1)     # BN(G), ReLU, Linear(G,G)
        return pointops.gva_aggregate(value, peb, weight, reference_index, self.groups)   # :110,:119-128


class Block(nn.Module):
    def __init__(self, embed_channels, groups, qkv_bias=True, pe_multiplier=False, pe_bias=True,
                 attn_drop_rate=0.0, drop_path_rate=0.0, enable_checkpoint=False):
        super().__init__()
        self.attn = GroupedVectorAttention(embed_channels=embed_channels, groups=groups, qkv_bias=qkv_bias,
                                           attn_drop_rate=attn_drop_rate, pe_multiplier=pe_multiplier,
                                           pe_bias=pe_bias)
        self.fc1 = nn.Linear(embed_channels, embed_channels, bias=False)
        self.fc3 = nn.Linear(embed_channels, embed_channels, bias=False)
        self.norm1 = PointBatchNorm(embed_channels)
        self.norm2 = PointBatchNorm(embed_channels)
        self.norm3 = PointBatchNorm(embed_channels)
        self.act = nn.ReLU(inplace=True)
        self.enable_checkpoint = enable_checkpoint
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()

    def forward(self, points, reference_index, pos=None, pos_moments=None):
        coord, feat, offset = points
        identity = feat
        feat = pointops.linear_bn_act(feat, self.fc1, self.norm1.norm, relu=True)
        if self.enable_checkpoint:
            from torch.utils.checkpoint import checkpoint

            feat = checkpoint(self.attn, feat, coord, reference_index, pos, pos_moments, use_reentrant=False)
        else:
            feat = self.attn(feat, coord, reference_index, pos, pos_moments)
        # norm2 + ReLU written in the dtype fc3 consumes; norm3 + DropPath + residual + ReLU in one pass (:194-197)
        feat = pointops.bn_act(feat, self.norm2.norm, relu=True, out_dtype=_autocast_dtype())
        row_scale = None
        if isinstance(self.drop_path, DropPath) and self.drop_path.drop_prob > 0.0 and self.training:
            keep = 1.0 - self.drop_path.drop_prob
            row_scale = torch.empty(feat.shape[0], dtype=torch.float32, device=feat.device).bernoulli_(keep).div_(keep)
        feat = pointops.linear_bn_act(feat, self.fc3, self.norm3.norm, relu=True, residual=identity, row_scale=row_scale)
        return [coord, feat, offset]


class BlockSequence(nn.Module):
    def __init__(self, depth, embed_channels, groups, neighbours=16, qkv_bias=True, pe_multiplier=False,
                 pe_bias=True, attn_drop_rate=0.0, drop_path_rate=0.0, enable_checkpoint=False):
        super().__init__()
        if isinstance(drop_path_rate, list):
            drop_path_rates = drop_path_rate
            assert len(drop_path_rates) == depth
        elif isinstance(drop_path_rate, float):
            drop_path_rates = [deepcopy(drop_path_rate) for _ in range(depth)]
        else:
            drop_path_rates = [0.0 for _ in range(depth)]
        self.neighbours = neighbours
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(embed_channels=embed_channels, groups=groups, qkv_bias=qkv_bias,
                                     pe_multiplier=pe_multiplier, pe_bias=pe_bias, attn_drop_rate=attn_drop_rate,
                                     drop_path_rate=drop_path_rates[i], enable_checkpoint=enable_checkpoint))
        self.knn_cache = None  # set by PointTransformerV2: {(coord ptr, n, offset ptr, k): (coord, offset, idx, pos, moments)}
        self.search_neighbours = neighbours   # PointTransformerV2 raises it to the largest k used on this level

    def _neighbour_list(self, coord, offset, feat):
        """(idx, pos, pos moments) of this sequence's k, shared by every BlockSequence that sees the same coordinate
        tensor.  A level that is used with two different k (ScanNet cfg: patch-embed k=8, last decoder k=16 on the
        level-0 coordinates) is searched ONCE with the larger k: the result is ordered by (dist2, idx), so the smaller
        list is its first k columns."""
        k = self.neighbours
        cache = self.knn_cache

        def lookup(kk):
            hit = None if cache is None else cache.get((coord.data_ptr(), coord.shape[0], offset.data_ptr(), kk))
            # the entry pins its tensors: a data_ptr match alone could be a recycled allocation
            return hit if hit is not None and hit[0] is coord and hit[1] is offset else None

        hit = lookup(k)
        if hit is not None:
            return hit[2:]
        ks = max(self.search_neighbours, k)
        wide = lookup(ks) if ks != k else None
        if wide is not None:
            idx_wide = wide[2]
        else:
            idx_wide, _ = pointops.knn_query(ks, coord, offset)                          # :223
        training = torch.is_grad_enabled() and feat.requires_grad
        out = None
        for kk in sorted({k, ks}, reverse=True):
            if kk == ks and wide is not None:
                continue
            idx = idx_wide if kk == ks else idx_wide[:, :kk].contiguous()
            # relative coordinates of the neighbours (:109,:111) are the same for every block of the sequence
            pos = pointops.group_xyz(idx, coord)
            if training:
                # transposed neighbour graph of the atomic-free backward passes: built now, on a side stream,
                # under the forward kernels instead of at the head of the backward pass
                pointops.prefetch_csr(idx, coord.shape[0], 0)
            # Σp, Σppᵀ of pos: the closed-form BatchNorm statistics of every block's fused positional MLP
            mom = pointops.pos_moments(pos) if (fused_pe_enabled() and self.training) else None
            entry = (coord, offset, idx, pos, mom)
            if cache is not None:
                cache[(coord.data_ptr(), coord.shape[0], offset.data_ptr(), kk)] = entry
            if kk == k:
                out = entry[2:]
        return out

    def forward(self, points):
        coord, feat, offset = points
        reference_index, pos, mom = self._neighbour_list(coord, offset, feat)
        for block in self.blocks:
            points = block(points, reference_index, pos, mom)
        return points


class GridPool(nn.Module):
    """Partition-based Pooling (Grid Pooling) (…v2m2_base.py:229-269)."""

    def __init__(self, in_channels, out_channels, grid_size, bias=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.grid_size = grid_size
        self.fc = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = PointBatchNorm(out_channels)
        self.act = nn.ReLU(inplace=True)

    def forward(self, points, start=None):
        coord, feat, offset = points
        feat = pointops.linear_bn_act(feat, self.fc, self.norm.norm, relu=True, out_dtype=torch.float32)
        (coord, feat, offset), cluster, part = pointops.grid_pool(coord, feat.float().contiguous(), offset,
                                                                  self.grid_size, start, return_partition=True)
        cluster._aopt_c32 = part.cluster32      # lets UnpoolWithSkip("map") reuse the partition as its CSR
        return [coord, feat, offset], cluster


class UnpoolWithSkip(nn.Module):
    """Map / interpolation unpooling with skip connection (…v2m2_base.py:272-316)."""

    def __init__(self, in_channels, skip_channels, out_channels, bias=True, skip=True, backend="map"):
        super().__init__()
        self.in_channels = in_channels
        self.skip_channels = skip_channels
        self.out_channels = out_channels
        self.skip = skip
        self.backend = backend
        assert self.backend in ["map", "interp"]
        self.proj = nn.Sequential(nn.Linear(in_channels, out_channels, bias=bias), PointBatchNorm(out_channels),
                                  nn.ReLU(inplace=True))
        self.proj_skip = nn.Sequential(nn.Linear(skip_channels, out_channels, bias=bias),
                                       PointBatchNorm(out_channels), nn.ReLU(inplace=True))

    def forward(self, points, skip_points, cluster=None):
        coord, feat, offset = points
        skip_coord, skip_feat, skip_offset = skip_points
        if self.backend == "map" and cluster is not None:
            feat = pointops.unpool_map(run_seq(self.proj, feat), cluster)
        else:
            feat = pointops.interpolation(coord, skip_coord, run_seq(self.proj, feat, torch.float32).float().contiguous(),
                                          offset, skip_offset)
        if self.skip:
            feat = feat + run_seq(self.proj_skip, skip_feat)
        return [skip_coord, feat, skip_offset]


class Encoder(nn.Module):
    def __init__(self, depth, in_channels, embed_channels, groups, grid_size=None, neighbours=16, qkv_bias=True,
                 pe_multiplier=False, pe_bias=True, attn_drop_rate=None, drop_path_rate=None,
                 enable_checkpoint=False):
        super().__init__()
        self.down = GridPool(in_channels=in_channels, out_channels=embed_channels, grid_size=grid_size)
        self.blocks = BlockSequence(depth=depth, embed_channels=embed_channels, groups=groups, neighbours=neighbours,
                                    qkv_bias=qkv_bias, pe_multiplier=pe_multiplier, pe_bias=pe_bias,
                                    attn_drop_rate=attn_drop_rate if attn_drop_rate is not None else 0.0,
                                    drop_path_rate=drop_path_rate if drop_path_rate is not None else 0.0,
                                    enable_checkpoint=enable_checkpoint)

    def forward(self, points):
        points, cluster = self.down(points)
        return self.blocks(points), cluster


class Decoder(nn.Module):
    def __init__(self, in_channels, skip_channels, embed_channels, groups, depth, neighbours=16, qkv_bias=True,
                 pe_multiplier=False, pe_bias=True, attn_drop_rate=None, drop_path_rate=None,
                 enable_checkpoint=False, unpool_backend="map"):
        super().__init__()
        self.up = UnpoolWithSkip(in_channels=in_channels, out_channels=embed_channels, skip_channels=skip_channels,
                                 backend=unpool_backend)
        self.blocks = BlockSequence(depth=depth, embed_channels=embed_channels, groups=groups, neighbours=neighbours,
                                    qkv_bias=qkv_bias, pe_multiplier=pe_multiplier, pe_bias=pe_bias,
                                    attn_drop_rate=attn_drop_rate if attn_drop_rate is not None else 0.0,
                                    drop_path_rate=drop_path_rate if drop_path_rate is not None else 0.0,
                                    enable_checkpoint=enable_checkpoint)

    def forward(self, points, skip_points, cluster):
        points = self.up(points, skip_points, cluster)
        return self.blocks(points)


class GVAPatchEmbed(nn.Module):
    def __init__(self, depth, in_channels, embed_channels, groups, neighbours=16, qkv_bias=True,
                 pe_multiplier=False, pe_bias=True, attn_drop_rate=0.0, drop_path_rate=0.0,
                 enable_checkpoint=False):
        super().__init__()
        self.in_channels = in_channels
        self.embed_channels = embed_channels
        self.proj = nn.Sequential(nn.Linear(in_channels, embed_channels, bias=False), PointBatchNorm(embed_channels),
                                  nn.ReLU(inplace=True))
        self.blocks = BlockSequence(depth=depth, embed_channels=embed_channels, groups=groups, neighbours=neighbours,
                                    qkv_bias=qkv_bias, pe_multiplier=pe_multiplier, pe_bias=pe_bias,
                                    attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate,
                                    enable_checkpoint=enable_checkpoint)

    def forward(self, points):
        coord, feat, offset = points
        feat = run_seq(self.proj, feat)
        return self.blocks([coord, feat, offset])


class PointTransformerV2(nn.Module):
    """PT-v2m2 (…v2m2_base.py:447-576).  Defaults equal the reference's; S3DIS_CFG below is
    configs/s3dis/semseg-pt-v2m2-0-base.py:10-36."""

    def __init__(self, in_channels, num_classes, patch_embed_depth=1, patch_embed_channels=48,
                 patch_embed_groups=6, patch_embed_neighbours=8, enc_depths=(2, 2, 6, 2),
                 enc_channels=(96, 192, 384, 512), enc_groups=(12, 24, 48, 64), enc_neighbours=(16, 16, 16, 16),
                 dec_depths=(1, 1, 1, 1), dec_channels=(48, 96, 192, 384), dec_groups=(6, 12, 24, 48),
                 dec_neighbours=(16, 16, 16, 16), grid_sizes=(0.06, 0.12, 0.24, 0.48), attn_qkv_bias=True,
                 pe_multiplier=False, pe_bias=True, attn_drop_rate=0.0, drop_path_rate=0, enable_checkpoint=False,
                 unpool_backend="map"):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.num_stages = len(enc_depths)
        for t in (dec_depths, enc_channels, dec_channels, enc_groups, dec_groups, enc_neighbours, dec_neighbours,
                  grid_sizes):
            assert self.num_stages == len(t)
        self.patch_embed = GVAPatchEmbed(in_channels=in_channels, embed_channels=patch_embed_channels,
                                         groups=patch_embed_groups, depth=patch_embed_depth,
                                         neighbours=patch_embed_neighbours, qkv_bias=attn_qkv_bias,
                                         pe_multiplier=pe_multiplier, pe_bias=pe_bias, attn_drop_rate=attn_drop_rate,
                                         enable_checkpoint=enable_checkpoint)
        enc_dp_rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(enc_depths))]
        dec_dp_rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(dec_depths))]
        enc_channels = [patch_embed_channels] + list(enc_channels)
        dec_channels = list(dec_channels) + [enc_channels[-1]]
        self.enc_stages = nn.ModuleList()
        self.dec_stages = nn.ModuleList()
        for i in range(self.num_stages):
            self.enc_stages.append(Encoder(
                depth=enc_depths[i], in_channels=enc_channels[i], embed_channels=enc_channels[i + 1],
                groups=enc_groups[i], grid_size=grid_sizes[i], neighbours=enc_neighbours[i], qkv_bias=attn_qkv_bias,
                pe_multiplier=pe_multiplier, pe_bias=pe_bias, attn_drop_rate=attn_drop_rate,
                drop_path_rate=enc_dp_rates[sum(enc_depths[:i]): sum(enc_depths[: i + 1])],
                enable_checkpoint=enable_checkpoint))
            self.dec_stages.append(Decoder(
                depth=dec_depths[i], in_channels=dec_channels[i + 1], skip_channels=enc_channels[i],
                embed_channels=dec_channels[i], groups=dec_groups[i], neighbours=dec_neighbours[i],
                qkv_bias=attn_qkv_bias, pe_multiplier=pe_multiplier, pe_bias=pe_bias, attn_drop_rate=attn_drop_rate,
                drop_path_rate=dec_dp_rates[sum(dec_depths[:i]): sum(dec_depths[: i + 1])],
                enable_checkpoint=enable_checkpoint, unpool_backend=unpool_backend))
        self.seg_head = (nn.Sequential(nn.Linear(dec_channels[0], dec_channels[0]), PointBatchNorm(dec_channels[0]),
                                       nn.ReLU(inplace=True), nn.Linear(dec_channels[0], num_classes))
                         if num_classes > 0 else nn.Identity())
        self.hoist_pyramid = True
        # neighbour lists are shared between every BlockSequence that sees the same (coords, k)
        self._knn_cache = {}
        for m in self.modules():
            if isinstance(m, BlockSequence):
                m.knn_cache = self._knn_cache
        # largest k per level: level 0 = patch embed + last decoder, level i+1 = encoder i + decoder i+1
        seqs = [[self.patch_embed.blocks, self.dec_stages[0].blocks]]
        for i in range(self.num_stages):
            seqs.append([self.enc_stages[i].blocks] + ([self.dec_stages[i + 1].blocks] if i + 1 < self.num_stages else []))
        for level in seqs:
            kmax = max(q.neighbours for q in level)
            for q in level:
                q.search_neighbours = kmax

    def forward(self, data_dict):
        coord = data_dict["coord"]
        feat = data_dict["feat"]
        offset = data_dict["offset"].int()
        self._knn_cache.clear()
        try:
            points = [coord, feat, offset]
            if self.hoist_pyramid:
                # all voxel partitions + coarse coordinates first: they depend on coordinates only, and they hold
                # the forward pass's host synchronisations (pointops.prepare_pyramid)
                pointops.prepare_pyramid(coord, offset, [enc.down.grid_size for enc in self.enc_stages],
                                         knn=[enc.blocks.search_neighbours for enc in self.enc_stages],
                                         interp_k=3 if self.dec_stages[0].up.backend == "interp" else None)
            points = self.patch_embed(points)
            skips = [[points]]
            for i in range(self.num_stages):
                points, cluster = self.enc_stages[i](points)
                skips[-1].append(cluster)
                skips.append([points])
            points = skips.pop(-1)[0]
            for i in reversed(range(self.num_stages)):
                skip_points, cluster = skips.pop(-1)
                points = self.dec_stages[i](points, skip_points, cluster)
            coord, feat, offset = points
            # logits in fp32: the loss upcasts them anyway, and the 13- / 19- / 20-output Linear has its own kernels
            return run_seq(self.seg_head, feat, torch.float32) if isinstance(self.seg_head, nn.Sequential) else self.seg_head(feat)
        finally:
            self._knn_cache.clear()   # the idx tensors stay alive through autograd; do not pin them here


S3DIS_CFG = dict(  # configs/s3dis/semseg-pt-v2m2-0-base.py:10-36
    in_channels=6, num_classes=13, patch_embed_depth=2, patch_embed_channels=48, patch_embed_groups=6,
    patch_embed_neighbours=16, enc_depths=(2, 6, 2), enc_channels=(96, 192, 384), enc_groups=(12, 24, 48),
    enc_neighbours=(16, 16, 16), dec_depths=(1, 1, 1), dec_channels=(48, 96, 192), dec_groups=(6, 12, 24),
    dec_neighbours=(16, 16, 16), grid_sizes=(0.1, 0.2, 0.4), attn_qkv_bias=True, pe_multiplier=False,
    pe_bias=True, attn_drop_rate=0.0, drop_path_rate=0.3, enable_checkpoint=False, unpool_backend="interp")

SCANNET_CFG = dict(  # configs/scannet/semseg-pt-v2m2-0-base.py
    in_channels=9, num_classes=20, patch_embed_depth=1, patch_embed_channels=48, patch_embed_groups=6,
    patch_embed_neighbours=8, enc_depths=(2, 2, 6, 2), enc_channels=(96, 192, 384, 512),
    enc_groups=(12, 24, 48, 64), enc_neighbours=(16, 16, 16, 16), dec_depths=(1, 1, 1, 1),
    dec_channels=(48, 96, 192, 384), dec_groups=(6, 12, 24, 48), dec_neighbours=(16, 16, 16, 16),
    grid_sizes=(0.06, 0.15, 0.375, 0.9375), attn_qkv_bias=True, pe_multiplier=False, pe_bias=True,
    attn_drop_rate=0.0, drop_path_rate=0.3, enable_checkpoint=False, unpool_backend="map")

KITTI_CFG = dict(SCANNET_CFG, in_channels=4, num_classes=19,   # configs/semantic_kitti/semseg-pt-v2m2-0-base.py
                 grid_sizes=(0.15, 0.375, 0.9375, 2.34375))
