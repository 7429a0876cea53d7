"""Drop-in replacement for the reference `pointops` package
(/root/reference/libs/pointops/functions/__init__.py:1-14): same public names and signatures for
the PTv2m2 hot path, backed by hand-written sm_100a kernels behind a C ABI (include/ao_pointops.h).

Names outside the hot path (ball_query, random_ball_query, farthest_point_sampling,
attention_*_step, ball_query_and_group) exist but raise NotImplementedError.
New fused operators for the PTv2 caller: group_xyz, gva_relation, gva_aggregate, grid_pool,
voxel_partition, unpool_map, interpolation_weights, knn_query_raw, pe_bias_mlp (fused positional-bias MLP),
pos_moments, vote_accumulate (the tester's softmax + fragment vote), bn_act (training-mode BatchNorm + ReLU / DropPath / residual on (rows, C)), we_tail (weight-encoding tail).
"""
from .query import knn_query, knn_query_raw, knn_query_sets, prefetch_knn, ball_query, random_ball_query
from .sampling import farthest_point_sampling
from .grouping import grouping, grouping2
from .interpolation import interpolation, interpolation2, interpolation_weights
from .subtraction import subtraction
from .aggregation import aggregation
from .attention import (
    attention_relation_step,
    attention_fusion_step,
    group_xyz,
    gva_relation,
    gva_aggregate,
)
from .pooling import grid_pool, voxel_partition, unpool_map, VoxelPartition, prepare_pyramid, pool_coord
from .utils import (
    query_and_group,
    knn_query_and_group,
    ball_query_and_group,
    batch2offset,
    offset2batch,
)
from .pe_mlp import pe_bias_mlp, pe_mlp_supported, pos_moments
from .dense import linear, linear_bn_act, col_sum, qkv_bn, qkv_usable, bn_act, bn_act_supported, bn_act_usable, bn_fusable, we_tail, we_tail_supported, we_tail_usable, fused_dense_enabled
from ._csr import get_csr, build_csr, prefetch_csr
from .vote import vote_accumulate
