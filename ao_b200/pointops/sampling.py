"""farthest_point_sampling is used by PTv1 / Stratified Transformer only
(/root/reference/pointcept/models/point_transformer/point_transformer_seg.py:101); it is outside the
PTv2m2 hot path and not built."""


def farthest_point_sampling(*args, **kwargs):
    raise NotImplementedError("ao_b200.pointops.farthest_point_sampling: outside the PTv2m2 hot path (not built)")
