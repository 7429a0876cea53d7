"""Generates tests/golden/fps_ref_cuda.npz ON THE GPU BOX: outputs of the UNMODIFIED reference farthest point
sampling launcher (oracle/_ref/libpointops_ref.so, built from
/root/reference/libs/pointops/src/sampling/sampling_cuda_kernel.cu by oracle/Makefile.ref) on seeded inputs from
ao_b200.scenes.  Inputs are regenerated from the seeds by the tests; only the outputs are stored.  Run
  gpurun -- 'python tests/golden/make_fps_golden_gpu.py gpurun_out/fps_ref_cuda.npz'
then copy the file to tests/golden/.  tests/test_oracle.py checks oracle/fps_oracle.c against it on CPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from ao_b200 import scenes

CASES = {
    # name: (generator, kwargs, stride)   new scene size = max(size // stride, 1)
    "small_s4": ("small_batch", dict(seed=31, sizes=(700, 5, 1300, 257)), 4),
    "small_dup_s2": ("small_batch", dict(seed=32, sizes=(300, 64, 9), dup=40), 2),      # duplicated points: exact ties
    "small_all": ("small_batch", dict(seed=33, sizes=(130, 17)), 1),                    # every point is selected
    "room_s4": ("s3dis_batch", dict(n_rooms=2, n_points=6000), 4),
    "room_s16": ("s3dis_batch", dict(n_rooms=1, n_points=5000), 16),
}


def inputs(name):
    gen, kw, stride = CASES[name]
    coord, feat, offset = getattr(scenes, gen)(**kw)
    sizes = np.diff(np.concatenate([[0], offset]))
    new_offset = np.cumsum(np.maximum(sizes // stride, 1)).astype(np.int32)
    return np.ascontiguousarray(coord), offset.astype(np.int32), new_offset


if __name__ == "__main__":
    import torch

    from oracle import ref_cuda

    out = {}
    for name in CASES:
        xyz, off, noff = inputs(name)
        d = [torch.from_numpy(a).cuda() for a in (xyz, off, noff)]
        idx, tmp = ref_cuda.farthest_point_sampling(*d)
        out[name + "_idx"] = idx.cpu().numpy()
        out[name + "_tmpbits"] = tmp.cpu().numpy().view(np.uint32)
    np.savez_compressed(sys.argv[1], **out)
    print("wrote", sys.argv[1], {k: v.shape for k, v in out.items()})
