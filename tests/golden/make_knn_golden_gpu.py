"""Generates tests/golden/knn_ref_cuda.npz ON THE GPU BOX: outputs of the UNMODIFIED reference kNN
launcher (oracle/_ref/libpointops_ref.so, built from
/root/reference/libs/pointops/src/knn_query/knn_query_cuda_kernel.cu by oracle/Makefile.ref) on seeded
inputs from ao_b200.scenes.  The inputs are regenerated from the seeds by the tests, so only the
outputs are stored.  Run:  gpurun -- 'python tests/golden/make_knn_golden_gpu.py gpurun_out/knn_ref_cuda.npz'
then copy the file to tests/golden/.  tests/test_oracle.py checks the C oracle against it on CPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from ao_b200 import scenes
from oracle import ref_cuda

CASES = {
    # name: (generator, kwargs, k, cross)
    "small_k16": ("small_batch", dict(seed=21, sizes=(700, 5, 1300, 257)), 16, False),
    "small_k3_cross": ("small_batch", dict(seed=22, sizes=(400, 90, 33)), 3, True),
    "small_k1_cross": ("small_batch", dict(seed=23, sizes=(400, 90, 33)), 1, True),
    "room_k16": ("s3dis_batch", dict(n_rooms=2, n_points=4000), 16, False),
    "room_k32": ("s3dis_batch", dict(n_rooms=1, n_points=3000), 32, False),
}


def inputs(name):
    gen, kw, k, cross = CASES[name]
    coord, feat, offset = getattr(scenes, gen)(**kw)
    if not cross:
        return k, coord, offset, coord, offset
    # cross-set: queries = a jittered copy of every 2nd..: deterministic "fine" set per scene
    rng = np.random.default_rng(1234)
    q, qo, s = [], [], 0
    for e in offset:
        pts = coord[s:e]
        rep = np.repeat(pts, 3, axis=0) + rng.normal(0, 0.05, (3 * len(pts), 3)).astype(np.float32)
        q.append(rep.astype(np.float32))
        qo.append(len(rep))
        s = e
    return k, coord, offset, np.concatenate(q), np.cumsum(qo).astype(np.int32)


if __name__ == "__main__":
    out = {}
    for name in CASES:
        k, xyz, off, q, qoff = inputs(name)
        d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (xyz, off, q, qoff)]
        idx, d2 = ref_cuda.knn_query(k, d[0], d[1], d[2], d[3])
        out[name + "_idx"] = idx.cpu().numpy()
        out[name + "_d2bits"] = d2.cpu().numpy().view(np.uint32)
    np.savez_compressed(sys.argv[1], **out)
    print("wrote", sys.argv[1], {k: v.shape for k, v in out.items()})
