"""world_size-2 gloo test of the N>1 host logic: scene sharding needs no data-path collective
(per-rank neighbour search on the shard == the global search), timings are max-reduced, and the
gradient all-reduce sums.  CPU only: the per-shard search is done by the C oracle (test
infrastructure), which is exactly the point — the shard boundaries, not the kernel, are under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ao_b200 import scenes, sharding
    from oracle import torch_ref

    coord, feat, offset = scenes.small_batch(5, sizes=(300, 7, 450, 120, 260))
    t_coord, t_feat, t_off = torch.from_numpy(coord), torch.from_numpy(feat), torch.from_numpy(offset)
    c_r, f_r, o_r, base = sharding.shard_batch(t_coord, t_feat, t_off, rank, world)
    idx_r, d2_r = torch_ref.knn_query(8, c_r.numpy(), o_r.numpy(), rule="lex")
    idx_g, d2_g = torch_ref.knn_query(8, coord, offset, rule="lex")
    lo, hi = base, base + c_r.shape[0]
    glob = np.where(idx_r >= 0, idx_r + base, -1)
    ok = np.array_equal(glob, idx_g[lo:hi]) and np.array_equal(d2_r.view(np.uint32), d2_g[lo:hi].view(np.uint32))
    # every point is owned by exactly one rank
    owned = torch.zeros(coord.shape[0])
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    ok = ok and bool((owned == 1).all())
    # the one collective of the path: gradient sum
    g = torch.full((1000,), float(rank + 1))
    dist.all_reduce(g)
    ok = ok and bool((g == sum(range(1, world + 1))).all())
    ok = ok and sharding.max_over_ranks(float(rank), "cpu") == float(world - 1)
    # ddp_wrap (the reference's create_ddp_model, engines/defaults.py:30-43) on a host model built from the PTv2 module
    # classes: several small buckets, gradients = mean over ranks of the per-rank gradients (DDP's contract)
    from ao_b200 import ptv2

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 32), ptv2.PointBatchNorm(32), torch.nn.ReLU(), torch.nn.Linear(32, 32),
                              ptv2.PointBatchNorm(32), torch.nn.ReLU(), torch.nn.Linear(32, 5))
    ref = __import__("copy").deepcopy(net)
    ddp = sharding.ddp_wrap(net, None, bucket_cap_mb=0.002)
    xs = [torch.randn(40 + 10 * r, 6, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    ddp(xs[rank]).square().mean().backward()
    want = [torch.zeros_like(p_) for p_ in ref.parameters()]
    for r in range(world):
        ref.zero_grad()
        ref(xs[r]).square().mean().backward()
        for w_, p_ in zip(want, ref.parameters()):
            w_ += p_.grad / world
    ok = ok and all(torch.allclose(p_.grad, w_, rtol=1e-5, atol=1e-6) for p_, w_ in zip(net.parameters(), want))
    ok = ok and len(ddp.reducer._get_bucket_tensors() if hasattr(ddp.reducer, "_get_bucket_tensors") else [0, 0]) >= 2
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_scene_range_partitions():
    from ao_b200 import sharding

    for n in (1, 4, 7, 64):
        for world in (1, 2, 3, 8):
            spans = [sharding.scene_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.scene_range(4, 2, 2)


def test_two_rank_gloo_scene_sharding(tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]
