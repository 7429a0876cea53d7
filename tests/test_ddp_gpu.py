"""The DDP-wrapped PTv2m2 step: ao_b200.sharding.ddp_wrap == the reference's create_ddp_model
(/root/reference/pointcept/engines/defaults.py:30-43, engines/train.py:209-213).  One NCCL rank per visible GPU
(two when the box has them, else a single-rank group): gradients after `backward()` through the wrapper equal the
mean over ranks of the gradients of the un-wrapped model on every rank's scene shard."""
import copy
import os
import socket

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import sys

    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ao_b200 import ptv2, scenes, sharding

    coord, feat, offset = scenes.s3dis_batch(2 * world, n_points=3000)
    t = [torch.from_numpy(x).to(dev) for x in (coord, feat, offset)]
    shards = [sharding.shard_batch(t[0], t[1], t[2], r, world) for r in range(world)]
    torch.manual_seed(0)
    model = ptv2.PointTransformerV2(**dict(ptv2.S3DIS_CFG, drop_path_rate=0.0)).to(dev).train()
    plain = copy.deepcopy(model)
    ddp = sharding.ddp_wrap(model, rank, bucket_cap_mb=1.0)

    def loss_of(net, shard):
        c, f, o, _ = shard
        logits = net(dict(coord=c, feat=f, offset=o))
        target = (torch.arange(c.shape[0], device=dev) % 13)
        return torch.nn.functional.cross_entropy(logits, target)

    loss_of(ddp, shards[rank]).backward()
    want = [torch.zeros_like(p) for p in plain.parameters()]
    for r in range(world):
        plain.zero_grad(set_to_none=True)
        loss_of(plain, shards[r]).backward()
        for w, p in zip(want, plain.parameters()):
            w += p.grad / world
    worst = max(float((p.grad - w).abs().max() / (w.abs().max() + 1e-6)) for p, w in zip(model.parameters(), want))
    n_buckets = len(ddp.reducer._get_bucket_tensors()) if hasattr(ddp.reducer, "_get_bucket_tensors") else -1
    open(os.path.join(tmp, f"ok{rank}"), "w").write(f"{worst:.3e} {n_buckets}")
    dist.destroy_process_group()


def test_ddp_wrapped_ptv2_gradients(tmp_path):
    import torch.multiprocessing as mp

    world = 2 if torch.cuda.device_count() >= 2 else 1
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        worst, n_buckets = open(tmp_path / f"ok{r}").read().split()
        # fp32 end to end; kernels are deterministic, so the only difference is the all-reduce's summation order
        assert float(worst) < 1e-4, f"rank {r}: DDP gradient differs from the mean of the per-shard gradients ({worst})"
        assert int(n_buckets) != 1, "expected several buckets (overlap with backward), got one"
