"""Per-step device time of the full PTv2m2 (S3DIS cfg) training step, 10 steps, to separate warm-up
(allocator growth, cuBLAS heuristics) from the steady state.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import torch

from ao_b200 import ptv2, scenes

dev = torch.device("cuda", 0)
coord_np, feat_np, off_np = scenes.s3dis_batch(int(os.environ.get('ROOMS', 4)), 80000)
coord, feat, offset = (torch.from_numpy(a).to(dev) for a in (coord_np, feat_np, off_np))
torch.manual_seed(0)
model = ptv2.PointTransformerV2(**ptv2.S3DIS_CFG).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=os.environ.get('ADAMW_FUSED', '1') == '1')
target = torch.randint(0, 13, (coord.shape[0],), device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(dict(coord=coord, feat=feat, offset=offset))
    loss = torch.nn.functional.cross_entropy(logits.float(), target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for i in range(10):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: device {e0.elapsed_time(e1):8.2f} ms   wall {(time.perf_counter()-w0)*1e3:8.2f} ms   fused_pe={ptv2.fused_pe_enabled()}")
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
