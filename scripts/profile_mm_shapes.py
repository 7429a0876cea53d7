"""Which aten::mm / addmm / bmm calls of one PTv2m2 training step are slow: input shapes, strides and CUDA time per call
(torch profiler with record_shapes), sorted by CUDA time.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from ao_b200 import ptv2, scenes

dev = torch.device("cuda", 0)
coord_np, feat_np, off_np = scenes.s3dis_batch(int(os.environ.get("ROOMS", 4)), 80000)
coord, feat, offset = (torch.from_numpy(a).to(dev) for a in (coord_np, feat_np, off_np))
torch.manual_seed(0)
model = ptv2.PointTransformerV2(**ptv2.S3DIS_CFG).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
target = torch.randint(0, 13, (coord.shape[0],), device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(dict(coord=coord, feat=feat, offset=offset))
    loss = torch.nn.functional.cross_entropy(logits.float(), target)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.name in ("aten::mm", "aten::addmm", "aten::bmm"):
        cuda_us = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        rows.append((cuda_us, e.name, str(e.input_shapes)))
rows.sort(reverse=True)
print("top GEMM calls by CUDA time (us, op, input shapes):")
for r in rows[:40]:
    print(f"{r[0]:9.1f}  {r[1]:12s} {r[2]}")
print("total GEMM calls", len(rows), "total us", sum(r[0] for r in rows))
