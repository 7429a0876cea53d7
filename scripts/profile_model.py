"""torch.profiler breakdown of one full PTv2m2 (S3DIS cfg) training step on 4 x 80k points:
which kernels (ours vs torch/cuBLAS) the 100+ ms go to.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from ao_b200 import ptv2, scenes

dev = torch.device("cuda", 0)
coord_np, feat_np, off_np = scenes.s3dis_batch(4, 80000)
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


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", 45)), max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=int(os.environ.get("CPU_ROWS", 30)), max_name_column_width=70))
