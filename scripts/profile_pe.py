"""pe_mlp forward + backward once at level-0 shapes (320k x 16 rows, C=48), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from ao_b200 import pointops, ptv2
c = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = 320000 if c == 48 else 50000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pos = 0.1 * torch.randn(n, 16, 3, device=dev)
mlp = nn.Sequential(nn.Linear(3, c), ptv2.PointBatchNorm(c), nn.ReLU(inplace=True), nn.Linear(c, c)).to(dev).train()
mom = pointops.pos_moments(pos)
def run():
    y = pointops.pe_bias_mlp(pos, mlp, mom)
    torch.autograd.grad(y, list(mlp.parameters()), torch.ones_like(y), allow_unused=True)
    torch.cuda.synchronize()
run()
torch.cuda.profiler.start()
run()
torch.cuda.profiler.stop()
