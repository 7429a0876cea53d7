"""One call of every hot-path operator at BASELINE.json configs[1] level-0 shapes (4 rooms x 80k, k=16,
C=48, G=6) plus the level-0→1 pool and level-1→0 interpolation, bracketed by cudaProfilerStart/Stop
so that `ncu --profile-from-start off` captures exactly these launches (scripts/gpu_ncu.sh).  Round 2 adds the
dense.cu operators (bn_act, we_tail) in front.
Optional argv[1] = level (0..3) whose block shapes to use."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ao_b200 import pointops, scenes

level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
C = (48, 96, 192, 384)
G = (6, 12, 24, 48)
GRID = (0.1, 0.2, 0.4)
dev = torch.device("cuda", 0)
coord_np, feat_np, off_np = scenes.s3dis_batch(4, 80000)
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
for li in range(level):
    (coord, _, offset), _ = pointops.grid_pool(coord, coord.clone(), offset, GRID[li])
    offset = offset.int()
n, k, c, g = coord.shape[0], 16, C[level], G[level]
torch.manual_seed(0)
key, query, value = (torch.randn(n, c, device=dev, requires_grad=True) for _ in range(3))
peb = torch.randn(n, k, c, device=dev, requires_grad=True)
logits = torch.randn(n, k, g, device=dev, requires_grad=True)
g_rel, g_out = torch.randn(n, k, c, device=dev), torch.randn(n, c, device=dev)
c_next = C[min(level + 1, 3)]
pool_in = torch.relu(torch.randn(n, c_next, device=dev)).requires_grad_(True)


mlp = None
if pointops.pe_mlp_supported(c):
    mlp = torch.nn.Sequential(torch.nn.Linear(3, c), torch.nn.BatchNorm1d(c), torch.nn.ReLU(inplace=True),
                              torch.nn.Linear(c, c)).to(dev).train()
    aux_w = torch.randn(g, c, device=dev, requires_grad=True)


bn_c = torch.nn.BatchNorm1d(c).to(dev).train()
bn_g = torch.nn.BatchNorm1d(g).to(dev).train()
lin_g = torch.nn.Linear(g, g).to(dev)
x_bf = torch.randn(n, c, device=dev).bfloat16().requires_grad_(True)
rel_g = torch.randn(n, k, g, device=dev, requires_grad=True)
upe_g = torch.randn(n, k, g, device=dev, requires_grad=True)


def run():
    # dense.cu: BatchNorm + ReLU on (n, c) bf16 and the weight-encoding tail on (n*k, g), forward + backward
    yb = pointops.bn_act(x_bf, bn_c, relu=True)
    torch.autograd.grad(yb, [x_bf, bn_c.weight, bn_c.bias], torch.ones_like(yb))
    if pointops.we_tail_supported(g):
        yl = pointops.we_tail(rel_g, upe_g, None, bn_g, lin_g)
        torch.autograd.grad(yl, [rel_g, upe_g, lin_g.weight], torch.ones_like(yl))
    idx, _ = pointops.knn_query(k, coord, offset)
    pos = pointops.group_xyz(idx, coord)
    if mlp is not None:      # fused positional-bias MLP (tcgen05 forward) with the auxiliary head, forward + backward
        mom = pointops.pos_moments(pos)
        y, u = pointops.pe_bias_mlp(pos, mlp, mom, aux_weight=aux_w)
        torch.autograd.grad([y, u], list(mlp.parameters()) + [aux_w], [torch.ones_like(y), torch.ones_like(u)])
    rel = pointops.gva_relation(key, query, idx)
    out = pointops.gva_aggregate(value, peb, logits, idx, g)
    torch.autograd.grad(rel, [key, query], g_rel)
    torch.autograd.grad(out, [value, peb, logits], g_out)
    if level < 3:
        (nc, nf, noff), cluster = pointops.grid_pool(coord, pool_in, offset, GRID[level])
        torch.autograd.grad(nf, [pool_in], torch.randn_like(nf))
        src = torch.randn(nc.shape[0], c, device=dev, requires_grad=True)
        up = pointops.interpolation(nc, coord, src, noff.int(), offset, k=3)
        torch.autograd.grad(up, [src], torch.randn_like(up))
    torch.cuda.synchronize()


run()
torch.cuda.profiler.start()
run()
torch.cuda.profiler.stop()
print("profile_ops: level", level, "n", n, "c", c)
