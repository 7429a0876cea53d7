"""Experiment: gva_backward_value at level-0 shapes, optional Morton presort of the points (argv: presort|plain).
Under ncu (--profile-from-start off) only the bracketed launch is captured."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ao_b200 import _lib, pointops, scenes

presort = len(sys.argv) > 1 and sys.argv[1] == "presort"
dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
if presort:
    order, s0 = [], 0
    for e0 in off_np:
        c = coord_np[s0:e0]
        cell = np.floor((c - c.min(0)) / 0.1).astype(np.int64)
        key = np.zeros(len(c), np.int64)
        for bit in range(10):
            for a in range(3):
                key |= ((cell[:, a] >> bit) & 1) << (3 * bit + a)
        order.append(s0 + np.argsort(key, kind="stable"))
        s0 = e0
    coord_np = np.ascontiguousarray(coord_np[np.concatenate(order)])
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
n, k, c, g = coord.shape[0], 16, 48, 6
idx, _ = pointops.knn_query(k, coord, offset)
csr = pointops.get_csr(idx, n)
deg = (csr.rowptr[1:] - csr.rowptr[:-1]).float()
print("presort", presort, "in-degree mean %.2f max %d" % (deg.mean().item(), int(deg.max().item())))
# index distance between a source and its in-edge queries (locality of the gather)
q = (csr.perm.long() >> 4)
src = torch.repeat_interleave(torch.arange(n, device=dev), (csr.rowptr[1:] - csr.rowptr[:-1]).long())
dd = (q - src).abs().float()
print("median |q - j| = %.0f, 90%% = %.0f" % (dd.median().item(), dd.quantile(0.9).item()))
torch.manual_seed(0)
g_out = torch.randn(n, c, device=dev)
prob = torch.softmax(torch.randn(n, k, g, device=dev), 1).contiguous()
gval = torch.empty(n, c, device=dev)
lib = _lib.load()


def run():
    lib.aopt_gva_backward_value(n, k, c, g, g_out.data_ptr(), prob.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(), gval.data_ptr(), _lib.stream())


for _ in range(3):
    run()
torch.cuda.synchronize()
best = 1e9
for _ in range(7):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) * 1e3)
print("gva_backward_value %.1f us  (impl=%s)" % (best, os.environ.get("AOPT_BV_IMPL", "default")))
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
