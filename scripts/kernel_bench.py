"""Per-operator timing at BASELINE.json configs[1] level shapes (4 rooms x 80k; levels 0..3), CUDA events,
best of 7 after warm-up.  Prints us, algorithmic GB/s and the fraction of the measured HBM peak.
  python scripts/kernel_bench.py [--presort] [--levels 0,1]
Kernel variants are selected with the AOPT_* environment variables read by the library."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ao_b200 import _lib, pointops, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--presort", action="store_true")
ap.add_argument("--levels", default="0")
args = ap.parse_args()
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
if args.presort:
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
C = (48, 96, 192, 384)
G = (6, 12, 24, 48)
GRID = (0.1, 0.2, 0.4)
K = 16
levels = [(torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev))]
for gs in GRID:
    c, o = levels[-1]
    (nc, _, no), _ = pointops.grid_pool(c, c.clone(), o, gs)
    levels.append((nc.contiguous(), no.int()))


def timeit(fn, reps=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3)
    return best


def row(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:34s} {us:9.1f} us {nbytes/1e6:9.1f} MB {gbs:8.0f} GB/s {gbs/PEAK:6.1%}")


for li in [int(x) for x in args.levels.split(",")]:
    coord, offset = levels[li]
    n, c, g, k = coord.shape[0], C[li], G[li], K
    print(f"--- level {li}: n={n} c={c} g={g} k={k} presort={args.presort}")
    torch.manual_seed(0)
    idx, _ = pointops.knn_query(k, coord, offset)
    key, query, value = (torch.randn(n, c, device=dev, requires_grad=True) for _ in range(3))
    peb = torch.randn(n, k, c, device=dev, requires_grad=True)
    logits = torch.randn(n, k, g, device=dev, requires_grad=True)
    g_rel, g_out = torch.randn(n, k, c, device=dev), torch.randn(n, c, device=dev)
    row("knn_query (self)", timeit(lambda: pointops.knn_query_raw(k, coord, offset)), 24.0 * n + 8.0 * n * k)
    from ao_b200.pointops._csr import build_csr
    row("csr_build", timeit(lambda: build_csr(idx, n)), 8.0 * n * k + 4.0 * (n + 1))
    row("group_xyz", timeit(lambda: pointops.group_xyz(idx, coord)), 24.0 * n + 16.0 * n * k)
    row("gather_sub fwd", timeit(lambda: pointops.gva_relation(key, query, idx)), 4.0 * n * k + 8.0 * n * c + 4.0 * n * k * c)
    rel = pointops.gva_relation(key, query, idx)
    lib = _lib.load()
    csr = pointops.get_csr(idx, n)
    gk = torch.empty(n, c, device=dev)
    row("segmented_sum (grad_key)", timeit(lambda: lib.aopt_grouping_backward(n, c, g_rel.data_ptr(), c, csr.rowptr.data_ptr(), csr.perm.data_ptr(), 1.0, gk.data_ptr(), _lib.stream())),
        4.0 * n * k * c + 4.0 * n * k + 4.0 * (n + 1) + 4.0 * n * c)
    row("sum_over_k (grad_query)", timeit(lambda: lib.aopt_sum_over_k(n, k, c, g_rel.data_ptr(), -1.0, gk.data_ptr(), _lib.stream())), 4.0 * n * k * c + 4.0 * n * c)
    gq = torch.empty(n, c, device=dev)
    row("relation_backward (fused pair)", timeit(lambda: lib.aopt_relation_backward(n, k, c, g_rel.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(), gk.data_ptr(), gq.data_ptr(), _lib.stream())),
        4.0 * n * k * c + 4.0 * n * k + 4.0 * (n + 1) + 8.0 * n * c)
    out = torch.empty(n, c, device=dev)
    prob = torch.empty(n, k, g, device=dev)
    row("gva_forward", timeit(lambda: lib.aopt_gva_forward(n, k, c, g, value.data_ptr(), peb.data_ptr(), logits.data_ptr(), idx.data_ptr(), out.data_ptr(), prob.data_ptr(), _lib.stream())),
        8.0 * n * c + 4.0 * n * k * c + 8.0 * n * k * g + 4.0 * n * k)
    gpeb, glog, gval = torch.empty(n, k, c, device=dev), torch.empty(n, k, g, device=dev), torch.empty(n, c, device=dev)
    row("gva_backward_query", timeit(lambda: lib.aopt_gva_backward_query(n, k, c, g, g_out.data_ptr(), value.data_ptr(), peb.data_ptr(), prob.data_ptr(), idx.data_ptr(), gpeb.data_ptr(), glog.data_ptr(), _lib.stream())),
        8.0 * n * c + 8.0 * n * k * c + 8.0 * n * k * g + 4.0 * n * k)
    row("gva_backward_value", timeit(lambda: lib.aopt_gva_backward_value(n, k, c, g, g_out.data_ptr(), prob.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(), gval.data_ptr(), _lib.stream())),
        4.0 * n * k * g + 8.0 * n * c + 4.0 * (n + 1) + 4.0 * n * k)
    row("gva_backward (fused q+v)", timeit(lambda: lib.aopt_gva_backward(n, k, c, g, g_out.data_ptr(), value.data_ptr(), peb.data_ptr(), prob.data_ptr(), idx.data_ptr(), csr.rowptr.data_ptr(), csr.perm.data_ptr(), gpeb.data_ptr(), glog.data_ptr(), gval.data_ptr(), _lib.stream())),
        8.0 * n * c + 8.0 * n * k * c + 8.0 * n * k * g + 8.0 * n * k + 4.0 * (n + 1) + 4.0 * n * c)
    if pointops.pe_mlp_supported(c):
        import torch.nn as nn
        from ao_b200 import ptv2 as _ptv2
        mlp = nn.Sequential(nn.Linear(3, c), _ptv2.PointBatchNorm(c), nn.ReLU(inplace=True), nn.Linear(c, c)).to(dev).train()
        pos = pointops.group_xyz(idx, coord)
        row("pos_moments", timeit(lambda: pointops.pos_moments(pos)), 12.0 * n * k)
        mom = pointops.pos_moments(pos)
        row("pe_mlp forward", timeit(lambda: pointops.pe_bias_mlp(pos, mlp, mom)), 12.0 * n * k + 4.0 * n * k * c)
        y = pointops.pe_bias_mlp(pos, mlp, mom)
        gy = torch.randn_like(y)
        row("pe_mlp backward", timeit(lambda: torch.autograd.grad(y, list(mlp.parameters()), gy, retain_graph=True, allow_unused=True)),
            12.0 * n * k + 4.0 * n * k * c)
        # device time of the C entry points alone (CUDA events around each call; best of 5)
        best = {}
        for _ in range(5):
            tr = _lib.trace_start()
            y2 = pointops.pe_bias_mlp(pos, mlp, mom)
            torch.autograd.grad(y2, list(mlp.parameters()), gy, allow_unused=True)
            torch.cuda.synchronize()
            _lib.trace_stop()
            for name, _a, s0, e0 in tr:
                best[name] = min(best.get(name, 1e9), s0.elapsed_time(e0) * 1e3)
        for name, us in best.items():
            row("  " + name, us, 12.0 * n * k + 4.0 * n * k * c)
        del y, y2, gy, pos
    # ---- dense.cu: BatchNorm + ReLU on (n, c) and the weight-encoding tail on (n*k, g); device time of the C entry points
    # (stats pass + partial reduce + apply pass each).  Algorithmic bytes count every tensor ONCE: the second pass over x
    # (and grad / out in the backward) is expected from the 126 MB L2 when the tensor fits.
    import torch.nn as nn
    for xdt, nm, esz in ((torch.bfloat16, "bf16", 2.0), (torch.float32, "f32", 4.0)):
        bn = nn.BatchNorm1d(c).to(dev).train()
        xin = torch.randn(n, c, device=dev).to(xdt).requires_grad_(True)
        best = {}
        for _ in range(5):
            tr = _lib.trace_start()
            yb = pointops.bn_act(xin, bn, relu=True)
            yb.backward(torch.ones_like(yb))
            torch.cuda.synchronize()
            _lib.trace_stop()
            for name, _a, s0, e0 in tr:
                best[name] = min(best.get(name, 1e9), s0.elapsed_time(e0) * 1e3)
        row(f"bn_act forward  ({nm}, relu)", best["aopt_bn_act_forward"], 2.0 * esz * n * c)
        row(f"bn_act backward ({nm}, relu)", best["aopt_bn_act_backward"], 4.0 * esz * n * c)
        del xin, yb
    if pointops.we_tail_supported(g):
        bn = nn.BatchNorm1d(g).to(dev).train()
        lin = nn.Linear(g, g).to(dev)
        relg = torch.randn(n, k, g, device=dev, requires_grad=True)
        upeg = torch.randn(n, k, g, device=dev, requires_grad=True)
        cstg = torch.randn(g, device=dev)
        best = {}
        for _ in range(5):
            tr = _lib.trace_start()
            yl = pointops.we_tail(relg, upeg, cstg, bn, lin)
            yl.backward(torch.ones_like(yl))
            torch.cuda.synchronize()
            _lib.trace_stop()
            for name, _a, s0, e0 in tr:
                best[name] = min(best.get(name, 1e9), s0.elapsed_time(e0) * 1e3)
        row("we_tail forward  (rel+upe -> logits)", best["aopt_we_tail_forward"], 12.0 * n * k * g)
        row("we_tail backward", best["aopt_we_tail_backward"], 16.0 * n * k * g)
        del relg, upeg, yl
    if li < 3:
        c2 = C[li + 1]
        pin = torch.relu(torch.randn(n, c2, device=dev)).requires_grad_(True)
        (nc, nf, noff), cluster, part = pointops.grid_pool(coord, pin, offset, GRID[li], return_partition=True)
        nv = nc.shape[0]
        row("voxel_partition (keys+sort+part)", timeit(lambda: pointops.voxel_partition(coord, offset, GRID[li])), 20.0 * n)
        of, am, oc = torch.empty(nv, c2, device=dev), torch.empty(nv, c2, dtype=torch.int32, device=dev), torch.empty(nv, 3, device=dev)
        row("pool_forward", timeit(lambda: lib.aopt_pool_forward(nv, c2, pin.data_ptr(), coord.data_ptr(), part.order.data_ptr(), part.idx_ptr.data_ptr(), of.data_ptr(), am.data_ptr(), oc.data_ptr(), _lib.stream())),
            4.0 * n * c2 + 16.0 * n + 4.0 * (nv + 1) + 8.0 * nv * c2 + 12.0 * nv)
        gp, gf = torch.randn(nv, c2, device=dev), torch.empty(n, c2, device=dev)
        row("pool_backward", timeit(lambda: lib.aopt_pool_backward(n, c2, gp.data_ptr(), am.data_ptr(), part.cluster32.data_ptr(), gf.data_ptr(), _lib.stream())),
            8.0 * nv * c2 + 4.0 * n + 4.0 * n * c2)
        src = torch.randn(nv, c, device=dev, requires_grad=True)
        i3, d3 = pointops.knn_query_raw(3, nc, noff.int(), coord, offset)
        w3 = pointops.interpolation_weights(d3)
        row("knn_query (cross k=3)", timeit(lambda: pointops.knn_query_raw(3, nc, noff.int(), coord, offset)), 12.0 * nv + 12.0 * n + 24.0 * n)
        up = torch.empty(n, c, device=dev)
        row("interpolation_forward", timeit(lambda: lib.aopt_interpolation_forward(n, c, 3, nv, src.data_ptr(), i3.data_ptr(), w3.data_ptr(), up.data_ptr(), _lib.stream())),
            24.0 * n + 4.0 * nv * c + 4.0 * n * c)
        csr3 = pointops.get_csr(i3, nv, 1)
        gs = torch.empty(nv, c, device=dev)
        row("interpolation_backward", timeit(lambda: lib.aopt_interpolation_backward(nv, c, 3, g_out.data_ptr(), w3.data_ptr(), csr3.rowptr.data_ptr(), csr3.perm.data_ptr(), gs.data_ptr(), _lib.stream())),
            4.0 * n * c + 24.0 * n + 4.0 * (nv + 1) + 4.0 * nv * c)
