"""profiles/traffic.json from an ncu summary table (scripts/ncu_summary.py output): DRAM read + write bytes per
level-0 launch of the top HBM kernels.  bench.py copies the dominant kernel's entry into `roofline.traffic`.
Usage: python scripts/make_traffic_json.py profiles/r01q_ops_L0_ncu_summary.md > profiles/traffic.json"""
import json
import sys

ALG = {  # level-0 algorithmic MB: N=320000, k=16, C=48, G=6; pool: C'=96, N'=50139
    "aopt_gather_sub_forward": ("gather_sub_ns_kernel", 1126.4),
    "aopt_gva_forward": ("gva_forward_ns_kernel", 1372.2),
    "aopt_relation_backward": ("relation_backward_vec_kernel", 1127.7),
    "aopt_gva_backward": ("gva_backward_fused_ns_kernel", 2438.4),
    "aopt_gva_backward_query": ("gva_backward_query_ns_kernel", 2355.2),
    "aopt_pe_mlp_forward": ("pe_mlp_forward_tc_kernel", 1052.2),
    "aopt_pe_mlp_backward": ("pe_mlp_backward_tc_kernel", 1052.2),
    "aopt_gva_backward_value": ("csr_walk_kernel<8, BvPolicy", 267.5),
    "aopt_pool_forward": ("pool_forward_kernel<4>", 167.3),
    "aopt_pool_backward": ("pool_backward_kernel<4>", 162.7),
    "aopt_interpolation_forward": ("interp_forward_kernel<4>", 78.7),
}
src = sys.argv[1]
rows = [l.split("|") for l in open(src) if l.startswith("|") and not l.startswith("|---") and not l.startswith("| #")]
out = {}
for entry, (pat, alg) in ALG.items():
    for r in rows:
        name = r[2].strip()
        if pat in name:
            rd, wr = float(r[6]), float(r[7])
            out[entry] = {
                "dram_bytes_per_launch": int(round((rd + wr) * 1e6)),
                "source": f"{src} row {name}: dram__bytes_read.sum {rd} MB + dram__bytes_write.sum {wr} MB (ncu --set full "
                          f"--clock-control none, level-0 launch: N=320000, k=16, C=48, G=6; algorithmic {alg} MB)",
            }
            break
print(json.dumps(out, indent=2))
