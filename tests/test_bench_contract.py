"""bench.py output contract, checked on CPU: the reference arm prints exactly ONE JSON line on stdout with the keys the
driver reads; under torchrun only rank 0 prints; the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ, AOPT_BENCH_CPU_BUDGET="2", OMP_NUM_THREADS="4")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=env, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ptv2_pointops_fwd_bwd_throughput" and d["unit"] == "Mpoints/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = _run(["--steps", "1", "--warmup", "0"])
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr


@pytest.mark.parametrize("cfg,points", [("s3dis4", 320000), ("s3dis8", 640000), ("scannet150k", 450000), ("kitti120k", None)])
def test_committed_bench_lines_keep_the_contract_and_add_up(cfg, points):
    """The end-of-round lines under profiles/ (written on a B200 by scripts/gpu_r04f.sh) carry every key of the contract and
    their numbers are consistent with each other: value = points / time, roofline.frac = achieved / peak with the
    algorithmic bytes of the dominant launch, e2e measured with real host copies, launches counted."""
    path = os.path.join(ROOT, "profiles", f"r04f_bench_{cfg}.json")
    d = json.loads([l for l in open(path).read().splitlines() if l.strip()][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "ptv2_pointops_fwd_bwd_throughput" and d["unit"] == "Mpoints/s" and d["dtype"] == "f32"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    n0 = d["level_sizes"][0]
    if points is not None:
        assert n0 == points
    assert d["value"] == pytest.approx(n0 / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], abs=2e-4)
    assert r["achieved"] == pytest.approx(r["alg_bytes_per_launch"] / (r["us_per_launch"] * 1e-6) / 1e9, rel=2e-3)
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] >= 100 * d["steps"]          # ~170-190 library kernels per schedule step
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < 1.0
