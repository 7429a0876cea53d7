"""Host-side check of the kNN register top-k insert (ao_b200/csrc/knn_common.cuh): the struct is compiled for the
host by nvcc and compared with a stable sort by (d2, idx) on random streams full of exact ties, sentinel-sized and
NaN distances, and fewer than k candidates.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_topk_insert_matches_sorted_reference(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "topk_host_test")
    src = os.path.join(ROOT, "tests", "host", "topk_host_test.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout
