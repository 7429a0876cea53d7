"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the header declares,
and the Python binding declares exactly the same set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ao_pointops.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aopt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_hot_path():
    fns = header_functions()
    for must in ("aopt_knn_query", "aopt_grouping_forward", "aopt_grouping_backward", "aopt_gva_forward",
                 "aopt_gva_backward_query", "aopt_gva_backward_value", "aopt_pool_forward", "aopt_pool_backward",
                 "aopt_interpolation_forward", "aopt_interpolation_backward", "aopt_csr_build"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from ao_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/ao_pointops.h but not exported"


def test_binding_matches_header():
    from ao_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert b"sm_100a" in lib.aopt_version()
    assert lib.aopt_status_string(0) == b"ok"
    assert lib.aopt_status_string(2) == b"workspace missing or too small"


def test_workspace_queries_are_pure_host_functions():
    from ao_b200 import _lib

    lib = _lib.load()
    assert lib.aopt_knn_workspace_bytes(320000, 320000, 4, 16, _lib.KNN_TILE) == 0
    assert lib.aopt_knn_workspace_bytes(320000, 320000, 4, 16, _lib.KNN_GRID) > 320000 * 16
    assert lib.aopt_knn_workspace_bytes(1000, 1000, 4, 16, _lib.KNN_AUTO) == 0      # small scenes → TILE
    assert lib.aopt_csr_workspace_bytes(1000, 16000) >= 4 * (1001 + 16000)


def test_argument_errors_are_reported_without_a_gpu():
    from ao_b200 import _lib

    lib = _lib.load()
    # nsample out of the reference's [1,128] range → invalid argument, before any launch
    assert lib.aopt_knn_query(10, 0, 10, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) == 1
    assert lib.aopt_knn_query(10, 129, 10, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) == 1
    assert lib.aopt_gva_forward(10, 16, 50, 6, 0, 0, 0, 0, 0, 0, 0) == 1            # c % g != 0
    assert lib.aopt_csr_build(10, 100, 0, 2, 0, 0, 0, 0, 0) == 1                     # bad negative_mode


def test_cpu_tensors_are_rejected():
    import pytest
    import torch

    from ao_b200 import pointops

    xyz = torch.zeros(8, 3)
    off = torch.tensor([8], dtype=torch.int32)
    with pytest.raises(ValueError):
        pointops.knn_query(3, xyz, off)
    with pytest.raises(NotImplementedError):
        pointops.ball_query(3, 1.0, 0.0, xyz, off)


def test_drop_in_names():
    """Every public name of the reference package exists (libs/pointops/functions/__init__.py:1-14)."""
    from ao_b200 import pointops

    for name in ("knn_query", "ball_query", "random_ball_query", "farthest_point_sampling", "grouping", "grouping2",
                 "interpolation", "interpolation2", "subtraction", "aggregation", "attention_relation_step",
                 "attention_fusion_step", "query_and_group", "knn_query_and_group", "ball_query_and_group",
                 "batch2offset", "offset2batch"):
        assert hasattr(pointops, name), name
    import ao_b200

    p = ao_b200.install_as_pointops()
    import pointops as q

    assert q is p
