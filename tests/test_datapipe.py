"""Per-sample data transforms.

CPU part: oracle/datapipe_ref.py against tests/golden/datapipe_ref.npz, which tests/golden/make_datapipe_golden.py
wrote by running the reference's own GridSample / SphereCrop / CenterShift / NormalizeColor classes.
GPU part: ao_b200.datapipe (csrc/datapipe.cu) against the oracle, bit for bit (both use stable sorts and the same
host random draws)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


def _gen():
    spec = importlib.util.spec_from_file_location("make_datapipe_golden", os.path.join(GOLD, "make_datapipe_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "datapipe_ref.npz"))


@pytest.fixture(scope="module")
def dref():
    from oracle import datapipe_ref

    return datapipe_ref


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_oracle_grid_sample_pinned_by_reference_classes(gold, dref, case):
    gen = _gen()
    cs = gen.CASES[case]
    data = gen.cloud(cs["seed"])
    assert str(gold[f"{case}_scaled_dtype"]) == "float64"            # the golden run used NumPy >= 2 semantics
    vh = dref.voxel_hash(data["coord"], cs["grid"], cs["hash_type"], division="float64")
    assert np.array_equal(vh["key"], gold[f"{case}_key"])            # per-point keys, uint64, bit for bit
    assert np.array_equal(vh["uniq"], gold[f"{case}_uniq"]) and np.array_equal(vh["count"], gold[f"{case}_count"])
    assert np.array_equal(vh["min_coord"].reshape(1, 3), gold[f"{case}_min_coord"])
    # train mode with the same seeded host generator: one point per voxel, voxels in ascending key order; which
    # point of a voxel is taken depends on numpy's unstable argsort, so membership is what is pinned
    np.random.seed(100 + cs["seed"])
    r = np.random.randint(0, vh["count"].max(), vh["count"].size)
    idx, _ = dref.grid_sample_train(data["coord"], cs["grid"], r, cs["hash_type"])
    ref_disc = gold[f"{case}_train_discrete"]
    assert idx.shape[0] == ref_disc.shape[0] == vh["uniq"].shape[0]
    assert np.array_equal(vh["discrete"][idx], ref_disc)             # same voxel at every output position
    same = (data["coord"][idx] == gold[f"{case}_train_coord"]).all(1)
    single = vh["count"] == 1
    assert same[single].all()                                        # voxels holding one point: identical output
    assert same.mean() > 0.3
    # test mode: same number of parts, every point covered, part 0 holds one point of every voxel
    parts, _ = dref.grid_sample_test(data["coord"], cs["grid"], cs["hash_type"])
    assert len(parts) == int(gold[f"{case}_test_nparts"])
    assert np.array_equal(np.unique(np.concatenate(parts)), gold[f"{case}_test_union"])
    assert np.array_equal(vh["key"][parts[0]], vh["key"][gold[f"{case}_test_index0"]])


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_oracle_sphere_crop_pinned_by_reference_class(gold, dref, case):
    gen = _gen()
    cs = gen.CASES[case]
    data = gen.cloud(cs["seed"])
    np.random.seed(200 + cs["seed"])
    center = data["coord"][np.random.randint(data["coord"].shape[0])]
    idx, d2 = dref.sphere_crop_index(data["coord"], center, 2500)
    ref = gold[f"{case}_crop_coord"]
    strict = np.concatenate([[True], np.diff(d2[idx]) > 0]) & np.concatenate([np.diff(d2[idx]) > 0, [d2[idx][-1] < np.sort(d2)[2500]]])
    assert np.array_equal(data["coord"][idx][strict], ref[strict])   # tie-free positions are fully determined
    assert np.array_equal(np.sort(d2[idx]), np.sort(np.sum(np.square(ref - center), 1)))
    if case == "a":
        assert np.array_equal(dref.center_shift(data["coord"], True), gold["a_shift_z"])
        assert np.array_equal(dref.center_shift(data["coord"], False), gold["a_shift_noz"])
        assert np.array_equal(dref.normalize_color(data["color"]), gold["a_color"])


def test_oracle_float32_division_differs_only_at_cell_boundaries(dref):
    """NumPy 1.x evaluated coord / np.array(grid) in fp32: the two semantics disagree on a handful of boundary points."""
    data = _gen().cloud(5)
    a = dref.voxel_hash(data["coord"], 0.04, division="float64")["discrete"] + 0
    b = dref.voxel_hash(data["coord"], 0.04, division="float32")["discrete"] + 0
    assert (np.abs(a - b) <= 1).all() and (a != b).any(1).mean() < 0.01


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("division", ["float64", "float32"])
@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_gpu_grid_sample_equals_oracle(dref, case, division):
    from ao_b200 import datapipe

    gen = _gen()
    cs = gen.CASES[case]
    data = gen.cloud(cs["seed"])
    vh = dref.voxel_hash(data["coord"], cs["grid"], cs["hash_type"], division)
    g = datapipe.voxel_hash(torch.from_numpy(data["coord"]).cuda(), cs["grid"], cs["hash_type"], division)
    assert g.n_vox == vh["uniq"].shape[0] and g.count_max == int(vh["count"].max())
    assert np.array_equal(g.cell.cpu().numpy(), vh["discrete"])
    assert np.array_equal(g.order.cpu().numpy(), vh["idx_sort"])
    assert np.array_equal(g.count.cpu().numpy(), vh["count"])
    assert np.array_equal(g.min_coord.cpu().numpy(), vh["min_coord"])
    # train mode through the class, same seeded host generator as the oracle / the reference
    np.random.seed(100 + cs["seed"])
    r = np.random.randint(0, vh["count"].max(), vh["count"].size)
    ref_idx, _ = dref.grid_sample_train(data["coord"], cs["grid"], r, cs["hash_type"], division)
    np.random.seed(100 + cs["seed"])
    out = datapipe.GridSample(grid_size=cs["grid"], hash_type=cs["hash_type"], mode="train", keys=("coord", "color", "segment"),
                              return_discrete_coord=True, return_min_coord=True, division=division)(dict(data))
    assert out["coord"].is_cuda and np.array_equal(out["coord"].cpu().numpy(), data["coord"][ref_idx])
    assert np.array_equal(out["color"].cpu().numpy(), data["color"][ref_idx])
    assert out["segment"].dtype == torch.int64 and np.array_equal(out["segment"].cpu().numpy(), data["segment"][ref_idx])
    assert np.array_equal(out["discrete_coord"].cpu().numpy(), vh["discrete"][ref_idx])
    # test mode
    parts = datapipe.GridSample(grid_size=cs["grid"], hash_type=cs["hash_type"], mode="test", keys=("coord", "segment"),
                                division=division)(dict(data))
    ref_parts, _ = dref.grid_sample_test(data["coord"], cs["grid"], cs["hash_type"], division)
    assert len(parts) == len(ref_parts)
    for p, rp in zip(parts, ref_parts):
        assert np.array_equal(p["index"].cpu().numpy(), rp) and np.array_equal(p["coord"].cpu().numpy(), data["coord"][rp])


@pytest.mark.gpu
def test_gpu_sphere_crop_and_pipeline_equal_oracle(dref, gold):
    from ao_b200 import datapipe

    gen = _gen()
    data = gen.cloud(11)
    np.random.seed(211)
    center = data["coord"][np.random.randint(data["coord"].shape[0])]
    ref_idx, ref_d2 = dref.sphere_crop_index(data["coord"], center, 2500)
    idx = datapipe.sphere_crop_index(torch.from_numpy(data["coord"]).cuda(), center, 2500)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    np.random.seed(211)
    out = datapipe.SphereCrop(point_max=2500, mode="random")(dict(data))
    assert np.array_equal(out["coord"].cpu().numpy(), data["coord"][ref_idx])
    assert np.array_equal(out["segment"].cpu().numpy(), data["segment"][ref_idx])
    # the S3DIS training pipeline of configs/s3dis/semseg-pt-v2m2-0-base.py:72-109 (the transforms built here)
    cfg = [dict(type="CenterShift", apply_z=True), dict(type="GridSample", grid_size=0.04, hash_type="fnv", mode="train",
                                                        keys=("coord", "color", "segment"), return_discrete_coord=True),
           dict(type="SphereCrop", point_max=1500, mode="random"), dict(type="CenterShift", apply_z=False),
           dict(type="NormalizeColor"), dict(type="ToTensor"),
           dict(type="Collect", keys=("coord", "segment"), feat_keys=["coord", "color"])]
    np.random.seed(7)
    sample = datapipe.Compose(cfg)(dict(data))
    # same pipeline on the host with the oracle
    np.random.seed(7)
    c0 = dref.center_shift(data["coord"], True)
    vh = dref.voxel_hash(c0, 0.04)
    r = np.random.randint(0, vh["count"].max(), vh["count"].size)
    i1, _ = dref.grid_sample_train(c0, 0.04, r)
    c1, col1, seg1 = c0[i1], data["color"][i1], data["segment"][i1]
    center = c1[np.random.randint(c1.shape[0])]
    i2, _ = dref.sphere_crop_index(c1, center, 1500)
    c2 = dref.center_shift(c1[i2], False)
    assert np.array_equal(sample["coord"].cpu().numpy(), c2)
    assert np.array_equal(sample["segment"].cpu().numpy(), seg1[i2])
    assert np.array_equal(sample["feat"].cpu().numpy(), np.concatenate([c2, dref.normalize_color(col1[i2])], 1))
    assert sample["offset"].tolist() == [1500]
    batch = datapipe.collate_fn([sample, sample])
    assert batch["offset"].tolist() == [1500, 3000] and batch["coord"].shape[0] == 3000


@pytest.mark.gpu
def test_gpu_grid_sample_full_size_raw_room():
    """A raw-room-sized cloud (600k points): properties that do not need the oracle — one pick per voxel, picks inside
    their voxel, voxels in ascending key order, test-mode parts cover every point exactly count times."""
    from ao_b200 import datapipe

    rng = np.random.default_rng(0)
    coord = (rng.random((600000, 3)) * np.array([10.0, 8.0, 3.0]) - 2.0).astype(np.float32)
    x = torch.from_numpy(coord).cuda()
    vh = datapipe.voxel_hash(x, 0.04)
    cnt = vh.count.long()
    assert int(cnt.sum()) == coord.shape[0] and int(cnt.min()) >= 1 and vh.count_max == int(cnt.max())
    np.random.seed(3)
    pick = datapipe.voxel_pick(vh, np.random.randint(0, vh.count_max, vh.n_vox))
    cells = vh.cell.long()
    ckey = (cells[:, 0] * 4096 + cells[:, 1]) * 4096 + cells[:, 2]
    assert torch.unique(ckey).numel() == vh.n_vox == torch.unique(ckey[pick]).numel()
    seg = torch.repeat_interleave(torch.arange(vh.n_vox, device="cuda"), cnt)
    assert torch.equal(ckey[vh.order.long()], ckey[pick][seg])         # every sorted point lies in its voxel's pick cell
    ref_cell = torch.floor(x.double() / 0.04).long()
    assert torch.equal(cells, ref_cell - ref_cell.min(0).values)
