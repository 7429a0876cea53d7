"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by ao_b200/).

numpy restatement of the reference's per-sample transforms, each step citing
/root/reference/pointcept/datasets/transform.py:
  grid_sample_train / grid_sample_test   GridSample.__call__ :804-862, fnv_hash_vec :881-896, ravel_hash_vec :864-878
  sphere_crop_index                      SphereCrop.__call__ :968-975
  center_shift, normalize_color          :128-142, :100-104
with ONE stated difference: every argsort is kind="stable" (the reference's default quicksort leaves the order of
equal keys to the numpy build).  Pinned by tests/golden/datapipe_ref.npz, written by tests/golden/make_datapipe_golden.py
which RUNS the reference classes (imported from /root/reference with a stub registry) on seeded clouds: voxel keys,
unique keys, counts, min_coord and crop distances must be identical; the picked points must fall in the same voxels /
inside the same radius.
"""
import numpy as np


def fnv_hash_vec(arr):                                   # :881-896
    arr = arr.copy().astype(np.uint64, copy=False)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h


def ravel_hash_vec(arr):                                 # :864-878
    arr = arr.copy()
    arr -= arr.min(0)
    arr = arr.astype(np.uint64, copy=False)
    arr_max = arr.max(0).astype(np.uint64) + 1
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= arr_max[j + 1]
    keys += arr[:, -1]
    return keys


def voxel_hash(coord, grid_size, hash_type="fnv", division="float64"):
    g = np.array(grid_size)
    if division == "float32":                            # NumPy 1.x value-based casting of the 0-d divisor
        scaled = coord.astype(np.float32) / g.astype(np.float32)
    else:                                                # NumPy >= 2: fp32 array / 0-d fp64 array -> fp64
        scaled = coord.astype(np.float64) / g.astype(np.float64)
    discrete = np.floor(scaled).astype(int)              # :807
    min_cell = discrete.min(0)
    min_coord = min_cell * np.array(grid_size)           # :808
    discrete = discrete - min_cell                       # :809
    key = (fnv_hash_vec if hash_type == "fnv" else ravel_hash_vec)(discrete)   # :810
    idx_sort = np.argsort(key, kind="stable")            # :811 (stable: see header)
    key_sort = key[idx_sort]
    uniq, inverse, count = np.unique(key_sort, return_inverse=True, return_counts=True)   # :812
    return dict(discrete=discrete, min_coord=min_coord, key=key, idx_sort=idx_sort, uniq=uniq, count=count, scaled=scaled)


def grid_sample_train(coord, grid_size, r, hash_type="fnv", division="float64"):
    """r = the reference's np.random.randint(0, count.max(), count.size) draw (:815)."""
    vh = voxel_hash(coord, grid_size, hash_type, division)
    count = vh["count"]
    idx_select = np.cumsum(np.insert(count, 0, 0)[0:-1]) + r % count      # :814-816
    return vh["idx_sort"][idx_select], vh


def grid_sample_test(coord, grid_size, hash_type="fnv", division="float64"):
    vh = voxel_hash(coord, grid_size, hash_type, division)
    count = vh["count"]
    parts = []
    for i in range(count.max()):                                          # :838-843
        idx_select = np.cumsum(np.insert(count, 0, 0)[0:-1]) + i % count
        parts.append(vh["idx_sort"][idx_select])
    return parts, vh


def sphere_crop_index(coord, center, point_max):
    d2 = np.sum(np.square(coord - center), 1)                             # :973-975
    return np.argsort(d2, kind="stable")[:point_max], d2


def center_shift(coord, apply_z=True):                                    # :128-142
    x_min, y_min, z_min = coord.min(axis=0)
    x_max, y_max, _ = coord.max(axis=0)
    shift = [(x_min + x_max) / 2, (y_min + y_max) / 2, z_min if apply_z else 0]
    return coord - np.array(shift, dtype=coord.dtype)


def normalize_color(color):                                               # :100-104
    return color / 127.5 - 1
