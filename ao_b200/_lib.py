"""ctypes binding of libao_pointops.so (C ABI declared in include/ao_pointops.h).

This is the ONLY compute path of the package: there is no CPU or torch fallback.  If the shared
library is missing or does not export a symbol, importing/using the ops fails loudly.

The reference binds its kernels through a pybind11 torch extension `pointops._C`
(/root/reference/libs/pointops/src/pointops_api.cpp:15-32) whose functions take at::Tensor and
forward `data_ptr()` to `extern "C"` launchers; here the same raw-pointer launch boundary is the
public C ABI and Python forwards `tensor.data_ptr()` + the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_ulonglong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AO_POINTOPS_LIB", os.path.join(_HERE, "lib", "libao_pointops.so"))

P = c_void_p  # device pointer / stream

# name -> (restype, argtypes); must list every function of include/ao_pointops.h
SIGNATURES = {
    "aopt_version": (c_char_p, []),
    "aopt_status_string": (c_char_p, [c_int]),
    "aopt_last_cuda_error": (c_char_p, []),
    "aopt_kernel_launches": (c_ulonglong, []),
    "aopt_set_tuning": (c_int, [c_char_p, c_int]),
    "aopt_offset2batch": (c_int, [c_int, c_int, P, P, P]),
    "aopt_knn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "aopt_knn_query": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_int, P, c_size_t, P]),
    "aopt_knn_query_multi": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, P, c_size_t, P]),
    "aopt_farthest_point_sampling": (c_int, [c_int, c_int, P, P, P, P, P, P]),
    "aopt_grid_sample_keys": (c_int, [c_int, P, c_double, c_double, c_double, c_int, c_int, P, P, P, P]),
    "aopt_voxel_pick": (c_int, [c_int, P, P, P, c_longlong, P, P]),
    "aopt_sphere_dist2": (c_int, [c_int, P, c_float, c_float, c_float, P, P]),
    "aopt_select_rows": (c_int, [c_longlong, c_int, P, P, P, P]),
    "aopt_csr_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "aopt_csr_build": (c_int, [c_int, c_int64, P, c_int, P, P, P, c_size_t, P]),
    "aopt_grouping_forward": (c_int, [c_int, c_int, c_int, P, P, P, c_int, P]),
    "aopt_grouping_backward": (c_int, [c_int, c_int, P, c_int, P, P, c_float, P, P]),
    "aopt_relation_backward": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "aopt_group_xyz": (c_int, [c_int, c_int, P, P, P, P, c_int, P]),
    "aopt_gather_sub_forward": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
    "aopt_sum_over_k": (c_int, [c_int, c_int, c_int, P, c_float, P, P]),
    "aopt_gva_forward": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "aopt_gva_backward_query": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P]),
    "aopt_gva_backward_value": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "aopt_gva_backward": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P]),
    "aopt_segment_min3": (c_int, [c_int, c_int, P, P, P, P]),
    "aopt_voxel_keys": (c_int, [c_int, c_int, P, P, P, c_float, P, P, P]),
    "aopt_voxel_partition_workspace_bytes": (c_size_t, [c_int]),
    "aopt_voxel_partition": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "aopt_voxel_grid_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aopt_voxel_grid": (c_int, [c_int, c_int, P, P, P, c_float, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "aopt_pool_forward": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P]),
    "aopt_pool_backward": (c_int, [c_int, c_int, P, P, P, P, P]),
    "aopt_interp_weights": (c_int, [c_int, c_int, P, P, P]),
    "aopt_interpolation_forward": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "aopt_interpolation_backward": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "aopt_vote_accumulate": (c_int, [c_int, c_int, c_longlong, P, P, P, P, P]),
    "aopt_pe_mlp_supported": (c_int, [c_int]),
    "aopt_pos_moments_workspace_bytes": (c_size_t, []),
    "aopt_pos_moments": (c_int, [c_int64, P, P, P, c_size_t, P]),
    "aopt_pe_mlp_state_bytes": (c_size_t, [c_int]),
    "aopt_pe_mlp_forward": (c_int, [c_int64, c_int, P, P, P, P, P, P, P, P, c_float, c_int, P, P, P, P, c_int, P, P,
                                    c_size_t, P]),
    "aopt_pe_mlp_stats": (c_int, [c_int, P, P, P]),
    "aopt_pe_mlp_backward_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "aopt_pe_mlp_backward": (c_int, [c_int64, c_int, P, P, P, P, c_int, P, P, P, P, P, P, P, P, c_int, P, P, P,
                                     c_size_t, P]),
    "aopt_dense_workspace_bytes": (c_size_t, [c_int]),
    "aopt_bn_act_supported": (c_int, [c_int]),
    "aopt_bn_act_forward": (c_int, [c_int64, c_int, P, c_int64, c_int, P, P, c_float, P, P, c_int, P, c_int, P, P, P, c_float, P, P, P,
                                    c_size_t, P]),
    "aopt_bn_act_backward": (c_int, [c_int64, c_int, P, P, c_int, P, c_int64, c_int, P, P, P, P, c_int64, P, P, P, P, c_size_t, P]),
    "aopt_we_tail_supported": (c_int, [c_int]),
    "aopt_we_tail_forward": (c_int, [c_int64, c_int, P, P, P, P, c_int, P, P, P, P, c_float, P, P, P, P, P, P, c_float, P, P,
                                     c_size_t, P]),
    "aopt_we_tail_backward": (c_int, [c_int64, c_int, P, P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "aopt_col_sum": (c_int, [c_int64, c_int, P, c_int64, c_int, P, P, c_size_t, P]),
    "aopt_copy_cols": (c_int, [c_int64, c_int, P, c_int64, c_int, P, P, c_int64, c_int, P]),
    "aopt_skinny_wgrad_supported": (c_int, [c_int, c_int]),
    "aopt_skinny_linear": (c_int, [c_int64, c_int, c_int, P, c_int64, c_int, P, P, P, P]),
    "aopt_skinny_dgrad": (c_int, [c_int64, c_int, c_int, P, P, P, c_int64, c_int, P]),
    "aopt_skinny_wgrad": (c_int, [c_int64, c_int, c_int, P, c_int, P, c_int64, c_int, P, P, c_size_t, P]),
    "aopt_aggregation_forward": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "aopt_aggregation_backward": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P]),
    "aopt_subtraction_forward": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
}

KNN_AUTO, KNN_TILE, KNN_GRID = 0, 1, 2
KNN_SQRT_DIST = 0x100   # OR-ed into the method: the kernel writes sqrt(dist2)

_lib = None
_trace = None  # list of (entry point, args, start event, end event) while bench.py's profiler is on
_UNTRACED = {"aopt_set_tuning", "aopt_version", "aopt_status_string", "aopt_last_cuda_error", "aopt_kernel_launches",
             "aopt_knn_workspace_bytes", "aopt_csr_workspace_bytes", "aopt_voxel_partition_workspace_bytes", "aopt_voxel_grid_workspace_bytes",
             "aopt_pe_mlp_supported", "aopt_pos_moments_workspace_bytes", "aopt_pe_mlp_state_bytes",
             "aopt_pe_mlp_backward_workspace_bytes", "aopt_dense_workspace_bytes", "aopt_bn_act_supported",
             "aopt_we_tail_supported", "aopt_skinny_wgrad_supported"}


class _Entry:
    """One C entry point; when tracing, brackets the call with CUDA events on the launching stream."""

    __slots__ = ("name", "fn", "traced")

    def __init__(self, name, fn):
        self.name, self.fn, self.traced = name, fn, name not in _UNTRACED

    def __call__(self, *args):
        if _trace is None or not self.traced:
            return self.fn(*args)
        s = _event_pool.pop() if _event_pool else torch.cuda.Event(enable_timing=True)
        e = _event_pool.pop() if _event_pool else torch.cuda.Event(enable_timing=True)
        s.record()
        status = self.fn(*args)
        e.record()
        _trace.append((self.name, args, s, e))
        return status


class _Library:
    """Attribute access to the declared entry points of libao_pointops.so."""

    def __init__(self, cdll):
        self._cdll = cdll


_event_pool = []


def trace_prepare(n_events: int = 1024) -> None:
    """Creates the CUDA events of a traced step ahead of time (outside the timed region): creating ~270 events
    inside the step costs more host time than recording them."""
    while len(_event_pool) < n_events:
        _event_pool.append(torch.cuda.Event(enable_timing=True))


def trace_start() -> list:
    """Starts recording every kernel-launching C-ABI call (name, args, CUDA events); returns the list."""
    global _trace
    _trace = []
    return _trace


def trace_stop() -> list:
    global _trace
    t, _trace = _trace, None
    return t or []


def kernel_launches() -> int:
    """Kernels enqueued by the library so far in this process."""
    return int(load().aopt_kernel_launches())


def load() -> "_Library":
    """Loads the shared library (once) and declares every signature.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"ao_b200: {LIB_PATH} not found. Build it with `make -C ao_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path."
        )
    cdll = ctypes.CDLL(LIB_PATH)
    lib = _Library(cdll)
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(cdll, name)
        except AttributeError as e:  # pragma: no cover - build/ABI mismatch
            raise ImportError(f"ao_b200: {LIB_PATH} does not export {name}") from e
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(lib, name, _Entry(name, fn))
    _lib = lib
    return lib


def set_tuning(name: str, value: int) -> None:
    """A/B switch of the library (include/ao_pointops.h aopt_set_tuning); results never depend on it."""
    check(load().aopt_set_tuning(name.encode(), int(value)), "set_tuning")


def version() -> str:
    return load().aopt_version().decode()


def check(status: int, what: str) -> None:
    if status != 0:
        lib = load()
        msg = lib.aopt_status_string(status).decode()
        if status == 3:
            msg += f" ({lib.aopt_last_cuda_error().decode()})"
        raise RuntimeError(f"ao_b200.{what} failed: {msg}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """The caller's current CUDA stream (the reference launches on the legacy default stream).
    Raw handle of torch's current stream on the current device: torch.cuda.current_stream() builds a Python
    Stream object per call (~3 us x ~120 calls per step on a host-bound path)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _NullCtx:
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def on_device(dev):
    """`with on_device(dev):` — torch.cuda.device(dev) only when dev is not already current (the context
    manager costs ~10 us of host time per call; the step issues ~250 calls and is host-bound below ~8 ms)."""
    idx = dev.index if isinstance(dev, torch.device) else int(dev)
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(dev)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors) -> torch.device:
    """All tensors must live on one CUDA device — the product path has no CPU implementation."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("ao_b200.pointops: expected CUDA tensors (there is no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("ao_b200.pointops: tensors are on different devices")
    if dev is None:
        raise ValueError("ao_b200.pointops: no tensor arguments")
    return dev


# ---- side streams --------------------------------------------------------------------------------------------
# Experiment kept behind switches (default: everything on the caller's stream).  Three kinds of work that leave the
# HBM pipe idle were issued on a side stream next to the HBM-bound streaming kernels; device ms per step of the
# configs[1] schedule on one B200 (scripts/exp_overlap.py, same box, single stream first):
#   walk  GVA backward-value CSR walk beside gva_backward_query            8.52 -> 8.48  (nothing)
#   geom  CSR build prefetched under the forward                           8.52 -> 8.95  (a loss)
#   knn   neighbour searches of levels 1..3 + interpolation beside the
#         level-0 blocks (pointops.prepare_pyramid(..., knn=k))            8.30 -> 8.49  (a loss)
# The streaming kernels run 12 warps per SM, each with ~32 copies in flight; any co-resident kernel that takes
# issue slots (kNN), L1 (the walk: the streaming kernel's 3 x 64 KB of shared memory leaves it almost none) or L2
# atomics (CSR build) slows them by more than it gains.  AOPT_OVERLAP=1 or AOPT_OVERLAP_ROLES=knn,walk,geom turn
# roles on; results are bit-identical either way (tests/test_streams_gpu.py).
_side = {}
# mode: None = per-role defaults (_roles_on), True = every role on, False = everything on the caller's stream
_mode = {"1": True, "0": False}.get(os.environ.get("AOPT_OVERLAP", ""), None)
_roles_on = {r for r in os.environ.get("AOPT_OVERLAP_ROLES", "").split(",") if r}   # roles: knn, walk, geom


def overlap(on=None, role: str = None) -> bool:
    """overlap(role=r): is role r issued on its side stream?  overlap(True / False): force every role on / off
    (bench.py forces off on the steps that carry per-kernel CUDA events, tests A/B both)."""
    global _mode
    if on is not None:
        _mode = bool(on)
    if _mode is not None:
        return _mode
    return (role in _roles_on) if role is not None else bool(_roles_on)


def overlap_mode(*set_to):
    """Get (no argument) or restore (one argument: None / True / False) the forced mode."""
    global _mode
    if set_to:
        _mode = set_to[0]
    return _mode


def side_stream(device, role: str = "aux") -> "torch.cuda.Stream":
    key = (torch.device(device).index, role)
    st = _side.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side[key] = st
    return st


def workspace(nbytes: int, device) -> torch.Tensor:
    """Scratch from torch's caching allocator (the library itself never allocates)."""
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
