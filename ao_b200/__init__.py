"""ao_b200 — B200-native (sm_100a) point operators for the PointTransformer V2 hot path.

Layout:
  ao_b200/csrc/      hand-written CUDA kernels + the C ABI (include/ao_pointops.h)
  ao_b200/lib/       built libao_pointops.so (git-ignored; `make -C ao_b200/csrc`)
  ao_b200/pointops/  drop-in Python package with the reference `libs/pointops` API
  ao_b200/ptv2.py    PTv2m2 backbone (reference state_dict keys) routed through these ops
  ao_b200/scenes.py  synthetic S3DIS / ScanNet / SemanticKITTI-shaped inputs

`ao_b200.install_as_pointops()` registers the package under the name `pointops`, so reference
call sites (`import pointops; pointops.knn_query(...)`) run unmodified.
"""
import sys

__version__ = "0.1.0"


def install_as_pointops():
    """Makes `import pointops` resolve to ao_b200.pointops (drop-in for libs/pointops)."""
    from . import pointops as _p
    from .pointops import _C

    sys.modules["pointops"] = _p
    sys.modules["pointops._C"] = _C
    return _p


def install_native_only():
    """Registers only the native module `pointops._C`: the reference's own Python package
    (libs/pointops/functions/*.py, imported as `pointops`) keeps running unmodified and its
    `from pointops._C import knn_query_cuda, …` lines resolve to the B200 kernels."""
    from .pointops import _C

    sys.modules["pointops._C"] = _C
    return _C
