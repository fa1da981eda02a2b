"""percept_b200 — B200-native (sm_100a) implementation of Percept's reference-mesh CG smoother hot path.

The compute path is the CUDA library percept_b200/lib/libpms.so behind the C ABI of
include/percept_smoother_c.h.  There is no CPU fallback: importing works without a GPU (so the
ABI can be inspected), but creating a context without a CUDA device raises SmootherError.
"""
import ctypes as _C
import os as _os

from .capi import AOS, SOA, TET4, HEX8, Api, Ctx, SmootherError, State  # noqa: F401

_HERE = _os.path.dirname(_os.path.abspath(__file__))
# PMS_LIB: an A/B build of the same library (tools/build_variant.sh); default the in-tree build
LIB_PATH = _os.environ.get("PMS_LIB") or _os.path.join(_HERE, "lib", "libpms.so")


def _load():
    if not _os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C percept_b200/csrc). "
            "percept_b200 has no CPU fallback.")
    return _C.CDLL(LIB_PATH)


# `lib` (the ctypes handle of libpms.so) and `api` (its bound entry points) are created on first use, so that
# importing a submodule that needs no device code (percept_b200.capi, .partition, .exodus) does not map the CUDA
# library into the process — bench.py's CPU reference arm relies on that.
def __getattr__(name):
    if name in ("lib", "api"):
        g = globals()
        if "lib" not in g:
            g["lib"] = _load()
            g["api"] = Api(g["lib"], "pms_", product=True)
        return g[name]
    raise AttributeError(f"module 'percept_b200' has no attribute {name!r}")


def create(device=0):
    """New smoother context on CUDA device `device`."""
    return __getattr__("api").create(device)
