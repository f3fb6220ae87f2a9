"""gigapaxos_b200 -- B200-native batched-Paxos engine for gigapaxos's phase-2 hot path.

The product is the CUDA library ``libgpx.so`` (hand-written sm_100a kernels behind the C ABI
of ``include/gpx.h``); this package is the host-side mirror of the reference's interface for
that path.  There is no CPU fallback: loading fails loudly when the library is missing and
engine creation fails with GPX_ENOGPU when no CUDA device is present.
"""
from __future__ import annotations

import os

from . import abi
from .abi import Config, Engine, GpxError, Library

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpx.so")
_lib = None


def load_library() -> Library:
    """Load the CUDA engine library (built in-tree by ``gigapaxos_b200.build``)."""
    global _lib, LIB_PATH
    if _lib is None:
        LIB_PATH = os.environ.get("GPX_LIB", LIB_PATH)  # alternative builds of the same CUDA library (tuning)
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m gigapaxos_b200.build` (nvcc, sm_100a). "
                "The engine has no CPU fallback.")
        _lib = Library(LIB_PATH, "gpx_")
    return _lib


def default_config(**overrides) -> Config:
    cfg = load_library().config_defaults()
    for k, v in overrides.items():
        if k == "lane_node":
            for i, x in enumerate(v):
                cfg.lane_node[i] = int(x)
        else:
            setattr(cfg, k, v)
    return cfg


def create_engine(cfg: Config | None = None, **overrides) -> Engine:
    lib = load_library()
    if cfg is None:
        cfg = default_config(**overrides)
    return Engine(lib, cfg)


__all__ = ["abi", "Config", "Engine", "GpxError", "Library", "load_library", "default_config", "create_engine",
           "LIB_PATH"]
