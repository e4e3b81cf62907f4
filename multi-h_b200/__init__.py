"""multih_b200 — B200-native (sm_100a) hot path of Multi-H behind the reference's pipeline surface.

The product is `libmultih_b200.so` (hand-written CUDA + host C++ behind the C ABI of include/multih_b200.h).
This Python package is only the host-side binding used by tests and bench.py:

  multih_b200.capi     ctypes binding of every C-ABI entry point (raises if the library or a GPU is missing —
                       there is no CPU fallback)
  multih_b200.MultiH   mirror of the reference's `class MultiH` (MultiH/MultiH/MultiH.h:20-149)
  multih_b200.scenes   synthetic multi-plane scenes + the reference's correspondence text format
  multih_b200.dist     correspondence sharding + collectives over torch.distributed (NCCL / gloo)
"""
from . import scenes  # noqa: F401
from . import capi  # noqa: F401
from . import dist  # noqa: F401
from .capi import Context, MHError, build_library, library_path  # noqa: F401
from .multih import MultiH  # noqa: F401

__all__ = ["capi", "scenes", "dist", "Context", "MHError", "MultiH", "build_library", "library_path"]
