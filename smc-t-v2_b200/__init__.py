"""smct_b200 — B200-native SMC-Transformer particle-filter forward pass.

The directory is named ``smc-t-v2_b200`` (repository layout contract); it is imported under the module
name ``smct_b200`` through the loader ``smct_b200.py`` at the repository root.

Layers:
  csrc/ + libsmct_b200.so   hand-written sm_100a CUDA kernels behind the C-ABI of include/smct.h
  _lib, functional          ctypes binding; torch tensors are containers only
  models/, train/, algos/   host-side mirror of the reference's Keras-level call surface
"""
from . import build as _build_mod  # noqa: F401
from ._lib import SmctError, load as load_library, lib_path  # noqa: F401
from .build import build  # noqa: F401

__all__ = ["build", "load_library", "lib_path", "SmctError"]
