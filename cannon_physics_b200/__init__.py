"""cannon_physics_b200 — B200-native (sm_100a) drop-in for the per-step hot path of cannon_physics.

The package is a thin host-side mirror of the reference's World/Body/Shape/Material/Broadphase/Solver
API (``api.py``) over the C ABI of ``libcannon_cuda.so`` (``include/cannon_cuda.h``).  There is no CPU
fallback: importing the package fails loudly when the CUDA library has not been built.
"""
from __future__ import annotations

import os

from . import _ffi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcannon_cuda.so")
_lib = None


def load_library():
    """Return the bound libcannon_cuda.so (cached). Raises ImportError when it is missing."""
    global _lib
    if _lib is None:
        try:
            _lib = _ffi.bind(LIB_PATH)
        except (OSError, AttributeError) as e:
            raise ImportError(
                f"libcannon_cuda.so is missing or incomplete ({e}); build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback") from e
    return _lib


lib = load_library()

from .engine import Context, DeviceWorld, SceneSpec  # noqa: E402
from . import scenes  # noqa: E402
from .api import *  # noqa: E402,F401,F403
