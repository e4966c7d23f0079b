"""Python binding of the product library libjmme_cuda.so (hand-written sm_100a kernels behind the
C ABI of include/jmme.h).  There is no CPU fallback: importing works anywhere, but `load()` raises
if the CUDA library has not been built, and every compute call fails without a CUDA device."""
from __future__ import annotations

import pathlib

from . import abi, synth
from .abi import *  # noqa: F401,F403  (constants, Lib, Context, JmmeError)

PKG_DIR = pathlib.Path(__file__).resolve().parent.parent
CUDA_LIB = PKG_DIR / "csrc" / "libjmme_cuda.so"

_lib = None


def load() -> abi.Lib:
    """Load libjmme_cuda.so (built by __graft_entry__.build() / `make -C .../csrc`)."""
    global _lib
    if _lib is None:
        if not CUDA_LIB.exists():
            raise RuntimeError(f"{CUDA_LIB} is missing: build it with `python -c 'import __graft_entry__ as g; "
                               f"g.build()'` — there is no CPU fallback")
        _lib = abi.Lib(CUDA_LIB)
        if not _lib.backend().startswith("cuda"):
            raise RuntimeError(f"{CUDA_LIB} is not the CUDA backend: {_lib.backend()}")
    return _lib
