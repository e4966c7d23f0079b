"""The C-ABI libraries load and export every symbol include/jmme.h declares (no compute calls)."""
import pathlib
import re

import numpy as np
import pytest

from jmme import abi

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    hdr = (ROOT / "include" / "jmme.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(jmme_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(abi.EXPORTS)


def test_oracle_exports_every_symbol(oracle):
    for s in declared_symbols():
        assert hasattr(oracle.dll, s), s


def test_cuda_library_exports_every_symbol_and_fails_loudly_without_gpu():
    import jmme
    import torch
    if not jmme.CUDA_LIB.exists():
        pytest.fail(f"{jmme.CUDA_LIB} not built: run __graft_entry__.build()")
    lib = jmme.load()
    for s in declared_symbols():
        assert hasattr(lib.dll, s), s
    assert lib.backend() == "cuda-sm_100a"
    assert lib.lambda_factor(28, 0) == 6 << 16                       # host-only helper
    if not torch.cuda.is_available():
        # no CPU fallback: creating a context must fail with NODEVICE, and so must the leaves
        with pytest.raises(abi.JmmeError) as e:
            lib.context(width=32, height=32, search_range=4)
        assert e.value.code == abi.ERR_NODEVICE
        with pytest.raises(abi.JmmeError) as e:
            lib.satd(np.zeros((1, 16), np.int16))
        assert e.value.code in (abi.ERR_NODEVICE, abi.ERR_CUDA)


def test_mbresult_layout_matches_header():
    assert abi.MBRESULT_DTYPE.itemsize == 372
    assert abi.MBRESULT_DTYPE.fields["cost"][1] == 164 and abi.MBRESULT_DTYPE.fields["ref_idx"][1] == 328
    assert len(abi.block_table()) == 41
