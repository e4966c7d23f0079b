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


def _struct_fields(name):
    """(field, array length or 1) of `typedef struct <name> { ... } <name>;` in include/jmme.h, in order."""
    hdr = (ROOT / "include" / "jmme.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(int32_t|int16_t|int8_t|int64_t)\s+(.*)$", decl, flags=re.S)
        assert m, decl
        for item in m.group(2).split(","):
            fm = re.match(r"\s*(\w+)\s*((?:\[\w+\])*)\s*$", item)
            n = 1
            for dim in re.findall(r"\[(\w+)\]", fm.group(2)):
                n *= int(dim) if dim.isdigit() else {"JMME_MAX_GPUS": abi.MAX_GPUS, "JMME_BLOCKS_PER_MB": abi.BLOCKS_PER_MB}[dim]
            out.append((fm.group(1), n))
    return out


def test_params_and_tuning_structs_match_the_header():
    """Field by field, in order: jmme_params / jmme_tuning of include/jmme.h against the ctypes structures (an
    int32 more or less in one of them shifts every later field silently), and the ABI version constant."""
    import ctypes as C
    for name, cls in (("jmme_params", abi.Params), ("jmme_tuning", abi.Tuning)):
        want = _struct_fields(name)
        got = [(n, (C.sizeof(t) // 4)) for n, t in cls._fields_]
        assert got == want, (name, got, want)
        assert C.sizeof(cls) == 4 * sum(n for _, n in want)
    hdr = (ROOT / "include" / "jmme.h").read_text()
    assert int(re.search(r"#define JMME_ABI_VERSION\s+(\d+)", hdr).group(1)) == abi.ABI_VERSION
