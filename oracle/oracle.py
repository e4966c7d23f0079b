"""Loader of the CPU oracle (oracle/libjmme_oracle.so).  TEST INFRASTRUCTURE ONLY — imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
from jmme import abi  # noqa: E402

ORACLE_LIB = HERE / "libjmme_oracle.so"
_lib = None


def build(force=False):
    src = HERE / "jmme_oracle.c"
    if force or not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "-B" if force else "-s", "libjmme_oracle.so"])


def load() -> abi.Lib:
    global _lib
    if _lib is None:
        if not ORACLE_LIB.exists():
            build()
        _lib = abi.Lib(ORACLE_LIB)
        assert _lib.backend() == "cpu-oracle"
    return _lib
