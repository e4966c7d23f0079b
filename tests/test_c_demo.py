"""The C ABI from plain C89 host code: examples/jmme_demo.c compiles with -std=c89 -pedantic -Werror against
include/jmme.h, runs against the CPU oracle here and against libjmme_cuda.so on the GPU box, and both print
the same motion-field checksums (the drop-in boundary of SURVEY.md §8(b) exercised the way JM would use it)."""
import pathlib
import subprocess

import pytest

from jmme import abi

ROOT = pathlib.Path(__file__).resolve().parent.parent


def build_and_run(tmp_path, libdir, libname):
    exe = tmp_path / f"demo_{libname}"
    cmd = ["/usr/bin/gcc", "-std=c89", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{ROOT / 'include'}",
           str(ROOT / "examples" / "jmme_demo.c"), f"-L{libdir}", f"-l{libname}", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=300).stdout.strip().splitlines()
    assert len(out) == 3, out
    return out


def test_c89_demo_against_the_oracle(oracle, tmp_path):
    out = build_and_run(tmp_path, ROOT / "oracle", "jmme_oracle")
    assert out[0].startswith(f"backend cpu-oracle, ABI {abi.ABI_VERSION}")
    assert "policy 0: 24 MBs" in out[1] and "policy 3: 24 MBs" in out[2]
    # the planted integer displacement is found by the inner macroblocks
    assert int(out[1].split("found in ")[1].split(",")[0]) >= 8


@pytest.mark.gpu
def test_c89_demo_against_the_cuda_library(cuda, oracle, tmp_path):
    ref = build_and_run(tmp_path, ROOT / "oracle", "jmme_oracle")
    got = build_and_run(tmp_path, ROOT / "h264-jm-commentary_b200" / "csrc", "jmme_cuda")
    assert got[0].startswith(f"backend cuda-sm_100a, ABI {abi.ABI_VERSION}")
    assert got[1:] == ref[1:]
