"""Host logic of the union kernel's layout builder (csrc/spmm_union_build.cu), checked without a GPU: the builder is pure
C++, so tests/union_layout_emul.cpp replays spmm_union_kernel's data movement on the CPU — loads applied as far ahead as
the kernel's barriers allow — and compares with the CSR multiply (SparseMatrixFatVectorMultiply.cpp:17-28)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "sparsematrixmultiplicationmpi_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_union_layout_emulation(tmp_path):
    exe = str(tmp_path / "union_layout_emul")
    subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-I", CSRC, "-o", exe,
                    os.path.join(ROOT, "tests", "union_layout_emul.cpp"), os.path.join(CSRC, "spmm_union_build.cu")],
                   check=True, capture_output=True, timeout=300)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "0 failures" in res.stdout
