"""The drop-in boundary itself on the GPU: the reference's main.cpp + utils.cpp, unchanged, linked against
libspmm_entry.so (bin/spmm_main), and the four C++ entry points called with C++ SparseMatrix / FatVector objects on
P rank-threads (spmm_entry_run), held to the oracle. SURVEY.md section 8(b): main.cpp:78,162,205,248 are the only callers
and areMatricesEqual(..., 1e-6) (main.cpp:184,227,270) the reference's only check."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import sparsematrixmultiplicationmpi_b200 as spmm
from sparsematrixmultiplicationmpi_b200 import generators as gen
from conftest import GOLDEN, REL_TOL, ROOT, assert_close_rel, random_csr

pytestmark = pytest.mark.gpu

PKG = os.path.join(ROOT, "sparsematrixmultiplicationmpi_b200")
MAIN = os.path.join(PKG, "bin", "spmm_main")
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "ref_main")
TIMING = re.compile(r"(Execution time: )\S+")


def run_main(binary, P, k, path):
    r = subprocess.run([binary, "-np", str(P), str(k), path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.fixture(scope="module")
def cfg1_mtx(tmp_path_factory):
    """BASELINE.json configs[0] written as a MatrixMarket file (what the reference's loader ingests)."""
    n, nc, r, c, v, sym = gen.uniform_random(10_000, 10, seed=1)
    path = str(tmp_path_factory.mktemp("mtx") / "cfg1.mtx")
    gen.write_matrix_market(path, n, nc, r, c, v, sym)
    return path


@pytest.mark.parametrize("P", [1, 4])
@pytest.mark.parametrize("case", ["golden_random_symmetric", "cfg1"])
def test_reference_main_runs_unchanged_on_the_gpu(P, case, cfg1_mtx):
    if not os.path.exists(MAIN):
        pytest.skip("bin/spmm_main not built (needs /root/reference at build time)")
    path, k = (os.path.join(GOLDEN, "random_symmetric.mtx"), 6) if case == "golden_random_symmetric" else (cfg1_mtx, 4)
    out = run_main(MAIN, P, k, path)
    # the reference's own verdict, three times (main.cpp:186,229,272)
    for label in ("Row-wise", "Column-wise", "Non-zero Elements"):
        assert f"{label}: Results are the same!" in out, out
    assert not [l for l in out.splitlines() if "different" in l and not l.startswith("PETSc")], out  # (PETSc is a compile-time stub)
    if os.path.exists(REF_MAIN):
        # same labelled lines as the reference build of the same main.cpp (timings aside): the CSV scrapers still work
        ref = run_main(REF_MAIN, P, k, path)
        strip = lambda s: [TIMING.sub(r"\1T", l) for l in s.splitlines() if not l.startswith("PETSc")]
        assert strip(out) == strip(ref)


# ---- the four entry points with C++ objects, P rank-threads, against the oracle ----
def entry_lib():
    lib = C.CDLL(os.path.join(PKG, "libspmm_entry.so"))
    lib.spmm_entry_run.restype = C.c_int
    lib.spmm_entry_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.c_char_p, C.c_int]
    return lib


def entry_run(strategy, P, rp, ci, va, n_cols, B, k, steps=0):
    lib = entry_lib()
    n = len(rp) - 1
    rp, ci, va = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32), np.ascontiguousarray(va, np.float64)
    B = np.ascontiguousarray(B, np.float64)
    out = np.full((n, k), np.nan)
    first, mean = C.c_double(), C.c_double()
    err = C.create_string_buffer(512)
    rc = lib.spmm_entry_run(strategy, P, n, n_cols, len(va), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, k,
                            B.ctypes.data, out.ctypes.data, 1 if steps else 0, steps, C.byref(first), C.byref(mean), err, 512)
    if rc:
        raise RuntimeError(err.value.decode() or f"spmm_entry_run status {rc}")
    return out, first.value, mean.value


STRATEGY = {0: "seq", 1: "row", 2: "seq", 3: "nnz"}  # column blocks sum partials: compared to seq within the tolerance


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8, 11])
@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [4, 7])
def test_cxx_entry_points_vs_oracle(oracle, strategy, P, k):
    if strategy == 0 and P > 1:
        pytest.skip("the sequential function runs on rank 0 only (main.cpp:77-81)")
    n = 3001
    rp, ci, va = random_csr(71, n, n, 9, long_row=5000, empty_every=6, positive=True)  # hub row: cut by several nnz ranges
    B = np.random.default_rng(k).integers(1, 101, (n, k)).astype(np.float64)
    got, _, _ = entry_run(strategy, P, rp, ci, va, n, B, k)
    ref = oracle.spmm(rp, ci, va, B, k, STRATEGY[strategy], P)
    assert_close_rel(got, ref, tol=REL_TOL)
    assert spmm.areMatricesEqual(got, oracle.spmm(rp, ci, va, B, k), 1e-6)  # the reference's own runtime check


def test_cxx_entry_points_more_ranks_than_rows_and_empty_matrix(oracle):
    rp, ci, va = random_csr(5, 5, 5, 2, positive=True)
    B = np.random.default_rng(1).integers(1, 101, (5, 3)).astype(np.float64)
    ref = oracle.spmm(rp, ci, va, B, 3)
    for strategy in (1, 2, 3):
        got, _, _ = entry_run(strategy, 6, rp, ci, va, 5, B, 3)  # P > N: trailing ranks own nothing
        assert_close_rel(got, ref, tol=REL_TOL)
    rp = np.zeros(41, np.int32)
    for strategy in (0, 1, 2, 3):
        got, _, _ = entry_run(strategy, 3, rp, np.zeros(0, np.int32), np.zeros(0), 40, np.ones((40, 2)), 2)
        assert not got.any()


def test_cxx_entry_points_cop20k_shape(oracle):
    """BASELINE.json configs[1] through the C++ objects: k = 64, P = 1 and 8."""
    n, nc, r, c, v, sym = gen.cop20k_A_shaped()
    rp, ci, va = oracle.csr_from_coo(n, r, c, v, sym)
    k = 64
    B = np.random.default_rng(64).integers(1, 101, (n, k)).astype(np.float64)
    ref = oracle.spmm(rp, ci, va, B, k)
    for strategy, P in ((0, 1), (1, 8), (2, 8), (3, 8)):
        got, first, mean = entry_run(strategy, P, rp, ci, va, n, B, k, steps=2)
        assert_close_rel(got, ref, tol=REL_TOL)
        assert first > 0 and mean > 0
