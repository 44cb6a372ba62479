"""The C-ABI library loads and exports every symbol include/spmm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "spmm_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spmm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ["spmm_csr_create_host", "spmm_csr_from_coo_host", "spmm_multiply_host", "spmm_multiply_device",
                 "spmm_multiply_rows_device", "spmm_multiply_nnz_range_device", "spmm_csr_column_block",
                 "spmm_partition_rows", "spmm_partition_cols", "spmm_partition_nnz", "spmm_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from sparsematrixmultiplicationmpi_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_prototypes_cover_the_header():
    from sparsematrixmultiplicationmpi_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()
    _cabi.lib()  # binds every prototype


def test_no_device_means_error_not_fallback():
    import torch
    from sparsematrixmultiplicationmpi_b200 import _cabi
    if torch.cuda.is_available():
        return
    n = ctypes.c_int(-1)
    rc = _cabi.lib().spmm_device_count(ctypes.byref(n))
    assert rc != 0 or n.value == 0
    h = ctypes.c_void_p()
    rc = _cabi.lib().spmm_gen_banded(0, 16, 2, 4, 1, ctypes.byref(h))
    assert rc != 0 and _cabi.lib().spmm_last_error()


def test_entry_point_library_keeps_the_reference_cxx_signatures():
    """libspmm_entry.so exports the four C++ entry points with the reference's mangled names."""
    path = os.path.join(ROOT, "sparsematrixmultiplicationmpi_b200", "libspmm_entry.so")
    lib = ctypes.CDLL(path)
    sig = "RK12SparseMatrixRKSt6vectorIS2_IdSaIdEESaIS4_EEi"
    for name in ["_Z29sparseMatrixFatVectorMultiply", "_Z36sparseMatrixFatVectorMultiplyRowWise",
                 "_Z39sparseMatrixFatVectorMultiplyColumnWise", "_Z43sparseMatrixFatVectorMultiplyNonZeroElement"]:
        assert hasattr(lib, name + sig), name
