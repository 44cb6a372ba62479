"""ctypes doors into the oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. The product package never
imports this module.

`Oracle`   -> oracle/spmm_oracle.c, the C restatement ("port").
`Reference`-> the reference's own sources compiled from /root/reference by
              oracle/Makefile ("reference"); `exact` build for parity bits
              (-O2 -ffp-contract=off), `fast` build for CPU timing (-O3 x86-64-v3).
All dense operands are flat row-major float64, indices int32.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """make -C oracle (liboracle.so always; _ref only when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _as(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """C restatement of the reference path (oracle/spmm_oracle.c)."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        common = [C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int]
        L.oracle_spmm_seq.argtypes = common + [_f64p]
        L.oracle_spmm_seq.restype = None
        for name in ("oracle_spmm_rowwise", "oracle_spmm_colwise", "oracle_spmm_nnz"):
            getattr(L, name).argtypes = common + [C.c_int, _f64p]
        L.oracle_spmm_nnz.restype = C.c_int
        for name in ("oracle_partition_rows", "oracle_partition_cols", "oracle_partition_nnz"):
            getattr(L, name).argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_generate_fatvector.argtypes = [C.c_int, C.c_int, _f64p]
        L.oracle_are_equal.argtypes = [_f64p, _f64p, C.c_size_t, C.c_double]
        L.oracle_csr_from_coo.argtypes = [C.c_int, C.c_long, _i32p, _i32p, _f64p, C.c_int, _i32p, _i32p, _f64p]
        L.oracle_csr_from_coo.restype = C.c_long
        L.oracle_read_mtx.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 2 + [C.POINTER(C.c_long)] + \
            [C.POINTER(C.c_int)] * 2 + [C.c_void_p] * 3

    # -- multiply --
    def spmm(self, rowptr, colidx, vals, B, k, strategy="seq", P=1):
        rowptr, colidx, vals = _as(rowptr, np.int32), _as(colidx, np.int32), _as(vals, np.float64)
        B = _as(B, np.float64).reshape(-1)
        n_rows = rowptr.size - 1
        Cm = np.empty(n_rows * k, dtype=np.float64)
        a = (n_rows, rowptr, colidx, vals, B, k)
        if strategy == "seq":
            self.lib.oracle_spmm_seq(*a, Cm)
        elif strategy == "row":
            self.lib.oracle_spmm_rowwise(*a, P, Cm)
        elif strategy == "col":
            self.lib.oracle_spmm_colwise(*a, P, Cm)
        elif strategy == "nnz":
            if self.lib.oracle_spmm_nnz(*a, P, Cm) != 0:
                raise MemoryError("oracle_spmm_nnz")
        else:
            raise ValueError(strategy)
        return Cm.reshape(n_rows, k)

    def partition(self, kind, total, P, r):
        s, e = C.c_int(), C.c_int()
        getattr(self.lib, f"oracle_partition_{kind}")(total, P, r, C.byref(s), C.byref(e))
        return s.value, e.value

    def generate_fatvector(self, n, k):
        out = np.empty(n * k, dtype=np.float64)
        self.lib.oracle_generate_fatvector(n, k, out)
        return out.reshape(n, k)

    def are_equal(self, a, b, tol):
        a, b = _as(a, np.float64).reshape(-1), _as(b, np.float64).reshape(-1)
        return a.size == b.size and bool(self.lib.oracle_are_equal(a, b, a.size, tol))

    # -- CSR construction --
    def csr_from_coo(self, n_rows, rows, cols, vals, symmetric=False):
        rows, cols, vals = _as(rows, np.int32), _as(cols, np.int32), _as(vals, np.float64)
        cap = rows.size * (2 if symmetric else 1)
        rowptr = np.empty(n_rows + 1, dtype=np.int32)
        colidx = np.empty(max(cap, 1), dtype=np.int32)
        ov = np.empty(max(cap, 1), dtype=np.float64)
        nnz = self.lib.oracle_csr_from_coo(n_rows, rows.size, rows, cols, vals, int(symmetric), rowptr, colidx, ov)
        if nnz < 0:
            raise MemoryError("oracle_csr_from_coo")
        return rowptr, colidx[:nnz].copy(), ov[:nnz].copy()

    def read_mtx_coo(self, path):
        nr, nc, sym, pat = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ne = C.c_long()
        hdr = (C.byref(nr), C.byref(nc), C.byref(ne), C.byref(sym), C.byref(pat))
        rc = self.lib.oracle_read_mtx(path.encode(), *hdr, None, None, None)
        if rc == 0:
            rows = np.empty(max(ne.value, 1), dtype=np.int32)
            cols = np.empty(max(ne.value, 1), dtype=np.int32)
            vals = np.empty(max(ne.value, 1), dtype=np.float64)
            rc = self.lib.oracle_read_mtx(path.encode(), *hdr, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data)
        if rc == 1:
            raise RuntimeError("Unable to open file: " + path)           # utils.cpp:77
        if rc == 2:
            raise RuntimeError("Failed to read matrix dimensions from file: " + path)  # utils.cpp:114
        if rc == 3:
            raise RuntimeError("Failed to read data from file: " + path)  # utils.cpp:140
        n = ne.value
        return nr.value, nc.value, rows[:n], cols[:n], vals[:n], bool(sym.value), bool(pat.value)

    def read_mtx(self, path):
        """-> (n_rows, n_cols, rowptr, colidx, vals) with the loader's exact CSR semantics."""
        nr, nc, rows, cols, vals, sym, _ = self.read_mtx_coo(path)
        rowptr, colidx, v = self.csr_from_coo(nr, rows, cols, vals, sym)
        return nr, nc, rowptr, colidx, v


class Reference:
    """The reference's own code (compiled from /root/reference into oracle/_ref)."""

    STRATEGY = {"seq": 0, "row": 1, "col": 2, "nnz": 3}

    def __init__(self, flavour: str = "exact"):
        path = os.path.join(HERE, "_ref", f"libref_{flavour}.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (needs /root/reference at build time)")
        self.flavour = flavour
        self.lib = L = C.CDLL(path)
        L.ref_spmm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int,
                               C.c_void_p, C.POINTER(C.c_double)]
        L.ref_read_mtx.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 3 + [C.c_void_p] * 3
        L.ref_generate_fatvector.argtypes = [C.c_int, C.c_int, _f64p]
        L.ref_serialize_is_rowmajor.argtypes = [_f64p, C.c_int, C.c_int]
        L.ref_are_equal.argtypes = [_f64p, _f64p, C.c_int, C.c_int, C.c_double]
        L.ref_last_error.restype = C.c_char_p

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libref_exact.so"))

    def spmm(self, n_cols, rowptr, colidx, vals, B, k, strategy="seq", P=1, want_result=True):
        """-> (C or None, seconds of the multiply call on rank 0)."""
        rowptr, colidx, vals = _as(rowptr, np.int32), _as(colidx, np.int32), _as(vals, np.float64)
        B = _as(B, np.float64).reshape(-1)
        n_rows = rowptr.size - 1
        out = np.empty(n_rows * k, dtype=np.float64) if want_result else None
        sec = C.c_double()
        rc = self.lib.ref_spmm(self.STRATEGY[strategy], P, n_rows, n_cols, rowptr, colidx, vals, B, k,
                               out.ctypes.data if want_result else None, C.byref(sec))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return (out.reshape(n_rows, k) if want_result else None), sec.value

    def read_mtx(self, path):
        nr, nc, nz = C.c_int(), C.c_int(), C.c_int()
        if self.lib.ref_read_mtx(path.encode(), C.byref(nr), C.byref(nc), C.byref(nz), None, None, None):
            raise RuntimeError(self.lib.ref_last_error().decode())
        rowptr = np.empty(nr.value + 1, dtype=np.int32)
        colidx = np.empty(max(nz.value, 1), dtype=np.int32)
        vals = np.empty(max(nz.value, 1), dtype=np.float64)
        if self.lib.ref_read_mtx(path.encode(), C.byref(nr), C.byref(nc), C.byref(nz),
                                 rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return nr.value, nc.value, rowptr, colidx[:nz.value], vals[:nz.value]

    def generate_fatvector(self, n, k):
        out = np.empty(n * k, dtype=np.float64)
        self.lib.ref_generate_fatvector(n, k, out)
        return out.reshape(n, k)

    def serialize_is_rowmajor(self, flat, n, k):
        return bool(self.lib.ref_serialize_is_rowmajor(_as(flat, np.float64).reshape(-1), n, k))

    def are_equal(self, a, b, tol):
        a, b = _as(a, np.float64), _as(b, np.float64)
        return bool(self.lib.ref_are_equal(a.reshape(-1), b.reshape(-1), a.shape[0], a.shape[1], tol))
