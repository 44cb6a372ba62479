#!/usr/bin/env python
"""Where the end-to-end time of one cfg2 k=64 multiply goes, per host path (B200 box):
flat pinned / flat pageable through spmm_multiply_host, row pointers through spmm_multiply_host_rows, the C++ entry point
(vector<vector<double>> in and out), the Python mirror. Prints one JSON line per path."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import _cabi  # noqa: E402


def mean_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n, nc, r, c, v, sym = bench.build_workload(k)
    A = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
    host = A.download()
    A.build_tiles(-1, 0, k)
    B = np.random.default_rng(1).integers(1, 101, (n, k)).astype(np.float64)
    out = {"k": k, "host_threads": _cabi.lib().spmm_host_threads()}
    Bp, Cp = torch.from_numpy(B).pin_memory(), torch.empty((n, k), dtype=torch.float64).pin_memory()
    out["flat_pinned_ms"] = mean_ms(lambda: A.multiply_host(Bp.numpy(), k, "auto", Cp.numpy()))
    Cq = np.empty((n, k))
    out["flat_pageable_ms"] = mean_ms(lambda: A.multiply_host(B, k, "auto", Cq))
    rows_in = [np.ascontiguousarray(B[i]) for i in range(n)]
    rows_out = [np.empty(k) for _ in range(n)]
    pin = (C.c_void_p * n)(*[x.ctypes.data for x in rows_in])
    pout = (C.c_void_p * n)(*[x.ctypes.data for x in rows_out])
    L = _cabi.lib()
    out["row_pointers_ms"] = mean_ms(lambda: _cabi.check(L.spmm_multiply_host_rows(A.handle, pin, k, pout, 0)))
    for slabs in (1, 2, 4):
        _cabi.tune("host.slabs", slabs)
        out[f"row_pointers_slabs{slabs}_ms"] = mean_ms(lambda: _cabi.check(L.spmm_multiply_host_rows(A.handle, pin, k, pout, 0)))
        out[f"flat_pinned_slabs{slabs}_ms"] = mean_ms(lambda: A.multiply_host(Bp.numpy(), k, "auto", Cp.numpy()))
    _cabi.tune("reset", 0)
    lib = bench.entry_lib()
    first, mean, _ = bench.entry_run(lib, 0, 1, host, B, k, 20)
    out["cxx_entry_first_call_ms"], out["cxx_entry_ms"] = first * 1e3, mean * 1e3
    for strategy, name in ((1, "row"), (2, "col"), (3, "nnz")):
        for P in (1, min(8, max(1, torch.cuda.device_count()))):
            first, mean, _ = bench.entry_run(lib, strategy, P, host, B, k, 5)
            out[f"cxx_{name}_P{P}_first_ms"], out[f"cxx_{name}_P{P}_ms"] = first * 1e3, mean * 1e3
    out["python_seq_ms"] = mean_ms(lambda: spmm.sparseMatrixFatVectorMultiply(host, B, k), 10)
    out["python_row_ms"] = mean_ms(lambda: spmm.sparseMatrixFatVectorMultiplyRowWise(host, B, k), 5)
    out["python_col_ms"] = mean_ms(lambda: spmm.sparseMatrixFatVectorMultiplyColumnWise(host, B, k), 5)
    out["python_nnz_ms"] = mean_ms(lambda: spmm.sparseMatrixFatVectorMultiplyNonZeroElement(host, B, k), 5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
