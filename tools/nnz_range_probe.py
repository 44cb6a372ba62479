#!/usr/bin/env python
"""cfg3 (R-MAT 2^22, k=32): the kernel time of each of the P equal non-zero ranges (NonZeroElement.cpp:24-39) on ONE GPU —
which range bounds the P-GPU time of the non-zero strategy."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparsematrixmultiplicationmpi_b200 as spmm

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:  # e.g. merge.items=512
    from sparsematrixmultiplicationmpi_b200 import _cabi
    for kv in sys.argv[2:]:
        key, val = kv.split("=")
        _cabi.tune(key, int(val))
A = spmm.DeviceCSR.rmat(22, 16 << 22, seed=11)
n, k = A.n_rows, 32
B = torch.randint(1, 101, (n, k), device="cuda").double()
s = torch.cuda.current_stream().cuda_stream
rp = A.download().rowPtr
for r in range(P):
    b, e = spmm.partition_nnz(A.nnz, P, r)
    first, last = A.nnz_range_rows(b, e)
    C = torch.empty((last - first + 1, k), dtype=torch.float64, device="cuda")
    for _ in range(3):
        A.multiply_nnz_range(b, e, first, last, B.data_ptr(), k, C.data_ptr(), "auto", s)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        A.multiply_nnz_range(b, e, first, last, B.data_ptr(), k, C.data_ptr(), "auto", s)
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"rank": r, "nnz": e - b, "rows": last - first + 1, "ms": t0.elapsed_time(t1) / 10}), flush=True)
    del C
