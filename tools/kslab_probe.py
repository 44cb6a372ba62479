#!/usr/bin/env python
"""cfg4 (banded 2^25 x 32 per row, half-bandwidth 4096): one multiply with k = 16 against k-slabs of 8 / 4 / 2 columns whose B slab
is CONTIGUOUS (ld = slab width), so that an L1 line holds 2 / 4 / 8 rows of the slab instead of one row of B."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparsematrixmultiplicationmpi_b200 as spmm

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = 1 << lg
A = spmm.DeviceCSR.banded(n, 32, 4096, seed=7)
s = torch.cuda.current_stream().cuda_stream


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for k in (16, 8, 4, 2):
    B = torch.randint(1, 101, (n, k), device="cuda").double()
    C = torch.empty((n, k), dtype=torch.float64, device="cuda")
    ms = timed(lambda: A.multiply(B.data_ptr(), k, C.data_ptr(), "auto", s))
    print(json.dumps({"k": k, "ms": ms, "passes_for_16": 16 // k, "ms_for_16_columns": ms * (16 // k)}), flush=True)
    del B, C
B16 = torch.randint(1, 101, (n, 16), device="cuda").double()
ms = timed(lambda: [B16[:, :8].contiguous(), B16[:, 8:].contiguous()])
print(json.dumps({"split_B_into_two_slabs_ms": ms}))
