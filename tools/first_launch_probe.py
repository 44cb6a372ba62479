#!/usr/bin/env python
"""What the first multiply of a process pays for (lazy module loading, schedule, allocations):
    python tools/first_launch_probe.py <warm_k or 0> <k>
times, in one fresh process, an optional warm-up multiply with `warm_k` columns on a tiny matrix and then the first three
multiplies with k columns on cfg2 (device-resident operands, synchronised wall clock)."""
import json
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sparsematrixmultiplicationmpi_b200 as spmm
from sparsematrixmultiplicationmpi_b200 import generators as gen

warm_k, k = int(sys.argv[1]), int(sys.argv[2])
kernel = sys.argv[3] if len(sys.argv) > 3 else "auto"
torch.cuda.init()
torch.zeros(1, device="cuda")
torch.cuda.synchronize()
out = {"warm_k": warm_k, "k": k, "kernel": kernel}
n, nc, r, c, v, sym = gen.cop20k_A_shaped()
A = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
torch.cuda.synchronize()
if warm_k:
    t = spmm.DeviceCSR.from_coo_host(8, 8, np.arange(8, dtype=np.int32), np.arange(8, dtype=np.int32), np.ones(8), False, device=0)
    b = torch.ones((8, warm_k), dtype=torch.float64, device="cuda")
    cc = torch.empty((8, warm_k), dtype=torch.float64, device="cuda")
    t0 = time.perf_counter()
    t.multiply(b.data_ptr(), warm_k, cc.data_ptr(), "rows")
    torch.cuda.synchronize()
    out["warm_ms"] = (time.perf_counter() - t0) * 1e3
B = torch.ones((n, k), dtype=torch.float64, device="cuda")
C = torch.empty((n, k), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ts = []
for i in range(3):
    t0 = time.perf_counter()
    A.multiply(B.data_ptr(), k, C.data_ptr(), kernel)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
out["multiply_ms"] = ts
print(json.dumps(out))
