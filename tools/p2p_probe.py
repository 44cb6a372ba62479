#!/usr/bin/env python
"""One process, P rank-threads, one GPU each: the C++ row-wise entry point on cfg2 k=64 a few times (spmm_entry_run).
Meant to be run under `ncu --metrics nvl...` on a 2-GPU box: the kernels of rank 1 store their C rows straight into rank 0's
buffer over NVLink (spmm_multiply_scatter_device with one peer destination)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
k = 64
n, nc, r, c, v, sym = bench.build_workload(k)
with spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0) as A:
    host = A.download()
B = np.random.default_rng(1).integers(1, 101, (n, k)).astype(np.float64)
first, mean, C = bench.entry_run(bench.entry_lib(), 1, P, host, B, k, 3, want_result=True, warmup=1)
print(f"row-wise P={P}: first {first * 1e3:.2f} ms, mean {mean * 1e3:.2f} ms, checksum {C.sum():.6e}")
