import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import sparsematrixmultiplicationmpi_b200 as spmm
k = 64
n, nc, r, c, v, sym = bench.build_workload(k)
A = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
host = A.download()
B = np.random.default_rng(1).integers(1, 101, (n, k)).astype(np.float64)
lib = bench.entry_lib()
for rep in range(4):
    spmm.clear_cache()
    first, mean, _ = bench.entry_run(lib, 0, 1, host, B, k, 3)
    print(json.dumps({"rep": rep, "first_ms": first * 1e3, "mean_ms": mean * 1e3}), flush=True)
