#!/usr/bin/env python
"""Wall-clock cost of spmm_csr_build_tiles_for_k on cfg2 (first call of the process, then repeats on fresh handles)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparsematrixmultiplicationmpi_b200 as spmm
from sparsematrixmultiplicationmpi_b200 import generators as gen

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, nc, r, c, v, sym = gen.cop20k_A_shaped()
first = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
host = first.download()
torch.cuda.synchronize()
for rep in range(4):
    A = spmm.DeviceCSR.from_host(host, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = A.build_tiles(-1, 0, k)
    torch.cuda.synchronize()
    print(json.dumps({"rep": rep, "build_ms": (time.perf_counter() - t0) * 1e3, "rows_per_tile": info["rows_per_tile"]}), flush=True)
    A.close()
