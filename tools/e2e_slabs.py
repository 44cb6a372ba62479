"""Host-buffer multiply (spmm_multiply_host) with 1/2/4 k-slabs in the PCIe pipeline: time and check (cfg2, k=64)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsematrixmultiplicationmpi_b200 as spmm
from sparsematrixmultiplicationmpi_b200 import _cabi, generators as gen

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, nc, r, c, v, sym = gen.cop20k_A_shaped()
A = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
host = A.download()
Bh = torch.randint(1, 101, (n, k)).double().pin_memory()
Ch = torch.empty((n, k), dtype=torch.float64).pin_memory()
ref = None
for slabs in (1, 2, 4):
    if k % (slabs * 2):
        continue
    _cabi.tune("reset", 0)
    _cabi.tune("host.slabs", slabs)
    for _ in range(3):
        spmm.sparseMatrixFatVectorMultiply(host, Bh.numpy(), k, out=Ch.numpy())
    t0 = time.perf_counter()
    for _ in range(20):
        spmm.sparseMatrixFatVectorMultiply(host, Bh.numpy(), k, out=Ch.numpy())
    dt = (time.perf_counter() - t0) / 20
    out = Ch.numpy().copy()
    if ref is None:
        ref = out
    print(f"k={k} slabs={slabs}: {dt*1e3:.3f} ms  {2.0*host.nnz*k/dt/1e9:.1f} GFLOP/s  max|diff vs 1 slab|={np.abs(out-ref).max():.3g}")
_cabi.tune("reset", 0)
