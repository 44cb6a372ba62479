#!/usr/bin/env python
"""Quick parity check of the union kernel against the CSR row kernel on small matrices (run under `timeout`)."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import _cabi  # noqa: E402
from sparsematrixmultiplicationmpi_b200.matrix import SparseMatrix  # noqa: E402


def banded(n, per_row, half_bw, seed, hub=0):
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for i in range(n):
        m = per_row if not (hub and i % 97 == 5) else hub
        c = np.unique(np.clip(i + rng.integers(-half_bw, half_bw + 1, size=m), 0, n - 1))
        rows.append(np.full(c.size, i))
        cols.append(c)
    r, c = np.concatenate(rows), np.concatenate(cols)
    rp = np.zeros(n + 1, dtype=np.int32)
    np.add.at(rp, r + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    return SparseMatrix(0.5 + rng.random(r.size), c.astype(np.int32), rp, n, n)


def main():
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    worst = 0.0
    for (n, per_row, bw, hub) in ((64, 5, 8, 0), (1000, 12, 30, 0), (5000, 20, 60, 300), (20011, 24, 200, 0)):
        m = banded(n, per_row, bw, n, hub)
        for R, sl, ncw in ((2, 4, 6), (2, 4, 4), (2, 4, 8), (2, 8, 4), (4, 4, 4)):
            for k in (64, 32, 40, 2, 96):
                A = spmm.DeviceCSR.from_host(m, 0, 0)
                _cabi.tune("reset", 0)
                _cabi.tune("union.slots", sl)
                _cabi.tune("tiled.ncw", ncw)
                try:
                    info = A.build_union(R, k)
                finally:
                    _cabi.tune("reset", 0)
                B = torch.randint(1, 101, (n, k), device=dev).double()
                ref = torch.empty((n, k), dtype=torch.float64, device=dev)
                out = torch.full((n, k), float("nan"), dtype=torch.float64, device=dev)
                A.multiply(B.data_ptr(), k, ref.data_ptr(), "rows", stream)
                A.multiply(B.data_ptr(), k, out.data_ptr(), "union", stream)
                torch.cuda.synchronize()
                err = ((out - ref).abs() / ref.abs().clamp_min(1e-300)).nan_to_num(nan=float("inf")).max().item()
                worst = max(worst, err)
                print(f"n={n} R={R} SL={sl} NCW={ncw} k={k}: max rel err {err:.3e} {info}", flush=True)
                A.close()
    print("WORST", worst)
    sys.exit(0 if worst <= 1e-12 else 1)


if __name__ == "__main__":
    main()
