#!/usr/bin/env python
"""Row-kernel schedule sweep on the large BASELINE shapes (cfg4 banded 2^25 x 32 k=16, cfg5 uniform 2^23 x 32 k=64):
round-robin row-tile height x side-by-side non-zeros x unroll. One JSON line per setting.
    python tools/tune_rows.py cfg4 [log2 rows]"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import _cabi  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else (25 if cfg == "cfg4" else 23)
KERNEL = "merge" if cfg == "cfg3" else "rows"
if cfg == "cfg3":
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    A = spmm.DeviceCSR.rmat(lg, 16 << lg, seed=11)
    n, k = A.n_rows, 32
else:
    n, k = 1 << lg, (16 if cfg == "cfg4" else 64)
    A = spmm.DeviceCSR.banded(n, 32, 4096 if cfg == "cfg4" else n // 2, seed=7)
B = torch.randint(1, 101, (n, k), device="cuda").double()
C = torch.empty((n, k), dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream


def run(tune, iters=5):
    _cabi.tune("reset", 0)
    for key, val in tune.items():
        _cabi.tune(key, val)
    for _ in range(2):
        A.multiply(B.data_ptr(), k, C.data_ptr(), KERNEL, s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        A.multiply(B.data_ptr(), k, C.data_ptr(), KERNEL, s)
    b.record()
    torch.cuda.synchronize()
    _cabi.tune("reset", 0)
    return a.elapsed_time(b) / iters


print(json.dumps({"setting": "default", "ms": run({})}), flush=True)
if cfg == "cfg3":
    for (kl, nv), u, items in itertools.product(((8, 2), (16, 1), (4, 4), (8, 1)), (1, 2, 4, 8), (256, 512, 2048)):
        tune = {"rows.kl": kl, "rows.nv": nv, "rows.unroll": u, "merge.items": items}
        try:
            print(json.dumps({"setting": tune, "ms": run(tune)}), flush=True)
        except Exception as e:
            print(json.dumps({"setting": tune, "error": str(e)[:100]}), flush=True)
    sys.exit(0)
tiles = (8, 16, 32, 48, 64, 96, 128) if cfg == "cfg4" else (16, 64, 128, 512)
for tile, np_, u in itertools.product(tiles, (1, 2, 4) if cfg == "cfg4" else (1,), (1, 2, 4)):
    tune = {"rows.tile": tile, "rows.unroll": u}
    if cfg == "cfg4":
        tune["rows.np"] = np_
    try:
        print(json.dumps({"setting": tune, "ms": run(tune)}), flush=True)
    except Exception as e:
        print(json.dumps({"setting": tune, "error": str(e)[:100]}), flush=True)
