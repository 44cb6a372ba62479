#!/usr/bin/env python
"""How much of the tiled kernel's advantage on the cop20k_A shape comes from the generator's natural x-fastest grid order?
The same matrix under other symmetric orderings (rows and columns permuted alike): what the tile layout reuses and how long
the AUTO multiply takes at k = 64, beside the CSR row kernel. Real cop20k_A is an unstructured mesh in a solver's order
(typically bandwidth-reducing); the orderings below bracket that.

    natural          the generator's order (x fastest on a 49 x 49 x 51 grid)
    scramble-512     unknowns shuffled inside consecutive windows of 512 (local disorder, same bandwidth)
    rcm-of-random    a random global permutation followed by reverse Cuthill-McKee (what a solver would do to a mesh)
    random           a random global permutation (no locality left: the layout must be refused and the row kernels take over)

Prints one JSON line per ordering."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402
from scipy.sparse.csgraph import reverse_cuthill_mckee  # noqa: E402

import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import generators as gen  # noqa: E402


def timed(fn, iters=40, warm=6):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


BOX = int(sys.argv[1]) if len(sys.argv) > 1 else 0  # box height of the tile layout (0 = the default, 16 rows)


def main():
    k = 64
    n, nc, r, c, v, sym = gen.cop20k_A_shaped()
    rng = np.random.default_rng(5)
    orders = {"natural": np.arange(n)}
    p = np.arange(n)
    for s in range(0, n, 512):
        p[s:s + 512] = s + rng.permutation(min(512, n - s))
    orders["scramble-512"] = p
    rnd = rng.permutation(n)
    A = sp.coo_matrix((np.ones(r.size), (rnd[r], rnd[c])), shape=(n, n)).tocsr()
    A = (A + A.T).tocsr()
    rcm = reverse_cuthill_mckee(A, symmetric_mode=True)  # new position i holds old (randomised) unknown rcm[i]
    inv = np.empty(n, dtype=np.int64)
    inv[rcm] = np.arange(n)
    orders["rcm-of-random"] = inv[rnd]
    orders["random"] = rnd
    stream = torch.cuda.current_stream().cuda_stream
    for name, perm in orders.items():
        rr, cc = perm[r].astype(np.int32), perm[c].astype(np.int32)
        lo, hi = np.minimum(rr, cc), np.maximum(rr, cc)  # keep the lower-triangle storage of a symmetric file
        sets = []
        for _ in range(3):
            Ad = spmm.DeviceCSR.from_coo_host(n, nc, hi, lo, v, sym, device=0)
            info = Ad.build_tiles(-1, BOX, k)
            sets.append((Ad, torch.randint(1, 101, (n, k), device="cuda").double(),
                         torch.empty((n, k), dtype=torch.float64, device="cuda")))
        host = sets[0][0].download()
        bw = int(np.max(np.abs(np.repeat(np.arange(n), np.diff(host.rowPtr)) - host.colIndices)))

        def run(kernel):
            return timed(lambda i: sets[i % 3][0].multiply(sets[i % 3][1].data_ptr(), k, sets[i % 3][2].data_ptr(), kernel, stream))
        t_auto, t_rows = run("auto"), run("rows")
        from sparsematrixmultiplicationmpi_b200 import _cabi
        print(json.dumps({"ordering": name, "box_rows": BOX or 16, "bandwidth": bw, "tiles": info, "auto_us": t_auto, "rows_us": t_rows,
                          "auto_kernel": (_cabi.lib().spmm_last_kernel_name() or b"").decode()}), flush=True)
        for Ad, _, _ in sets:
            Ad.close()


if __name__ == "__main__":
    main()
