#!/usr/bin/env python
"""Layout search of the union kernel without a device: compiles csrc/spmm_union_build.cu with g++ and
reports, for the cop20k_A-shaped matrix, how many B rows each parameter set stages per pass.

    g++ -x c++ -std=c++17 -O2 -DSPMM_UNION_PROBE -fPIC -shared -o /tmp/ub/libub.so sparsematrixmultiplicationmpi_b200/csrc/spmm_union_build.cu
    python tools/union_layout_probe.py /tmp/ub/libub.so
"""
import ctypes as C
import sys
import os
import time
import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsematrixmultiplicationmpi_b200 import generators as gen  # noqa: E402


def cfg2_csr():
    n, _, r, c, v, sym = gen.cop20k_A_shaped()
    off = r != c
    rr = np.concatenate([r, c[off]])
    cc = np.concatenate([c, r[off]])
    vv = np.concatenate([v, v[off]])
    m = sp.csr_matrix((vv, (rr, cc)), shape=(n, n))
    m.sort_indices()
    return n, m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.float64)


def main():
    lib = C.CDLL(sys.argv[1])
    n, rp, ci, va = cfg2_csr()
    nnz = int(rp[-1])
    print("n", n, "nnz", nnz)
    stats = (C.c_longlong * 12)()
    err = C.create_string_buffer(256)
    smem = 232448 - 2048
    for R, KT, D, nch, split in [(2, 32, 8, 74, 48), (2, 32, 6, 74, 48), (2, 32, 10, 74, 48), (2, 32, 12, 74, 48),
                                 (4, 32, 6, 74, 64), (2, 16, 16, 37, 48), (2, 16, 12, 37, 48), (2, 16, 24, 37, 48),
                                 (4, 16, 12, 37, 64), (2, 32, 8, 148, 48), (2, 32, 8, 74, 40), (2, 32, 8, 74, 64)]:
        t0 = time.time()
        rc = lib.spmm_union_layout_probe(n, n, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, R, KT, D, nch, smem, split, 0,
                                         stats, err, 256)
        s = list(stats)
        print(f"R={R} KT={KT} D={D} chunks={nch} split={split}: rc={rc} {err.value.decode()} items={s[0]} NG={s[1]} ring={s[2]} "
              f"union/nnz={s[3] / nnz:.3f} slotsteps/union={s[4] / max(1, s[3]):.3f} staged/N={s[5] / n:.2f} maxblob={s[6]} "
              f"maxsteps={s[7]} maxgroups={s[8]} splits={s[9]} blobMB={s[10] / 1e6:.1f} t={time.time() - t0:.2f}s")


if __name__ == "__main__":
    main()
