#!/usr/bin/env python
"""Write BASELINE.json's cfg1 / cfg2 synthetic matrices as MatrixMarket files (what the reference's loader and bin/spmm_main ingest)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sparsematrixmultiplicationmpi_b200 import generators as gen  # noqa: E402

name, path = sys.argv[1], sys.argv[2]
n, nc, r, c, v, sym = gen.uniform_random() if name == "cfg1" else gen.cop20k_A_shaped()
with open(path, "w") as f:
    f.write(f"%%MatrixMarket matrix coordinate real {'symmetric' if sym else 'general'}\n{n} {nc} {len(r)}\n")
    np.savetxt(f, np.column_stack([r + 1, c + 1, v]), fmt="%d %d %.17g")
print("wrote", path, len(r), "records")
