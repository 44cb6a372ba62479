"""Regenerates tests/golden/*.npz and *.mtx by running the REFERENCE'S OWN code.

Needs oracle/_ref (the reference sources compiled from /root/reference by oracle/Makefile), so it
only runs in the build container; the fixtures it writes are committed and travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

LOADER_CASES = {
    # name: (banner qualifiers, size line, records)  -- SURVEY.md Appendix A.2
    "general_unsorted": ("real general", "3 3 4", ["3 1 5.5", "1 3 2", "1 1 1", "2 2 -3e-1"]),
    "symmetric": ("real symmetric", "3 3 4", ["1 1 1", "2 1 2", "3 1 3", "3 3 4"]),
    "pattern": ("pattern general", "3 3 3", ["1 2", "2 3", "3 1"]),
    "duplicates": ("real general", "2 2 4", ["1 2 9", "1 2 1", "1 1 7", "2 2 1"]),
    "skew_symmetric": ("real skew-symmetric", "2 2 1", ["2 1 5"]),
    "rectangular": ("real general", "2 4 3", ["1 4 1", "2 1 2", "2 3 3"]),
    "pattern_symmetric": ("pattern symmetric", "4 4 5", ["1 1", "3 1", "4 2", "4 4", "2 1"]),
    "empty_rows": ("real general", "5 5 3", ["5 5 1.25", "1 2 -2", "5 1 1e300"]),
}


def random_case(seed, n, mean, k, long_row=None, empty_every=0):
    rng = np.random.default_rng(seed)
    lens = rng.poisson(mean, n)
    if empty_every:
        lens[::empty_every] = 0
    if long_row is not None:
        lens[n // 2] = long_row
    rowptr = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
    nnz = int(rowptr[-1])
    colidx = np.empty(nnz, dtype=np.int32)
    for i in range(n):
        row = np.sort(rng.integers(0, n, lens[i]))  # duplicates allowed, ascending like the loader
        colidx[rowptr[i]:rowptr[i + 1]] = row
    vals = rng.standard_normal(nnz)  # mixed sign
    B = rng.integers(1, 101, size=(n, k)).astype(np.float64)
    return rowptr, colidx, vals, B


def main():
    ref = pyoracle.Reference("exact")
    out = {}
    # A.1: the report's CSR example with the reference's own B generator
    rowptr = np.array([0, 2, 3, 3, 4], np.int32)
    colidx = np.array([0, 2, 2, 3], np.int32)
    vals = np.array([1., 2., 3., 4.])
    B = ref.generate_fatvector(4, 3)
    out["kat_rowptr"], out["kat_colidx"], out["kat_vals"], out["kat_B"] = rowptr, colidx, vals, B
    out["kat_C"] = ref.spmm(4, rowptr, colidx, vals, B, 3)[0]
    out["fatvec_7x5"] = ref.generate_fatvector(7, 5)

    for name, (seed, n, mean, k, long_row, empty_every) in {
        "small": (11, 60, 5, 4, None, 7),
        "hub": (12, 200, 8, 3, 1500, 9),
        "k1": (13, 150, 12, 1, None, 0),
        "k8": (14, 150, 20, 8, 300, 11),
    }.items():
        rowptr, colidx, vals, B = random_case(seed, n, mean, k, long_row, empty_every)
        out[f"{name}_rowptr"], out[f"{name}_colidx"], out[f"{name}_vals"], out[f"{name}_B"] = rowptr, colidx, vals, B
        out[f"{name}_C_seq"] = ref.spmm(n, rowptr, colidx, vals, B, k, "seq")[0]
        for P in (2, 3, 7):
            out[f"{name}_C_row_P{P}"] = ref.spmm(n, rowptr, colidx, vals, B, k, "row", P)[0]
            out[f"{name}_C_col_P{P}"] = ref.spmm(n, rowptr, colidx, vals, B, k, "col", P)[0]
            out[f"{name}_C_nnz_P{P}"] = ref.spmm(n, rowptr, colidx, vals, B, k, "nnz", P)[0]
    np.savez_compressed(os.path.join(HERE, "multiply.npz"), **out)

    loader = {}
    for name, (qual, size, recs) in LOADER_CASES.items():
        path = os.path.join(HERE, f"{name}.mtx")
        with open(path, "w") as f:
            f.write(f"%%MatrixMarket matrix coordinate {qual}\n% golden loader case {name}\n{size}\n")
            f.write("\n".join(recs) + "\n")
        nr, nc, rp, ci, va = ref.read_mtx(path)
        loader[f"{name}_shape"] = np.array([nr, nc])
        loader[f"{name}_rowptr"], loader[f"{name}_colidx"], loader[f"{name}_vals"] = rp, ci, va
    # a larger symmetric file with duplicates and unsorted records, written with %.17g
    rng = np.random.default_rng(5)
    n, m = 300, 2500
    r = rng.integers(0, n, m)
    c = rng.integers(0, n, m)
    r, c = np.maximum(r, c), np.minimum(r, c)
    v = np.round(rng.standard_normal(m), 3)  # rounded so equal (row, col, value) triples and value ties occur
    path = os.path.join(HERE, "random_symmetric.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n")
        f.write(f"{n} {n} {m}\n")
        for a, b, x in zip(r, c, v):
            f.write(f"{a + 1} {b + 1} {x:.17g}\n")
    nr, nc, rp, ci, va = ref.read_mtx(path)
    loader["random_symmetric_shape"] = np.array([nr, nc])
    loader["random_symmetric_rowptr"], loader["random_symmetric_colidx"], loader["random_symmetric_vals"] = rp, ci, va
    np.savez_compressed(os.path.join(HERE, "loader.npz"), **loader)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
