"""The N>1 host logic of the three strategies on CPU: world_size 2 and 3 over gloo.

The package has no CPU engine; these tests inject one built on the oracle (test infrastructure)
so that partitioning, shard construction, the boundary-row exchange and the gathers are exercised
without a GPU. Results are checked against the oracle's restatement of each reference strategy.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_csr


class OracleCompute:
    """Test-only engine: same interface as CudaCompute, CPU tensors, arithmetic by oracle/spmm_oracle.c."""

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        self.o = pyoracle.Oracle()
        self.device = torch.device("cpu")

    class Shard:
        def __init__(self, m):
            self.m, self.n_rows, self.n_cols, self.nnz = m, m.numRows, m.numCols, m.nnz

    def upload(self, m):
        return OracleCompute.Shard(m)

    def multiply(self, A, B, k, out=None):
        m = A.m
        C = self.o.spmm(m.rowPtr, m.colIndices, m.values, B.numpy(), k) if m.numRows else np.zeros((0, k))
        t = torch.from_numpy(np.ascontiguousarray(C))
        if out is None:
            return t
        out.copy_(t)
        return out

    def column_span(self, A):
        return (int(A.m.colIndices.min()), int(A.m.colIndices.max())) if A.m.nnz else (0, -1)

    def multiply_window(self, A, window, first_row, k, out=None):
        full = np.full((A.m.numCols, k), np.nan)  # rows outside the window must never be read with a non-zero weight
        full[first_row:first_row + window.shape[0]] = window.numpy()
        m = A.m
        C = self.o.spmm(m.rowPtr, m.colIndices, m.values, np.nan_to_num(full, nan=0.0), k) if m.numRows else np.zeros((0, k))
        touched = np.zeros(m.numCols, bool)
        touched[m.colIndices] = True
        assert not np.isnan(full[touched]).any(), "the window misses a row of B the block reads"
        t = torch.from_numpy(np.ascontiguousarray(C))
        if out is None:
            return t
        out.copy_(t)
        return out

    def multiply_rows(self, A, row_begin, row_end, B, k, out):
        m = A.m
        if row_end > row_begin:
            blk = m.row_block(row_begin, row_end)
            out.copy_(torch.from_numpy(self.o.spmm(blk.rowPtr, blk.colIndices, blk.values, B.numpy(), k)))
        return out

    def multiply_slab(self, A, B, k, k_begin, k_count, out):
        m = A.m
        if k_count:
            Bs = np.ascontiguousarray(B.numpy()[:, k_begin:k_begin + k_count])
            out[:, k_begin:k_begin + k_count] = torch.from_numpy(self.o.spmm(m.rowPtr, m.colIndices, m.values, Bs, k_count))
        return out


def _worker(rank, world, port, case, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparsematrixmultiplicationmpi_b200 as spmm
        seed, n, mean, k, long_row, empty_every = case
        rowptr, colidx, vals = random_csr(seed, n, n, mean, long_row=long_row, empty_every=empty_every)
        m = spmm.SparseMatrix(vals, colidx, rowptr, n, n)
        B = np.random.default_rng(seed).integers(1, 101, (n, k)).astype(np.float64)
        Bt = torch.from_numpy(B)
        eng = OracleCompute()
        out = {}
        row = spmm.RowWise.from_host(eng, m, k)
        out["row"] = row.run(Bt)
        out["row_allgather"] = row.all_gather(row.multiply_local(Bt))
        blk = spmm.ColumnBlocks.from_host(eng, m, k)
        out["colblk"] = blk.run(blk.local_B(Bt))
        out["colslab"] = spmm.ColumnSlabs.from_host(eng, m, k).run(Bt)
        extra = {"colblk_overlap": blk.multiply_reduce_scatter_overlapped(blk.local_B(Bt), chunks=3),
                 "colblk_plain": blk.reduce_scatter(blk.multiply_local(blk.local_B(Bt))),
                 "row_overlap": row.multiply_all_gather_overlapped(Bt, chunks=3)}
        assert torch.equal(extra["colblk_overlap"], extra["colblk_plain"])  # same sums in the same order
        # the exchange fused into the multiply (on the GPU: peer stores; here: its host mirror): every row block lands in its
        # owner's slot `sender`, the slots are added in ascending rank order
        pushed = blk.multiply_reduce_scatter_push(blk.local_B(Bt))
        partials = [torch.empty((blk.block * world, k), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(partials, blk.multiply_local(blk.local_B(Bt)))
        want = partials[0][rank * blk.block:(rank + 1) * blk.block].clone()
        for q in range(1, world):
            want += partials[q][rank * blk.block:(rank + 1) * blk.block]
        assert torch.equal(pushed, want)
        out["nnz"] = spmm.NonZeroRanges.from_host(eng, m, k).run(Bt)
        # B sharded by rows like C: halo exchange, then the block multiply on the window; gathered for the check
        bs, be = spmm.partition_rows(n, world, rank)
        out["row_sharded"] = row.gather(row.multiply_sharded(Bt[bs:be].clone()))
        window, own = row.alloc_window(torch.device("cpu"))
        own.copy_(Bt[bs:be])
        w0 = row.exchange_halo_inplace(window)  # no local copy: the rank's rows already sit inside the window buffer
        again = row.gather(eng.multiply_window(row.A, window, w0, k))
        assert again is None if rank else torch.equal(again, out["row_sharded"])
        results[f"row_overlap_{rank}"] = extra["row_overlap"].numpy().copy()
        if rank == 0:
            results.update({name: t.numpy().copy() for name, t in out.items()})
        else:
            assert all(out[name] is None for name in ("row", "colblk", "colslab", "nnz", "row_sharded"))  # FatVector{} off-root
            results[f"allgather_{rank}"] = out["row_allgather"].numpy().copy()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = [
    (2, (21, 90, 6, 4, None, 5)),
    (2, (22, 64, 4, 3, 500, 7)),    # hub row cut by the non-zero range boundary
    (3, (23, 50, 3, 5, 400, 0)),    # hub row spanning all three ranks
    (3, (24, 7, 2, 2, None, 3)),    # fewer rows than a comfortable split
]


@pytest.mark.parametrize("world,case", CASES)
def test_strategies_world(oracle, world, case):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), case, results), nprocs=world, join=True)
    seed, n, mean, k, long_row, empty_every = case
    rowptr, colidx, vals = random_csr(seed, n, n, mean, long_row=long_row, empty_every=empty_every)
    B = np.random.default_rng(seed).integers(1, 101, (n, k)).astype(np.float64)
    seq = oracle.spmm(rowptr, colidx, vals, B, k, "seq")
    # row-wise and the reference's column slabs keep the sequential accumulation order: bit-identical
    assert np.array_equal(results["row"], oracle.spmm(rowptr, colidx, vals, B, k, "row", world))
    assert np.array_equal(results["row"], seq)
    assert np.array_equal(results["colslab"], oracle.spmm(rowptr, colidx, vals, B, k, "col", world))
    assert np.array_equal(results["row_allgather"], seq)
    assert np.array_equal(results["row_sharded"], seq)  # halo exchange instead of a replicated B: same rows, same order
    for r in range(1, world):
        assert np.array_equal(results[f"allgather_{r}"], seq)
    for r in range(world):
        assert np.array_equal(results[f"row_overlap_{r}"], seq)
    # non-zero ranges: partial sums of cut rows added in rank order == the oracle's rank-order reduce
    assert np.array_equal(results["nnz"], oracle.spmm(rowptr, colidx, vals, B, k, "nnz", world))
    # column blocks regroup the sum by column block: equal up to FP64 summation order
    scale = np.abs(vals).max() * 100 * max(np.diff(rowptr).max(), 1)
    assert np.allclose(results["colblk"], seq, rtol=1e-12, atol=1e-12 * scale)


def test_single_rank_degenerate_case(oracle):
    import sparsematrixmultiplicationmpi_b200 as spmm
    rowptr, colidx, vals = random_csr(31, 40, 40, 5, long_row=90, empty_every=6)
    m = spmm.SparseMatrix(vals, colidx, rowptr, 40, 40)
    B = torch.from_numpy(np.random.default_rng(3).integers(1, 101, (40, 3)).astype(np.float64))
    eng = OracleCompute()
    seq = oracle.spmm(rowptr, colidx, vals, B.numpy(), 3)
    assert np.array_equal(spmm.RowWise.from_host(eng, m, 3).run(B).numpy(), seq)
    blk = spmm.ColumnBlocks.from_host(eng, m, 3)
    assert np.array_equal(blk.run(blk.local_B(B)).numpy(), seq)
    assert np.array_equal(spmm.ColumnSlabs.from_host(eng, m, 3).run(B).numpy(), seq)
    assert np.array_equal(spmm.NonZeroRanges.from_host(eng, m, 3).run(B).numpy(), seq)
