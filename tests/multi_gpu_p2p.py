"""torchrun script (not collected by pytest): RowWise with the gather fused into the multiply, checked against
the NCCL path and the oracle on every rank. Run: torchrun --nproc-per-node N tests/multi_gpu_p2p.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import generators as gen  # noqa: E402
from sparsematrixmultiplicationmpi_b200.strategies import ColumnBlocks, CudaCompute, RowWise  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, P = dist.get_rank(), dist.get_world_size()
    oracle = pyoracle.Oracle()
    n, nc, r, c, v, sym = gen.cop20k_A_shaped(n=30_011, nnz=630_001, nx=31, ny=31, seed=9)
    rp, ci, va = oracle.csr_from_coo(n, r, c, v, sym)
    host = spmm.SparseMatrix(va, ci, rp, n, nc)
    eng = CudaCompute(local, kernel="rows")  # one kernel family on every path: the fused stores must be bit-identical to NCCL's copy
    for k in (64, 6):
        B = np.random.default_rng(k).integers(1, 101, (n, k)).astype(np.float64)
        ref = oracle.spmm(rp, ci, va, B, k)
        plan = RowWise.from_host(eng, host, k)
        dB = torch.from_numpy(B).cuda()
        full = plan.multiply_all_gather_p2p(dB)
        torch.cuda.synchronize()
        got = full.cpu().numpy()
        assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref)), f"rank {rank}: all-gather p2p mismatch at k={k}"
        nccl = plan.all_gather(plan.multiply_local(dB)).cpu().numpy()
        assert np.array_equal(got, nccl), f"rank {rank}: p2p differs from the NCCL all-gather at k={k}"
        dist.barrier()
        root = plan.run_p2p(dB)
        torch.cuda.synchronize()
        if rank == 0:
            assert np.array_equal(root.cpu().numpy(), nccl), "gather-to-root p2p mismatch"
        else:
            assert root is None
        dist.barrier()
        # column blocks: partial C reduced by P2P loads in rank order against NCCL reduce-scatter
        cb = ColumnBlocks.from_host(eng, host, k)
        Bl = torch.from_numpy(np.ascontiguousarray(B[cb.col_start:cb.col_end])).cuda()
        mine_nccl = cb.reduce_scatter(cb.multiply_local(Bl)).cpu().numpy()
        for _ in range(2):  # twice: the symmetric partial buffer is reused
            mine_p2p = cb.multiply_reduce_scatter_p2p(Bl).cpu().numpy()
        r0 = rank * cb.block
        want = ref[r0:r0 + cb.block]
        assert np.all(np.abs(mine_p2p[:want.shape[0]] - want) <= 1e-12 * np.abs(want)), f"rank {rank}: p2p reduce mismatch k={k}"
        assert np.all(np.abs(mine_p2p - mine_nccl) <= 1e-12 * np.abs(mine_nccl) + 1e-300)
        # the exchange fused into the multiply (peer stores into the owner's slots, local rank-order sum): the same bits
        for _ in range(2):
            mine_push = cb.multiply_reduce_scatter_push(Bl).cpu().numpy()
        assert np.array_equal(mine_push, mine_p2p), f"rank {rank}: fused peer-store reduce differs from the pull variant at k={k}"
        dist.barrier()
    if rank == 0:
        print("p2p ok: world", P)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
