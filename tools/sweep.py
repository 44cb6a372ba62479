#!/usr/bin/env python
"""Sweep + CSV of the four entry points: the local equivalent of the reference's batch_test.sh + get_csv_all.sh
(/root/reference "Source Code/scripts/batch_test.sh", "get_csv_all.sh":7).

For every matrix x k it does what main.cpp does (main.cpp:60-280): load or generate the matrix, generate the fat vector
with generateLargeFatVector, time the sequential multiply and the three strategies wall-clock around the call (host
buffers in, host FatVector out), compare every strategy with the sequential result (areMatricesEqual, 1e-6) and print the
reference's own lines, then writes one CSV row per run with the reference's columns plus GFLOP/s and the achieved
fraction of the HBM roofline of the sequential (single GPU) multiply. "Average Computation Time" = CUDA-event time of the
rank's multiply kernels, "Average Communication Time" = the rest of the call (copies, collectives), both averaged over
the ranks as the reference's debug build does (RowWise.cpp:89-98).

    python tools/sweep.py --matrices cfg1,cfg2 --k 4,64 --out gpurun_out/results.csv
    torchrun --nproc-per-node 4 tools/sweep.py --matrices cfg2,path/to/file.mtx --k 1,8,32,64
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sparsematrixmultiplicationmpi_b200 as spmm  # noqa: E402
from sparsematrixmultiplicationmpi_b200 import generators as gen, strategies  # noqa: E402

HEADER = ["file Name", "Cores Number", "Sparse Matrix", "Fat Vector", "Serial Algo Execution time",
          "Row-wise Average Communication Time", "Row-wise Average Computation Time", "Row-wise Execution time", "Row-wise Result",
          "Column-wise Average Communication Time", "Column-wise Average Computation Time", "Column-wise Execution time",
          "Column-wise Result", "Non-zero elements Average Communication Time", "Non-zero elements Average Computation Time",
          "Non-zero Elements Execution time", "Non-zero Elements Result", "PETSc Execution time", "PETSc Result",
          # additions
          "GPUs", "nnz", "Serial GFLOP/s", "Serial kernel us", "Serial HBM fraction"]


class KernelClock:
    """Accumulates the CUDA-event time of every multiply the engine launches (the 'computation' of a strategy)."""

    def __init__(self):
        self.pairs = []
        for name in ("multiply", "multiply_rows", "multiply_slab"):
            inner = getattr(strategies.CudaCompute, name)

            def wrapped(eng, *a, _inner=inner, **kw):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = _inner(eng, *a, **kw)
                e.record()
                self.pairs.append((s, e))
                return out
            setattr(strategies.CudaCompute, name, wrapped)

    def take(self) -> float:
        torch.cuda.synchronize()
        t = sum(s.elapsed_time(e) for s, e in self.pairs) * 1e-3
        self.pairs.clear()
        return t


def load(name: str, dev: int) -> tuple[str, spmm.SparseMatrix]:
    if name == "cfg1":
        n, nc, r, c, v, sym = gen.uniform_random()
    elif name == "cfg2":
        n, nc, r, c, v, sym = gen.cop20k_A_shaped()
    else:
        return os.path.basename(name), spmm.readMatrixMarketFile(name, dev)
    with spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=dev) as A:
        return name, A.download()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrices", default="cfg1,cfg2", help="cfg1, cfg2 or MatrixMarket files")
    ap.add_argument("--k", default="4,64")
    ap.add_argument("--repeat", type=int, default=3, help="calls per entry point; the fastest is reported (the first one uploads A)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "results.csv"))
    ap.add_argument("--append", action="store_true", help="add the rows to an existing CSV (one file over several launches at 1, 2, 4, 8 GPUs)")
    args = ap.parse_args()
    P, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if P > 1:
        dist.init_process_group("nccl")
    dev = torch.cuda.current_device()
    clock = KernelClock()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    rows = []
    for name in args.matrices.split(","):
        label, M = load(name, dev)
        for k in [int(x) for x in args.k.split(",")]:
            # pinned host buffers (what a caller that cares about the PCIe rate passes; the reference's FatVector is pageable)
            v = torch.from_numpy(spmm.generateLargeFatVector(M.numCols, k)).pin_memory().numpy()
            serial_out = torch.empty((M.numRows, k), dtype=torch.float64).pin_memory().numpy()
            if rank == 0:
                print(f"World size: {P}\nSparse matrix: {label}\nMatrix size: {M.numRows}x{M.numCols}\nVector size: {M.numCols}x{k}",
                      flush=True)

            def timed(fn):
                best, comp, out = None, 0.0, None
                for _ in range(args.repeat):
                    if P > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    clock.take()
                    t0 = time.perf_counter()
                    out = fn(M, v, k)
                    torch.cuda.synchronize()
                    t = time.perf_counter() - t0
                    c = clock.take()
                    if best is None or t < best:
                        best, comp = t, c
                tt = torch.tensor([best, comp, best - comp], dtype=torch.float64, device="cuda")
                if P > 1:
                    mx = tt.clone()
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                    dist.all_reduce(tt, op=dist.ReduceOp.SUM)
                    return out, mx[0].item(), tt[1].item() / P, tt[2].item() / P
                return out, best, comp, best - comp

            serial = serial_t = None
            kernel_us = 0.0
            # the sequential multiply runs on rank 0 only (main.cpp:70-84); keep the other ranks in step
            if rank == 0:
                best = None
                for _ in range(args.repeat):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    serial = spmm.sparseMatrixFatVectorMultiply(M, v, k, out=serial_out)
                    t = time.perf_counter() - t0
                    best = t if best is None else min(best, t)
                serial_t = best
                # device-resident kernel time of the same multiply, for the roofline columns
                with spmm.DeviceCSR.from_host(M, dev) as A:
                    dB = torch.from_numpy(v).cuda()
                    dC = torch.empty((M.numRows, k), dtype=torch.float64, device="cuda")
                    st = torch.cuda.current_stream().cuda_stream
                    for _ in range(12):  # AUTO builds a handle's tile layout on its 9th multiply: outside the timing
                        A.multiply(dB.data_ptr(), k, dC.data_ptr(), "auto", st)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    for _ in range(20):
                        A.multiply(dB.data_ptr(), k, dC.data_ptr(), "auto", st)
                    e.record()
                    torch.cuda.synchronize()
                    kernel_us = s.elapsed_time(e) / 20 * 1e3
                print(f"Serial Algo Execution time: {serial_t}s", flush=True)
            res = {}
            for tag, fn, line in (("Row-wise", spmm.sparseMatrixFatVectorMultiplyRowWise, "Row-wise"),
                                  ("Column-wise", spmm.sparseMatrixFatVectorMultiplyColumnWise, "Column-wise"),
                                  ("Non-zero elements", spmm.sparseMatrixFatVectorMultiplyNonZeroElement, "Non-zero Elements")):
                out, t, comp, comm = timed(fn)
                same = ""
                if rank == 0:
                    same = "same" if spmm.areMatricesEqual(serial, out, 1e-6) else "different"
                    # the reference's labels (main.cpp:168-277; the two averages are its debug build's, RowWise.cpp:89-98)
                    print(f"{tag} Average Communication Time: {comm}s\n{tag} Average Computation Time: {comp}s\n"
                          f"{line} Execution time: {t}s\n{line}: Results are {'the same!' if same == 'same' else 'different!'}",
                          flush=True)
                res[tag] = (comm, comp, t, same)
            if rank == 0:
                flops = 2.0 * M.nnz * k
                algo = M.nnz * 12 + (M.numRows + 1) * 4 + (M.numCols + M.numRows) * k * 8
                rows.append([f"{label}_k{k}_gpus{P}", P, f"{M.numRows}x{M.numCols}", f"{M.numCols}x{k}", serial_t,
                             *res["Row-wise"], *res["Column-wise"], *res["Non-zero elements"], "", "not built (PETSc is out of scope)",
                             P, M.nnz, flops / serial_t / 1e9, kernel_us, algo / (kernel_us * 1e-6) / 1e9 / peak])
        spmm.clear_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        fresh = not (args.append and os.path.exists(args.out))
        with open(args.out, "w" if fresh else "a", newline="") as f:
            w = csv.writer(f)
            if fresh:
                w.writerow(HEADER)
            w.writerows(rows)
        print("wrote", args.out)
    if P > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
