#!/usr/bin/env python
"""bench.py — the headline measurement of the SpMM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--k 64]

One "step" = one pass C = A*B of the cop20k_A-shaped k=64 SpMM (BASELINE.json configs[1], the
configuration the north_star target is quoted on) over one resident operand set. With N > 1 (one
process per GPU under torchrun) every rank owns one cop20k_A-shaped diagonal block of a
(N*121,192)-row matrix — the row-wise partition (RowWise.cpp:26-29) with B replicated and C left
row-sharded, no collective inside the timed region (weak scaling); the NCCL broadcast of B and
the all-gather of C are timed separately and reported under "collectives".

Printed by rank 0: ONE JSON line with metric / value (GFLOP/s = 2*nnz*k/t, whole job) plus
roofline (algorithmic bytes / kernel time vs the measured HBM copy peak), e2e (the same multiply
through the reference-shaped host entry point, host buffers, copies inside the timed region) and
cpu_baseline (the reference's own CPU code on the host cores).

--impl reference times the reference's own CPU implementation (oracle/_ref, compiled from the
reference sources; else the oracle port) on the same workload, all host threads, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ROWS, NNZ = 121_192, 2_624_331
OPERAND_SETS = 4  # rotating resident copies of (A, B, C): 4 x ~175 MB >> 126 MB L2


def algorithmic_bytes(n_rows: int, nnz: int, k: int) -> int:
    """A's CSR arrays + B once + C once."""
    return nnz * 12 + (n_rows + 1) * 4 + 2 * n_rows * k * 8


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def recorded_traffic(k: int):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"cop20k_k{k}")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~5 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(k: int, seed: int = 20):
    """cop20k_A-shaped records (host) — the same for both arms."""
    from sparsematrixmultiplicationmpi_b200 import generators as gen
    return gen.cop20k_A_shaped(n=N_ROWS, nnz=NNZ, seed=seed)


def cpu_reference_run(host, B, k, steps, warmup, max_seconds=60.0):
    """The reference's own CPU code on the host cores: row-wise strategy at P = all hardware threads
    (its "mpirun" path on compat MPI) and the sequential function; returns the faster one."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    cores = os.cpu_count() or 1
    if pyoracle.Reference.available():
        ref, kind = pyoracle.Reference("fast"), "reference"

        def run(strategy, P):
            return ref.spmm(host.numCols, host.rowPtr, host.colIndices, host.values, B, k, strategy, P,
                            want_result=False)[1]
    else:
        orc, kind = pyoracle.Oracle(), "port"
        cores = 1

        def run(strategy, P):
            t0 = time.perf_counter()
            orc.spmm(host.rowPtr, host.colIndices, host.values, B, k, "seq", 1)
            return time.perf_counter() - t0
    best = {}
    t_begin = time.perf_counter()
    for strategy, P in (("row", cores), ("seq", 1)) if kind == "reference" else (("seq", 1),):
        for _ in range(max(1, warmup)):
            run(strategy, P)
        times = []
        for _ in range(steps):
            times.append(run(strategy, P))
            if time.perf_counter() - t_begin > max_seconds:
                break
        best[(strategy, P)] = statistics.mean(times)
    (strategy, P), t = min(best.items(), key=lambda kv: kv[1])
    flops = 2.0 * host.nnz * k
    return {"value": flops / t / 1e9, "unit": "GFLOP/s", "cores": P, "kind": kind,
            "sample": f"full cop20k_A-shaped k={k} multiply, reference {strategy} strategy at P={P} "
                      f"(-O3 -march=x86-64-v3, compat MPI rank-threads), mean of {len(times)} calls",
            "seconds_per_step": t,
            "sequential_gflops": flops / best[("seq", 1)] / 1e9 if ("seq", 1) in best else None,
            "rowwise_all_cores_gflops": flops / best[("row", cores)] / 1e9 if ("row", cores) in best else None,
            "host_cores": os.cpu_count()}


def main():
    # stdout carries exactly ONE JSON line: everything else this process or its libraries print
    # (NCCL's version banner, torchrun notices) is sent to stderr.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    k = args.k
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "SpMM GFLOP/s and HBM GB/s (% roofline) at 1/2/4/8 B200 vs ref CPU MPI"
    config = {"workload": f"cop20k_A-shaped synthetic FEM matrix {N_ROWS}x{N_ROWS}, {NNZ} nnz, k={k}, FP64 "
                          f"(BASELINE.json configs[1]); one such diagonal block per GPU, row-wise partition",
              "n_rows_per_gpu": N_ROWS, "nnz_per_gpu": NNZ, "k": k,
              "l2": f"{OPERAND_SETS} rotating resident operand sets (> 126 MB L2) so every step reads cold operands"}

    if args.impl == "reference":
        if rank != 0:
            return
        import sparsematrixmultiplicationmpi_b200  # noqa: F401  (generators only; no GPU work on this arm)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        n, nc, r, c, v, sym = build_workload(k)
        rp, ci, va = pyoracle.Oracle().csr_from_coo(n, r, c, v, sym)  # the reference loader's CSR assembly
        from sparsematrixmultiplicationmpi_b200.matrix import SparseMatrix
        host = SparseMatrix(va, ci, rp, n, nc)
        B = np.random.default_rng(1).integers(1, 101, (n, k)).astype(np.float64)
        steps = min(args.steps, 20)
        res = cpu_reference_run(host, B, k, steps, min(args.warmup, 2))
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": res["seconds_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import sparsematrixmultiplicationmpi_b200 as spmm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SpMM path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- operands resident in HBM: OPERAND_SETS copies of this rank's diagonal block ----
    n, nc, r, c, v, sym = build_workload(k, seed=20 + rank)
    first = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=local_rank)
    host = first.download()
    sets = []
    for s in range(OPERAND_SETS):
        A = first if s == 0 else spmm.DeviceCSR.from_host(host, local_rank)
        if args.kernel in ("auto", "tiled"):
            A.build_tiles(-1, 0, k)  # what AUTO does by itself on its first k>=16 multiply; done here so it is outside any timing
        Bd = torch.randint(1, 101, (n, k), device=dev).double()
        Cd = torch.empty((n, k), dtype=torch.float64, device=dev)
        sets.append((A, Bd, Cd))
    tiles = sets[0][0].tile_info()
    launches_per_step = 2 if (args.kernel == "merge") else 1

    def step(i):
        A, Bd, Cd = sets[i % OPERAND_SETS]
        A.multiply(Bd.data_ptr(), k, Cd.data_ptr(), args.kernel, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    flops_per_step = 2.0 * NNZ * k * world
    value = flops_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (this rank's launch) ----
    peak, peak_kind = measured_peak_gbs()
    abytes = algorithmic_bytes(N_ROWS, NNZ, k)
    achieved = abytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(k), "peak_kind": peak_kind, "algorithmic_bytes_per_launch": abytes,
                "frac_of_8TBs_nominal": achieved / 8000.0,
                "kernel": "spmm_tiled_kernel" if tiles["rows_per_tile"] and k >= 8 and k % 2 == 0 and args.kernel in ("auto", "tiled")
                else ("spmm_merge_kernel" if args.kernel == "merge" else "spmm_rows_kernel"),
                "tiles": tiles}

    # ---- e2e: the reference-shaped host entry point, pinned host buffers, copies inside the timed region ----
    Bh = torch.randint(1, 101, (n, k)).double().pin_memory()
    Ch = torch.empty((n, k), dtype=torch.float64).pin_memory()
    Bh_np, Ch_np = Bh.numpy(), Ch.numpy()
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(3):
        spmm.sparseMatrixFatVectorMultiply(host, Bh_np, k, out=Ch_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        spmm.sparseMatrixFatVectorMultiply(host, Bh_np, k, out=Ch_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": flops_per_step / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": n * k * 8,
           "d2h_bytes_per_step": n * k * 8, "ms_per_step": e2e_s * 1e3,
           "api": "sparseMatrixFatVectorMultiply(M, B_host, k) with A cached in HBM after the first call"}

    # ---- N > 1: the collectives of the row-wise strategy, timed on their own ----
    collectives = None
    if world > 1:
        plan_counts = [N_ROWS] * world
        Bfull = torch.empty((N_ROWS * world, k), dtype=torch.float64, device=dev)
        Call = torch.empty((N_ROWS * world, k), dtype=torch.float64, device=dev)
        out = {}
        for name, fn in (("broadcast_B", lambda: dist.broadcast(Bfull, src=0)),
                         ("all_gather_C", lambda: dist.all_gather_into_tensor(Call, sets[0][2]))):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name + "_ms"] = float(t.item())
        out["bytes"] = N_ROWS * world * k * 8

        # the same all-gather fused into the multiply (C rows stored from registers into every peer's buffer
        # over NVLink, spmm_multiply_scatter_device) against multiply + NCCL all-gather, both timed as one unit
        def timed_unit(fn, iters=10):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        A0, B0, C0 = sets[0]
        out["multiply+all_gather_nccl_ms"] = timed_unit(
            lambda: (A0.multiply(B0.data_ptr(), k, C0.data_ptr(), args.kernel, stream),
                     dist.all_gather_into_tensor(Call, C0)))
        try:
            import torch.distributed._symmetric_memory as symm_mem
            sym = symm_mem.empty((N_ROWS * world, k), dtype=torch.float64, device=dev)
            hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
            ptrs = [int(hdl.buffer_ptrs[(rank + i) % world]) + rank * N_ROWS * k * 8 for i in range(world)]
            out["multiply+all_gather_fused_p2p_ms"] = timed_unit(
                lambda: (A0.multiply_scatter(B0.data_ptr(), k, ptrs, args.kernel if args.kernel in ("auto", "rows", "merge", "tiled") else "auto", stream),
                         hdl.barrier(channel=0)))
            torch.cuda.synchronize()
            out["fused_p2p_matches_nccl"] = bool(torch.equal(sym, Call))
        except Exception as e:  # symmetric memory unavailable on this box: report, do not fail the bench line
            out["multiply+all_gather_fused_p2p_error"] = str(e)[:200]
        collectives = out
        del plan_counts

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = np.random.default_rng(1).integers(1, 101, (n, k)).astype(np.float64)
        cpu = cpu_reference_run(host, Bc, k, steps=5, warmup=1, max_seconds=30.0)

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": wall,
                "hbm_gbs_per_gpu": achieved, "kernel_arg": args.kernel}
        if collectives:
            line["collectives"] = collectives
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
