import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import sparsematrixmultiplicationmpi_b200 as spmm
from sparsematrixmultiplicationmpi_b200 import generators as gen, _cabi
def run(scale):
    n0, nnz0 = 121192, 2624331
    if scale == 1:
        n, nc, r, c, v, sym = gen.cop20k_A_shaped()
    else:
        n, nc, r, c, v, sym = gen.cop20k_A_shaped(n=n0*scale, nnz=nnz0*scale+ (scale%2), nx=49*2, ny=49*2 if scale==4 else 49)
    A0 = spmm.DeviceCSR.from_coo_host(n, nc, r, c, v, sym, device=0)
    host = A0.download(); A0.close()
    nnz = host.nnz
    for k in (1, 8):
        sets = []
        for s in range(6 if scale == 1 else 3):
            A = spmm.DeviceCSR.from_host(host, 0, 0)
            B = torch.randint(1, 101, (n, k), device='cuda').double(); C = torch.empty((n, k), dtype=torch.float64, device='cuda')
            sets.append((A, B, C))
        st = torch.cuda.current_stream().cuda_stream
        for kern in (('auto', 'stream') if k == 1 else ('auto',)):
            def fn(i):
                A, B, C = sets[i % len(sets)]; A.multiply(B.data_ptr(), k, C.data_ptr(), kern, st)
            for i in range(12): fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(60): fn(i)
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 60 * 1e3
            by = nnz * 12 + (n + 1) * 4 + 2 * n * k * 8
            print(f"scale={scale} n={n} nnz={nnz} k={k} {kern}: {us:.1f} us  {by/us/1e3:.0f} GB/s algorithmic", flush=True)
        for A, _, _ in sets: A.close()
run(1); run(4)
