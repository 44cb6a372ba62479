import torch
dev="cuda"
def t(fn, iters=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/iters*1e3
for mb in (17, 34, 78, 156):
    n=mb*1000*1000//8
    srcs=[torch.rand(n,dtype=torch.float64,device=dev) for _ in range(max(2, 400//mb))]
    dst=[torch.empty(n//2,dtype=torch.float64,device=dev) for _ in range(len(srcs))]
    # copy half -> read n/2*8 + write n/2*8 = mb MB of traffic; sum -> read mb MB
    us_copy=t(lambda i: dst[i%len(srcs)].copy_(srcs[i%len(srcs)][:n//2]))
    us_sum=t(lambda i: srcs[i%len(srcs)].sum())
    print(f"{mb} MB: copy (read+write = {mb} MB) {us_copy:.1f} us = {mb/us_copy*1e3:.0f} GB/s ; sum (read {mb} MB) {us_sum:.1f} us = {mb/us_sum*1e3:.0f} GB/s", flush=True)
