import torch, numpy as np
dev=torch.device('cuda:0')
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for gb in (1.26, 4.0):
    n=int(gb*1e9)//8
    a=torch.empty(n,dtype=torch.int64,device=dev); b=torch.empty(n,dtype=torch.int64,device=dev)
    t=tm(lambda: a.fill_(7)); print(f'fill  {gb} GB: {t:.3f} ms = {n*8/t/1e6:.0f} GB/s written')
    t=tm(lambda: a.zero_()); print(f'zero  {gb} GB: {t:.3f} ms = {n*8/t/1e6:.0f} GB/s written')
    t=tm(lambda: b.copy_(a)); print(f'copy  {gb} GB: {t:.3f} ms = {2*n*8/t/1e6:.0f} GB/s read+write')
    t=tm(lambda: a.sum()); print(f'read  {gb} GB: {t:.3f} ms = {n*8/t/1e6:.0f} GB/s read')
    del a,b
