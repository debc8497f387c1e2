import torch
dev=torch.device("cuda:0")
for mb in (110, 440, 1024):
    n = mb*1024*1024//4
    bufs=[torch.empty(n, device=dev) for _ in range(3)]
    for b in bufs: b.fill_(1.0)
    torch.cuda.synchronize()
    a,b_=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(30): bufs[r%3].fill_(float(r))
    b_.record(); torch.cuda.synchronize()
    us=1e3*a.elapsed_time(b_)/30
    print("fill %4d MB: %8.2f us  %7.1f GB/s" % (mb, us, n*4/us/1e3))
    src=torch.empty(n, device=dev)
    a.record()
    for r in range(30): bufs[r%3].copy_(src)
    b_.record(); torch.cuda.synchronize()
    us=1e3*a.elapsed_time(b_)/30
    print("copy %4d MB: %8.2f us  %7.1f GB/s (r+w)" % (mb, us, 2*n*4/us/1e3))
    del bufs, src
