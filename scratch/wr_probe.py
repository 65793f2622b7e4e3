import torch, time
x = torch.empty(1301*1024*1024//4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
nb = x.numel()*4
for name, fn, bytes_ in [("fill_", lambda: x.fill_(1.0), nb), ("zero_", lambda: x.zero_(), nb), ("copy_", lambda: y.copy_(x), 2*nb)]:
    ms = t(fn); print(f"{name}: {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s")
