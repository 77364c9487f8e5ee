import torch, time
dev = torch.device("cuda", 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
af = a.view(torch.float32)
def t(fn, iters=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
ms = t(lambda: a.zero_()); print(f"write-only (memset 1 GiB): {n/ms/1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy 1 GiB (read+write)  : {2*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: af.sum()); print(f"read-only (sum 1 GiB fp32): {n/ms/1e6:8.1f} GB/s")
for mb in (42, 126, 168, 336):
    m = mb << 20
    ms = t(lambda: a[:m].zero_()); print(f"memset {mb:4d} MB: {ms*1e3:7.1f} us {m/ms/1e6:8.1f} GB/s")
    ms = t(lambda: b[:m].copy_(a[:m])); print(f"copy   {mb:4d} MB: {ms*1e3:7.1f} us {2*m/ms/1e6:8.1f} GB/s")
