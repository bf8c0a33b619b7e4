"""Reference points for the roofline: device memset and copy bandwidth on this GPU."""
import torch
x = torch.empty(3 * 1024**3, dtype=torch.float32, device="cuda")   # 12 GB
y = torch.empty_like(x)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n / 1e3
gb = x.numel() * 4 / 1e9
print("memset (zero_)   GB/s written:", round(gb / t(lambda: x.zero_())))
print("fill_(1.0)       GB/s written:", round(gb / t(lambda: x.fill_(1.0))))
print("copy_ y<-x       GB/s (r+w)  :", round(2 * gb / t(lambda: y.copy_(x))))
print("sum (read only)  GB/s read   :", round(gb / t(lambda: x.sum())))
