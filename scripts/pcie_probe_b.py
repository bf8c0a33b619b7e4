"""Chunked bidirectional copies with torch only: is ~74 GB/s (sum) inherent to chunking / threads?"""
import time, threading, torch
N = 1367 * 1000 * 1000 // 4
x = torch.empty(N, dtype=torch.float32).pin_memory(); d = torch.empty(N, dtype=torch.float32, device="cuda")
x2 = torch.empty(N, dtype=torch.float32).pin_memory(); d2 = torch.empty(N, dtype=torch.float32, device="cuda")
def chunks(mb):
    step = mb * (1 << 20) // 4
    return [(i, min(i + step, N)) for i in range(0, N, step)]
def h2d(mb, streams, kern):
    for k, (a, b) in enumerate(chunks(mb)):
        with torch.cuda.stream(streams[k & 1]):
            d[a:b].copy_(x[a:b], non_blocking=True)
            if kern: d[a:a + 1024].add_(1.0)
    for s in streams: s.synchronize()
def d2h(mb, streams, kern):
    for k, (a, b) in enumerate(chunks(mb)):
        with torch.cuda.stream(streams[k & 1]):
            if kern: d2[a:a + 1024].add_(1.0)
            x2[a:b].copy_(d2[a:b], non_blocking=True)
    for s in streams: s.synchronize()
sa = [torch.cuda.Stream(), torch.cuda.Stream()]; sb = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(mb, kern, threads, n=4):
    def once():
        if threads:
            t = threading.Thread(target=h2d, args=(mb, sa, kern)); t.start(); d2h(mb, sb, kern); t.join()
        else:
            cs = chunks(mb)
            for k, (a, b) in enumerate(cs):
                with torch.cuda.stream(sa[k & 1]): d[a:b].copy_(x[a:b], non_blocking=True)
                with torch.cuda.stream(sb[k & 1]): x2[a:b].copy_(d2[a:b], non_blocking=True)
            torch.cuda.synchronize()
    once(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): once()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"chunk {mb:5d} MB kern={kern} threads={threads}: {dt*1e3:.1f} ms, sum {2*N*4/dt/1e9:.1f} GB/s", flush=True)
for mb in (2048, 512, 128, 32):
    for kern in (False, True):
        for threads in (False, True):
            run(mb, kern, threads)
# one direction at a time for reference
for fn, name in ((h2d, "h2d"), (d2h, "d2h")):
    fn(128, sa, True); t0 = time.perf_counter()
    for _ in range(4): fn(128, sa, True)
    dt = (time.perf_counter() - t0) / 4
    print(f"{name} alone chunk 128: {dt*1e3:.1f} ms, {N*4/dt/1e9:.1f} GB/s")
