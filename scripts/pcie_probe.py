"""PCIe rates vs. the host-buffer C-ABI calls, staged vs. zero-copy (MGD_HOST_ZEROCOPY)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for mode, mb in (("0", 128), ("0", 64), ("0", 32), ("0", 16)):
        env = dict(os.environ, MGD_HOST_ZEROCOPY=mode if mode != "raw" else "0", MGD_HOST_CHUNK_MB=str(mb))
        subprocess.run([sys.executable, __file__, mode], env=env)
    sys.exit(0)
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from multigriddet_b200 import engine, synth
mode = sys.argv[1]
S, C, B = 608, 80, 512
if mode == "raw":
    x = torch.empty(1 << 28, dtype=torch.float32).pin_memory(); d = torch.empty_like(x, device="cuda")
    x2 = torch.empty(1 << 28, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(d)
    def rate(fn, nbytes, n=4):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return nbytes * n / (time.perf_counter() - t0) / 1e9
    print("H2D pinned GB/s", round(rate(lambda: d.copy_(x, non_blocking=True), x.numel() * 4), 1))
    print("D2H pinned GB/s", round(rate(lambda: x.copy_(d, non_blocking=True), x.numel() * 4), 1))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
        with torch.cuda.stream(s2): x2.copy_(d2, non_blocking=True)
    print("H2D+D2H concurrent GB/s (sum)", round(rate(both, 2 * x.numel() * 4), 1), flush=True)
    sys.exit(0)
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, B, 100, S, C)
y = [torch.empty((B, g, g, 88), dtype=torch.float32).pin_memory().numpy() for g in (19, 38, 76)]
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
ref_y = [t.cpu().numpy() for t in yt]
preds = [p.cpu().pin_memory().numpy() for p in synth.planted_head_outputs(yt, 3, 1)]
dense = [torch.randn((B, g, g, 88)).pin_memory().numpy() for g in (19, 38, 76)]
hw = synth.image_shapes(0, B)
kw = dict(confidence=0.001, nms_threshold=0.45, want=("boxes_xyxy", "scores", "classes"))
enc = lambda: engine.encode_targets(boxes, (S, S), anchors, C, out=y)
dec = lambda: engine.decode_nms(preds, hw, (S, S), anchors, C, **kw)
decd = lambda: engine.decode_nms(dense, hw, (S, S), anchors, C, **kw)
pool = ThreadPoolExecutor(2)
def both():
    a, b = pool.submit(enc), pool.submit(dec); a.result(); return b.result()
def timeit(name, fn, n=4):
    fn(); fn(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    dt = (time.perf_counter() - t0) / n
    print(f"[zerocopy={mode} chunk={os.environ.get('MGD_HOST_CHUNK_MB')}MB] {name}: {dt*1e3:.1f} ms per {B} images = {B/dt:.0f} img/s", flush=True)
    return r
timeit("encode host", enc)
assert all(np.array_equal(a, b) for a, b in zip(y, ref_y)), "encode host output differs from device output"
r = timeit("decode host (planted)", dec)
print("   detections", int(r["counts"].sum()))
timeit("decode host (dense random, every cell a candidate)", decd, n=2)
timeit("encode || decode", both)

