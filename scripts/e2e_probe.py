"""Where does the host-buffer (e2e) path spend its time?  PCIe rates vs. the C-ABI calls."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C, B = 608, 80, 512
anchors = synth.coco_anchors(np.float32)
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory()      # 1 GiB pinned
d = torch.empty_like(x, device="cuda")
def rate(fn, nbytes, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return nbytes * n / (time.perf_counter() - t0) / 1e9
print("H2D pinned GB/s", round(rate(lambda: d.copy_(x, non_blocking=True), x.numel() * 4), 1))
print("D2H pinned GB/s", round(rate(lambda: x.copy_(d, non_blocking=True), x.numel() * 4), 1))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d); x2 = torch.empty(1 << 28, dtype=torch.float32).pin_memory()
def both():
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrent GB/s (sum)", round(rate(both, 2 * x.numel() * 4), 1))
boxes = synth.synth_boxes(1, B, 100, S, C)
y = [torch.empty((B, g, g, 88), dtype=torch.float32).pin_memory().numpy() for g in (19, 38, 76)]
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = [p.cpu().pin_memory().numpy() for p in synth.planted_head_outputs(yt, 3, 1)]
hw = synth.image_shapes(0, B)
for name, fn in (("encode host->host", lambda: engine.encode_targets(boxes, (S, S), anchors, C, out=y)),
                 ("decode host->host", lambda: engine.decode_nms(preds, hw, (S, S), anchors, C, confidence=0.001, nms_threshold=0.45))):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt*1e3:.1f} ms per {B} images = {B/dt:.0f} img/s = {B*2.67e6/dt/1e9:.1f} GB/s")
for Bq in (64, 128, 256, 512):
    yq = [a[:Bq] for a in y]
    t0 = time.perf_counter()
    for _ in range(3): engine.encode_targets(boxes[:Bq], (S, S), anchors, C, out=yq)
    dt = (time.perf_counter() - t0) / 3
    print(f"encode host B={Bq}: {dt*1e3:.2f} ms  -> {Bq*2.67e6/dt/1e9:.1f} GB/s")
