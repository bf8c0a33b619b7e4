import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(3, 2, 100, S, C)
y = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = [p.cpu().numpy()[:1].copy() for p in synth.planted_head_outputs(y, 3, seed=2)]
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45)
for _ in range(20): engine.decode_nms(preds, [(480, 640)], (S, S), anchors, C, **kw)
t0 = time.perf_counter()
for _ in range(200): engine.decode_nms(preds, [(480, 640)], (S, S), anchors, C, **kw)
print("pageable ms/call", (time.perf_counter() - t0) / 200 * 1e3)
from multigriddet_b200 import _lib
pin = [_lib.pinned.empty(p.shape, np.float32) for p in preds]
for a, b in zip(pin, preds): a[...] = b
for _ in range(20): engine.decode_nms(pin, [(480, 640)], (S, S), anchors, C, **kw)
t0 = time.perf_counter()
for _ in range(200): engine.decode_nms(pin, [(480, 640)], (S, S), anchors, C, **kw)
print("pinned ms/call", (time.perf_counter() - t0) / 200 * 1e3)
t0 = time.perf_counter()
for _ in range(200): engine.decode_nms(pin, [(480, 640)], (S, S), anchors, C, zerocopy=True, **kw)
print("pinned zerocopy ms/call", (time.perf_counter() - t0) / 200 * 1e3)
d = [torch.from_numpy(p).cuda() for p in preds]
hw = torch.tensor([[480, 640]], dtype=torch.int32).cuda()
for _ in range(20): engine.decode_nms(d, hw, (S, S), anchors, C, **kw)
t0 = time.perf_counter()
for _ in range(200): engine.decode_nms(d, hw, (S, S), anchors, C, **kw)
print("device ms/call", (time.perf_counter() - t0) / 200 * 1e3)
os.environ["MGD_TRACE"] = "1"
