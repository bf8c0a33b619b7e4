import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
for B in (64, 1024):
    boxes = synth.synth_boxes(3, B, 100, S, C)
    yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
    yp = synth.planted_head_outputs(yt, 3, 1)
    out = engine.ignore_masks(yp, yt, anchors, (S, S), C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = engine.ignore_masks(yp, yt, anchors, (S, S), C, sync=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    pos = sum(int((t[..., 4] > 0.5).sum()) for t in yt)
    print(f"ignore mask B={B}: {ms:.3f} ms = {B/ms*1e3:.0f} img/s; positive cells/img {pos/B:.0f}; ignored cells/img {sum(float(o[0].sum()) for o in out)/B:.0f}")
