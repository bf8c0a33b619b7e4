"""Decode + NMS on a dense random head (every cell a candidate: SURVEY 8d (ii))."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
for B in (16, 256):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    preds = [torch.randn((B, s, s, 88), device="cuda", generator=g) for s in (19, 38, 76)]
    hw = torch.from_numpy(synth.image_shapes(0, B, mixed=True)).cuda()
    for method in ("diou", "standard"):
        kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method=method)
        out = engine.decode_nms(preds, hw, (S, S), anchors, C, return_stats=True, **kw)
        engine.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): engine.decode_nms(preds, hw, (S, S), anchors, C, sync=False, **kw)
        e1.record(); torch.cuda.synchronize()
        pr = engine.profile_end()
        ms = e0.elapsed_time(e1) / 3
        print(f"dense random B={B} {method}: {ms:.2f} ms = {B/ms*1e3:.0f} img/s (decode {pr['decode_compact'][0]/3:.2f} ms, nms {pr['nms'][0]/3:.2f} ms); "
              f"candidates/img {out['stats']['n_candidates']/B:.0f}, detections/img {out['stats']['n_detections']/B:.0f}")
