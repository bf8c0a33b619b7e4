"""Decode + NMS time against the number of candidates per image (background objectness raised)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C, B = 608, 80, int(os.environ.get("B", 256))
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(3, B, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
hw = torch.from_numpy(synth.image_shapes(0, B, mixed=True)).cuda()
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
for neg in (-8.0, -6.0, -5.0, -4.0, -3.0, -1.0):
    preds = synth.planted_head_outputs(yt, 3, 1, obj_neg=(neg, 1.5))
    out = engine.decode_nms(preds, hw, (S, S), anchors, C, return_stats=True, **kw)
    engine.profile_begin()
    for _ in range(3): engine.decode_nms(preds, hw, (S, S), anchors, C, sync=False, **kw)
    torch.cuda.synchronize()
    pr = engine.profile_end()
    print(f"B={B} obj_neg mean {neg:+.0f}: candidates/img {out['stats']['n_candidates']/B:6.0f}  decode {pr['decode_compact'][0]/3:.3f} ms  nms {pr['nms'][0]/3:.3f} ms")
