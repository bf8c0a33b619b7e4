import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C, B = 608, 80, 256
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(3, B, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = synth.planted_head_outputs(yt, 3, 1)
hw = torch.from_numpy(synth.image_shapes(0, B, mixed=True)).cuda()
for method in ("diou", "soft", "wbf"):
    kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method=method)
    out = engine.decode_nms(preds, hw, (S, S), anchors, C, return_stats=True, **kw)
    engine.profile_begin()
    for _ in range(3): engine.decode_nms(preds, hw, (S, S), anchors, C, sync=False, **kw)
    torch.cuda.synchronize()
    pr = engine.profile_end()
    print(f"{method}: decode {pr['decode_compact'][0]/3:.3f} ms, nms-stage {pr['nms'][0]/3:.3f} ms per {B} images; candidates/img {out['stats']['n_candidates']/B:.0f}, detections/img {out['stats']['n_detections']/B:.0f}")
