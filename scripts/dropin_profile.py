"""Where one MultiGridDecoder.postprocess call (pageable NumPy in / out, one image: the
evaluator's usage, evaluator.py:254-289) spends its host time."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from multigriddet_b200.postprocess import MultiGridDecoder

S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(3, 8, 100, S, C)
y = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = [p.cpu().numpy() for p in synth.planted_head_outputs(y, 3, seed=2)]
dec = MultiGridDecoder(anchors, C, (S, S))
imgs = [[p[i:i + 1].copy() for p in preds] for i in range(8)]
shapes = synth.image_shapes(0, 8)
def one(i):
    return dec.postprocess(imgs[i % 8], tuple(int(v) for v in shapes[i % 8]), (S, S), max_boxes=100,
                           confidence=0.001, nms_threshold=0.45)
for i in range(50): one(i)
t0 = time.perf_counter()
for i in range(400): one(i)
print("ms per image:", (time.perf_counter() - t0) / 400 * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(400): one(i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
