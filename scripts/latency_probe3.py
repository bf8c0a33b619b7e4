import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, 4, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
one = [p[:1].cpu().numpy() for p in synth.planted_head_outputs(yt, 3, 1)]
for i in range(4):
    t0 = time.perf_counter()
    engine.decode_nms(one, (480, 640), (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)
    print("call %d: %.3f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
