import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C, B = 608, 80, 512
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, B, 100, S, C)
y = [torch.empty((B, g, g, 88), dtype=torch.float32).pin_memory().numpy() for g in (19, 38, 76)]
for Bq in (512, 64, 128, 64, 512, 300):
    yq = [a[:Bq] for a in y]
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); engine.encode_targets(boxes[:Bq], (S, S), anchors, C, out=yq); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"encode host B={Bq}: " + " ".join(f"{t:.1f}" for t in ts) + " ms")
