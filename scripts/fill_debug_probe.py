import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C, B = 608, 80, 4096
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, 512, 100, S, C)
boxes = np.tile(boxes, (B // 512, 1, 1))
d_boxes = torch.from_numpy(boxes).cuda()
y = [torch.empty((B, g, g, 88), device="cuda") for g in (19, 38, 76)]
for _ in range(3): engine.encode_targets(d_boxes, (S, S), anchors, C, out=y, sync=False)
torch.cuda.synchronize()
engine.profile_begin()
for _ in range(10): engine.encode_targets(d_boxes, (S, S), anchors, C, out=y, sync=False)
torch.cuda.synchronize()
pr = engine.profile_end()
ms = pr["encode_fill"][0] / pr["encode_fill"][1]
print(os.environ.get("MGD_FILL_DEBUG", "0"), "fill ms per launch", round(ms, 4), "launches", pr["encode_fill"][1], "TB/s", round(B / 4 * 2670512 / ms / 1e9, 3) if pr["encode_fill"][1] == 40 else "")
