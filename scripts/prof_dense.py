"""Single-launch dense-random decode workload for ncu (every row takes the exact path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
preds = synth.dense_random_head_outputs(B, S, 3, C, seed=5, device="cuda")
hw = torch.from_numpy(synth.image_shapes(1, B)).cuda()
torch.cuda.synchronize()
for it in range(3):
    det = engine.decode_nms(preds, hw, (S, S), anchors, C, max_boxes=100, confidence=0.001, nms_threshold=0.45)
torch.cuda.synchronize()
print("ok", int(det["counts"].sum()))
