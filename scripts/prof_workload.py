"""Small single-launch workload for ncu: one encode + one decode/NMS call on device tensors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, min(B, 256), 100, S, C)
boxes = np.tile(boxes, ((B + 255) // 256, 1, 1))[:B]
d_boxes = torch.from_numpy(boxes).cuda()
yt = engine.encode_targets(d_boxes, (S, S), anchors, C)
preds = synth.planted_head_outputs(yt, 3, seed=1)
hw = torch.from_numpy(synth.image_shapes(0, B)).cuda()
torch.cuda.synchronize()
for it in range(3):
    engine.encode_targets(d_boxes, (S, S), anchors, C, out=yt)
    det = engine.decode_nms(preds, hw, (S, S), anchors, C, max_boxes=100, confidence=0.001, nms_threshold=0.45)
torch.cuda.synchronize()
print("ok", int(det["counts"].sum()))
