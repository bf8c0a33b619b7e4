import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from multigriddet_b200 import engine, synth, _lib
S, C, B = 608, 80, 256
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(3, B, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = synth.planted_head_outputs(yt, 3, 1)
out = engine.decode_dense(preds, anchors, C, (S, S))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = engine.decode_dense(preds, anchors, C, (S, S))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byt = B * 7581 * (88 * 4 + 85 * 8)
print(f"decode_dense B={B}: {ms:.3f} ms = {B/ms*1e3:.0f} img/s, {byt/ms/1e6:.0f} GB/s (read 352 B + write 680 B per cell)", type(out), tuple(out.shape))
