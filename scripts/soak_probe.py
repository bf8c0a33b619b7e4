"""Leak check: thousands of mixed calls; device memory and host RSS must stay flat."""
import os, sys, resource
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from multigriddet_b200.postprocess import MultiGridDecoder
from multigriddet_b200.data import preprocess_true_boxes
S, C, B = 416, 20, 8
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, B, 20, S, C)
d_boxes = torch.from_numpy(boxes).cuda()
yt = engine.encode_targets(d_boxes, (S, S), anchors, C)
d_preds = synth.planted_head_outputs(yt, 3, 1)
h_preds = [p.cpu().numpy() for p in d_preds]
dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
def rss(): return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss // 1024
def snap(tag):
    torch.cuda.synchronize()
    print(f"{tag}: device free {torch.cuda.mem_get_info()[0] >> 20} MB, host max RSS {rss()} MB", flush=True)
def round_():
    for _ in range(500):
        engine.encode_targets(d_boxes, (S, S), anchors, C, sync=False)
        engine.decode_nms(d_preds, None, (S, S), anchors, C, sync=False, confidence=0.05)
        preprocess_true_boxes(boxes, (S, S), anchors, C, False)
        dec.postprocess([p[:1] for p in h_preds], (375, 500), (S, S), confidence=0.05)
        engine.decode_nms(h_preds, None, (S, S), anchors, C, nms_method="soft", confidence=0.05)
    engine.poll_status(0)
round_(); snap("after 500 iterations")
f0 = torch.cuda.mem_get_info()[0]; r0 = rss()
for i in range(3): round_()
snap("after 2000 iterations")
assert f0 - torch.cuda.mem_get_info()[0] < (64 << 20), "device memory grew"
assert rss() - r0 < 200, "host memory grew"
print("ok")
