import sys, threading; sys.path.insert(0, '.')
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 416, 20
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1, 16, 20, S, C)
def work():
    torch.cuda.set_device(0)
    y = engine.encode_targets(boxes, (S, S), anchors, C)
    assert y[0].shape[0] == 16
free0 = torch.cuda.mem_get_info()[0]
for i in range(40):
    t = threading.Thread(target=work); t.start(); t.join()
    if i == 4: free5 = torch.cuda.mem_get_info()[0]
free1 = torch.cuda.mem_get_info()[0]
print("free MB before", free0 >> 20, "after 5 threads", free5 >> 20, "after 40 threads", free1 >> 20)
assert free5 - free1 < (64 << 20), "staging of exited threads leaked"
print("ok")
