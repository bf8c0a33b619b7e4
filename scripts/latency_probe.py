"""Small-batch latency of the drop-in calls (the reference's own call sizes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from multigriddet_b200.data import preprocess_true_boxes
from multigriddet_b200.postprocess import MultiGridDecoder
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
boxes64 = synth.synth_boxes(1, 64, 100, S, C)
d64 = torch.from_numpy(boxes64).cuda()
yt = engine.encode_targets(d64, (S, S), anchors, C)
preds_dev = synth.planted_head_outputs(yt, 3, 1)
dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
one_host = [p[:1].cpu().numpy() for p in preds_dev]
one_dev = [p[:1].contiguous() for p in preds_dev]
b256 = [torch.cat([p] * 4) for p in preds_dev]
print("encode B=64 host numpy -> numpy   : %.3f ms" % t(lambda: preprocess_true_boxes(boxes64, (S, S), anchors, C, False)))
print("encode B=64 device tensors (sync) : %.3f ms" % t(lambda: engine.encode_targets(d64, (S, S), anchors, C)))
out = [torch.empty_like(y) for y in yt]
print("encode B=64 device, preallocated  : %.3f ms" % t(lambda: engine.encode_targets(d64, (S, S), anchors, C, out=out)))
print("postprocess 1 image host arrays   : %.3f ms" % t(lambda: dec.postprocess(one_host, (480, 640), (S, S), confidence=0.001, nms_threshold=0.45)))
print("decode_nms 1 image device (sync)  : %.3f ms" % t(lambda: engine.decode_nms(one_dev, (480, 640), (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)))
print("decode_nms B=256 device (sync)    : %.3f ms" % t(lambda: engine.decode_nms(b256, None, (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)))
one_pin = [torch.from_numpy(p).pin_memory().numpy() for p in one_host]
print("decode_nms 1 image host pageable  : %.3f ms" % t(lambda: engine.decode_nms(one_host, (480, 640), (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)))
print("decode_nms 1 image host pinned    : %.3f ms" % t(lambda: engine.decode_nms(one_pin, (480, 640), (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)))
b8 = [np.concatenate([p] * 8) for p in one_host]
print("decode_nms 8 images host pageable : %.3f ms" % t(lambda: engine.decode_nms(b8, (480, 640), (S, S), anchors, C, confidence=0.001, nms_threshold=0.45)))
import ctypes
lib = engine._lib.load()
print("mgd_device_count call             : %.4f ms" % t(lambda: lib.mgd_device_count(), 200))
