"""Per-image latency of the evaluator-style call (one image per postprocess call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from multigriddet_b200.postprocess import MultiGridDecoder
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(4, 4, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
d_preds = synth.planted_head_outputs(yt, 3, 4)
one_pg = [p[:1].cpu().numpy() for p in d_preds]
one_pin = [p[:1].cpu().pin_memory().numpy() for p in d_preds]
one_dev = [p[:1].contiguous() for p in d_preds]
dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
kw = dict(confidence=0.001, nms_threshold=0.45)
def timeit(name, fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter()-t0)/n*1e3:.3f} ms")
timeit("MultiGridDecoder.postprocess pageable numpy", lambda: dec.postprocess(one_pg, (480, 640), (S, S), **kw))
timeit("MultiGridDecoder.postprocess pinned numpy", lambda: dec.postprocess(one_pin, (480, 640), (S, S), **kw))
hw = np.array([[480, 640]], np.int32)
timeit("engine.decode_nms pageable numpy", lambda: engine.decode_nms(one_pg, hw, (S, S), anchors, C, max_boxes=100, **kw))
timeit("engine.decode_nms pinned numpy", lambda: engine.decode_nms(one_pin, hw, (S, S), anchors, C, max_boxes=100, **kw))
d_hw = torch.from_numpy(hw).cuda()
timeit("engine.decode_nms device tensors (sync)", lambda: engine.decode_nms(one_dev, d_hw, (S, S), anchors, C, max_boxes=100, **kw))
timeit("engine.decode_nms device tensors (async enqueue)", lambda: engine.decode_nms(one_dev, d_hw, (S, S), anchors, C, max_boxes=100, sync=False, **kw))
os.environ["MGD_TRACE"] = "1"
engine.decode_nms(one_pg, hw, (S, S), anchors, C, max_boxes=100, **kw)
