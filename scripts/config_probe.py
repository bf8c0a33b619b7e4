"""BASELINE.json configs 2-4 through the public entry points: pageable NumPy, pinned NumPy
and device tensors.  Prints images/s (these configs are parity-test cases, not bench lines)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from multigriddet_b200.data import preprocess_true_boxes
from multigriddet_b200.postprocess import MultiGridDecoder

def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
# config 2: training-target encoding, batch 64, <= 100 boxes
B = 64
boxes = synth.synth_boxes(3, B, 100, S, C)
d_boxes = torch.from_numpy(boxes).cuda()
pin_y = [torch.empty((B, g, g, 88), dtype=torch.float32).pin_memory().numpy() for g in (19, 38, 76)]
dev_y = [torch.empty((B, g, g, 88), dtype=torch.float32, device="cuda") for g in (19, 38, 76)]
t = timeit(lambda: preprocess_true_boxes(boxes, (S, S), anchors, C, False))
print(f"config2 encode B=64 numpy in -> fresh numpy out (drop-in): {t*1e3:.2f} ms = {B/t:.0f} img/s")
t = timeit(lambda: engine.encode_targets(boxes, (S, S), anchors, C, out=pin_y))
print(f"config2 encode B=64 numpy in -> pinned out: {t*1e3:.2f} ms = {B/t:.0f} img/s")
t = timeit(lambda: engine.encode_targets(d_boxes, (S, S), anchors, C, out=dev_y, sync=False), n=50)
print(f"config2 encode B=64 device tensors: {t*1e3:.3f} ms = {B/t:.0f} img/s")
# config 3: eval decode + NMS, batch 256, conf 0.001
B = 256
boxes = synth.synth_boxes(4, B, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
d_preds = synth.planted_head_outputs(yt, 3, 4)
pg_preds = [p.cpu().numpy() for p in d_preds]
pin_preds = [p.cpu().pin_memory().numpy() for p in d_preds]
hw = synth.image_shapes(0, B, mixed=True)
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
for per_class in (False, True):
    t = timeit(lambda: engine.decode_nms(pg_preds, hw, (S, S), anchors, C, per_class=per_class, **kw))
    print(f"config3 decode+NMS B=256 per_class={per_class} pageable numpy: {t*1e3:.2f} ms = {B/t:.0f} img/s")
    t = timeit(lambda: engine.decode_nms(pin_preds, hw, (S, S), anchors, C, per_class=per_class, **kw))
    print(f"config3 decode+NMS B=256 per_class={per_class} pinned numpy: {t*1e3:.2f} ms = {B/t:.0f} img/s")
    t = timeit(lambda: engine.decode_nms(d_preds, torch.from_numpy(hw).cuda(), (S, S), anchors, C, per_class=per_class, sync=False, **kw), n=50)
    print(f"config3 decode+NMS B=256 per_class={per_class} device tensors: {t*1e3:.3f} ms = {B/t:.0f} img/s")
dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
one = [p[:1] for p in pg_preds]
t = timeit(lambda: dec.postprocess(one, (480, 640), (S, S), confidence=0.001, nms_threshold=0.45), n=50)
print(f"config3 MultiGridDecoder.postprocess, one image per call (the evaluator's usage): {t*1e3:.3f} ms/image")
# config 4: multi-scale shapes, mosaic-dense labels (<= 300 boxes)
for S4 in (320, 416, 512, 608):
    B = 64
    boxes = synth.synth_boxes(5, B, 300, S4, C, layout="mosaic", corners="frac")
    d = torch.from_numpy(boxes).cuda()
    out, st = engine.encode_targets(d, (S4, S4), anchors, C, return_stats=True)
    t = timeit(lambda: engine.encode_targets(d, (S4, S4), anchors, C, out=out, sync=False), n=50)
    print(f"config4 encode S={S4} B=64 N<=300 mosaic: {t*1e3:.3f} ms = {B/t:.0f} img/s; skipped writes {st['n_skipped_writes']}, positive cells {st['n_positive_cells']}")
