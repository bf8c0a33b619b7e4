"""BASELINE configs[0]: VOC 20-class 416x416, batch 8 -- CUDA path vs the CPU port, same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
from oracle import mgd_oracle as O
S, C, B, N = 416, 20, 8, 20
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(0, B, N, S, C)
kw = dict(max_boxes=100, confidence=0.1, nms_threshold=0.45, nms_method="diou")
t0 = time.perf_counter(); y_ref = O.encode_targets(boxes, (S, S), anchors, C); t_enc_cpu = time.perf_counter() - t0
preds = [p.numpy() for p in synth.planted_head_outputs([torch.from_numpy(a) for a in y_ref], 3, 0)]
hw = np.tile(np.array([[S, S]]), (B, 1))
t0 = time.perf_counter(); O.postprocess_batch(preds, hw, (S, S), anchors, C, **kw); t_dec_cpu = time.perf_counter() - t0
def best(fn, n=20):
    fn(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)
t_enc = best(lambda: engine.encode_targets(boxes, (S, S), anchors, C))
t_dec = best(lambda: engine.decode_nms(preds, hw, (S, S), anchors, C, **kw))
print(f"configs[0] VOC 416 B=8: encode CPU port {t_enc_cpu*1e3:.1f} ms vs CUDA (NumPy in/out) {t_enc*1e3:.3f} ms; "
      f"decode+NMS CPU port {t_dec_cpu*1e3:.1f} ms vs CUDA {t_dec*1e3:.3f} ms")
