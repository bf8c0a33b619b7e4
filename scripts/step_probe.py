"""Step timing: separate calls on one stream vs. the fused mgd_encode_decode_nms entry."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from multigriddet_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
anchors, boxes_np, d_boxes, preds = bench.make_device_inputs(B, dev, seed=1)
S, C, D = bench.S, bench.C, bench.D
y_out = [torch.empty((B, g, g, D), dtype=torch.float32, device=dev) for g in (19, 38, 76)]
d_hw = torch.from_numpy(synth.image_shapes(0, B)).to(dev)
POST = bench.POST
keep = []
def sep():
    engine.encode_targets(d_boxes, (S, S), anchors, C, out=y_out, sync=False)
    keep[:] = [engine.decode_nms(preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **POST)]
def fused():
    keep[:] = [engine.grid_step(d_boxes, y_out, preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **POST)]
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {"batch": B, "separate_ms": timeit(sep), "fused_ms": timeit(fused)}
engine.profile_begin()
res["fused_profiled_ms"] = timeit(fused)
res["spans"] = engine.profile_end()
engine.profile_begin()
res["separate_profiled_ms"] = timeit(sep)
engine.profile_end()
res["fused_again_ms"] = timeit(fused)
ref = engine.decode_nms(preds, d_hw, (S, S), anchors, C, want=("boxes_xyxy", "scores", "classes"), **POST)
got = engine.grid_step(d_boxes, y_out, preds, d_hw, (S, S), anchors, C, want=("boxes_xyxy", "scores", "classes"), **POST)
y_ref = engine.encode_targets(d_boxes, (S, S), anchors, C)
res["same"] = bool(all(torch.equal(ref[k], got[k]) for k in ("boxes_xyxy", "scores", "classes", "counts")) and all(torch.equal(a, b) for a, b in zip(y_ref, y_out)))
print(json.dumps(res))
