"""CUDA-graph capture of the library's device-memory calls (small batches are launch-bound):
capture mgd_encode_targets / mgd_decode_nms / mgd_encode_decode_nms once, replay, compare with
the eager calls and time both."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth

S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
torch.cuda.set_device(0)
res = {}
for B in [int(v) for v in (sys.argv[1:] or ["1", "64", "256"])]:
    boxes = torch.from_numpy(synth.synth_boxes(3, B, 100, S, C)).cuda()
    y = engine.encode_targets(boxes, (S, S), anchors, C)
    preds = synth.planted_head_outputs(y, 3, seed=2)
    hw = torch.from_numpy(synth.image_shapes(0, B)).cuda()
    y_out = [torch.empty_like(t) for t in y]
    ref = engine.decode_nms(preds, hw, (S, S), anchors, C, **kw)
    out = {k: torch.empty_like(v) for k, v in ref.items() if hasattr(v, "data_ptr")}

    def step():
        engine.grid_step(boxes, y_out, preds, hw, (S, S), anchors, C, sync=False, out=out, **kw)

    def timeit(f, n=200):
        for _ in range(10): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    step(); torch.cuda.synchronize()
    eager_us = timeit(step)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for k in out: out[k].zero_()
    for t in y_out: t.zero_()
    with torch.cuda.graph(g):
        step()
    g.replay(); torch.cuda.synchronize()
    same = all(torch.equal(out[k], ref[k]) for k in out) and all(torch.equal(a, b) for a, b in zip(y_out, y))
    graph_us = timeit(g.replay)
    res[B] = {"eager_us": round(eager_us, 1), "graph_us": round(graph_us, 1), "same": bool(same)}
print(json.dumps(res))
