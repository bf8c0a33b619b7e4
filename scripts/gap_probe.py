import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
import bench
B = 4096
anchors, boxes_np, d_boxes, preds = bench.make_device_inputs(B, torch.device("cuda", 0), seed=1)
y_out = [torch.empty((B, g, g, 88), dtype=torch.float32, device="cuda") for g in (19, 38, 76)]
d_hw = torch.from_numpy(synth.image_shapes(0, B, mixed=True)).cuda()
S, C = 608, 80
def enc(): engine.encode_targets(d_boxes, (S, S), anchors, C, out=y_out, sync=False)
def dec(): return engine.decode_nms(preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **bench.POST)
def timeit(name, fn, n=10, prof=False):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if prof: engine.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    p = engine.profile_end() if prof else None
    print(f"{name}: gpu {e0.elapsed_time(e1)/n:.3f} ms/iter, host enqueue {th/n*1e3:.3f} ms/iter", {k: round(v[0]/n, 3) for k, v in p.items()} if p else "")
timeit("encode", enc)
timeit("decode", dec)
timeit("decode prof", dec, prof=True)
timeit("encode+decode", lambda: (enc(), dec()))
timeit("encode+decode prof", lambda: (enc(), dec()), prof=True)
