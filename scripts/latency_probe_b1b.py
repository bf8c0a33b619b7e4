import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
NB = int(os.environ.get('NB', 64))
boxes = synth.synth_boxes(4, min(NB, 512), 100, S, C)
boxes = np.tile(boxes, ((NB + 511) // 512, 1, 1))[:NB]
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
parts = [synth.planted_head_outputs([y[i:i+256] for y in yt], 3, 4 + i) for i in range(0, NB, 256)]
d_preds = [torch.cat([p[j] for p in parts]) for j in range(3)]
kw = dict(confidence=0.001, nms_threshold=0.45, max_boxes=100)
for B in [int(v) for v in os.environ.get('BS', '1,4,16,64').split(',')]:
    p = [t[:B].contiguous() for t in d_preds]
    hw = torch.from_numpy(synth.image_shapes(0, B, mixed=True)).cuda()
    for _ in range(10): engine.decode_nms(p, hw, (S, S), anchors, C, sync=False, **kw)
    torch.cuda.synchronize()
    engine.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): engine.decode_nms(p, hw, (S, S), anchors, C, sync=False, **kw)
    e1.record(); torch.cuda.synchronize()
    pr = engine.profile_end()
    t0 = time.perf_counter()
    for _ in range(50): engine.decode_nms(p, hw, (S, S), anchors, C, sync=True, **kw)
    lat = (time.perf_counter() - t0) / 50
    print(f"B={B}: gpu {e0.elapsed_time(e1)/20*1e3:.0f} us/call (decode {pr['decode_compact'][0]/20*1e3:.0f}, nms {pr['nms'][0]/20*1e3:.0f}); synchronous call latency {lat*1e3:.3f} ms")
