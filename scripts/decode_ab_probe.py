"""Decode kernel A/B: time mgd_decode_nms' decode_compact span for the planted and the
dense-random workload (env MGD_DECODE_IMPL / MGD_DECODE_PREFETCH select the kernel).
usage: python scripts/decode_ab_probe.py [batch] [dense_batch]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multigriddet_b200 import engine, synth, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
BD = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S, C = 608, 80
anchors = synth.coco_anchors(np.float32)
torch.cuda.set_device(0)
if os.environ.get("L2_FETCH"):
    import ctypes, glob
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL(sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[0])
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2_FETCH"])))   # cudaLimitMaxL2FetchGranularity
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("MaxL2FetchGranularity rc", rc, "now", v.value, file=sys.stderr)
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")

def timed(preds, shapes, n=5):
    for _ in range(2):
        engine.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
    torch.cuda.synchronize()
    engine.profile_begin()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        det = engine.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
    e1.record(); torch.cuda.synchronize()
    spans = engine.profile_end()
    return e0.elapsed_time(e1) / n, spans, det

out = {"impl": os.environ.get("MGD_DECODE_IMPL", "ws"), "shape": os.environ.get("MGD_DECODE_SHAPE", "0")}
chunk = 512
boxes = synth.synth_boxes(11, chunk, 100, S, C)
preds = [torch.empty((B, g, g, 88), device="cuda") for g in (19, 38, 76)]
for c0 in range(0, B, chunk):
    yt = engine.encode_targets(torch.from_numpy(np.roll(boxes, c0, 0)).cuda(), (S, S), anchors, C)
    pp = synth.planted_head_outputs(yt, 3, seed=3 + c0)
    for p, q in zip(preds, pp):
        p[c0:c0 + chunk] = q[: min(chunk, B - c0)]
    del yt, pp
shapes = torch.from_numpy(synth.image_shapes(0, B)).cuda()
ms, spans, det = timed(preds, shapes)
out["planted"] = {"batch": B, "ms_total": ms, "spans": spans,
                  "dets": int(det["counts"].sum().item())}
del preds
dp = synth.dense_random_head_outputs(BD, S, 3, C, seed=5, device="cuda")
ms, spans, det = timed(dp, torch.from_numpy(synth.image_shapes(1, BD)).cuda())
out["dense"] = {"batch": BD, "ms_total": ms, "spans": spans, "dets": int(det["counts"].sum().item())}
print(json.dumps(out))
