"""Host time per fused call and device-memory growth over the first calls."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from multigriddet_b200 import engine, synth
B = 4096
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
anchors, boxes_np, d_boxes, preds = bench.make_device_inputs(B, dev, seed=1)
S, C, D = bench.S, bench.C, bench.D
y_out = [torch.empty((B, g, g, D), dtype=torch.float32, device=dev) for g in (19, 38, 76)]
d_hw = torch.from_numpy(synth.image_shapes(0, B)).to(dev)
keep = []
mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
def fused():
    keep[:] = [engine.grid_step(d_boxes, y_out, preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **bench.POST)]
def sep():
    engine.encode_targets(d_boxes, (S, S), anchors, C, out=y_out, sync=False)
    keep[:] = [engine.decode_nms(preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **bench.POST)]
fn = fused if mode == "fused" else sep
torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
rows = []
for i in range(30):
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter()
    rows.append((round((t1 - t0) * 1e3, 2), round((free0 - torch.cuda.mem_get_info()[0]) / 2**30, 2)))
torch.cuda.synchronize()
print(mode, json.dumps(rows))
