"""Does NVML polling during the timed region slow the step down? (host-side driver contention)"""
import os, sys, json, time, threading
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
def fused():
    keep[:] = [engine.grid_step(d_boxes, y_out, preds, d_hw, (S, S), anchors, C, sync=False, want=("boxes_xyxy", "scores", "classes"), **bench.POST)]
def timeit(n=20):
    for _ in range(3): fused()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n): fused()
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3), round((t1 - t0) / n * 1e3, 3)
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
res = {"no_sampler": timeit()}
for name, interval, reasons in (("5ms_both", 0.005, True), ("5ms_clock_only", 0.005, False), ("25ms_both", 0.025, True)):
    stop = [False]; cost = []
    def run():
        while not stop[0]:
            t = time.perf_counter()
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if reasons:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            cost.append(time.perf_counter() - t)
            time.sleep(interval)
    th = threading.Thread(target=run, daemon=True); th.start()
    res[name] = timeit() + (round(1e3 * sum(cost) / max(len(cost), 1), 3),)
    stop[0] = True; th.join()
print(json.dumps(res))
