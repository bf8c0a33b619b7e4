"""torchrun --nproc-per-node N scripts/e2e_multi_probe.py: host-buffer path under N-way
contention for the host interface: bare copies vs encode alone / decode alone / both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from concurrent.futures import ThreadPoolExecutor
from multigriddet_b200 import engine, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, C, B = 608, 80, 512
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(1 + rank, B, 100, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = [p.cpu().pin_memory().numpy() for p in synth.planted_head_outputs(yt, 3, 1)]
y = [torch.empty((B, g, g, 88), dtype=torch.float32).pin_memory() for g in (19, 38, 76)]
ny = [t.numpy() for t in y]
hw = synth.image_shapes(0, B)
d_in = [torch.empty(p.shape, dtype=torch.float32, device="cuda") for p in preds]
d_out = [torch.empty_like(t, device="cuda") for t in y]
tp = [torch.from_numpy(p) for p in preds]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
kw = dict(confidence=0.001, nms_threshold=0.45, want=("boxes_xyxy", "scores", "classes"))
def up():
    with torch.cuda.stream(s_up):
        for d, h in zip(d_in, tp): d.copy_(h, non_blocking=True)
def dn():
    with torch.cuda.stream(s_dn):
        for h, d in zip(y, d_out): h.copy_(d, non_blocking=True)
enc = lambda: engine.encode_targets(boxes, (S, S), anchors, C, out=ny)
dec = lambda: engine.decode_nms(preds, hw, (S, S), anchors, C, **kw)
pool = ThreadPoolExecutor(2)
def both():
    a, b = pool.submit(enc), pool.submit(dec); a.result(); b.result()
def measure(name, fn, n=4):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"[{world} ranks] {name}: {dt.item()*1e3:.1f} ms per 512 images per rank", flush=True)
measure("bare H2D", up)
measure("bare D2H", dn)
measure("bare H2D || D2H", lambda: (up(), dn()))
measure("encode (host)", enc)
measure("decode (host)", dec)
measure("encode || decode", both)
def both_seq_enqueue():
    # same two calls, but the decode call is issued ~1 ms after the encode call has its copies queued
    a = pool.submit(enc); time.sleep(0.002); b = pool.submit(dec); a.result(); b.result()
measure("encode, then decode 2 ms later", both_seq_enqueue)
def both_rev():
    b = pool.submit(dec); time.sleep(0.002); a = pool.submit(enc); a.result(); b.result()
measure("decode, then encode 2 ms later", both_rev)
dist.barrier()
if rank == 0:
    os.environ["MGD_TRACE"] = "1"
both()
dist.barrier()
dist.destroy_process_group()
