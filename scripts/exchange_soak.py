"""Soak of the detection exchange: two processes (one GPU or two), hundreds of mirrored calls
with changing NMS methods and alternating entry points; every call's whole-batch result is
compared with a single-process decode, barrier timeouts must stay 0.
usage: python scripts/exchange_soak.py [iterations] [--two-gpus]"""
import os, sys, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def rank_main(rank, world, port, iters, two_gpus, q):
    import torch
    import torch.distributed as dist
    from multigriddet_b200 import engine, sharding, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank if two_gpus else 0)
    S, C, N, B = 416, 20, 12, 37
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(5, B, N, S, C)
    y = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
    preds = synth.planted_head_outputs(y, 3, seed=5)
    shapes = synth.image_shapes(5, B)
    hw = torch.from_numpy(shapes).cuda()
    keys = ("boxes_xywh", "boxes_xyxy", "scores", "classes", "index", "counts")
    ex = sharding.DetectionExchange(B, 50)
    path = sharding.ShardedGridPath(anchors, C, (S, S))
    lo, hi = sharding.shard_bounds(B, rank, world)
    y_out = [torch.empty_like(t[lo:hi]) for t in y]
    gt = torch.from_numpy(boxes[lo:hi]).cuda()
    refs = {}
    bad = 0
    methods = ["diou", "standard", "soft", "diou"]
    for it in range(iters):
        m = methods[it % 4]
        kw = dict(max_boxes=50, confidence=[0.05, 0.001, 0.3][it % 3], nms_threshold=0.45, nms_method=m,
                  per_class=bool(it % 2) and m != "soft")
        key = (m, kw["confidence"], kw["per_class"])
        if key not in refs:
            refs[key] = engine.decode_nms(preds, hw, (S, S), anchors, C, **kw)
        if it % 5 == 4:
            engine.grid_step(gt, y_out, [p[lo:hi].contiguous() for p in preds], hw[lo:hi].contiguous(),
                             (S, S), anchors, C, out=ex.local(), sync=(it % 2 == 0), **kw)
            full = ex.full()
        else:
            full = path.decode_nms(preds, shapes, gather="exchange", exchange=ex, sync=(it % 2 == 0), **kw)
        torch.cuda.synchronize()
        if not all(torch.equal(full[k], refs[key][k]) for k in keys):
            bad += 1
        if it % 7 == 0:                                   # skew the ranks against each other
            torch.cuda._sleep(int(2e6) * (1 + rank))
    q.put((rank, bad, ex.timeouts()))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    iters = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300
    two = "--two-gpus" in sys.argv
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=rank_main, args=(r, 2, port, iters, two, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=900) for _ in ps)
    for p in ps: p.join(timeout=60)
    print("iterations", iters, "two_gpus", two, "-> (rank, mismatching calls, barrier timeouts):", res)
    sys.exit(0 if all(b == 0 and t == 0 for _, b, t in res) else 1)
