"""torchrun --nproc-per-node N scripts/nccl_gather_check.py: image-sharded decode/NMS on N
GPUs, detections gathered device-to-device with NCCL, compared with rank 0 doing the whole
batch alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from multigriddet_b200 import engine, sharding, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, C, B = 608, 80, 37                      # not divisible by the world size
anchors = synth.coco_anchors(np.float32)
boxes = synth.synth_boxes(9, B, 60, S, C)
yt = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
preds = synth.planted_head_outputs(yt, 3, 9)
shapes = synth.image_shapes(1, B, mixed=True)
kw = dict(max_boxes=100, confidence=0.001, nms_threshold=0.45, nms_method="diou")
path = sharding.ShardedGridPath(anchors, C, (S, S))
y_local = path.encode(torch.from_numpy(boxes).cuda())
lo, hi = sharding.shard_bounds(B, rank, world)
assert all(torch.equal(a, b[lo:hi]) for a, b in zip(y_local, yt))
got = path.decode_nms(preds, shapes, gather="device", **kw)
ref = engine.decode_nms(preds, torch.from_numpy(shapes).cuda(), (S, S), anchors, C, **kw)
for k in ("boxes_xyxy", "scores", "classes", "index", "counts"):
    assert got[k].device.index == local and torch.equal(got[k], ref[k]), k
host = path.decode_nms(preds, shapes, **kw)
assert np.array_equal(host["index"], ref["index"].cpu().numpy())
dist.barrier()
if rank == 0:
    print(f"ok: {world} ranks, {int(ref['counts'].sum())} detections gathered over NCCL == single-GPU result")
dist.destroy_process_group()
