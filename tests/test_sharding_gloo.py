"""The N > 1 path on the CPU: world_size 2, gloo backend.

Each rank takes its contiguous slice of a global batch, runs the hot path on it (the
oracle stands in for the CUDA kernels -- this tests the sharding / gather plumbing,
not the arithmetic) and the host-side gather must reproduce the single-process
result in image order.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleCompute:
    """Same call surface as multigriddet_b200.engine, computed by the C oracle."""

    @staticmethod
    def encode_targets(boxes, input_shape, anchors, num_classes, **kw):
        from oracle import c_oracle
        return c_oracle.encode_targets(boxes, input_shape, anchors, num_classes)

    @staticmethod
    def decode_nms(preds, shapes, model_size, anchors, num_classes, **kw):
        from oracle import c_oracle
        if shapes is None:
            shapes = [model_size]
        r = c_oracle.decode_nms(preds, shapes, model_size, anchors, num_classes, **kw)
        return {k: r[k] for k in ("boxes_xywh", "boxes_xyxy", "scores", "classes", "index", "counts")}


def _inputs():
    from multigriddet_b200 import synth
    from oracle import c_oracle
    S, C, B = 160, 20, 7            # odd batch: ranks get 4 and 3 images
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(4, B, 10, S, C)
    yt = c_oracle.encode_targets(boxes, (S, S), anchors, C)
    preds = [p.numpy() for p in synth.planted_head_outputs([torch.from_numpy(y) for y in yt], 3, 4)]
    shapes = synth.image_shapes(2, B)
    return S, C, anchors, boxes, preds, shapes


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multigriddet_b200 import sharding
    S, C, anchors, boxes, preds, shapes = _inputs()
    path = sharding.ShardedGridPath(anchors, C, (S, S), compute=OracleCompute)
    assert (path.rank, path.world_size) == (rank, world)
    y_local = path.encode(boxes)
    lo, hi = sharding.shard_bounds(boxes.shape[0], rank, world)
    assert y_local[0].shape[0] == hi - lo
    kw = dict(max_boxes=20, confidence=0.05, nms_threshold=0.45, nms_method="diou")
    everyone = path.decode_nms(preds, shapes, **kw)                 # all_gather
    only0 = path.decode_nms(preds, shapes, dst=0, **kw)             # gather to rank 0
    assert (only0 is None) == (rank != 0)
    # tensor gather (NCCL over NVLink on the GPU box, gloo here): unequal shards 4 + 3
    local = path.decode_nms(preds, shapes, gather=False, **kw)
    dev = sharding.gather_detections_device({k: torch.from_numpy(np.ascontiguousarray(v))
                                             for k, v in local.items()}, n_total=boxes.shape[0])
    for k, v in everyone.items():
        assert np.array_equal(dev[k].numpy(), v), k
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), y0=y_local[0], lo=lo, hi=hi,
             **{"all_" + k: v for k, v in everyone.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_reproduces_single_process(tmp_path):
    from oracle import c_oracle
    c_oracle.build()
    import socket
    with socket.socket() as sk:                               # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    S, C, anchors, boxes, preds, shapes = _inputs()
    full_y = c_oracle.encode_targets(boxes, (S, S), anchors, C)
    full = OracleCompute.decode_nms(preds, shapes, (S, S), anchors, C, max_boxes=20, confidence=0.05,
                                    nms_threshold=0.45, nms_method="diou")
    covered = []
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        covered.append((lo, hi))
        assert np.array_equal(z["y0"], full_y[0][lo:hi])           # y_true shard stays local
        for k, v in full.items():                                  # gathered detections, image order
            assert np.array_equal(z["all_" + k], v), k
    assert covered == [(0, 4), (4, 7)]
