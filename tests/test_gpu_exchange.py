"""Detection exchange (mgd_exchange_*, sharding.DetectionExchange): detections mirrored into
every rank's tensors by the NMS kernels themselves.

The driver's GPU box has one device, so the multi-rank case runs as TWO PROCESSES ON THE SAME
GPU: the buffers still cross process boundaries through CUDA IPC and every remote row is a
peer store through the mapping, exactly the code path of a 2-GPU node (where the mapping
leads over NVLink; `bench.py --scaling strong` under torchrun measures that)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

S, C, N = 416, 20, 12
KW = dict(max_boxes=50, confidence=0.05, nms_threshold=0.45, nms_method="diou")
KEYS = ("boxes_xywh", "boxes_xyxy", "scores", "classes", "index", "counts")


def _inputs(B, seed=5):
    import torch
    from multigriddet_b200 import engine, synth
    anchors = synth.voc_anchors(np.float32) if hasattr(synth, "voc_anchors") else synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(seed, B, N, S, C)
    y = engine.encode_targets(torch.from_numpy(boxes).cuda(), (S, S), anchors, C)
    preds = synth.planted_head_outputs(y, 3, seed=seed)
    shapes = synth.image_shapes(seed, B)
    return anchors, preds, shapes


def test_exchange_world_of_one_matches_plain_call():
    import torch
    from multigriddet_b200 import engine, sharding
    B = 9
    anchors, preds, shapes = _inputs(B)
    ref = engine.decode_nms(preds, shapes, (S, S), anchors, C, **KW)
    ex = sharding.DetectionExchange(B, KW["max_boxes"])
    try:
        got = engine.decode_nms(preds, shapes, (S, S), anchors, C, out=ex.local(), **KW)
        full = ex.full()
        for k in KEYS:
            assert got[k].data_ptr() == full[k].data_ptr()
            assert torch.equal(full[k], ref[k]), k
        assert ex.timeouts() == 0
        # outputs partly outside the exchange are refused, not silently half-mirrored
        mixed = dict(ex.local())
        mixed["scores"] = torch.empty_like(mixed["scores"])
        with pytest.raises(ValueError, match="exchange"):
            engine.decode_nms(preds, shapes, (S, S), anchors, C, out=mixed, **KW)
    finally:
        ex.close()


def _rank_main(rank, world, port, B, method, q):
    sys.path.insert(0, ROOT)
    if method.endswith("+chunks"):                            # several internal chunks per mirrored call
        method = method.split("+")[0]
        os.environ["MGD_DECODE_CHUNK_IMAGES"] = "2"
        os.environ["MGD_ENCODE_CHUNK_IMAGES"] = "2"
    import torch
    import torch.distributed as dist
    from multigriddet_b200 import engine, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    try:
        anchors, preds, shapes = _inputs(B)
        kw = dict(KW, nms_method=method)
        ref = engine.decode_nms(preds, shapes, (S, S), anchors, C, **kw)
        ex = sharding.DetectionExchange(B, kw["max_boxes"])
        path = sharding.ShardedGridPath(anchors, C, (S, S))
        ok = True
        for it in range(3):                                   # repeated calls: epochs advance
            full = path.decode_nms(preds, shapes, gather="exchange", exchange=ex, **kw)
            torch.cuda.synchronize()
            ok = ok and all(torch.equal(full[k], ref[k]) for k in KEYS)
            if it == 0:                                       # wipe, so the next round must rewrite it
                dist.barrier()
                for k in KEYS:
                    full[k].zero_()
                torch.cuda.synchronize()
                dist.barrier()
        # the fused entry (encode + decode/NMS in one call) mirrors the same way
        from multigriddet_b200 import synth
        lo, hi = sharding.shard_bounds(B, rank, world)
        for k in KEYS:
            ex.full()[k].zero_()
        torch.cuda.synchronize()
        dist.barrier()
        boxes = torch.from_numpy(synth.synth_boxes(9, hi - lo, N, S, C)).cuda()
        y_ref = engine.encode_targets(boxes, (S, S), anchors, C)
        y_out = [torch.empty_like(y) for y in y_ref]
        hw = torch.from_numpy(np.ascontiguousarray(shapes[lo:hi])).cuda()
        engine.grid_step(boxes, y_out, [p[lo:hi].contiguous() for p in preds], hw, (S, S), anchors, C,
                         out=ex.local(), **kw)
        torch.cuda.synchronize()
        dist.barrier()
        ok = ok and all(torch.equal(ex.full()[k], ref[k]) for k in KEYS)
        ok = ok and all(torch.equal(a, b) for a, b in zip(y_out, y_ref))
        q.put((rank, bool(ok), ex.timeouts()))
        dist.barrier()
        ex.close()
    except Exception as e:                                    # pragma: no cover
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method", ["diou", "soft", "diou+chunks"])
def test_two_ranks_mirror_their_shards_into_each_other(method):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, B = 2, 11                                          # uneven shards: 6 + 5 images
    import socket
    with socket.socket() as sk:                               # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, B, method, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, timeouts in sorted(res):
        assert ok is True, (rank, timeouts)
        assert timeouts == 0, (rank, timeouts)


def test_device_calls_are_cuda_graph_capturable():
    """Device-memory calls neither synchronise nor allocate outside the stream: the fused step
    can be captured into a CUDA graph and replayed (scripts/graph_probe.py times it; replay is
    not faster than the eager calls, which are not launch-bound)."""
    import torch
    from multigriddet_b200 import engine, synth
    B = 5
    anchors, preds, shapes = _inputs(B, seed=8)
    hw = torch.from_numpy(shapes).cuda()
    boxes = torch.from_numpy(synth.synth_boxes(4, B, N, S, C)).cuda()
    y_ref = engine.encode_targets(boxes, (S, S), anchors, C)
    ref = engine.decode_nms(preds, hw, (S, S), anchors, C, **KW)
    y_out = [torch.zeros_like(t) for t in y_ref]
    out = {k: torch.zeros_like(ref[k]) for k in KEYS}

    def step():
        engine.grid_step(boxes, y_out, preds, hw, (S, S), anchors, C, sync=False, out=out, **KW)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                            # warm-up: lazy one-time allocations
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for t in list(out.values()) + y_out:
        t.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize()
    assert int(out["counts"].sum()) == 0                  # captured, not executed
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(out[k], ref[k]) for k in KEYS)
    assert all(torch.equal(a, b) for a, b in zip(y_out, y_ref))


def test_host_zerocopy_flag_gives_the_staged_result():
    """MGD_FLAG_HOST_ZEROCOPY: page-locked host tensors accessed in place by the kernels (the
    decoder reads only the sectors it asks for, the writer stores y_true over the link) -- same
    bits as the staged path; pageable arrays silently take the staged path."""
    import torch
    from multigriddet_b200 import _lib, engine, synth
    B = 7
    anchors, preds, shapes = _inputs(B, seed=12)
    pinned = [_lib.pinned.empty(tuple(p.shape), np.float32) for p in preds]
    for dst, p in zip(pinned, preds):
        dst[...] = p.cpu().numpy()
    ref = engine.decode_nms(pinned, shapes, (S, S), anchors, C, **KW)
    got = engine.decode_nms(pinned, shapes, (S, S), anchors, C, zerocopy=True, **KW)
    pageable = [np.array(p) for p in pinned]
    got_pageable = engine.decode_nms(pageable, shapes, (S, S), anchors, C, zerocopy=True, **KW)
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k
        assert np.array_equal(got_pageable[k], ref[k]), k
    boxes = synth.synth_boxes(6, B, N, S, C)
    y_ref = engine.encode_targets(boxes, (S, S), anchors, C)
    y_out = [_lib.pinned.empty(y.shape, np.float32) for y in y_ref]
    for y in y_out:
        y[...] = -7.0
    engine.encode_targets(boxes, (S, S), anchors, C, out=y_out, zerocopy=True)
    assert all(np.array_equal(a, b) for a, b in zip(y_out, y_ref))
