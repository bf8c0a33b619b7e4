"""Image-sharded multi-GPU driver.

Every image is independent in both directions (the reference encoder loops
``for b`` with no cross-image state, generators.py:3427; the evaluator
post-processes image by image, evaluator.py:254-273), so N GPUs split the batch
dimension into contiguous slices and run the same kernels.  There is no
data-path collective: ``y_true`` shards stay on the rank that will consume them
(data-parallel training wants them local) and the only exchange is the host-side
gather of the small per-rank detection lists (<= max_boxes x 28 B per image).

One process per GPU (``torchrun``); ``torch.distributed`` is used only for the
gather (any backend: NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of the batch owned by ``rank``.

    The first ``n_images % world_size`` ranks get one extra image."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(n_images), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(arrays: Sequence, rank: int, world_size: int) -> List:
    """Slice every array / tensor of ``arrays`` along dim 0 for ``rank``."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(a.shape[0], rank, world_size)
        out.append(a[lo:hi])
    return out


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def world() -> Tuple[int, int]:
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def to_host(det: Dict) -> Dict[str, np.ndarray]:
    """Detection dict of torch CUDA tensors / arrays -> plain NumPy arrays."""
    out = {}
    for k, v in det.items():
        if k.startswith("_") or k == "stats":
            continue
        out[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return out


def gather_detections(local: Dict, dst: Optional[int] = None) -> Optional[Dict[str, np.ndarray]]:
    """Host-side gather of per-rank padded detection tensors, in rank (= image) order.

    ``local``: the dict ``engine.decode_nms`` returns for this rank's shard.
    ``dst=None``: every rank gets the full result (all_gather); otherwise only
    ``dst`` does and the others get ``None``.
    """
    host = to_host(local)
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return host
    if dst is None:
        parts: List = [None] * d.get_world_size()
        d.all_gather_object(parts, host)
    else:
        parts = [None] * d.get_world_size() if d.get_rank() == dst else None
        d.gather_object(host, parts, dst=dst)
        if d.get_rank() != dst:
            return None
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def gather_detections_device(local: Dict, n_total: Optional[int] = None) -> Dict:
    """Device-side gather: every rank ends with the whole batch's padded detection
    tensors in ITS OWN device memory (the path's one exchange step, <= 2.8 KB
    per image over NVLink with the NCCL backend; CPU tensors with gloo).

    ONE collective per call: the rows of all tensors of an image are packed into one byte
    row, the packed shards are all-gathered (``all_gather_into_tensor``) and the result is
    viewed back -- nothing synchronises with the host.  ``local`` holds torch tensors of this
    rank's shard; shards follow ``shard_bounds`` (they differ by at most one image and are
    padded to the largest for the collective).  ``n_total``: images in the global batch;
    default ``world_size`` equal shards of the local size."""
    import torch
    d = _dist()
    keys = [k for k, v in local.items() if hasattr(v, "detach") and not k.startswith("_")]
    if d is None or d.get_world_size() == 1:
        return {k: local[k] for k in keys}
    ws, rank = d.get_world_size(), d.get_rank()
    n_local = int(local["counts"].shape[0])
    if n_total is None:
        n_total = n_local * ws
    sizes = [shard_bounds(n_total, r, ws)[1] - shard_bounds(n_total, r, ws)[0] for r in range(ws)]
    if sizes[rank] != n_local:
        raise ValueError(f"rank {rank} holds {n_local} images, shard_bounds({n_total}) says {sizes[rank]}")
    n_max = max(sizes)
    # byte layout of one image's row: the tensors in key order, each padded to 8 bytes
    spans, off = [], 0
    for k in keys:
        t = local[k]
        nbytes = t.element_size() * int(np.prod(t.shape[1:], dtype=np.int64))
        spans.append((k, off, nbytes))
        off += (nbytes + 7) & ~7
    row = off
    dev = local["counts"].device
    send = torch.zeros((n_max, row), dtype=torch.uint8, device=dev)
    for k, o, nb in spans:
        if n_local:
            send[:n_local, o:o + nb] = local[k].contiguous().view(torch.uint8).reshape(n_local, nb)
    recv = torch.empty((ws * n_max, row), dtype=torch.uint8, device=dev)
    d.all_gather_into_tensor(recv, send)
    recv = recv.view(ws, n_max, row)
    if all(sz == n_max for sz in sizes):
        full = recv.reshape(ws * n_max, row)
    else:
        full = torch.cat([recv[r, :sizes[r]] for r in range(ws)], 0)
    out = {}
    for k, o, nb in spans:
        t = local[k]
        out[k] = full[:, o:o + nb].contiguous().view(t.dtype).reshape((full.shape[0],) + tuple(t.shape[1:]))
    return out


class ShardedGridPath:
    """Runs the hot path on this rank's slice of a global batch.

    ``encode`` returns the local ``y_true`` shard (never gathered); ``decode_nms``
    returns the gathered detections of the whole batch (host NumPy arrays; torch tensors
    gathered device-to-device with ``gather="device"``; the local ones with ``gather=False``).  ``compute`` is injectable so the plumbing is testable on a
    CPU box (tests pass the oracle); by default it is ``multigriddet_b200.engine``.
    """

    def __init__(self, anchors, num_classes, input_shape, compute=None):
        if compute is None:
            from . import engine as compute
        self.compute = compute
        self.anchors = anchors
        self.num_classes = num_classes
        self.input_shape = tuple(int(v) for v in input_shape)
        self.rank, self.world_size = world()

    def local_slice(self, n_images: int) -> slice:
        lo, hi = shard_bounds(n_images, self.rank, self.world_size)
        return slice(lo, hi)

    def encode(self, global_boxes, **kw):
        sl = self.local_slice(global_boxes.shape[0])
        return self.compute.encode_targets(global_boxes[sl], self.input_shape, self.anchors,
                                           self.num_classes, **kw)

    def decode_nms(self, global_preds, global_image_shapes=None, gather=True, dst=None, **kw):
        n = global_preds[0].shape[0]
        sl = self.local_slice(n)
        shapes = None
        if global_image_shapes is not None:
            shapes = np.asarray(global_image_shapes).reshape(-1, 2)
            shapes = shapes[sl] if shapes.shape[0] == n else shapes
        det = self.compute.decode_nms([p[sl] for p in global_preds], shapes, self.input_shape,
                                      self.anchors, self.num_classes, **kw)
        if gather == "device":
            return gather_detections_device({k: v for k, v in det.items() if k != "stats"}, n)
        return gather_detections(det, dst) if gather else det
