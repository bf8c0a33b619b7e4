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


_EXCHANGE_SPEC = (("boxes_xywh", "float64", 4), ("scores", "float64", 1), ("boxes_xyxy", "int32", 4),
                  ("classes", "int32", 1), ("index", "int32", 1))


def exchange_layout(n_total: int, max_boxes: int) -> Tuple[Dict[str, Tuple[int, Tuple[int, ...], str]], int]:
    """Byte layout of the whole-batch detection tensors inside an exchange buffer:
    ``{name: (offset, shape, dtype)}`` and the total size.  Every tensor starts on a
    256-byte boundary (the kernels send remote rows as 16-byte stores)."""
    off, out = 0, {}
    for name, dt, width in _EXCHANGE_SPEC:
        shape = (n_total, max_boxes, 4) if width == 4 else (n_total, max_boxes)
        out[name] = (off, shape, dt)
        off += (int(np.prod(shape)) * np.dtype(dt).itemsize + 255) & ~255
    out["counts"] = (off, (n_total,), "int32")
    off += (n_total * 4 + 255) & ~255
    return out, off


class _DeviceSpan:
    """A span of device memory as a ``__cuda_array_interface__`` provider (keeps ``owner`` alive)."""

    def __init__(self, ptr, shape, dtype, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class DetectionExchange:
    """Whole-batch detection tensors that every rank ends up holding -- without a collective.

    SURVEY.md 8e: the path's one exchange step.  Each rank owns ``bytes`` of device memory laid
    out as the padded detection tensors of the WHOLE batch (``n_total`` images); the buffers
    of all ranks are mapped into every process (CUDA IPC, ``mgd_exchange_*`` in libmgd).  A
    rank decodes its shard with ``out=self.local()``: the NMS kernels store each image's rows
    into every rank's tensors while they run (peer stores over NVLink), so when the call's
    work completes on the stream ``self.full()`` holds all images on every rank.  Collective
    semantics: every rank makes the same sequence of mirrored calls.

    One process per GPU on one node; ``torch.distributed`` only carries the 64-byte IPC
    handles at construction.  World size 1 works without ``torch.distributed``."""

    def __init__(self, n_total: int, max_boxes: int, device: Optional[int] = None, group=None):
        import ctypes
        import torch
        from . import _lib
        lib = _lib.load()
        d = _dist()
        self.rank, self.world_size = (d.get_rank(group), d.get_world_size(group)) if d else (0, 1)
        self.n_total, self.max_boxes = int(n_total), int(max_boxes)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.layout, self.bytes = exchange_layout(self.n_total, self.max_boxes)
        self._lib, self._mod = lib, _lib
        self._h = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        _lib.raise_for_status(lib.mgd_exchange_create(self.device, self.world_size, self.rank, self.bytes,
                                                      ctypes.byref(self._h), handle))
        if self.world_size > 1:
            handles: List = [None] * self.world_size
            d.all_gather_object(handles, bytes(handle), group=group)
            blob = (ctypes.c_ubyte * (_lib.IPC_HANDLE_BYTES * self.world_size)).from_buffer_copy(b"".join(handles))
            err = None
            try:
                _lib.raise_for_status(lib.mgd_exchange_connect(self._h, blob))
            except Exception as e:                        # e.g. no peer access between two devices
                err = e
            # all ranks agree on the outcome (and nobody stores before everybody has mapped everything)
            oks: List = [None] * self.world_size
            d.all_gather_object(oks, err is None, group=group)
            if not all(oks):
                lib.mgd_exchange_destroy(self._h)
                self._h = None
                raise RuntimeError(f"detection exchange unavailable on ranks {[r for r, ok in enumerate(oks) if not ok]}"
                                   + (f": {err}" if err is not None else ""))
        base, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.raise_for_status(lib.mgd_exchange_buffer(self._h, ctypes.byref(base), ctypes.byref(nbytes)))
        dev = torch.device("cuda", self.device)
        self._full = {k: torch.as_tensor(_DeviceSpan(base.value + off, shape, dt, self), device=dev)
                      for k, (off, shape, dt) in self.layout.items()}

    def full(self) -> Dict:
        """The whole batch's tensors in this rank's memory (valid once a mirrored call completed)."""
        return dict(self._full)

    def local(self, keys: Optional[Sequence[str]] = None) -> Dict:
        """This rank's slice of every tensor (``shard_bounds`` order): pass as ``out=``."""
        lo, hi = shard_bounds(self.n_total, self.rank, self.world_size)
        if hi == lo:
            raise ValueError("every rank of a detection exchange needs at least one image "
                             f"({self.n_total} images over {self.world_size} ranks)")
        keys = list(self._full) if keys is None else list(keys) + ["counts"]
        return {k: self._full[k][lo:hi] for k in dict.fromkeys(keys)}

    def timeouts(self) -> int:
        """Device-side barrier timeouts so far (a peer that never made its call); 0 is healthy."""
        import ctypes
        import torch
        n = ctypes.c_int()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._mod.raise_for_status(self._lib.mgd_exchange_timeouts(self._h, ctypes.c_void_p(stream), ctypes.byref(n)))
        return int(n.value)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            self._full = {}
            self._lib.mgd_exchange_destroy(self._h)
            self._h = None


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

    def decode_nms(self, global_preds, global_image_shapes=None, gather=True, dst=None, exchange=None, **kw):
        """``gather``: True (host lists), "device" (one all-gather), "exchange" (the NMS kernels
        write every rank's tensors directly; needs ``exchange=DetectionExchange(n, max_boxes)``),
        False (local shard only)."""
        n = global_preds[0].shape[0]
        sl = self.local_slice(n)
        shapes = None
        if global_image_shapes is not None:
            shapes = np.asarray(global_image_shapes).reshape(-1, 2)
            shapes = shapes[sl] if shapes.shape[0] == n else shapes
        if gather == "exchange":
            if exchange is None or exchange.n_total != n:
                raise ValueError("gather='exchange' needs exchange=DetectionExchange(n_images, max_boxes)")
            self.compute.decode_nms([p[sl] for p in global_preds], shapes, self.input_shape,
                                    self.anchors, self.num_classes, out=exchange.local(kw.get("want")), **kw)
            return exchange.full()
        det = self.compute.decode_nms([p[sl] for p in global_preds], shapes, self.input_shape,
                                      self.anchors, self.num_classes, **kw)
        if gather == "device":
            return gather_detections_device({k: v for k, v in det.items() if k != "stats"}, n)
        return gather_detections(det, dst) if gather else det
