"""ctypes binding of ``oracle/mgd_oracle.c`` (test infrastructure, see
``oracle/__init__.py``).  ``build()`` compiles it with ``oracle/Makefile``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmgd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mgd_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_expf.restype = ctypes.c_float
        _lib.orc_expf.argtypes = [ctypes.c_float]
        _lib.orc_logf.restype = ctypes.c_float
        _lib.orc_logf.argtypes = [ctypes.c_float]
        _lib.orc_tanhf.restype = ctypes.c_float
        _lib.orc_tanhf.argtypes = [ctypes.c_float]
    return _lib


def _anchor_args(anchors):
    flat = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64)
                                                for a in anchors], 0))
    na = np.array([len(a) for a in anchors], dtype=np.int32)
    f64 = int(np.asarray(anchors[0]).dtype == np.float64)
    return flat, na, f64


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def encode_targets(true_boxes, input_shape, anchors, num_classes, grid_shapes=None,
                   return_stats=False):
    boxes = np.ascontiguousarray(np.asarray(true_boxes), dtype=np.float32)
    B, N = boxes.shape[0], boxes.shape[1]
    L = len(anchors)
    flat, na, f64 = _anchor_args(anchors)
    if grid_shapes is None:
        grid_shapes = [(input_shape[0] // s, input_shape[1] // s) for s in (32, 16, 8, 4, 2)[:L]]
    ghw = np.array([[int(g[0]), int(g[1])] for g in grid_shapes], dtype=np.int32)
    outs = [np.empty((B, int(ghw[l, 0]), int(ghw[l, 1]), 5 + int(na[l]) + num_classes),
                     dtype=np.float32) for l in range(L)]
    arr = (ctypes.POINTER(ctypes.c_float) * L)(*[_ptr(o, ctypes.c_float) for o in outs])
    stats = np.zeros(2, dtype=np.int64)
    rc = lib().orc_encode(_ptr(boxes, ctypes.c_float), B, N, _ptr(flat, ctypes.c_double),
                          _ptr(na, ctypes.c_int), L, f64, int(num_classes),
                          int(input_shape[0]), int(input_shape[1]), _ptr(ghw, ctypes.c_int),
                          arr, _ptr(stats, ctypes.c_longlong))
    if rc == 1:
        raise AssertionError("class id must be less than num_classes")
    if rc:
        raise RuntimeError(f"orc_encode failed: {rc}")
    if return_stats:
        return outs, {"n_valid_boxes": int(stats[0]), "n_skipped_writes": int(stats[1])}
    return outs


def decode_nms(preds, image_shapes, model_image_size, anchors, num_classes,
               max_boxes=100, confidence=0.1, nms_threshold=0.5, nms_method="diou",
               per_class=False, use_softmax=True, rescore_confidence=True,
               dense=False):
    preds = [np.ascontiguousarray(p, dtype=np.float32) for p in preds]
    B, L = preds[0].shape[0], len(preds)
    flat, na, f64 = _anchor_args(anchors)
    ghw = np.array([[p.shape[1], p.shape[2]] for p in preds], dtype=np.int32)
    cells = int(sum(int(g[0]) * int(g[1]) for g in ghw))
    ihw = np.ascontiguousarray(np.asarray(image_shapes, dtype=np.int32).reshape(-1, 2))
    if ihw.shape[0] == 1 and B > 1:
        ihw = np.ascontiguousarray(np.tile(ihw, (B, 1)))
    res = {
        "boxes_xywh": np.zeros((B, max_boxes, 4), np.float64),
        "boxes_xyxy": np.zeros((B, max_boxes, 4), np.int32),
        "scores": np.zeros((B, max_boxes), np.float64),
        "classes": np.zeros((B, max_boxes), np.int32),
        "index": np.zeros((B, max_boxes), np.int32),
        "counts": np.zeros((B,), np.int32),
        "n_candidates": np.zeros((B,), np.int32),
    }
    all_scores = np.zeros((B, cells), np.float32) if dense else None
    all_cls = np.zeros((B, cells), np.int32) if dense else None
    arr = (ctypes.POINTER(ctypes.c_float) * L)(*[_ptr(p, ctypes.c_float) for p in preds])
    method = {"diou": 1, "iou": 0, "standard": 0, "cluster": 0}[nms_method]
    rc = lib().orc_decode_nms(
        arr, B, L, _ptr(ghw, ctypes.c_int), _ptr(flat, ctypes.c_double),
        _ptr(na, ctypes.c_int), f64, int(num_classes), int(model_image_size[0]),
        int(model_image_size[1]), _ptr(ihw, ctypes.c_int), int(use_softmax),
        int(rescore_confidence), ctypes.c_double(confidence),
        ctypes.c_double(nms_threshold), method, int(per_class), int(max_boxes),
        _ptr(res["boxes_xywh"], ctypes.c_double), _ptr(res["boxes_xyxy"], ctypes.c_int),
        _ptr(res["scores"], ctypes.c_double), _ptr(res["classes"], ctypes.c_int),
        _ptr(res["index"], ctypes.c_int), _ptr(res["counts"], ctypes.c_int),
        _ptr(res["n_candidates"], ctypes.c_int),
        _ptr(all_scores, ctypes.c_float) if dense else None,
        _ptr(all_cls, ctypes.c_int) if dense else None)
    if rc:
        raise RuntimeError(f"orc_decode_nms failed: {rc}")
    if dense:
        res["all_scores"], res["all_cls"] = all_scores, all_cls
    return res
