"""C-ABI checks that need no GPU: the library loads, exports every symbol the
header declares, the ctypes structs match the C layout, argument validation and the
no-device path behave (and there is no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from multigriddet_b200 import _lib, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mgd.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    declared = re.findall(r"MGD_API\s+[\w\s\*]+?\b(mgd_\w+)\s*\(", text)
    assert len(declared) >= 12
    assert sorted(set(declared)) == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mgd_\w+)", nm))
    assert exported == set(declared)          # nothing else leaks out of the library


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mgd.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mgd_head_config),'
                   'offsetof(mgd_head_config, anchors), offsetof(mgd_head_config, anchors_f64),'
                   'sizeof(mgd_post_config), offsetof(mgd_post_config, nms_method),'
                   'offsetof(mgd_post_config, soft_score_threshold));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    H, P = _lib.HeadConfig, _lib.PostConfig
    assert got == [ctypes.sizeof(H), H.anchors.offset, H.anchors_f64.offset,
                   ctypes.sizeof(P), P.nms_method.offset, P.soft_score_threshold.offset]


def test_version_and_device_count(lib):
    assert lib.mgd_version() == 100
    assert lib.mgd_device_count() >= 0


def test_argument_validation_without_a_device(lib):
    anchors = synth.coco_anchors(np.float32)
    boxes = np.zeros((1, 2, 5), np.float32)
    with pytest.raises(NotImplementedError):           # non-square input
        engine.encode_targets(boxes, (608, 416), anchors, 80)
    with pytest.raises(ValueError):                    # no classes
        engine.encode_targets(boxes, (608, 608), anchors, 0)
    with pytest.raises(ValueError):
        engine.encode_targets(np.zeros((1, 2, 4), np.float32), (608, 608), anchors, 80)
    with pytest.raises(ValueError):                    # too many anchors per layer
        _lib.make_head_config([np.ones((9, 2), np.float32)], 80, (608, 608))
    preds = [np.zeros((1, g, g, 88), np.float32) for g in (19, 38, 76)]
    with pytest.raises(ValueError):                    # multigrid_decode.py:62-63
        engine.decode_nms(preds[:2], None, (608, 608), anchors, 80)
    with pytest.raises(ValueError):
        engine.decode_nms(preds, None, (608, 608), anchors, 80, max_boxes=0)
    with pytest.raises(NotImplementedError):
        engine.decode_nms(preds, None, (608, 608), anchors, 80, nms_method="fuse")
    with pytest.raises(ValueError):                    # wrong channel count
        engine.decode_nms(preds, None, (608, 608), anchors, 20)


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    if lib.mgd_device_count() > 0:
        pytest.skip("a CUDA device is present")
    anchors = synth.coco_anchors(np.float32)
    with pytest.raises(_lib.MgdError, match="no CPU fallback"):
        engine.encode_targets(np.zeros((1, 2, 5), np.float32), (608, 608), anchors, 80)
    preds = [np.zeros((1, g, g, 88), np.float32) for g in (19, 38, 76)]
    with pytest.raises(_lib.MgdError, match="no CPU fallback"):
        engine.decode_nms(preds, None, (608, 608), anchors, 80)
    with pytest.raises(_lib.MgdError, match="no CPU fallback"):
        engine.nms(np.zeros((3, 4)), np.zeros(3))
    with pytest.raises(_lib.MgdError, match="no CPU fallback"):
        engine.soft_nms(np.zeros((3, 4)), np.ones(3))
    with pytest.raises(_lib.MgdError, match="no CPU fallback"):
        engine.decode_dense(preds, anchors, 80, (608, 608))


def test_exchange_argument_validation_and_no_device(lib):
    """mgd_exchange_*: argument errors are reported before any device work; without a device
    the exchange fails like every other entry point."""
    ex = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
    bad = [(0, 0, 0, 1024), (0, 9, 0, 1024), (0, 2, 2, 1024), (0, 2, -1, 1024), (0, 1, 0, 0)]
    for dev, world, rank, nbytes in bad:
        rc = lib.mgd_exchange_create(dev, world, rank, nbytes, ctypes.byref(ex), handle)
        assert rc in (_lib.ERR_INVALID_ARGUMENT, _lib.ERR_UNSUPPORTED), (world, rank, nbytes, rc)
        assert not ex.value
    assert lib.mgd_exchange_create(0, 1, 0, 1024, None, handle) == _lib.ERR_INVALID_ARGUMENT
    assert lib.mgd_exchange_connect(None, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.mgd_exchange_destroy(None) == _lib.OK
    if lib.mgd_device_count() == 0:
        rc = lib.mgd_exchange_create(0, 1, 0, 1024, ctypes.byref(ex), handle)
        assert rc == _lib.ERR_NO_DEVICE and not ex.value
        with pytest.raises(_lib.MgdError, match="no CPU fallback"):
            _lib.raise_for_status(rc)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "multigriddet_b200")
    for dirpath, _, names in os.walk(pkg):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, n)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), n
                assert "mgd_oracle" not in text and "/root/reference" not in text, n


def test_dlpack_struct_abi():
    """The minimal DLTensor the library reads has the standard dlpack layout."""
    import torch
    t = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    cap = t.__dlpack__()
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    p = ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")

    class DLTensor(ctypes.Structure):
        _fields_ = [("data", ctypes.c_void_p), ("device_type", ctypes.c_int32),
                    ("device_id", ctypes.c_int32), ("ndim", ctypes.c_int32),
                    ("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16),
                    ("shape", ctypes.POINTER(ctypes.c_int64)), ("strides", ctypes.POINTER(ctypes.c_int64)),
                    ("byte_offset", ctypes.c_uint64)]
    d = DLTensor.from_address(p)
    assert (d.device_type, d.ndim, d.code, d.bits, d.lanes) == (1, 2, 2, 32, 1)
    assert [d.shape[0], d.shape[1]] == [2, 3] and d.data == t.data_ptr()
