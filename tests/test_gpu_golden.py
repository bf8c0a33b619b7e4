"""GPU: the CUDA path (through the C ABI and through the drop-in classes) against the
golden vectors generated from the reference itself (tests/golden, no reference
tree needed)."""
import ctypes

import numpy as np
import pytest

import golden_util as G
from multigriddet_b200 import engine, _lib
from multigriddet_b200.data import MultiGridTargetEncoder, MultiGridConfig, preprocess_true_boxes
from multigriddet_b200.postprocess import (ClusterNMS, DIoUNMS, MultiGridDecoder, SoftNMS, StandardNMS,
                                           WeightedBoxesFusion,
                                           multigriddet_postprocess_gpu, nms_boxes)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", G.files("tfencode"))
def test_tf_compat_encode_against_reference_golden(path):
    """MGD_FLAG_TF_COMPAT against the outputs of the reference's own TensorFlow encoder
    (generators.py:2696-3390, run over oracle/tf_shim.py when the fixtures were made); through
    the C ABI with host arrays and with device tensors.  Logarithms to 1e-5 (TensorFlow's log
    is Eigen's; the fixture's is glibc's logf)."""
    import torch
    from multigriddet_b200.data import tf_preprocess_true_boxes
    z = np.load(path)
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    ref = G.dense_y_true(z)
    got = engine.encode_targets(z["boxes"], (S, S), anchors, C, semantics="tf_compat")
    G.assert_encode_matches(got, ref, exact_floats=False)
    got = engine.encode_targets(torch.from_numpy(z["boxes"]).cuda(), (S, S), anchors, C, semantics="tf_compat")
    G.assert_encode_matches([g.cpu().numpy() for g in got], ref, exact_floats=False)
    got = tf_preprocess_true_boxes(z["boxes"], (S, S), anchors, C, False,
                                   [(S // 32,) * 2, (S // 16,) * 2, (S // 8,) * 2])
    G.assert_encode_matches([np.asarray(g) for g in got], ref, exact_floats=False)


@pytest.mark.parametrize("path", G.files("encode"))
def test_encode_against_reference_golden(path):
    z = np.load(path)
    if "boxes" not in z:
        return
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    got = preprocess_true_boxes(z["boxes"], (S, S), anchors, C, False)
    assert isinstance(got, list) and all(isinstance(g, np.ndarray) for g in got)
    G.assert_encode_matches(got, G.dense_y_true(z), exact_floats=False)


def test_encode_known_answers_through_the_dropin():
    z = np.load(G.GOLDEN + "/encode_known_answer.npz")
    anchors = [np.array(a, dtype=np.float32) for a in z["anchors"]]
    for i in range(3):
        got = preprocess_true_boxes(z[f"box{i}"], (608, 608), anchors, int(z[f"C{i}"]), False,
                                    grid_shapes=[(19, 19), (38, 38), (76, 76)])
        G.assert_encode_matches(got, G.dense_y_true(z, prefix=f"c{i}_"), exact_floats=False)
    enc = MultiGridTargetEncoder(MultiGridConfig(num_classes=1))     # small-first default anchors
    y = enc.encode_targets(z["box0"][0])
    G.assert_encode_matches([a[None] for a in y], G.dense_y_true(z, prefix="c0_"), exact_floats=False)


@pytest.fixture(params=["cta_per_image", "warp_per_image"])
def nms_kernel_choice(request):
    """Both NMS kernels against the reference's golden detections (the library would pick
    the CTA-per-image kernel for batches this small)."""
    import os
    os.environ["MGD_NMS_WARP_MIN_IMAGES"] = "1" if request.param == "warp_per_image" else "1000000"
    yield request.param
    del os.environ["MGD_NMS_WARP_MIN_IMAGES"]


@pytest.mark.parametrize("path", G.files("decode"))
def test_postprocess_against_reference_golden(path, nms_kernel_choice):
    z = np.load(path)
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    preds = G.preds_of(z)
    B = preds[0].shape[0]
    dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
    dense = dec.decode_predictions(preds)
    np.testing.assert_allclose(dense[:, ::37, 4:], z["dense_sample"][..., 4:], rtol=1e-5, atol=1e-30)
    np.testing.assert_allclose(dense[:, ::37, :4], z["dense_sample"][..., :4], rtol=1e-5, atol=1e-7)
    for k, kn in G.knobs_of(z):
        ishape = kn.pop("image_shape")
        wbf = kn["nms_method"] == "wbf"
        if wbf:
            kn = dict(kn, nms_method="diou", use_wbf=True)
        for b in range(B):
            one = [p[b:b + 1] for p in preds]
            boxes, classes, scores = dec.postprocess(one, ishape, (S, S), **kn)
            ref_s = z[f"k{k}_b{b}_scores"]
            assert len(scores) == len(ref_s)
            if len(ref_s) == 0:
                assert boxes.size == 0 and classes.size == 0
                continue
            # the reference's container types
            assert boxes.dtype == np.int32 and classes.dtype == np.int32 and scores.dtype == np.float64
            if kn["nms_method"] == "soft" or wbf:
                np.testing.assert_allclose(scores, ref_s, rtol=1e-5)   # depend on the boxes (1e-5 bar)
            else:
                assert np.array_equal(scores, ref_s)        # float32 scores reproduced bit-for-bit
            assert np.array_equal(classes, z[f"k{k}_b{b}_classes"])
            ref_xywh = z[f"k{k}_b{b}_xywh"].reshape(-1, 4)
            bx, _, _ = dec.postprocess(one, ishape, (S, S), return_xyxy=False, **kn)
            assert bx.dtype == np.float64
            np.testing.assert_allclose(bx, ref_xywh, rtol=1e-5, atol=1e-4)
            diff = boxes != z[f"k{k}_b{b}_xyxy"].reshape(-1, 4)
            if diff.any():                                   # only at a .5 rounding boundary
                xy = ref_xywh.copy(); xy[:, 2:] += xy[:, :2]
                assert np.all(np.abs((xy + 0.5) - np.round(xy + 0.5))[diff] < 1e-4)
        # batched entry point == B single calls
        kn2 = dict(kn)
        batch = dec.postprocess_batch(preds, [ishape] * B, (S, S), kn2.pop("max_boxes"),
                                      kn2.pop("confidence"), kn2.pop("nms_threshold"),
                                      "wbf" if wbf else kn2.pop("nms_method"))
        for b in range(B):
            np.testing.assert_allclose(batch[b][2], z[f"k{k}_b{b}_scores"], rtol=1e-5)


def test_nms_classes_against_reference_golden():
    z = np.load(G.GOLDEN + "/nms_cases.npz")
    i = 0
    while f"n{i}_boxes" in z:
        boxes, scores, classes = z[f"n{i}_boxes"], z[f"n{i}_scores"], z[f"n{i}_classes"]
        for name, cls in (("diou", DIoUNMS), ("standard", StandardNMS), ("cluster", ClusterNMS)):
            for thr in (0.3, 0.5):
                kb, kc, ks = cls().apply_nms(boxes, classes, scores, thr, 0.0)
                assert isinstance(kb, list) and len(kb) == 1
                assert np.array_equal(ks[0], z[f"n{i}_{name}_{thr}_scores"])
                assert np.array_equal(kb[0], z[f"n{i}_{name}_{thr}_boxes"])
        kb, kc, ks = nms_boxes(boxes, classes, scores, 0.5, use_diou=True)
        assert np.array_equal(ks[0], z[f"n{i}_diou_0.5_scores"])
        half = len(boxes) // 2
        for ct in ("avg", "max", "box_and_model_avg"):
            fb, fc, fs = WeightedBoxesFusion(iou_thr=0.4, skip_box_thr=0.05, conf_type=ct).fuse_boxes(
                [boxes[:half], boxes[half:]], [classes[:half], classes[half:]],
                [scores[:half], scores[half:]], (600, 600), weights=[1.0, 0.6])
            ref_b = z[f"n{i}_wbf_{ct}_boxes"]
            if len(ref_b) == 0:
                assert fb == []
                continue
            np.testing.assert_allclose(fb[0], ref_b, rtol=1e-12)
            np.testing.assert_allclose(fs[0], z[f"n{i}_wbf_{ct}_scores"], rtol=1e-12)
            assert np.array_equal(fc[0], z[f"n{i}_wbf_{ct}_classes"])
        for sigma in (0.5, 0.1):
            kb, kc, ks = SoftNMS(sigma=sigma).apply_nms(boxes, classes, scores, 0.5, 0.0)
            np.testing.assert_allclose(ks[0], z[f"n{i}_soft_{sigma}_scores"], rtol=1e-12)
            assert np.array_equal(kb[0], z[f"n{i}_soft_{sigma}_boxes"])
        i += 1


def test_postprocess_gpu_signature_and_layout():
    z = np.load(G.files("decode")[0])
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    preds = G.preds_of(z)
    B = preds[0].shape[0]
    boxes, scores, classes, valid = multigriddet_postprocess_gpu(
        preds, np.tile(np.array([[S, S]]), (B, 1)), anchors, C, (S, S), max_boxes=50,
        confidence=0.1, nms_threshold=0.45)
    assert boxes.shape == (B, 50, 4) and scores.shape == (B, 50) and valid.shape == (B,)
    assert boxes.dtype == np.float32 and (valid > 0).all()
    for b in range(B):
        k = valid[b]
        assert np.all(np.diff(scores[b, :k]) <= 0) and np.all(boxes[b, k:] == 0)
        assert np.all(boxes[b, :k, 2] >= boxes[b, :k, 0])


def test_dlpack_entry_points_zero_copy():
    """mgd_*_dlpack with torch-exported DLTensors (CUDA): same bytes as the pointer API."""
    import torch
    from multigriddet_b200 import synth
    lib = _lib.load()
    S, C, B, N = 608, 80, 4, 50
    anchors = synth.coco_anchors(np.float32)
    boxes = synth.synth_boxes(6, B, N, S, C)
    ref = engine.encode_targets(boxes, (S, S), anchors, C)
    cfg = _lib.make_head_config(anchors, C, (S, S))
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]

    def dl(t):
        cap = t.__dlpack__()
        return cap, ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")

    d_boxes = torch.from_numpy(boxes).cuda()
    outs = [torch.empty((B, g, g, 88), dtype=torch.float32, device="cuda") for g in (19, 38, 76)]
    caps = [dl(d_boxes)] + [dl(o) for o in outs]
    arr = (ctypes.c_void_p * 3)(*[c[1] for c in caps[1:]])
    stats = (ctypes.c_longlong * 4)()
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.mgd_encode_targets_dlpack(ctypes.byref(cfg), caps[0][1], arr, ctypes.c_void_p(stream),
                                       _lib.FLAG_SYNC, stats)
    _lib.raise_for_status(rc)
    for o, r in zip(outs, ref):
        assert np.array_equal(o.cpu().numpy(), r)
    # host tensors through DLPack as well (kDLCPU)
    h_boxes = torch.from_numpy(boxes)
    h_outs = [torch.empty((B, g, g, 88), dtype=torch.float32) for g in (19, 38, 76)]
    caps = [dl(h_boxes)] + [dl(o) for o in h_outs]
    arr = (ctypes.c_void_p * 3)(*[c[1] for c in caps[1:]])
    _lib.raise_for_status(lib.mgd_encode_targets_dlpack(ctypes.byref(cfg), caps[0][1], arr, None,
                                                        _lib.FLAG_SYNC, stats))
    for o, r in zip(h_outs, ref):
        assert np.array_equal(o.numpy(), r)
    # a wrong shape is rejected, not read out of bounds
    bad = torch.empty((B, 19, 19, 80), dtype=torch.float32, device="cuda")
    caps2 = [dl(bad), dl(outs[1]), dl(outs[2])]
    arr2 = (ctypes.c_void_p * 3)(*[c[1] for c in caps2])
    cap_b = dl(d_boxes)
    assert lib.mgd_encode_targets_dlpack(ctypes.byref(cfg), cap_b[1], arr2, ctypes.c_void_p(stream),
                                         _lib.FLAG_SYNC, stats) == _lib.ERR_INVALID_ARGUMENT


def test_threads_share_one_decoder():
    """evaluator.py:283-286 calls postprocess from up to 8 threads on one object."""
    from concurrent.futures import ThreadPoolExecutor
    z = np.load(G.files("decode")[-1])
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    preds = G.preds_of(z)
    dec = MultiGridDecoder(anchors, C, input_shape=(S, S))
    one = [p[0:1] for p in preds]
    ref = dec.postprocess(one, (S, S), (S, S), confidence=0.001, nms_threshold=0.45)
    with ThreadPoolExecutor(8) as ex:
        outs = list(ex.map(lambda _: dec.postprocess(one, (S, S), (S, S), confidence=0.001,
                                                     nms_threshold=0.45), range(64)))
    for o in outs:
        assert all(np.array_equal(a, b) for a, b in zip(o, ref))


def test_per_class_nms_against_reference_composition():
    """a-9 through the C ABI: per_class keeps == the reference's per-partition composition."""
    z = np.load(G.GOLDEN + "/nms_cases.npz")
    pc = np.load(G.GOLDEN + "/perclass_cases.npz")
    i = 0
    while f"n{i}_boxes" in z:
        boxes, scores, classes = z[f"n{i}_boxes"], z[f"n{i}_scores"], z[f"n{i}_classes"]
        for name in ("diou", "standard"):
            for thr in (0.3, 0.5):
                keep = engine.nms(boxes, scores, classes, thr, name, True)
                for mx in (1000, 20):
                    k = keep[:mx]
                    assert np.array_equal(scores[k], pc[f"n{i}_{name}_{thr}_{mx}_scores"])
                    assert np.array_equal(boxes[k], pc[f"n{i}_{name}_{thr}_{mx}_boxes"])
                    assert np.array_equal(classes[k], pc[f"n{i}_{name}_{thr}_{mx}_classes"])
        i += 1


def test_coco608_against_reference_golden(c_oracle, nms_kernel_choice):
    """COCO 608 (configs[2] geometry) against detections of the reference itself, in the
    class-agnostic and the per-class mode; inputs regenerated from seeds (SHA-256 checked)."""
    z = np.load(G.GOLDEN + "/coco608_detections.npz")
    S, C, anchors, preds, sha = G.coco608_inputs(c_oracle.encode_targets)
    assert sha == str(z["sha256"]), "regenerated head outputs differ from the generator's"
    B = len(preds[0])
    for k in range(int(z["n_knobs"])):
        kn = dict(max_boxes=int(z[f"k{k}_max_boxes"]), confidence=float(z[f"k{k}_confidence"]),
                  nms_threshold=float(z[f"k{k}_nms_threshold"]), nms_method="diou")
        ishape = tuple(int(v) for v in z[f"k{k}_image_shape"])
        for per_class, tag in ((False, ""), (True, "pc_")):
            got = engine.decode_nms(preds, ishape, (S, S), anchors, C, per_class=per_class,
                                    return_stats=True, **kn)
            if not per_class:
                assert got["stats"]["n_candidates"] == sum(int(z[f"k{k}_b{b}_candidates"]) for b in range(B))
            for b in range(B):
                ref_s = z[f"k{k}_b{b}_{tag}scores"]
                n = len(ref_s)
                assert int(got["counts"][b]) == n
                assert np.array_equal(got["scores"][b, :n], ref_s)          # bit-for-bit float32 scores
                assert np.array_equal(got["classes"][b, :n], z[f"k{k}_b{b}_{tag}classes"])
                diff = got["boxes_xyxy"][b, :n] != z[f"k{k}_b{b}_{tag}xyxy"].reshape(-1, 4)
                if diff.any():                               # only at a .5 rounding boundary
                    xy = got["boxes_xywh"][b, :n].copy(); xy[:, 2:] += xy[:, :2]
                    assert np.all(np.abs((xy + 0.5) - np.round(xy + 0.5))[diff] < 1e-4)
