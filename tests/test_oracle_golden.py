"""The oracle (NumPy and C restatements) against the reference's golden vectors.

tests/golden/*.npz were produced by running the reference itself
(oracle/gen_golden.py); these tests run everywhere (no GPU, no reference tree).
"""
import numpy as np
import pytest

import golden_util as G
from oracle import mgd_oracle as O


@pytest.mark.parametrize("path", G.files("encode"))
def test_encode_oracles_match_reference_golden(path, c_oracle):
    z = np.load(path)
    if "boxes" not in z:          # known-answer file, handled below
        return
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    ref = G.dense_y_true(z)
    got = O.encode_targets(z["boxes"], (S, S), anchors, C)
    G.assert_encode_matches(got, ref, exact_floats=G.numpy_pinned())
    got = O.encode_targets_parallel_scheme(z["boxes"], (S, S), anchors, C)
    G.assert_encode_matches(got, ref, exact_floats=G.numpy_pinned())
    got = c_oracle.encode_targets(z["boxes"], (S, S), anchors, C)
    G.assert_encode_matches(got, ref, exact_floats=True)


def test_encode_known_answers(c_oracle):
    z = np.load(G.GOLDEN + "/encode_known_answer.npz")
    anchors = [np.array(a, dtype=np.float32) for a in z["anchors"]]
    for i in range(3):
        ref = G.dense_y_true(z, prefix=f"c{i}_")
        for impl in (O.encode_targets, c_oracle.encode_targets):
            got = impl(z[f"box{i}"], (608, 608), anchors, int(z[f"C{i}"]))
            G.assert_encode_matches(got, ref, exact_floats=G.numpy_pinned() or impl is c_oracle.encode_targets)
    # the values SURVEY.md 8c quotes from the reference's own test inputs
    y = G.dense_y_true(z, prefix="c0_")
    np.testing.assert_allclose(y[2][0, 38, 38, :4], [0, 0, -0.14842002, -0.11778303], rtol=1e-6)
    y = G.dense_y_true(z, prefix="c1_")
    np.testing.assert_allclose(y[1][0, 14, 8, :4], [0.75, 0.375, 0.25489223, 0.28768212], rtol=1e-6)
    # test_9cell_alignment.py's box: stored xy stays in [-1, 2)
    y = G.dense_y_true(z, prefix="c2_")
    pos = [l for l in range(3) if y[l].any()]
    assert len(pos) == 1
    cells = y[pos[0]][y[pos[0]][..., 4] == 1]
    assert len(cells) == 9 and cells[:, :2].min() >= -1 and cells[:, :2].max() < 2


@pytest.mark.parametrize("path", G.files("decode"))
def test_decode_oracles_match_reference_golden(path, c_oracle):
    z = np.load(path, allow_pickle=False)
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    preds = G.preds_of(z)
    B = preds[0].shape[0]
    dense = O.decode_predictions(preds, anchors, (S, S), C)
    if G.numpy_pinned():
        assert np.array_equal(dense[:, ::37, :], z["dense_sample"])
    else:
        np.testing.assert_allclose(dense[:, ::37, :], z["dense_sample"], rtol=1e-6, atol=1e-30)
    for k, kn in G.knobs_of(z):
        ishape = kn.pop("image_shape")
        py = O.postprocess_batch(preds, np.tile(np.array(ishape), (B, 1)), (S, S), anchors, C, **kn)
        soft = kn["nms_method"] in ("soft", "wbf")      # restated in NumPy only
        cc = None if soft else c_oracle.decode_nms(preds, [ishape], (S, S), anchors, C, **kn)
        for b in range(B):
            ref_s = z[f"k{k}_b{b}_scores"]
            n = len(ref_s)
            assert len(py[b]["scores"]) == n and (soft or int(cc["counts"][b]) == n)
            if G.numpy_pinned():
                assert np.array_equal(py[b]["scores"], ref_s)
                assert np.array_equal(py[b]["boxes_xywh"], z[f"k{k}_b{b}_xywh"].reshape(-1, 4))
            else:
                np.testing.assert_allclose(py[b]["scores"], ref_s, rtol=1e-6)
                np.testing.assert_allclose(py[b]["boxes_xywh"], z[f"k{k}_b{b}_xywh"].reshape(-1, 4),
                                           rtol=1e-6, atol=1e-4)
            assert np.array_equal(py[b]["boxes_xyxy"], z[f"k{k}_b{b}_xyxy"].reshape(-1, 4))
            assert np.array_equal(py[b]["classes"], z[f"k{k}_b{b}_classes"])
            if soft:
                continue
            # the C restatement: bit-exact too (NumPy pinned to libm in conftest.py)
            assert np.array_equal(cc["scores"][b, :n], ref_s)
            assert np.array_equal(cc["boxes_xywh"][b, :n], z[f"k{k}_b{b}_xywh"].reshape(-1, 4))
            assert np.array_equal(cc["boxes_xyxy"][b, :n], z[f"k{k}_b{b}_xyxy"].reshape(-1, 4))
            assert np.array_equal(cc["classes"][b, :n], z[f"k{k}_b{b}_classes"])
            assert np.array_equal(cc["index"][b, :n], py[b]["index"])


def test_nms_oracle_matches_reference_golden():
    z = np.load(G.GOLDEN + "/nms_cases.npz")
    i = 0
    while f"n{i}_boxes" in z:
        boxes, scores = z[f"n{i}_boxes"], z[f"n{i}_scores"]
        for name, diou in (("diou", True), ("standard", False), ("cluster", False)):
            for thr in (0.3, 0.5):
                keep = O.greedy_nms(boxes, scores, thr, diou)
                assert np.array_equal(scores[keep], z[f"n{i}_{name}_{thr}_scores"])
                assert np.array_equal(boxes[keep], z[f"n{i}_{name}_{thr}_boxes"])
        half = len(boxes) // 2
        w = np.concatenate([np.full(half, 1.0), np.full(len(boxes) - half, 0.6)])
        for ct in ("avg", "max", "box_and_model_avg"):
            fb, fs, fc, _ = O.weighted_boxes_fusion(boxes, scores, z[f"n{i}_classes"], w, iou_thr=0.4,
                                                    skip_box_thr=0.05, conf_type=ct)
            assert np.array_equal(fb, z[f"n{i}_wbf_{ct}_boxes"])
            assert np.array_equal(fs, z[f"n{i}_wbf_{ct}_scores"])
            assert np.array_equal(fc, z[f"n{i}_wbf_{ct}_classes"])
        for sigma in (0.5, 0.1):
            keep, soft = O.soft_nms(boxes, scores, sigma=sigma)
            assert np.array_equal(soft, z[f"n{i}_soft_{sigma}_scores"])
            assert np.array_equal(boxes[keep], z[f"n{i}_soft_{sigma}_boxes"])
        i += 1
    assert i == 4


def test_per_class_nms_oracle_matches_reference_composition():
    """per-class NMS := the reference's greedy NMS per class partition, merged
    by score, top-k (perclass_cases.npz holds that composition run on reference code)."""
    z = np.load(G.GOLDEN + "/nms_cases.npz")
    pc = np.load(G.GOLDEN + "/perclass_cases.npz")
    i = 0
    while f"n{i}_boxes" in z:
        boxes, scores, classes = z[f"n{i}_boxes"], z[f"n{i}_scores"], z[f"n{i}_classes"]
        for name, diou in (("diou", True), ("standard", False)):
            for thr in (0.3, 0.5):
                keep = O.greedy_nms(boxes, scores, thr, diou, classes=classes, per_class=True)
                for mx in (1000, 20):
                    k = keep[:mx]
                    assert np.array_equal(scores[k], pc[f"n{i}_{name}_{thr}_{mx}_scores"])
                    assert np.array_equal(boxes[k], pc[f"n{i}_{name}_{thr}_{mx}_boxes"])
                    assert np.array_equal(classes[k], pc[f"n{i}_{name}_{thr}_{mx}_classes"])
        i += 1
    assert i == 4


def test_coco608_oracle_matches_reference_golden(c_oracle):
    """Full-size COCO head: class-agnostic and per-class detections of the reference."""
    z = np.load(G.GOLDEN + "/coco608_detections.npz")
    S, C, anchors, preds, sha = G.coco608_inputs(c_oracle.encode_targets)
    assert sha == str(z["sha256"]), "regenerated head outputs differ from the generator's"
    dense = O.decode_predictions(preds, anchors, (S, S), C)
    if G.numpy_pinned():
        assert np.array_equal(dense[:, ::97, :], z["dense_sample"])
    else:
        np.testing.assert_allclose(dense[:, ::97, :], z["dense_sample"], rtol=1e-6, atol=1e-30)
    for k in range(int(z["n_knobs"])):
        kn = dict(max_boxes=int(z[f"k{k}_max_boxes"]), confidence=float(z[f"k{k}_confidence"]),
                  nms_threshold=float(z[f"k{k}_nms_threshold"]), nms_method="diou")
        ishape = tuple(int(v) for v in z[f"k{k}_image_shape"])
        for per_class, tag in ((False, ""), (True, "pc_")):
            cc = c_oracle.decode_nms(preds, [ishape], (S, S), anchors, C, per_class=per_class, **kn)
            for b in range(len(preds[0])):
                ref_s = z[f"k{k}_b{b}_{tag}scores"]
                n = len(ref_s)
                assert int(cc["counts"][b]) == n
                assert np.array_equal(cc["scores"][b, :n], ref_s)
                assert np.array_equal(cc["classes"][b, :n], z[f"k{k}_b{b}_{tag}classes"])
                assert np.array_equal(cc["boxes_xyxy"][b, :n], z[f"k{k}_b{b}_{tag}xyxy"].reshape(-1, 4))


@pytest.mark.parametrize("path", G.files("tfencode"))
def test_tf_encoder_oracle_matches_reference_golden(path):
    """tests/golden/tfencode_*.npz: outputs of the reference's own tf_preprocess_true_boxes
    (generators.py:2696-3390) executed over oracle/tf_shim.py in the build container."""
    z = np.load(path)
    anchors = G.anchors_of(z)
    S, C = int(z["S"]), int(z["C"])
    grids = [(S // 32,) * 2, (S // 16,) * 2, (S // 8,) * 2]
    got = O.encode_targets_tf_compat(z["boxes"], (S, S), anchors, C, grids)
    G.assert_encode_matches(got, G.dense_y_true(z), exact_floats=G.numpy_pinned())
