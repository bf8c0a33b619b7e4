"""NumPy restatement of the reference's detection-head grid path (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): never imported by the
product package.  Parity status: PINNED against the reference executed from
``/root/reference`` (``tests/test_oracle_vs_reference.py``) and against the
golden vectors in ``tests/golden/`` (``tests/test_oracle_golden.py``).

Each function cites the reference lines it restates (paths relative to
``/root/reference``).  The arithmetic *dtype path* of the reference is kept on
purpose (float32 transcendentals and score product, float64 box maths, float32
letterbox constants) because the parity bar is bit-exact indices / keep sets.

Deviations, all where the reference is unspecified or undefined:

* ties: ``np.argsort`` is unstable on AVX hosts (``generators.py:2530``,
  ``nms.py:161``).  The oracle fixes the portable rule the CUDA kernels use:
  anchor ties -> lowest global anchor index; score ties -> lowest candidate
  index (layer-major, row-major: the order ``decode_predictions`` concatenates).
* ``nms_method='standard'`` raises ``NotImplementedError`` in the reference's
  ``handle_predictions`` (``multigrid_decode.py:296-297`` -> ``nms.py:39``);
  here it is routed to the IoU greedy NMS (``nms.py:83-148``).
* per-class NMS has no runnable reference (``gpu_postprocess.py:207`` is a TF
  op).  It is DEFINED as: partition candidates by argmax class, run the
  reference greedy NMS per partition, merge, order by (score desc, index asc),
  keep the first ``max_boxes``.
* batches: ``handle_predictions`` flattens the batch (``multigrid_decode.py:
  271-278``); every caller uses batch 1, so a batch here means B independent
  batch-1 calls, each with its own ``image_shape``.
"""
from __future__ import annotations

import numpy as np
from scipy.special import expit, softmax

_STRIDES = (32, 16, 8, 4, 2)


# --------------------------------------------------------------------------
# encode  (multigriddet/data/generators.py:2473-2544, 3393-3473)
# --------------------------------------------------------------------------

def default_grid_shapes(input_shape, num_layers):
    """generators.py:3423 -- ``input_shape // {32,16,8,4,2}[l]``."""
    shp = np.array(input_shape, dtype=np.int32)
    return [shp // _STRIDES[l] for l in range(num_layers)]


def anchor_iol(wh, anchor_table):
    """Rounded intersection-over-largest of one box against every anchor.

    generators.py:2486-2494 (``iol_common_center``) + ``np.round(.,3)`` at :2529.
    ``wh`` is a float32 (2,) vector; ``anchor_table`` is (K,2) float32 or
    float64 and decides the dtype everything after the box area runs in.
    """
    clipped = np.minimum(wh[np.newaxis, :], anchor_table)
    box_area = wh[0] * wh[1]                      # float32 product (:2489)
    anchor_area = anchor_table[:, 0] * anchor_table[:, 1]
    largest = np.maximum(np.array([box_area]), anchor_area)
    return np.round((clipped[:, 0] * clipped[:, 1]) / largest, 3)


def match_anchor(wh, anchors):
    """(layer, anchor-in-layer, rounded IoLs).  generators.py:2514-2544 with
    ``multi_anchor_assign=False`` (the only reachable branch, :3435).

    Tie rule: lowest global anchor index (portable ``argsort`` behaviour)."""
    table = np.concatenate([np.asarray(a) for a in anchors], axis=0)
    iols = anchor_iol(wh, table)
    best = int(np.argsort(-iols, kind="stable")[0])
    start = 0
    for layer, a in enumerate(anchors):
        if best < start + len(a):
            return layer, best - start, iols
        start += len(a)
    raise AssertionError("unreachable")


def encode_targets(true_boxes, input_shape, anchors, num_classes,
                   grid_shapes=None, return_stats=False):
    """Sequential multi-grid y_true encoder -- generators.py:3393-3473.

    true_boxes : (B, N, 5) ``[x1, y1, x2, y2, class]`` pixels, zero rows = padding
    returns    : list of L float32 arrays (B, Gh, Gw, 5 + A_l + C)
    """
    raw = np.asarray(true_boxes)
    if not (raw[..., 4] < num_classes).all():                       # :3409
        raise AssertionError("class id must be less than num_classes")
    tb = np.array(raw, dtype=np.float32)
    in_shape = np.array(input_shape, dtype=np.int32)
    num_layers = len(anchors)
    centre = (tb[..., 0:2] + tb[..., 2:4]) // 2                     # :3415 f32 floor-div
    extent = tb[..., 2:4] - tb[..., 0:2]                            # :3416
    cls_col = tb[..., 4]
    if grid_shapes is None:
        grid_shapes = default_grid_shapes(in_shape, num_layers)
    widths = [5 + len(anchors[l]) + num_classes for l in range(num_layers)]
    y_true = [np.zeros((tb.shape[0], int(grid_shapes[l][0]), int(grid_shapes[l][1]),
                        widths[l]), dtype=np.float32) for l in range(num_layers)]
    n_valid = 0
    n_skipped = 0
    for b in range(tb.shape[0]):
        for t in range(tb.shape[1]):
            bw = extent[b, t, 0]
            bh = extent[b, t, 1]
            if bw * bh <= 0.0:                                      # :3431
                continue
            n_valid += 1
            layer, k, _ = match_anchor(extent[b, t], anchors)
            g0 = grid_shapes[layer][0]
            g1 = grid_shapes[layer][1]
            cls = int(cls_col[b, t].astype(np.int32))               # :3437
            # :3438-3439  f32 * np.float64 -> float64 under NumPy 2 promotion
            gx = np.float64(centre[b, t, 0]) * (np.float64(g0) / np.float64(in_shape[0]))
            gy = np.float64(centre[b, t, 1]) * (np.float64(g1) / np.float64(in_shape[1]))
            col = int(gx)                                           # :3441 truncation
            row = int(gy)
            fx = float(gx - col)
            fy = float(gy - row)
            aw = anchors[layer][k][0]
            ah = anchors[layer][k][1]
            tw = np.log(max(bw / aw, 1e-3))                         # :3446-3449
            th = np.log(max(bh / ah, 1e-3))
            written = 0
            out = y_true[layer]
            for dx in (-1, 0, 1):                                   # :3454 x outer
                cc = col + dx
                for dy in (-1, 0, 1):                               # :3456 y inner
                    rr = row + dy
                    if cc < 0 or cc >= g0 or rr < 0 or rr >= g1:    # :3459-3462
                        continue
                    if out[b, rr, cc, 4] == 1 and written >= 3:     # :3463
                        n_skipped += 1
                        continue
                    out[b, rr, cc] *= 0
                    out[b, rr, cc, 0:4] = [-dx + fx, -dy + fy, tw, th]
                    out[b, rr, cc, 4] = 1.0
                    out[b, rr, cc, 5 + k] = 1.0
                    out[b, rr, cc, 5 + len(anchors[layer]) + cls] = 1.0
                    written += 1
    if return_stats:
        return y_true, {"n_valid_boxes": n_valid, "n_skipped_writes": n_skipped}
    return y_true


def encode_targets_tf_compat(true_boxes, input_shape, anchors, num_classes, grid_shapes=None,
                             return_stats=False):
    """The TensorFlow encoder's semantics -- generators.py:2696-3390
    (``tf_preprocess_true_boxes``, the reference's default training path).

    PINNED AGAINST THE REFERENCE'S OWN CODE OVER A TF-OP STAND-IN: TensorFlow is not installed
    in this image, so the reference function cannot run as it is; its source is executed
    statement by statement with ``oracle/tf_shim.py`` answering the ``tf.*`` calls in NumPy
    (``ref_loader.load_tf_encoder``), and this restatement equals it bit for bit on random and
    adversarial inputs (tests/test_oracle_vs_reference.py; fixtures tests/golden/tfencode_*.npz).
    What stays assumed is the semantics of the primitive ops listed in tf_shim's header
    (notably: duplicate scatter updates applied in order, TensorFlow's CPU behaviour, and the
    float32 log -- Eigen's in TensorFlow -- which is why tw/th are only claimed to 1e-5).
    It differs from the NumPy encoder above in every one of these points:

    * centre = (x1y1 + x2y2) / 2.0, no floor (:2730); valid <=> w*h > 0 (:2757);
    * IoL = inter / (max(area_box, area_anchor) + 1e-7), float32, no rounding (:2823-2825);
      layer = first argmax of the per-layer maxima, anchor = first argmax inside that
      layer (:2882-2931) == first global argmax;
    * cell: col = int(cx), row = int(cy) with cx = x * (grid_w / input_w) in float32
      (:2960-2979);
    * every in-bounds cell of the 3x3 block is written: the occupancy rule reads a
      tensor that is still all zero (:3245, one scatter per layer :3370) so it never fires;
    * stored xy = [-kj + ty, -ki + tx] with ki the ROW offset, kj the COLUMN offset,
      tx = frac(cx), ty = frac(cy) (:2978-2979, :3337-3339): channel 0 carries the column
      offset plus the fractional ROW position and vice versa;
    * duplicates inside one ``tensor_scatter_nd_update`` (:3370): last update wins on
      CPU, i.e. the highest box index covering the cell (GPU/XLA order is unspecified);
    * no class-range assertion: ``tf.one_hot`` (:3351) leaves the class channels zero
      for ids outside [0, C).
    """
    tb = np.array(np.asarray(true_boxes), dtype=np.float32)
    B, N = tb.shape[0], tb.shape[1]
    num_layers = len(anchors)
    in_h, in_w = int(input_shape[0]), int(input_shape[1])
    if grid_shapes is None:
        grid_shapes = default_grid_shapes(np.array(input_shape, dtype=np.int32), num_layers)
    f32 = np.float32
    centre = (tb[..., 0:2] + tb[..., 2:4]) / f32(2.0)                # :2730
    extent = tb[..., 2:4] - tb[..., 0:2]                             # :2731
    cls_all = tb[..., 4].astype(np.int32)                            # :2732 tf.cast truncates
    area = extent[..., 0] * extent[..., 1]
    valid = area > 0.0                                               # :2757
    table = np.concatenate([np.asarray(a, dtype=np.float32) for a in anchors], axis=0)
    counts = [len(a) for a in anchors]
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    inter = np.minimum(extent[:, :, None, :], table[None, None])     # :2811
    inter_area = inter[..., 0] * inter[..., 1]
    anchor_area = table[:, 0] * table[:, 1]
    largest = np.maximum(area[..., None], anchor_area[None, None])
    with np.errstate(invalid="ignore", divide="ignore"):
        iols = inter_area / (largest + f32(1e-7))                    # :2825 keras epsilon
    per_layer_max = np.stack([iols[..., s:s + c].max(-1) for s, c in zip(starts, counts)], -1)
    layer_of = per_layer_max.argmax(-1)                              # :2895 first maximum
    anchor_in = np.stack([iols[..., s:s + c].argmax(-1) for s, c in zip(starts, counts)], -1)
    widths = [5 + counts[l] + num_classes for l in range(num_layers)]
    y_true = [np.zeros((B, int(grid_shapes[l][0]), int(grid_shapes[l][1]), widths[l]),
                       dtype=np.float32) for l in range(num_layers)]
    n_valid = int(valid.sum())
    for b in range(B):
        for t in range(N):                                           # ascending: last write wins
            if not valid[b, t]:
                continue
            layer = int(layer_of[b, t])
            k = int(anchor_in[b, t, layer])
            gh, gw = int(grid_shapes[layer][0]), int(grid_shapes[layer][1])
            cx = centre[b, t, 0] * (f32(gw) / f32(in_w))             # :2960-2965 float32
            cy = centre[b, t, 1] * (f32(gh) / f32(in_h))
            col, row = int(cx), int(cy)                              # :2974-2975 truncation
            tx = f32(cx - f32(col))
            ty = f32(cy - f32(row))
            aw, ah = table[starts[layer] + k]
            tw = np.log(np.maximum(extent[b, t, 0] / aw, f32(1e-3))) # :3330-3333
            th = np.log(np.maximum(extent[b, t, 1] / ah, f32(1e-3)))
            cls = int(cls_all[b, t])
            for ki in (-1, 0, 1):                                    # :2993 row offset, outer
                for kj in (-1, 0, 1):                                # column offset, inner
                    rr, cc = row + ki, col + kj
                    if rr < 0 or rr >= gh or cc < 0 or cc >= gw:     # :3083-3091
                        continue
                    rowv = y_true[layer][b, rr, cc]
                    rowv[:] = 0.0                                    # the update replaces the row
                    rowv[0] = f32(-kj) + ty                          # :3337
                    rowv[1] = f32(-ki) + tx                          # :3338
                    rowv[2] = tw
                    rowv[3] = th
                    rowv[4] = 1.0
                    rowv[5 + k] = 1.0
                    if 0 <= cls < num_classes:                       # tf.one_hot :3351
                        rowv[5 + counts[layer] + cls] = 1.0
    if return_stats:
        return y_true, {"n_valid_boxes": n_valid, "n_skipped_writes": 0}
    return y_true


def encode_targets_parallel_scheme(true_boxes, input_shape, anchors, num_classes,
                                   grid_shapes=None):
    """The order-free formulation the CUDA encoder uses (encode):

    firstCover[cell] = min box index covering the cell; a box skips a candidate
    iff ``firstCover < t`` and it has already written >= 3 cells; owner[cell] =
    max box index among writers; every cell row is then produced once.  Kept in
    the oracle so the scheme itself is checked against the sequential rule on
    the CPU, independent of any kernel.
    """
    tb = np.array(np.asarray(true_boxes), dtype=np.float32)
    in_shape = np.array(input_shape, dtype=np.int32)
    num_layers = len(anchors)
    if grid_shapes is None:
        grid_shapes = default_grid_shapes(in_shape, num_layers)
    centre = (tb[..., 0:2] + tb[..., 2:4]) // 2
    extent = tb[..., 2:4] - tb[..., 0:2]
    widths = [5 + len(anchors[l]) + num_classes for l in range(num_layers)]
    y_true = [np.zeros((tb.shape[0], int(grid_shapes[l][0]), int(grid_shapes[l][1]),
                        widths[l]), dtype=np.float32) for l in range(num_layers)]
    big = np.iinfo(np.int32).max
    for b in range(tb.shape[0]):
        recs = []
        first = [np.full((int(g[1]), int(g[0])), big, dtype=np.int64) for g in grid_shapes]
        for t in range(tb.shape[1]):
            bw, bh = extent[b, t]
            if bw * bh <= 0.0:
                continue
            layer, k, _ = match_anchor(extent[b, t], anchors)
            g0, g1 = int(grid_shapes[layer][0]), int(grid_shapes[layer][1])
            gx = np.float64(centre[b, t, 0]) * (np.float64(g0) / np.float64(in_shape[0]))
            gy = np.float64(centre[b, t, 1]) * (np.float64(g1) / np.float64(in_shape[1]))
            col, row = int(gx), int(gy)
            cand = []
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    cc, rr = col + dx, row + dy
                    if 0 <= cc < g0 and 0 <= rr < g1:
                        cand.append((dx, dy, cc, rr))
                        first[layer][rr, cc] = min(first[layer][rr, cc], t)
            tw = np.log(max(bw / anchors[layer][k][0], 1e-3))
            th = np.log(max(bh / anchors[layer][k][1], 1e-3))
            recs.append((t, layer, k, float(gx - col), float(gy - row), tw, th,
                         int(tb[b, t, 4].astype(np.int32)), cand))
        owner = {}
        for (t, layer, k, fx, fy, tw, th, cls, cand) in recs:
            written = 0
            for (dx, dy, cc, rr) in cand:
                if first[layer][rr, cc] < t and written >= 3:
                    continue
                written += 1
                key = (layer, rr, cc)
                if key not in owner or owner[key][0] < t:
                    owner[key] = (t, dx, dy, k, fx, fy, tw, th, cls)
        for (layer, rr, cc), (t, dx, dy, k, fx, fy, tw, th, cls) in owner.items():
            rowv = y_true[layer][b, rr, cc]
            rowv[0:4] = [-dx + fx, -dy + fy, tw, th]
            rowv[4] = 1.0
            rowv[5 + k] = 1.0
            rowv[5 + len(anchors[layer]) + cls] = 1.0
    return y_true


# --------------------------------------------------------------------------
# decode  (multigriddet/postprocess/multigrid_decode.py)
# --------------------------------------------------------------------------

def decode_layer(pred, layer_anchors, input_shape, num_classes,
                 use_softmax=True, rescore_confidence=True):
    """One FPN level -> (B, Gh*Gw, 5+C).  multigrid_decode.py:100-183."""
    pred = np.asarray(pred)
    batch, gh, gw = pred.shape[0], pred.shape[1], pred.shape[2]
    n_anchor = len(layer_anchors)
    cols, rows = np.meshgrid(np.arange(gw), np.arange(gh))          # :119-121 int64
    cell_xy = np.stack([cols, rows], axis=-1).reshape(-1, gh, gw, 2)  # cell[i,j]=[j,i]
    t_xy = pred[..., 0:2]
    t_wh = pred[..., 2:4]
    t_obj = pred[..., 4:5]
    t_anchor = pred[..., 5:5 + n_anchor]
    t_class = pred[..., 5 + n_anchor:]
    if use_softmax:                                                 # :140-145
        p_anchor = softmax(t_anchor, axis=-1)
        p_class = softmax(t_class, axis=-1)
    else:
        p_anchor = expit(t_anchor)
        p_class = expit(t_class)
    p_obj = expit(t_obj)                                            # :147
    act_xy = np.tanh(0.15 * t_xy) + expit(0.15 * t_xy)              # :151
    xy = act_xy + cell_xy                                           # :154 -> float64
    xy /= (gh, gw)                                                  # :155
    which = np.argmax(p_anchor, axis=-1)                            # :158 first max
    wh = np.take(layer_anchors, which, axis=0) * np.exp(t_wh)       # :159-162
    wh /= input_shape                                               # :163 (in place)
    if rescore_confidence:                                          # :166-170
        p_obj = (p_obj * np.max(p_anchor, axis=-1, keepdims=True)
                 * np.max(p_class, axis=-1, keepdims=True))
    out = np.concatenate([xy, wh, p_obj, p_class], axis=-1)         # :173-175
    return out.reshape(batch, gh * gw, num_classes + 5)


def decode_predictions(preds, anchors, input_shape, num_classes,
                       use_softmax=True, rescore_confidence=True):
    """All levels, concatenated layer-major.  multigrid_decode.py:48-98."""
    if len(preds) != len(anchors):
        raise ValueError(f"Expected {len(anchors)} predictions, got {len(preds)}")
    parts = []
    batch = None
    for l, p in enumerate(preds):
        if p is None or np.size(p) == 0 or np.shape(p)[0] == 0:     # :71-72
            continue
        d = decode_layer(p, anchors[l], input_shape, num_classes,
                         use_softmax, rescore_confidence)
        if batch is None:
            batch = d.shape[0]
        elif d.shape[0] != batch:                                   # :83-86
            continue
        parts.append(d)
    if not parts:
        return np.zeros((batch or 0, 0, 5 + num_classes), dtype="float32")
    return np.concatenate(parts, axis=1)


def letterbox_constants(image_shape, model_image_size):
    """(offset_wh, scale_wh, image_wh) as float32.  multigrid_decode.py:205-216,226."""
    model = np.array(model_image_size, dtype="float32")
    image = np.array(image_shape, dtype="float32")
    fitted = np.round(image * np.min(model / image))
    offset = ((model - fitted) / 2.0 / model)[..., ::-1]
    scale = (model / fitted)[..., ::-1]
    return offset, scale, image[..., ::-1]


def correct_boxes(decoded, image_shape, model_image_size):
    """Undo the letterbox -> ``[x_min, y_min, w, h]`` in original-image pixels.

    multigrid_decode.py:185-235.  Does not modify its input (the reference
    scales ``box_wh`` through a view of its argument; the values are the same).
    """
    offset, scale, image_wh = letterbox_constants(image_shape, model_image_size)
    xy = (decoded[..., 0:2] - offset) * scale                       # :219
    wh = decoded[..., 2:4] * scale                                  # :220
    xy = xy - wh / 2.0                                              # :223
    xy = xy * image_wh                                              # :227
    wh = wh * image_wh                                              # :228
    return np.concatenate([xy, wh, decoded[..., 4:5], decoded[..., 5:]], axis=-1)


def pair_metric(box, others, use_diou):
    """IoU / DIoU of one xywh box against many.  nms.py:121-148, 189-231."""
    x1, y1, w1, h1 = box
    x2, y2, w2, h2 = others[:, 0], others[:, 1], others[:, 2], others[:, 3]
    iw = np.maximum(0.0, np.minimum(x1 + w1, x2 + w2) - np.maximum(x1, x2))
    ih = np.maximum(0.0, np.minimum(y1 + h1, y2 + h2) - np.maximum(y1, y2))
    inter = iw * ih
    union = w1 * h1 + w2 * h2 - inter
    iou = inter / (union + 1e-8)
    if not use_diou:
        return iou
    dist = ((x1 + w1 / 2) - (x2 + w2 / 2)) ** 2 + ((y1 + h1 / 2) - (y2 + h2 / 2)) ** 2
    diag = ((np.maximum(x1 + w1, x2 + w2) - np.minimum(x1, x2)) ** 2
            + (np.maximum(y1 + h1, y2 + h2) - np.minimum(y1, y2)) ** 2)
    return iou - dist / (diag + 1e-8)


def greedy_nms(boxes, scores, nms_threshold, use_diou=True, tiebreak=None,
               classes=None, per_class=False):
    """Greedy hard NMS; returns kept positions in descending-score order.

    nms.py:151-187 (DIoU), 83-119 (Standard), 320-356 (Cluster == Standard).
    ``tiebreak``: secondary ascending sort key for equal scores (candidate
    index); the reference leaves that order unspecified.
    """
    n = len(boxes)
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    if tiebreak is None:
        tiebreak = np.arange(n)
    order = np.lexsort((tiebreak, -np.asarray(scores, dtype=np.float64)))
    kept = []
    while len(order) > 0:
        cur = order[0]
        kept.append(cur)
        if len(order) == 1:
            break
        rest = order[1:]
        m = pair_metric(boxes[cur], boxes[rest], use_diou)
        survive = m < nms_threshold                                 # nms.py:180 / :112
        if per_class:
            survive = survive | (classes[rest] != classes[cur])
        order = rest[survive]
    return np.asarray(kept, dtype=np.int64)


def soft_nms(boxes, scores, sigma=0.5, score_threshold=0.001, tiebreak=None):
    """Gaussian SoftNMS; returns (kept positions in input order, decayed scores).

    nms.py:234-288."""
    n = len(boxes)
    if n == 0:
        return np.zeros((0,), dtype=np.int64), np.zeros((0,))
    if tiebreak is None:
        tiebreak = np.arange(n)
    order = np.lexsort((tiebreak, -np.asarray(scores, dtype=np.float64)))
    s = np.array(scores, dtype=np.float64).copy()
    for i in range(n):
        cur = order[i]
        if s[cur] < score_threshold:
            s[cur] = 0
            continue
        rest = order[i + 1:]
        if len(rest) == 0:
            break
        iou = pair_metric(boxes[cur], boxes[rest], use_diou=False)
        s[rest] *= np.exp(-iou ** 2 / sigma)
    keep = np.nonzero(s >= score_threshold)[0]
    return keep, s[keep]


def wbf_iou(a, b):
    """Scalar IoU of two xywh boxes as wbf.py:220-250 computes it (no epsilon)."""
    x0, y0 = max(a[0], b[0]), max(a[1], b[1])
    x1, y1 = min(a[0] + a[2], b[0] + b[2]), min(a[1] + a[3], b[1] + b[3])
    if x1 <= x0 or y1 <= y0:
        return 0.0
    inter = (x1 - x0) * (y1 - y0)
    union = a[2] * a[3] + b[2] * b[3] - inter
    return inter / union if union > 0 else 0.0


def weighted_boxes_fusion(boxes, scores, classes, box_weights=None, iou_thr=0.55,
                          skip_box_thr=0.0, conf_type="avg", tiebreak=None):
    """The reference's box fusion (wbf.py:38-218) on the concatenated boxes of all models.

    Per class (ascending id): boxes in descending score order; each unused box leads a
    cluster and absorbs the later unused boxes whose IoU with the LEADER is >= iou_thr;
    a cluster becomes its (score x weight)-weighted mean box and a fused confidence.
    Returns (fused boxes, fused scores, classes, leader positions)."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float64).reshape(-1)
    classes = np.asarray(classes).reshape(-1)
    n = len(boxes)
    weights = np.ones(n) if box_weights is None else np.asarray(box_weights, dtype=np.float64)
    if tiebreak is None:
        tiebreak = np.arange(n)
    live = np.nonzero(scores >= skip_box_thr)[0]                      # wbf.py:74
    out_b, out_s, out_c, out_lead = [], [], [], []
    for cls in np.unique(classes[live]):                              # wbf.py:98
        members = live[classes[live] == cls]
        order = members[np.lexsort((tiebreak[members], -scores[members]))]   # wbf.py:150
        used = np.zeros(len(order), dtype=bool)
        for i in range(len(order)):
            if used[i]:
                continue
            group = [order[i]]
            for j in range(i + 1, len(order)):
                if not used[j] and wbf_iou(boxes[order[i]], boxes[order[j]]) >= iou_thr:
                    group.append(order[j])
                    used[j] = True
            g = np.array(group)
            w = scores[g] * weights[g]                                # wbf.py:198-199
            w = w / np.sum(w)
            out_b.append(np.average(boxes[g], axis=0, weights=w))     # wbf.py:202
            if conf_type == "max":
                out_s.append(np.max(scores[g]))
            elif conf_type in ("box_and_model_avg", "absent_model_aware_avg"):
                out_s.append(np.mean(scores[g] * weights[g]))
            else:
                out_s.append(np.mean(scores[g]))
            out_c.append(cls)
            out_lead.append(order[i])
    if not out_b:
        return np.zeros((0, 4)), np.zeros((0,)), np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    return np.array(out_b), np.array(out_s), np.array(out_c), np.array(out_lead)


_GREEDY = {"diou": True, "cluster": False, "standard": False, "iou": False}


def postprocess_image(preds, image_shape, model_image_size, anchors, num_classes,
                      max_boxes=100, confidence=0.1, nms_threshold=0.5,
                      nms_method="diou", per_class=False, use_softmax=True,
                      rescore_confidence=True):
    """Full decode -> letterbox -> threshold -> NMS -> top-k -> xyxy for ONE image.

    multigrid_decode.py:347-395 (``postprocess``) with 237-345 and 397-422.
    ``preds`` is a list of L arrays (1, Gh, Gw, D).  Returns a dict with the
    reference outputs plus ``index`` = flat candidate index of each detection
    (position in the ``decode_predictions`` concatenation), used for keep-set
    parity.
    """
    dec = decode_predictions(preds, anchors, model_image_size, num_classes,
                             use_softmax, rescore_confidence)
    cor = correct_boxes(dec, image_shape, model_image_size)[0]
    score = cor[:, 4]
    cls_all = np.argmax(cor[:, 5:], axis=-1)                        # :268
    cand = np.nonzero(score >= confidence)[0]                       # :271
    empty = {"boxes_xywh": np.zeros((0, 4)), "boxes_xyxy": np.zeros((0, 4), np.int32),
             "classes": np.zeros((0,), np.int32), "scores": np.zeros((0,)),
             "index": np.zeros((0,), np.int64), "n_candidates": int(len(cand))}
    if len(cand) == 0:
        return empty
    boxes = cor[cand, 0:4]
    scores = score[cand]
    classes = cls_all[cand]
    if nms_method == "wbf":                                           # use_wbf=True, :281-287
        fb, fs, fc, lead = weighted_boxes_fusion(boxes, scores, classes, iou_thr=nms_threshold,
                                                 tiebreak=cand)
        if len(fb) > max_boxes:                                       # :336-345
            top = np.lexsort((np.arange(len(fs)), -fs))[:max_boxes]
            fb, fs, fc, lead = fb[top], fs[top], fc[top], lead[top]
        return {"boxes_xywh": fb, "boxes_xyxy": to_xyxy(fb, image_shape),
                "classes": fc.astype(np.int32), "scores": fs,
                "index": cand[lead].astype(np.int64), "n_candidates": int(len(cand))}
    if nms_method == "soft":
        keep, soft = soft_nms(boxes, scores, tiebreak=cand)
        # multigrid_decode.py:336-345: stays in input order unless > max_boxes
        if len(keep) > max_boxes:
            top = np.lexsort((cand[keep], -soft))[:max_boxes]
            keep, soft = keep[top], soft[top]
        out_scores = soft
    else:
        keep = greedy_nms(boxes, scores, nms_threshold, _GREEDY[nms_method],
                          tiebreak=cand, classes=classes, per_class=per_class)
        keep = keep[:max_boxes]                                     # :336-345
        out_scores = scores[keep]
    if len(keep) == 0:
        return empty
    kb = boxes[keep]
    return {"boxes_xywh": kb, "boxes_xyxy": to_xyxy(kb, image_shape),
            "classes": classes[keep].astype(np.int32), "scores": out_scores,
            "index": cand[keep].astype(np.int64), "n_candidates": int(len(cand))}


def to_xyxy(boxes_xywh, image_shape):
    """Clip to the image and round half up to int32.  multigrid_decode.py:397-422."""
    b = np.array(boxes_xywh, dtype=np.float64, copy=True)
    b[:, 2] = boxes_xywh[:, 0] + boxes_xywh[:, 2]
    b[:, 3] = boxes_xywh[:, 1] + boxes_xywh[:, 3]
    h, w = image_shape[0], image_shape[1]
    b[:, 0] = np.clip(b[:, 0], 0, w)
    b[:, 1] = np.clip(b[:, 1], 0, h)
    b[:, 2] = np.clip(b[:, 2], 0, w)
    b[:, 3] = np.clip(b[:, 3], 0, h)
    return np.floor(b + 0.5).astype("int32")


def postprocess_batch(preds, image_shapes, model_image_size, anchors, num_classes,
                      **kw):
    """B independent batch-1 calls.  ``image_shapes`` is (B,2) (h,w)."""
    image_shapes = np.asarray(image_shapes).reshape(-1, 2)
    out = []
    for b in range(np.shape(preds[0])[0]):
        one = [np.asarray(p)[b:b + 1] for p in preds]
        out.append(postprocess_image(one, tuple(int(v) for v in image_shapes[b]),
                                     model_image_size, anchors, num_classes, **kw))
    return out


# --------------------------------------------------------------------------
# parity helpers: which comparisons are fragile?  
# --------------------------------------------------------------------------

def ulp_distance_f32(a, b):
    ia = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def fragility_report(preds, image_shape, model_image_size, anchors, num_classes,
                     confidence, nms_threshold, use_diou=True, ulps=8,
                     metric_eps=1e-5, use_softmax=True, rescore_confidence=True):
    """Count comparisons whose outcome could flip under a few-ulp score change."""
    dec = decode_predictions(preds, anchors, model_image_size, num_classes,
                             use_softmax, rescore_confidence)
    cor = correct_boxes(dec, image_shape, model_image_size)[0]
    s32 = cor[:, 4].astype(np.float32)
    thr32 = np.float32(confidence)
    graze = int(np.sum(ulp_distance_f32(s32, np.full_like(s32, thr32)) <= ulps))
    cand = np.nonzero(cor[:, 4] >= confidence)[0]
    boxes = cor[cand, 0:4]
    sc = s32[cand]
    order = np.argsort(-sc, kind="stable")
    close_pairs = 0
    for a in range(len(order) - 1):
        i, j = order[a], order[a + 1]
        if ulp_distance_f32(sc[i], sc[j]) <= ulps:
            m = pair_metric(boxes[i], boxes[j:j + 1], use_diou)[0]
            if m >= nms_threshold - metric_eps:
                close_pairs += 1
    keep = greedy_nms(boxes, cor[cand, 4], nms_threshold, use_diou, tiebreak=cand)
    metric_graze = 0
    for kpos in keep[:200]:
        m = pair_metric(boxes[kpos], boxes, use_diou)
        metric_graze += int(np.sum(np.abs(m - nms_threshold) <= metric_eps))
    return {"threshold_grazing": graze, "fragile_score_pairs": close_pairs,
            "metric_grazing": metric_graze,
            "fragile": bool(graze or close_pairs or metric_graze)}


# --------------------------------------------------------------------------
# box-side pre-step of the encoder  (multigriddet/data/augmentation.py:112-164, 606-667)
# --------------------------------------------------------------------------

def reshape_boxes(boxes, src_shape, target_shape, padding_shape, offset,
                  horizontal_flip=False, vertical_flip=False):
    """``reshape_boxes`` (augmentation.py:112-164) WITHOUT the row shuffle of :146 and without
    modifying the caller's array.  int32 boxes keep NumPy's float -> int32 store (truncation
    toward zero) exactly where the reference assigns into the int array (:149-150)."""
    b = np.array(boxes)
    if len(b) == 0:
        return b
    src_w, src_h = src_shape
    target_w, target_h = target_shape
    padding_w, padding_h = padding_shape
    dx, dy = offset
    b[:, [0, 2]] = b[:, [0, 2]] * padding_w / src_w + dx
    b[:, [1, 3]] = b[:, [1, 3]] * padding_h / src_h + dy
    if horizontal_flip:
        b[:, [0, 2]] = target_w - b[:, [2, 0]]
    if vertical_flip:
        b[:, [1, 3]] = target_h - b[:, [3, 1]]
    b[:, 0:2][b[:, 0:2] < 0] = 0
    b[:, 2][b[:, 2] > target_w] = target_w
    b[:, 3][b[:, 3] > target_h] = target_h
    w = b[:, 2] - b[:, 0]
    h = b[:, 3] - b[:, 1]
    return b[np.logical_and(w > 1, h > 1)]


def merge_mosaic_bboxes(bboxes, crop_x, crop_y, image_size):
    """``merge_mosaic_bboxes`` (augmentation.py:606-667): quadrant order top-left,
    bottom-left, bottom-right, top-right."""
    bboxes = np.asarray(bboxes, dtype=np.float64)
    max_boxes = bboxes.shape[1]
    height, width = image_size
    merged = []
    for q in range(4):
        for box in bboxes[q]:
            x1, y1, x2, y2 = box[0], box[1], box[2], box[3]
            cut_y = y2 > crop_y and y1 < crop_y
            cut_x = x2 > crop_x and x1 < crop_x
            if q == 0:
                if y1 > crop_y or x1 > crop_x:
                    continue
                if cut_y: y2 = crop_y
                if cut_x: x2 = crop_x
            elif q == 1:
                if y2 < crop_y or x1 > crop_x:
                    continue
                if cut_y: y1 = crop_y
                if cut_x: x2 = crop_x
            elif q == 2:
                if y2 < crop_y or x2 < crop_x:
                    continue
                if cut_y: y1 = crop_y
                if cut_x: x1 = crop_x
            else:
                if y1 > crop_y or x2 < crop_x:
                    continue
                if cut_y: y2 = crop_y
                if cut_x: x1 = crop_x
            if abs(x2 - x1) < max(10, width * 0.01) or abs(y2 - y1) < max(10, height * 0.01):
                continue
            merged.append([x1, y1, x2, y2, box[4]])
    merged = merged[:max_boxes]
    out = np.zeros((max_boxes, 5))
    if merged:
        out[:len(merged)] = merged
    return out


# --------------------------------------------------------------------------
# tf.data box pre-step (generators.py:1859-1916, 227-256, 1963-2034).  TensorFlow is not
# installed; every op restated here is an IEEE float32 multiply / divide / add or an int
# truncation, in the order the TF graph applies them, and the restatement equals the reference
# functions' own source executed over oracle/tf_shim.py bit for bit
# (ref_loader.load_tf_box_prestep, tests/test_box_transforms.py, tests/golden/tfboxes_cases.npz).
# --------------------------------------------------------------------------

def tf_letterbox_params(src_h, src_w, input_shape, multiscale_shape=None):
    """(sx, sy, pad_left, pad_top) as float32 scalars, following tf_letterbox_resize (:167-209)
    and _preprocess_image_and_boxes (:1866-1916)."""
    f = np.float32
    th, tw = f(input_shape[0]), f(input_shape[1])
    sh, sw = f(src_h), f(src_w)
    if multiscale_shape is not None and multiscale_shape[0] > 0 and multiscale_shape[1] > 0:
        scale_h = f(multiscale_shape[0]) / th                       # :1872-1873 (base = input_shape)
        scale_w = f(multiscale_shape[1]) / tw
        scaled_h = int(f(sh * scale_h))                             # tf.cast(..., tf.int32) :1879-1880
        scaled_w = int(f(sw * scale_w))
        fh, fw = f(scaled_h), f(scaled_w)
        ls0 = min(tw / fw, th / fh)                                 # :186-187
        new_w, new_h = int(f(fw * ls0)), int(f(fh * ls0))           # :190-191
        ls = min(f(new_w) / fw, f(new_h) / fh)                      # :1887-1890
        sx, sy = f(scale_w * ls), f(scale_h * ls)                   # :1891-1892
    else:
        s0 = min(tw / sw, th / sh)
        new_w, new_h = int(f(sw * s0)), int(f(sh * s0))
        sc = min(f(new_w) / sw, f(new_h) / sh)                      # :1906-1909
        sx = sy = f(sc)
    pad_left = (int(input_shape[1]) - new_w) // 2                   # :197-200
    pad_top = (int(input_shape[0]) - new_h) // 2
    return sx, sy, f(pad_left), f(pad_top)


def tf_letterbox_boxes(boxes, counts, src_shapes, input_shape, max_boxes_per_image, expansion=1,
                       multiscale_shapes=None, hflip=None):
    """Batch form of the box side of build_tf_dataset up to the encoder's input:
    transform (:1891-1916), flip (:248-251), padded_batch (:1963-1976),
    _expand_box_capacity (:1983-2034).  (B, N, 5) float32 -> (B, max*expansion, 5) float32."""
    f = np.float32
    boxes = np.asarray(boxes, dtype=f)
    B, N = boxes.shape[0], boxes.shape[1]
    out = np.zeros((B, int(max_boxes_per_image) * int(expansion), 5), dtype=f)
    for b in range(B):
        ms = None if multiscale_shapes is None else multiscale_shapes[b]
        sx, sy, pl, pt = tf_letterbox_params(src_shapes[b][0], src_shapes[b][1], input_shape, ms)
        n = N if counts is None else int(counts[b])
        n = max(0, min(n, N, int(max_boxes_per_image)))
        bx = boxes[b, :n] * np.array([sx, sy, sx, sy, 1.0], dtype=f)
        bx = bx + np.array([pl, pt, pl, pt, 0.0], dtype=f)
        if hflip is not None and bool(hflip[b]):
            w = f(input_shape[1])
            bx = np.stack([w - bx[:, 2], bx[:, 1], w - bx[:, 0], bx[:, 3], bx[:, 4]], axis=1)
        out[b, :n] = bx
    return out
