"""Seeded synthetic inputs for the detection-head grid path.

Used by the parity tests, ``__graft_entry__.smoke()`` and ``bench.py``; there is
no dataset or checkpoint on the GPU box.  Two kinds of input:

* ground-truth boxes ``(B, N, 5)`` ``[x1, y1, x2, y2, class]`` in pixels for the
  encoder (uniform or mosaic-clustered centres, integer or fractional corners,
  zero padding rows at the end or interleaved);
* head outputs ``3 x (B, G, G, 5+A+C)``: *planted* (a trained-looking head: the
  encoder's targets pushed through the inverse activations plus noise) or
  *dense random* (every cell a candidate -- the NMS worst case).
"""
from __future__ import annotations

import numpy as np

# configs/yolov3_coco_anchor.txt of the reference: large -> small, layer 0 = stride 32
COCO_ANCHORS = (((112, 74), (149, 190), (370, 328)),
                ((28, 17), (56, 112), (57, 35)),
                ((9, 10), (13, 28), (28, 55)))


def coco_anchors(dtype=np.float32):
    """The reference's default anchors as a list of three (3, 2) arrays."""
    return [np.array(a, dtype=dtype) for a in COCO_ANCHORS]


def grid_sizes(input_size: int, num_layers: int = 3):
    return [input_size // s for s in (32, 16, 8, 4, 2)[:num_layers]]


def synth_boxes(seed, batch, max_boxes, input_size, num_classes, corners="int",
                layout="uniform", padding="tail", min_boxes=1, anchors=None,
                reject_iol_ties=True):
    """Ground-truth boxes, float32 (B, N, 5); rows beyond n_i are zero.

    corners: 'int' (legacy loader casts to int32) or 'frac' (letterbox/mosaic
    leave fractional corners).  layout: 'uniform' or 'mosaic' (four quadrants,
    clustered centres, sigma ~1.5 stride-8 cells: many shared 3x3 neighbourhoods).
    padding: 'tail' (valid rows first) or 'interleaved' (zero rows in between).
    reject_iol_ties: resample sizes whose two best rounded IoLs tie (the
    reference's argsort is host-dependent there); needs ``anchors``.
    """
    rng = np.random.default_rng(seed)
    s = float(input_size)
    out = np.zeros((batch, max_boxes, 5), dtype=np.float32)
    table = None
    if reject_iol_ties:
        table = np.concatenate(anchors if anchors is not None else coco_anchors(), 0)
        table = table.astype(np.float64)
    for b in range(batch):
        n = int(rng.integers(min_boxes, max_boxes + 1))
        if layout == "mosaic":
            n_clusters = 4 * int(rng.integers(1, 4))
            quad = rng.integers(0, 4, size=n_clusters)
            base = np.stack([(quad % 2) * s / 2, (quad // 2) * s / 2], -1)
            centres_c = base + rng.uniform(0.15, 0.85, size=(n_clusters, 2)) * s / 2
            which = rng.integers(0, n_clusters, size=n)
            centre = centres_c[which] + rng.normal(0.0, 12.0, size=(n, 2))
            centre = np.clip(centre, 1.0, s - 1.0)
        else:
            centre = rng.uniform(0.0, s, size=(n, 2))
        rows = []
        for i in range(n):
            for _ in range(64):
                wh = np.exp(rng.normal(np.log(60.0), 0.9, size=2))
                wh = np.clip(wh, 4.0, 0.9 * s)
                x1y1 = np.clip(centre[i] - wh / 2, 0.0, s)
                x2y2 = np.clip(centre[i] + wh / 2, 0.0, s)
                if corners == "int":
                    x1y1, x2y2 = np.floor(x1y1), np.ceil(x2y2)
                    x2y2 = np.minimum(x2y2, s)
                box = np.concatenate([x1y1, x2y2]).astype(np.float32)
                w, h = box[2] - box[0], box[3] - box[1]
                if not (w > 1.0 and h > 1.0):
                    continue
                if table is not None and _has_top2_tie(np.float64(w), np.float64(h), table):
                    continue
                rows.append(np.concatenate([box, [float(rng.integers(0, num_classes))]]))
                break
        rows = np.array(rows, dtype=np.float32).reshape(-1, 5)
        if padding == "interleaved":
            slots = np.sort(rng.choice(max_boxes, size=len(rows), replace=False))
            out[b, slots] = rows
        else:
            out[b, :len(rows)] = rows
    return out


def _has_top2_tie(w, h, table):
    iw = np.minimum(w, table[:, 0])
    ih = np.minimum(h, table[:, 1])
    iol = np.round(iw * ih / np.maximum(w * h, table[:, 0] * table[:, 1]), 3)
    top = np.sort(iol)[::-1]
    # within one rounding step also counts: f32 and f64 paths may round apart
    return bool(top[0] - top[1] < 1.5e-3)


def count_iol_ties(boxes, anchors):
    """How many valid boxes have their two best rounded IoLs within one step."""
    table = np.concatenate(anchors, 0).astype(np.float64)
    wh = (boxes[..., 2:4] - boxes[..., 0:2]).reshape(-1, 2).astype(np.float64)
    n = 0
    for w, h in wh:
        if w * h > 0 and _has_top2_tie(w, h, table):
            n += 1
    return n


# --------------------------------------------------------------------------
# head outputs (torch: the same code runs on the CPU here and on cuda in bench)
# --------------------------------------------------------------------------

def _inverse_xy_activation(v, iters=40):
    """Solve tanh(.15 x) + sigmoid(.15 x) = v for x by bisection (v in (-1, 2))."""
    import torch
    v = v.clamp(-0.97, 1.97)
    lo = torch.full_like(v, -40.0)
    hi = torch.full_like(v, 40.0)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        f = torch.tanh(0.15 * mid) + torch.sigmoid(0.15 * mid)
        below = f < v
        lo = torch.where(below, mid, lo)
        hi = torch.where(below, hi, mid)
    return 0.5 * (lo + hi)


def planted_head_outputs(y_true, num_anchors, seed, obj_pos=(4.0, 1.0),
                         obj_neg=(-8.0, 1.5)):
    """Turn encoder targets into plausible raw head logits.

    y_true: list of torch tensors (B, G, G, 5+A+C) on any device.  Positive
    cells get the inverse-activated xy, the stored log-ratios (+noise), a high
    objectness logit and peaked anchor/class logits; other cells get noise.
    About 770 candidates >= 0.001 per COCO image at 100 boxes.
    """
    import torch
    outs = []
    for l, yt in enumerate(y_true):
        gen = torch.Generator(device=yt.device)
        gen.manual_seed(int(seed) * 7919 + l)
        pos = yt[..., 4:5] > 0.5
        noise = torch.randn(yt.shape, generator=gen, device=yt.device, dtype=torch.float32)
        o = torch.empty_like(yt)
        o[..., 0:2] = torch.where(pos, _inverse_xy_activation(yt[..., 0:2]) + 0.05 * noise[..., 0:2],
                                  2.0 * noise[..., 0:2])
        o[..., 2:4] = torch.where(pos, yt[..., 2:4] + 0.03 * noise[..., 2:4],
                                  0.5 * noise[..., 2:4])
        o[..., 4:5] = torch.where(pos, obj_pos[0] + obj_pos[1] * noise[..., 4:5],
                                  obj_neg[0] + obj_neg[1] * noise[..., 4:5])
        a0, a1 = 5, 5 + num_anchors
        o[..., a0:a1] = 6.0 * yt[..., a0:a1] + noise[..., a0:a1]
        o[..., a1:] = 8.0 * yt[..., a1:] + noise[..., a1:]
        outs.append(o.contiguous())
    return outs


def dense_random_head_outputs(batch, input_size, num_anchors, num_classes, seed,
                              device="cpu", num_layers=3):
    """All channels N(0,1): every cell survives 0.001 (NMS worst case)."""
    import torch
    outs = []
    d = 5 + num_anchors + num_classes
    for l, g in enumerate(grid_sizes(input_size, num_layers)):
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) * 104729 + l)
        outs.append(torch.randn((batch, g, g, d), generator=gen, device=device,
                                dtype=torch.float32))
    return outs


LETTERBOX_SHAPES = ((608, 608), (480, 640), (427, 640), (640, 480), (375, 500), (1080, 1920))


def image_shapes(seed, batch, mixed=True, square=(608, 608)):
    """(B, 2) int32 original-image (h, w) per image."""
    if not mixed:
        return np.tile(np.array(square, dtype=np.int32), (batch, 1))
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(LETTERBOX_SHAPES), size=batch)
    return np.array(LETTERBOX_SHAPES, dtype=np.int32)[pick]
