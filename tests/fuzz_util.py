"""Random geometries / boxes shared by the CPU and GPU randomised parity sweeps."""
import numpy as np


def random_head(rng):
    L = int(rng.integers(1, 6))
    S = int(rng.choice([64, 96, 100, 128, 160, 224, 250]))
    strides = (32, 16, 8, 4, 2)[:L]
    if S // strides[0] < 1:
        S = 64
    A = [int(rng.integers(1, 9)) if rng.random() < 0.3 else 3 for _ in range(L)]
    C = int(rng.choice([1, 2, 3, 7, 20, 80]))
    dt = np.float64 if rng.random() < 0.3 else np.float32
    anchors = [np.sort(rng.uniform(4, S * 0.9, (a, 2)), axis=0).astype(dt) for a in A]
    grids = [(max(S // s, 1), max(S // s, 1)) for s in strides]
    return S, C, anchors, grids


def random_boxes(rng, B, N, S, C):
    boxes = np.zeros((B, N, 5), np.float32)
    for b in range(B):
        n = int(rng.integers(0, N + 1))
        rows = rng.permutation(N)[:n] if rng.random() < 0.5 else np.arange(n)
        c = rng.uniform(-0.1 * S, 1.1 * S, (n, 2))
        wh = np.exp(rng.normal(np.log(S / 6), 1.0, (n, 2))).clip(0.5, 1.5 * S)
        x1y1 = c - wh / 2
        x2y2 = c + wh / 2
        if rng.random() < 0.5:
            x1y1, x2y2 = np.floor(x1y1), np.ceil(x2y2)
        boxes[b, rows, :4] = np.concatenate([x1y1, x2y2], 1)
        boxes[b, rows, 4] = rng.integers(0, C, n)
    return boxes


def tie_free(boxes, anchors):
    """Drop boxes whose two best rounded IoLs tie (the reference's argsort is host-dependent
    there): zero their rows."""
    table = np.concatenate(anchors, 0).astype(np.float64)
    wh = (boxes[..., 2:4] - boxes[..., 0:2]).astype(np.float64)
    inter = np.minimum(wh[..., None, 0], table[:, 0]) * np.minimum(wh[..., None, 1], table[:, 1])
    big = np.maximum((wh[..., 0] * wh[..., 1])[..., None], table[:, 0] * table[:, 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        iol = np.round(inter / big, 3)
    top2 = np.sort(iol, -1)[..., -2:] if table.shape[0] > 1 else None
    if top2 is not None:
        tie = np.abs(top2[..., 1] - top2[..., 0]) < 2e-3
        boxes[tie] = 0
    return boxes


def tf_encoder_odd_case(rng):
    """A random head geometry and box set for the TF encoder's semantics: 1-4 layers, 1-4 anchors
    per layer, boxes that stick out of the image, negative sizes, class ids outside [0, C)."""
    L = int(rng.integers(1, 5))
    A = [int(rng.integers(1, 5)) for _ in range(L)]
    S = int(rng.choice([64, 96, 160, 224, 320, 416]))
    C = int(rng.choice([1, 3, 20, 80]))
    grids = [(max(1, S // max(32 >> l, 4)),) * 2 for l in range(L)]
    anchors = [np.abs(rng.normal(40 * (L - l), 20, (A[l], 2))).astype(np.float32) + 2 for l in range(L)]
    B, N = int(rng.integers(1, 4)), int(rng.integers(1, 30))
    x1 = rng.uniform(-10, S, (B, N))
    y1 = rng.uniform(-10, S, (B, N))
    w = rng.uniform(-5, S / 2, (B, N))
    h = rng.uniform(-5, S / 2, (B, N))
    boxes = np.stack([x1, y1, x1 + w, y1 + h, rng.integers(-1, C + 1, (B, N))], -1).astype(np.float32)
    if rng.integers(0, 3) == 0:
        boxes[..., :4] = np.round(boxes[..., :4])
    return S, C, anchors, grids, boxes
