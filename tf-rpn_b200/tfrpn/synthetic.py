"""Seeded synthetic VOC-like inputs of the shapes BASELINE.json names.  NumPy only;
used by bench.py and the tests (there is no network for datasets or checkpoints)."""
import numpy as np

F32 = np.float32

CONFIGS = {
    # id: (backbone, B, G, hyper-param overrides)
    "C1": ("vgg16", 1, 50, {}),
    "C2": ("vgg16", 64, 50, {}),
    "C3": ("mobilenet_v2", 128, 50, {}),
    "C4": ("vgg16", 32, 200, {"img_size": (800, 1333), "feature_map_shape": (50, 84)}),
}


def gt_batch(rng, B, G, size_lo=0.05, size_hi=0.6):
    """GT boxes (B,G,4) f32 zero padded + labels (B,G) i32, -1 padded (utils/data_utils.py:152-157)."""
    boxes = np.zeros((B, G, 4), F32)
    labels = np.full((B, G), -1, np.int32)
    for b in range(B):
        n = int(rng.integers(1, G + 1))
        c = rng.uniform(0.1, 0.9, size=(n, 2))
        s = rng.uniform(size_lo, size_hi, size=(n, 2))
        bx = np.concatenate([c - s / 2, c + s / 2], axis=1)
        boxes[b, :n] = np.clip(bx, 0, 1).astype(F32)
        labels[b, :n] = rng.integers(1, 21, size=n)
    return boxes, labels


def head_outputs(rng, B, fm_h, fm_w, A):
    """rpn_reg (B,fm_h,fm_w,4A) ~ N(0, 0.5^2); rpn_cls (B,fm_h,fm_w,A) = sigmoid(N(0, 2^2)) made
    distinct per image (ties are exercised by a dedicated test)."""
    reg = rng.normal(0, 0.5, size=(B, fm_h, fm_w, 4 * A)).astype(F32)
    N = fm_h * fm_w * A
    cls = (1.0 / (1.0 + np.exp(-rng.normal(0, 2, size=(B, N))))).astype(F32)
    for b in range(B):
        order = np.argsort(cls[b], kind="stable")
        vals = cls[b][order]
        # strictly increasing by bumping equal neighbours up one ulp at a time
        for i in np.flatnonzero(vals[1:] <= vals[:-1]):
            vals[i + 1] = np.nextafter(max(vals[i], vals[i + 1]), F32(2), dtype=F32)
        cls[b][order] = vals
    return reg, cls.reshape(B, fm_h, fm_w, A)


def nms_boxes(rng, B, K, size_lo=0.02, size_hi=0.3):
    """C5 stress inputs: K boxes/image drawn like GT boxes, distinct scores U(0,1)."""
    c = rng.uniform(0.1, 0.9, size=(B, K, 2))
    s = rng.uniform(size_lo, size_hi, size=(B, K, 2))
    boxes = np.clip(np.concatenate([c - s / 2, c + s / 2], axis=-1), 0, 1).astype(F32)
    scores = np.stack([rng.permutation(K) for _ in range(B)]).astype(F32)
    scores = ((scores + F32(0.5)) / F32(K)).astype(F32)
    return boxes, scores
