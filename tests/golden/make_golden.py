#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's OWN source files.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

How: ``oracle/tf_numpy_shim`` is put first on sys.path so ``import tensorflow as tf`` inside
/root/reference/utils/{bbox_utils,train_utils}.py resolves to the NumPy stand-in; the
reference modules are then imported unmodified and called on seeded synthetic inputs.  The
only patch is ``train_utils.randomly_select_xyz_mask`` (tf.random cannot be reproduced): for
the bit-exact vectors it is replaced by the counter-RNG selection of the oracle (same
semantics, utils/train_utils.py:50-65), and its inputs are recorded (= the pre-sampling
masks).  A second, unpatched run checks the reference's own sampler keeps exactly
min(#True, select) entries, the property the replacement preserves.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_numpy_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from utils import bbox_utils, train_utils  # noqa: E402  (the reference, unmodified)
from oracle import rpn_oracle  # noqa: E402

assert tf.__file__.startswith(ROOT), tf.__file__
assert bbox_utils.__file__.startswith("/root/reference"), bbox_utils.__file__

F32 = np.float32


def synth_gt(rng, B, G, size_lo=0.05, size_hi=0.6):
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


def np_(x):
    return np.asarray(x.a if hasattr(x, "a") else x)


def run_targets(hp, anchors, gt_boxes, gt_labels, seed, offset):
    """Reference calculate_rpn_actual_outputs with the sampler swapped for the counter RNG."""
    calls = []
    orig = train_utils.randomly_select_xyz_mask

    def patched(mask, select_xyz):
        stream = len(calls)
        calls.append(np_(mask).copy())
        sel = rpn_oracle.randomly_select_xyz_mask(np_(mask), np_(select_xyz), seed=seed,
                                                  offset=offset, stream=stream)
        return tf.Tensor(sel)

    train_utils.randomly_select_xyz_mask = patched
    try:
        deltas, labels = train_utils.calculate_rpn_actual_outputs(
            tf.constant(anchors), tf.constant(gt_boxes), tf.constant(gt_labels), hp)
    finally:
        train_utils.randomly_select_xyz_mask = orig
    assert len(calls) == 2
    return np_(deltas), np_(labels), calls[0], calls[1]


def main():
    out = {}
    rng = np.random.default_rng(20261017)

    # ---- anchors at the two reference configs (utils/train_utils.py:5-18) -------------
    for bb in ("vgg16", "mobilenet_v2"):
        hp = train_utils.get_hyper_params(bb)
        out["base_anchors_" + bb] = np_(bbox_utils.generate_base_anchors(hp))
        out["anchors_" + bb] = np_(bbox_utils.generate_anchors(hp))
    assert out["anchors_vgg16"].shape == (8649, 4) and out["anchors_vgg16"].dtype == F32

    # ---- elementwise box math on random inputs ----------------------------------------
    B, N, G = 3, 257, 7
    a = np.sort(rng.uniform(0, 1, size=(B, N, 2, 2)), axis=2).astype(F32)
    boxes = np.stack([a[..., 0, 0], a[..., 0, 1], a[..., 1, 0], a[..., 1, 1]], axis=-1)
    gt, gl = synth_gt(rng, B, G)
    gt[1, 0] = boxes[1, 5]                        # an exact-match pair (IoU == 1)
    gt[2, 1] = [0.2, 0.3, 0.2, 0.9]               # zero-height GT
    out["iou_boxes"], out["iou_gt"] = boxes, gt
    out["iou_map_batched"] = np_(bbox_utils.generate_iou_map(tf.constant(boxes), tf.constant(gt)))
    out["iou_map_unbatched"] = np_(bbox_utils.generate_iou_map(tf.constant(boxes[0]), tf.constant(gt)))
    deltas = rng.normal(0, 0.5, size=(B, N, 4)).astype(F32)
    out["dec_deltas"] = deltas
    out["dec_boxes_batched"] = np_(bbox_utils.get_bboxes_from_deltas(tf.constant(boxes), tf.constant(deltas)))
    out["dec_boxes_unbatched"] = np_(bbox_utils.get_bboxes_from_deltas(tf.constant(boxes[0]), tf.constant(deltas)))
    g2 = np.sort(rng.uniform(0, 1, size=(B, N, 2, 2)), axis=2).astype(F32)
    gtb = np.stack([g2[..., 0, 0], g2[..., 0, 1], g2[..., 1, 0], g2[..., 1, 1]], axis=-1)
    gtb[0, :9] = 0                                # all-zero GT rows (the "not positive" case)
    gtb[1, 3, 3] = gtb[1, 3, 1]                   # zero-width GT
    bz = boxes.copy()
    bz[2, 4, 2] = bz[2, 4, 0]                     # zero-height bbox -> 1e-3 substitution
    out["enc_boxes"], out["enc_gt"] = bz, gtb
    out["enc_deltas"] = np_(bbox_utils.get_deltas_from_bboxes(tf.constant(bz), tf.constant(gtb)))
    px = (boxes * F32(500)).astype(F32)
    out["norm_in"] = px
    out["norm_out"] = np_(bbox_utils.normalize_bboxes(tf.constant(px), 375, 500))
    out["denorm_out"] = np_(bbox_utils.denormalize_bboxes(tf.constant(boxes), 375, 500))

    # ---- target assignment: full vgg16 geometry, B=2, G=12 -----------------------------
    for tag, bb, Bt, Gt, seed, offset, kw in (
            ("t_vgg16", "vgg16", 2, 12, 1234, 0, {}),
            ("t_mnv2", "mobilenet_v2", 3, 6, 99, 7, {}),
            ("t_smallquota", "vgg16", 2, 5, 5, 1, {"total_pos_bboxes": 4, "total_neg_bboxes": 6})):
        hp = dict(train_utils.get_hyper_params(bb, **kw))
        anchors = np_(bbox_utils.generate_anchors(hp))
        gtb, gtl = synth_gt(rng, Bt, Gt)
        if tag == "t_smallquota":
            gtb[0, 0] = [0.40, 0.40, 0.41, 0.41]  # tiny GT: best IoU < 0.3 (KAT 7)
            gtl[0, 0] = 3
        if tag == "t_vgg16":
            gtb[1, 0] = [0, 0, 1, 1]              # whole-image GT: big tie group (KAT 6)
            gtl[1, 0] = 1
        d, l, pos_pre, neg_pre = run_targets(hp, anchors, gtb, gtl, seed, offset)
        out[tag + "_gt_boxes"], out[tag + "_gt_labels"] = gtb, gtl
        out[tag + "_seed_offset"] = np.asarray([seed, offset], np.int64)
        out[tag + "_quota"] = np.asarray([hp["total_pos_bboxes"], hp["total_neg_bboxes"]], np.int32)
        out[tag + "_labels"] = l.astype(np.int8)
        nz = np.flatnonzero(np.any(d != 0, axis=-1).reshape(-1))
        out[tag + "_delta_rows"] = nz.astype(np.int32)       # sparse: non-zero rows only
        out[tag + "_delta_vals"] = d.reshape(-1, 4)[nz]
        out[tag + "_delta_shape"] = np.asarray(d.shape, np.int32)
        out[tag + "_pos_pre"] = np.packbits(pos_pre, axis=-1)
        out[tag + "_neg_pre"] = np.packbits(neg_pre, axis=-1)
        # unpatched run: the reference's own sampler must keep min(#True, quota) per image
        tf.random.set_seed(7)
        d2, l2 = train_utils.calculate_rpn_actual_outputs(
            tf.constant(anchors), tf.constant(gtb), tf.constant(gtl), hp)
        l2 = np_(l2).reshape(Bt, -1)
        npos = (l2 == 1).sum(-1)
        nneg = (l2 == 0).sum(-1)
        want_pos = np.minimum(pos_pre.sum(-1), hp["total_pos_bboxes"])
        assert np.array_equal(npos, want_pos), (npos, want_pos)
        assert np.all(nneg <= hp["total_pos_bboxes"] + hp["total_neg_bboxes"] - npos)
        assert np.all(pos_pre[l2 == 1])
        out[tag + "_ref_pos_count"] = npos.astype(np.int32)

    # ---- predictor.py:52-60 (transcribed call sequence; the file itself is a script) ---
    hp = train_utils.get_hyper_params("vgg16")
    anchors = bbox_utils.generate_anchors(hp)
    Bp = 2
    reg = rng.normal(0, 0.5, size=(Bp, 31, 31, 36)).astype(F32)
    cls = (1 / (1 + np.exp(-rng.normal(0, 2, size=(Bp, 31, 31, 9))))).astype(F32)
    rpn_bbox_deltas = tf.reshape(tf.constant(reg), (Bp, -1, 4))          # :52
    rpn_labels = tf.reshape(tf.constant(cls), (Bp, -1))                  # :53
    rpn_bbox_deltas *= hp["variances"]                                   # :55
    rpn_bboxes = bbox_utils.get_bboxes_from_deltas(anchors, rpn_bbox_deltas)  # :56
    _, top_indices = tf.nn.top_k(rpn_labels, 10)                         # :58
    selected = tf.gather(rpn_bboxes, top_indices, batch_dims=1)          # :60
    out["pred_reg"], out["pred_cls"] = reg, cls
    out["pred_top10_idx"] = np_(top_indices)
    out["pred_top10_boxes"] = np_(selected)
    chk = np_(rpn_bboxes)
    out["pred_boxes_sample_idx"] = np.arange(0, 8649, 37, dtype=np.int32)
    out["pred_boxes_sample"] = chk[:, ::37]

    # ---- composed proposal stage via the reference NMS wrapper (bbox_utils.py:48-70) ---
    k, post = 600, 40
    clipped = tf.clip_by_value(rpn_bboxes, 0, 1)
    top_scores, top_idx = tf.nn.top_k(rpn_labels, k)
    top_boxes = tf.gather(clipped, top_idx, batch_dims=1)
    nb, ns, nc, nv = bbox_utils.non_max_suppression(
        tf.reshape(top_boxes, (Bp, k, 1, 4)), tf.reshape(top_scores, (Bp, k, 1)),
        max_output_size_per_class=post, max_total_size=post, iou_threshold=0.7)
    out["prop_k_post_thr"] = np.asarray([k, post, 0.7], np.float64)
    out["prop_boxes"], out["prop_scores"] = np_(nb), np_(ns)
    out["prop_classes"], out["prop_valid"] = np_(nc), np_(nv)

    # ---- standalone NMS: ragged / degenerate / thresholded ----------------------------
    K = 300
    c = rng.uniform(0.2, 0.8, size=(2, K, 2)); s = rng.uniform(0.05, 0.4, size=(2, K, 2))
    nbx = np.concatenate([c - s / 2, c + s / 2], axis=-1).astype(F32)
    nbx[0, 3] = nbx[0, 3][[2, 3, 0, 1]]          # flipped corners
    nbx[0, 4] = [0.5, 0.5, 0.5, 0.7]             # zero-area box
    nbx[1, 7] = [-0.2, 0.1, 0.4, 1.3]            # out of range: clipped only in the output
    nsc = rng.permutation(2 * K).reshape(2, K).astype(F32) / F32(2 * K)
    r = bbox_utils.non_max_suppression(tf.constant(nbx.reshape(2, K, 1, 4)), tf.constant(nsc.reshape(2, K, 1)),
                                       max_output_size_per_class=50, max_total_size=60,
                                       iou_threshold=0.3, score_threshold=0.25)
    out["nms_in_boxes"], out["nms_in_scores"] = nbx, nsc
    for name, v in zip(("boxes", "scores", "classes", "valid"), r):
        out["nms_out_" + name] = np_(v)

    # ---- losses (train_utils.py:146-185) -----------------------------
    yt = np.zeros((2, 50, 4), F32); yt[:, :9] = rng.normal(0, 1, size=(2, 9, 4))
    yp = rng.normal(0, 1, size=(2, 5, 5, 8)).astype(F32)
    out["loss_reg_true"], out["loss_reg_pred"] = yt, yp
    out["loss_reg"] = np_(train_utils.reg_loss(tf.constant(yt), tf.constant(yp)))
    ct = rng.integers(-1, 2, size=(2, 5, 5, 9)).astype(F32)
    cp = rng.uniform(0.01, 0.99, size=(2, 5, 5, 9)).astype(F32)
    out["loss_cls_true"], out["loss_cls_pred"] = ct, cp
    out["loss_cls"] = np_(train_utils.cls_loss(tf.constant(ct), tf.constant(cp)))

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
