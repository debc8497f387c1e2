#!/usr/bin/env python
"""Golden vectors for the SURVEY 8f ("next") rows, produced by running the reference's OWN source
through oracle/tf_numpy_shim, like make_golden.py (build container only; needs /root/reference):

    python tests/golden/make_golden_next.py      ->  tests/golden/next_vectors.npz

* predictor loop body: the statements of predictor.py:52-60 are read from the file and exec'd,
  unmodified, on seeded head outputs (reshape, *= variances, get_bboxes_from_deltas, tf.nn.top_k,
  tf.gather(batch_dims=1)).
* GT preprocessing: utils/data_utils.flip_horizontally (:54-68) and get_padding_values (:152-157);
  tensorflow_datasets / PIL, which that module imports but the two functions never touch, are
  stubbed with empty modules.
"""
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_numpy_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)
for name in ("tensorflow_datasets", "PIL", "PIL.Image"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["PIL"].Image = sys.modules["PIL.Image"]

import tensorflow as tf  # noqa: E402  (the shim)
from utils import bbox_utils, data_utils, train_utils  # noqa: E402  (the reference, unmodified)

assert tf.__file__.startswith(ROOT) and data_utils.__file__.startswith("/root/reference")
F32 = np.float32


def np_(x):
    return np.asarray(x.a if hasattr(x, "a") else x)


def main():
    rng = np.random.default_rng(20261017)
    out = {}

    # ---- predictor.py:52-60, exec'd verbatim -------------------------------------------------
    src = open("/root/reference/predictor.py").read().splitlines()
    body = textwrap.dedent("\n".join(src[51:60]))          # lines 52..60 (1-based)
    assert "tf.reshape(rpn_bbox_deltas" in body and "tf.gather(rpn_bboxes" in body, body
    hyper_params = train_utils.get_hyper_params("vgg16")
    anchors = bbox_utils.generate_anchors(hyper_params)
    B, F, A = 3, 31, 9
    reg = rng.normal(0, 0.5, size=(B, F, F, 4 * A)).astype(F32)
    cls = rng.permutation(B * F * F * A).reshape(B, F, F, A).astype(F32) / F32(B * F * F * A)
    cls[1, 7, 3, 2:6] = F32(0.9999)                         # a tie group inside the top 10: lower index first
    cls[1, 0, 0, 1] = F32(0.9999)
    ns = {"tf": tf, "bbox_utils": bbox_utils, "hyper_params": hyper_params, "anchors": anchors, "batch_size": B,
          "rpn_bbox_deltas": tf.constant(reg), "rpn_labels": tf.constant(cls)}
    exec(body, ns)                                          # noqa: S102  the reference's statements
    out["pred_reg"], out["pred_cls"] = reg, cls
    out["pred_top_indices"] = np_(ns["top_indices"]).astype(np.int32)
    out["pred_selected_bboxes"] = np_(ns["selected_rpn_bboxes"])
    out["pred_all_bboxes"] = np_(ns["rpn_bboxes"])

    # ---- data_utils.flip_horizontally + padding values -------------------------------------------
    boxes = np.sort(rng.uniform(0, 1, size=(7, 2, 2)), axis=1).astype(F32)
    boxes = np.stack([boxes[:, 0, 0], boxes[:, 0, 1], boxes[:, 1, 0], boxes[:, 1, 1]], axis=-1)
    img = rng.uniform(size=(4, 6, 3)).astype(F32)
    fimg, fboxes = data_utils.flip_horizontally(tf.constant(img), tf.constant(boxes))
    out["flip_in"], out["flip_out"] = boxes, np_(fboxes)
    assert np.array_equal(np_(fimg), img[:, ::-1])
    pv = data_utils.get_padding_values()
    out["pad_values"] = np.array([float(np_(pv[1])), float(np_(pv[2]))], F32)      # boxes 0, labels -1

    path = os.path.join(HERE, "next_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
