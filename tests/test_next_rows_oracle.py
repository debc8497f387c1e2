"""Oracle restatements of the callers either side of the path against the
vectors made by running the reference's own statements (tests/golden/make_golden_next.py)."""
import os

import numpy as np
import pytest

from oracle import rpn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


@pytest.fixture(scope="module")
def nxt():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "next_vectors.npz")))


def test_predictor_body_matches_reference(nxt):
    hp = O.get_hyper_params("vgg16")
    boxes, vals, idx = O.predictor_top_boxes(nxt["pred_reg"], nxt["pred_cls"], O.generate_anchors(hp), hp, k=10)
    assert np.array_equal(idx, nxt["pred_top_indices"])
    assert np.array_equal(boxes.view(np.uint32), nxt["pred_selected_bboxes"].view(np.uint32))
    assert np.array_equal(vals, np.take_along_axis(nxt["pred_cls"].reshape(3, -1), idx.astype(np.int64), axis=1))
    # the tie group planted by the generator: equal scores come out in ascending index order
    same = vals[1][:-1] == vals[1][1:]
    assert same.any() and np.all(idx[1][:-1][same] < idx[1][1:][same])


def test_flip_and_padding_match_reference(nxt):
    assert np.array_equal(O.flip_horizontally_boxes(nxt["flip_in"]).view(np.uint32), nxt["flip_out"].view(np.uint32))
    assert list(nxt["pad_values"]) == [0.0, -1.0]
    rng = np.random.default_rng(0)
    bl = [nxt["flip_in"][:3], np.zeros((0, 4), F32), nxt["flip_in"]]
    ll = [np.array([4, 5, 6]), np.zeros((0,), np.int64), np.arange(7)]
    boxes, labels = O.pad_gt_batch(bl, ll, flip=[True, False, False], label_add=1)
    assert boxes.shape == (3, 7, 4) and labels.shape == (3, 7) and labels.dtype == np.int32
    assert np.array_equal(boxes[0, :3], nxt["flip_out"][:3]) and not boxes[0, 3:].any() and not boxes[1].any()
    assert list(labels[0]) == [5, 6, 7, -1, -1, -1, -1] and list(labels[1]) == [-1] * 7 and list(labels[2]) == list(range(1, 8))
    b2, l2 = O.pad_gt_batch(bl, ll, max_boxes=2)          # truncation to G
    assert b2.shape == (3, 2, 4) and list(l2[2]) == [0, 1]
    del rng
