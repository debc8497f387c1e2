"""The two restatements (NumPy and C) must agree: integer outputs bit-exact, IEEE-only floats
bit-exact, exp/log floats within 1e-6 relative (glibc vs NumPy SIMD differ by <= 1 ulp)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import rpn_oracle as O

F32 = np.float32
pytestmark = pytest.mark.skipif(not CO.available(), reason="C oracle not built (make -C oracle)")


def close(a, b):
    return np.all(np.abs(a.astype(np.float64) - b) <= 1e-6 * np.maximum(1, np.abs(b)))


def bits(a):
    return np.ascontiguousarray(a, F32).view(np.uint32)


def test_c_oracle_matches_golden(golden):
    assert np.array_equal(bits(CO.iou_map(golden["iou_boxes"], golden["iou_gt"])), bits(golden["iou_map_batched"]))
    assert close(CO.decode(golden["iou_boxes"], golden["dec_deltas"]), golden["dec_boxes_batched"])
    assert close(CO.encode(golden["enc_boxes"], golden["enc_gt"]), golden["enc_deltas"])
    r = CO.nms(golden["nms_in_boxes"], golden["nms_in_scores"], 50, 60, 0.3, 0.25)
    assert np.array_equal(bits(r[0]), bits(golden["nms_out_boxes"])) and np.array_equal(r[3], golden["nms_out_valid"])
    assert np.array_equal(bits(r[1]), bits(golden["nms_out_scores"]))


@pytest.mark.parametrize("bb,B,G", [("vgg16", 3, 20), ("mobilenet_v2", 2, 50)])
def test_targets_numpy_vs_c(bb, B, G):
    from tfrpn import synthetic
    hp = O.get_hyper_params(bb)
    anchors = O.generate_anchors(hp)
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(B + G), B, G)
    d, l, dbg = O.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=7, offset=9, image_offset=2, return_debug=True)
    cd, cl, cdbg = CO.rpn_targets(anchors, gtb, gtl, hp, seed=7, offset=9, image_offset=2, debug=True)
    assert np.array_equal(cl, l.reshape(B, -1))
    assert np.array_equal(cdbg["argmax_row"], dbg["argmax_row"]) and np.array_equal(cdbg["argmax_col"], dbg["argmax_col"])
    assert np.array_equal(cdbg["pos_pre"].astype(bool), dbg["pos_pre"])
    assert np.array_equal(cdbg["neg_pre"].astype(bool), dbg["neg_pre"])
    assert np.array_equal(cd != 0, d != 0) and close(cd, d)


def test_select_topk_proposals_numpy_vs_c():
    from tfrpn import synthetic
    rng = np.random.default_rng(4)
    mask = rng.uniform(size=(3, 3000)) < 0.4
    assert np.array_equal(CO.select_mask(mask, [100, 5, 0], seed=3, offset=1, stream=1, image_offset=5),
                          O.randomly_select_xyz_mask(mask, [100, 5, 0], seed=3, offset=1, stream=1, image_offset=5))
    s = (rng.integers(0, 50, size=(2, 700)) / 50).astype(F32)
    v, i = CO.top_k(s, 300)
    ov, oi = O.top_k(s, 300)
    assert np.array_equal(i, oi) and np.array_equal(v, ov)
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    reg, cls = synthetic.head_outputs(rng, 2, 31, 31, 9)
    cb, cs, cv, ck = CO.proposals(reg, cls, anchors, hp)
    ob, os_, ov, ok = O.generate_proposals(reg, cls, anchors, hp)
    assert np.array_equal(cv, ov) and np.array_equal(ck, ok) and np.array_equal(cs, os_) and close(cb, ob)
