"""The NumPy oracle must reproduce, bit for bit, the vectors obtained by running the
reference's own source (tests/golden/make_golden.py) -- this is what pins the restatement."""
import numpy as np
import pytest

from oracle import rpn_oracle as O

F32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, dtype=F32).view(np.uint32)


def assert_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    if a.dtype.kind == "f":
        same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), "first mismatch at %s" % (np.argwhere(~same)[:3],)
    else:
        assert np.array_equal(a, b)


@pytest.mark.parametrize("bb", ["vgg16", "mobilenet_v2"])
def test_anchors_match_reference(golden, bb):
    hp = O.get_hyper_params(bb)
    assert_bits(O.generate_base_anchors(hp), golden["base_anchors_" + bb])
    assert_bits(O.generate_anchors(hp), golden["anchors_" + bb])


def test_iou_map_matches_reference(golden):
    assert_bits(O.generate_iou_map(golden["iou_boxes"], golden["iou_gt"]), golden["iou_map_batched"])
    assert_bits(O.generate_iou_map(golden["iou_boxes"][0], golden["iou_gt"]), golden["iou_map_unbatched"])


def test_decode_encode_match_reference(golden):
    assert_bits(O.get_bboxes_from_deltas(golden["iou_boxes"], golden["dec_deltas"]), golden["dec_boxes_batched"])
    assert_bits(O.get_bboxes_from_deltas(golden["iou_boxes"][0], golden["dec_deltas"]), golden["dec_boxes_unbatched"])
    assert_bits(O.get_deltas_from_bboxes(golden["enc_boxes"], golden["enc_gt"]), golden["enc_deltas"])


def test_normalize_denormalize_match_reference(golden):
    assert_bits(O.normalize_bboxes(golden["norm_in"], 375, 500), golden["norm_out"])
    assert_bits(O.denormalize_bboxes(golden["iou_boxes"], 375, 500), golden["denorm_out"])


def golden_targets(golden, tag):
    shape = tuple(golden[tag + "_delta_shape"])
    d = np.zeros((shape[0] * shape[1], 4), F32)
    d[golden[tag + "_delta_rows"]] = golden[tag + "_delta_vals"]
    N = shape[1]
    pos_pre = np.unpackbits(golden[tag + "_pos_pre"], axis=-1)[:, :N].astype(bool)
    neg_pre = np.unpackbits(golden[tag + "_neg_pre"], axis=-1)[:, :N].astype(bool)
    return d.reshape(shape), golden[tag + "_labels"].astype(F32), pos_pre, neg_pre


TARGET_CASES = [("t_vgg16", "vgg16"), ("t_mnv2", "mobilenet_v2"), ("t_smallquota", "vgg16")]


def target_case(golden, tag, bb):
    tp, tn = (int(v) for v in golden[tag + "_quota"])
    hp = O.get_hyper_params(bb, total_pos_bboxes=tp, total_neg_bboxes=tn)
    seed, offset = (int(v) for v in golden[tag + "_seed_offset"])
    return hp, golden[tag + "_gt_boxes"], golden[tag + "_gt_labels"], seed, offset


@pytest.mark.parametrize("tag,bb", TARGET_CASES)
def test_targets_match_reference(golden, tag, bb):
    hp, gtb, gtl, seed, offset = target_case(golden, tag, bb)
    anchors = O.generate_anchors(hp)
    d, l, dbg = O.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=seed, offset=offset,
                                               return_debug=True)
    gd, gl, pos_pre, neg_pre = golden_targets(golden, tag)
    assert np.array_equal(dbg["pos_pre"], pos_pre)
    assert np.array_equal(dbg["neg_pre"], neg_pre)
    assert_bits(l, gl.reshape(l.shape))
    assert_bits(d, gd)
    # the reference's own (tf.random) sampler kept the same number of positives
    assert np.array_equal(dbg["pos_count"], golden[tag + "_ref_pos_count"])


def test_predictor_sequence_matches_reference(golden):
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    reg, cls = golden["pred_reg"], golden["pred_cls"]
    B = reg.shape[0]
    deltas = reg.reshape(B, -1, 4) * np.asarray(hp["variances"], F32)
    boxes = O.get_bboxes_from_deltas(anchors, deltas)
    _, idx = O.top_k(cls.reshape(B, -1), 10)
    assert np.array_equal(idx, golden["pred_top10_idx"])
    assert_bits(np.take_along_axis(boxes, idx[..., None].astype(np.int64), axis=1), golden["pred_top10_boxes"])
    assert_bits(boxes[:, golden["pred_boxes_sample_idx"]], golden["pred_boxes_sample"])


def test_proposals_match_reference(golden):
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    k, post, thr = golden["prop_k_post_thr"]
    nb, ns, nv, keep = O.generate_proposals(golden["pred_reg"], golden["pred_cls"], anchors, hp,
                                            pre_nms_topn=int(k), post_nms_topn=int(post),
                                            nms_iou_threshold=float(thr))
    assert_bits(nb, golden["prop_boxes"])
    assert_bits(ns, golden["prop_scores"])
    assert np.array_equal(nv, golden["prop_valid"])
    assert (golden["prop_classes"] == 0).all()
    # keep indices (an addition over TF) must point at the kept scores
    sc = golden["pred_cls"].reshape(2, -1)
    for b in range(2):
        n = nv[b]
        assert np.array_equal(sc[b, keep[b, :n]], ns[b, :n]) and (keep[b, n:] == -1).all()


def test_nms_matches_reference(golden):
    bx, sc = golden["nms_in_boxes"], golden["nms_in_scores"]
    B, K = sc.shape
    nb, ns, nc, nv = O.combined_non_max_suppression(
        bx.reshape(B, K, 1, 4), sc.reshape(B, K, 1), max_output_size_per_class=50,
        max_total_size=60, iou_threshold=0.3, score_threshold=0.25)
    assert_bits(nb, golden["nms_out_boxes"])
    assert_bits(ns, golden["nms_out_scores"])
    assert_bits(nc, golden["nms_out_classes"])
    assert np.array_equal(nv, golden["nms_out_valid"])
