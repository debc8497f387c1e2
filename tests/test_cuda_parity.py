"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in Python functions) against
the oracle on the same seeded inputs and against the committed golden vectors.

Bars (BASELINE.json north star): integer / index / label / mask outputs bit-exact; IoU (IEEE-exact
ops only) bit-exact; deltas and decoded boxes (exp / log) within 1e-6 relative in fp32.
"""
import numpy as np
import pytest

from oracle import rpn_oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32
RTOL = 1e-6  # the tolerance BASELINE.json's north_star states for floating point


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, F32), np.ascontiguousarray(b, F32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= RTOL * np.maximum(1.0, np.abs(b))))


@pytest.fixture(scope="module")
def T(cuda_device):
    import torch
    import tfrpn
    from tfrpn.utils import bbox_utils, train_utils

    class Ns:
        pass
    ns = Ns()
    ns.torch, ns.bbox, ns.train, ns.tfrpn, ns.dev = torch, bbox_utils, train_utils, tfrpn, cuda_device
    ns.cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    ns.np = lambda t: t.detach().cpu().numpy()
    return ns


HP_C4 = dict(O.get_hyper_params("vgg16"), img_size=(800, 1333), feature_map_shape=(50, 84))


# ---------------------------------------------------------------- anchors
@pytest.mark.parametrize("hp", [O.get_hyper_params("vgg16"), O.get_hyper_params("mobilenet_v2"), HP_C4,
                                dict(O.get_hyper_params("vgg16"), feature_map_shape=7, anchor_scales=[64, 300],
                                     anchor_ratios=[1., 3., 0.25, 0.7])],
                         ids=["vgg16", "mobilenet_v2", "c4_1333x800", "odd"])
def test_anchors_bit_exact(T, hp):
    hp = dict(hp, anchor_count=len(hp["anchor_ratios"]) * len(hp["anchor_scales"]))
    assert bits_equal(T.np(T.bbox.generate_anchors(hp)), O.generate_anchors(hp))
    assert bits_equal(T.np(T.bbox.generate_base_anchors(hp)), O.generate_base_anchors(hp))


def test_anchors_match_golden(T, golden):
    for bb in ("vgg16", "mobilenet_v2"):
        assert bits_equal(T.np(T.bbox.generate_anchors(T.train.get_hyper_params(bb))), golden["anchors_" + bb])


# ---------------------------------------------------------------- IoU map
def rand_boxes(rng, *shape):
    a = np.sort(rng.uniform(0, 1, size=shape + (2, 2)), axis=-2).astype(F32)
    return np.stack([a[..., 0, 0], a[..., 0, 1], a[..., 1, 0], a[..., 1, 1]], axis=-1)


@pytest.mark.parametrize("B,N,G,batched", [(3, 257, 7, True), (3, 257, 7, False), (1, 1, 1, False), (2, 128, 50, True),
                                           (2, 8649, 50, False), (2, 1000, 300, False), (5, 129, 3, True),
                                           # G % 4 == 0 / G % 2 == 0: 4 / 2 GT columns per thread (16- / 8-byte stores)
                                           (2, 300, 200, True), (2, 513, 52, False), (3, 100, 6, True),
                                           (1, 40, 1024, False), (1, 33, 600, False), (1, 20, 1026, False),
                                           # G % 4 == 2: 16-byte stores over row pairs; odd N*G/2 shifts every other image by 8 bytes
                                           (3, 8649, 50, False), (4, 301, 6, True), (5, 257, 2, True), (3, 1000, 10, False),
                                           (2, 255, 510, False), (3, 511, 50, True), (2, 256, 50, False)])
def test_iou_map_bit_exact(T, B, N, G, batched):
    rng = np.random.default_rng(B * 1000 + N + G)
    boxes = rand_boxes(rng, B, N) if batched else rand_boxes(rng, N)
    gt = rand_boxes(rng, B, G)
    gt[0, 0] = 0                                  # padded GT row
    if N > 5:
        gt[-1, -1] = boxes[-1, 5] if batched else boxes[5]   # IoU == 1
    got = T.np(T.bbox.generate_iou_map(T.cu(boxes), T.cu(gt)))
    assert bits_equal(got, O.generate_iou_map(boxes, gt))


def test_iou_map_golden_and_nan(T, golden):
    got = T.np(T.bbox.generate_iou_map(T.cu(golden["iou_boxes"]), T.cu(golden["iou_gt"])))
    assert bits_equal(got, golden["iou_map_batched"])
    got = T.np(T.bbox.generate_iou_map(T.cu(golden["iou_boxes"][0]), T.cu(golden["iou_gt"])))
    assert bits_equal(got, golden["iou_map_unbatched"])
    z = np.zeros((1, 2, 4), F32)                   # both degenerate -> 0/0 = NaN, like the reference
    got = T.np(T.bbox.generate_iou_map(T.cu(z[0]), T.cu(z)))
    assert np.isnan(got).all() and np.isnan(O.generate_iou_map(z[0], z)).all()


def test_iou_map_accepts_numpy_host_buffers(T):
    rng = np.random.default_rng(5)
    boxes, gt = rand_boxes(rng, 300), rand_boxes(rng, 2, 9)
    got = T.bbox.generate_iou_map(boxes, gt)
    assert isinstance(got, np.ndarray) and bits_equal(got, O.generate_iou_map(boxes, gt))


# ---------------------------------------------------------------- encode / decode / scale
def test_encode_decode_golden(T, golden):
    enc = T.np(T.bbox.get_deltas_from_bboxes(T.cu(golden["enc_boxes"]), T.cu(golden["enc_gt"])))
    ref = golden["enc_deltas"]
    assert close(enc, ref)
    assert bits_equal(enc[..., :2], ref[..., :2])         # dy, dx use IEEE-exact ops only
    assert np.array_equal(enc == 0, ref == 0)
    dec = T.np(T.bbox.get_bboxes_from_deltas(T.cu(golden["iou_boxes"]), T.cu(golden["dec_deltas"])))
    assert close(dec, golden["dec_boxes_batched"])
    dec = T.np(T.bbox.get_bboxes_from_deltas(T.cu(golden["iou_boxes"][0]), T.cu(golden["dec_deltas"])))
    assert close(dec, golden["dec_boxes_unbatched"])


def test_encode_decode_kat(T):
    a = np.asarray([[[0, 0, .5, .5]]], F32)
    g = np.asarray([[[.25, .25, .75, .75]]], F32)
    assert np.array_equal(T.np(T.bbox.get_deltas_from_bboxes(T.cu(a), T.cu(g)))[0, 0], np.asarray([.5, .5, 0, 0], F32))
    assert np.array_equal(T.np(T.bbox.get_deltas_from_bboxes(T.cu(a), T.cu(g * 0)))[0, 0], np.zeros(4, F32))
    d = (np.asarray([[[5, 5, 0, 0]]], F32) * np.asarray([.1, .1, .2, .2], F32)).astype(F32)
    assert np.array_equal(T.np(T.bbox.get_bboxes_from_deltas(T.cu(a), T.cu(d))), g)
    assert np.array_equal(T.np(T.bbox.get_bboxes_from_deltas(T.cu(a), T.cu(d * 0))), a)


@pytest.mark.parametrize("B,N", [(1, 1), (2, 511), (3, 513), (64, 8649)])
def test_decode_encode_random_and_roundtrip(T, B, N):
    rng = np.random.default_rng(N)
    anchors = rand_boxes(rng, N)
    deltas = rng.normal(0, 0.5, size=(B, N, 4)).astype(F32)
    dec = T.np(T.bbox.get_bboxes_from_deltas(T.cu(anchors), T.cu(deltas)))
    assert close(dec, O.get_bboxes_from_deltas(anchors, deltas))
    # size-independent property: encode(decode(d)) ~ d wherever the anchor is not degenerate
    enc = T.np(T.bbox.get_deltas_from_bboxes(T.cu(anchors), T.cu(dec)))
    ok = ((anchors[:, 2] - anchors[:, 0]) > 1e-2) & ((anchors[:, 3] - anchors[:, 1]) > 1e-2)
    assert np.allclose(enc[:, ok], deltas[:, ok], atol=2e-3)
    assert close(enc, O.get_deltas_from_bboxes(anchors, dec))


def test_normalize_denormalize_bit_exact(T, golden):
    assert bits_equal(T.np(T.bbox.normalize_bboxes(T.cu(golden["norm_in"]), 375, 500)), golden["norm_out"])
    assert bits_equal(T.np(T.bbox.denormalize_bboxes(T.cu(golden["iou_boxes"]), 375, 500)), golden["denorm_out"])


# ---------------------------------------------------------------- target assignment
def check_targets(T, hp, anchors, gtb, gtl, seed, offset, image_offset=0):
    d, l, dbg = T.train.calculate_rpn_actual_outputs(T.cu(anchors), T.cu(gtb), T.cu(gtl), hp, seed=seed,
                                                     offset=offset, image_offset=image_offset, return_debug=True)
    d, l = T.np(d), T.np(l)
    dbg = {k: T.np(v) for k, v in dbg.items()}
    od, ol, odbg = O.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=seed, offset=offset,
                                                  image_offset=image_offset, return_debug=True)
    assert np.array_equal(dbg["argmax_row"], odbg["argmax_row"])
    assert np.array_equal(dbg["argmax_col"], odbg["argmax_col"])
    assert np.array_equal(dbg["max_iou"], odbg["max_iou"])          # numerically (+-0 compare equal)
    assert np.array_equal(dbg["pos_pre"].astype(bool), odbg["pos_pre"])
    assert np.array_equal(dbg["neg_pre"].astype(bool), odbg["neg_pre"])
    assert np.array_equal(dbg["pos_count"], odbg["pos_count"])
    assert np.array_equal(dbg["neg_count"], odbg["neg_count"])
    assert l.shape == ol.shape and bits_equal(l, ol)                  # labels {1,0,-1} bit-exact
    assert np.array_equal(d != 0, od != 0) and close(d, od)
    assert bits_equal(d[..., :2], od[..., :2])
    return d, l, dbg


@pytest.mark.parametrize("tag,bb", [("t_vgg16", "vgg16"), ("t_mnv2", "mobilenet_v2"), ("t_smallquota", "vgg16")])
def test_targets_golden(T, golden, tag, bb):
    tp, tn = (int(v) for v in golden[tag + "_quota"])
    hp = dict(O.get_hyper_params(bb, total_pos_bboxes=tp, total_neg_bboxes=tn))
    seed, offset = (int(v) for v in golden[tag + "_seed_offset"])
    anchors = golden["anchors_" + bb]
    d, l, _ = check_targets(T, hp, anchors, golden[tag + "_gt_boxes"], golden[tag + "_gt_labels"], seed, offset)
    shape = tuple(golden[tag + "_delta_shape"])
    gd = np.zeros((shape[0] * shape[1], 4), F32)
    gd[golden[tag + "_delta_rows"]] = golden[tag + "_delta_vals"]
    assert close(d, gd.reshape(shape))
    assert bits_equal(l.reshape(shape[0], -1), golden[tag + "_labels"].astype(F32).reshape(shape[0], -1))


@pytest.mark.parametrize("cfg,B", [("C1", 1), ("C2", 64), ("C3", 8), ("C4", 2)])
def test_targets_vs_oracle_synthetic(T, cfg, B):
    from tfrpn import synthetic
    bb, _, G, over = synthetic.CONFIGS[cfg]
    hp = dict(O.get_hyper_params(bb), **over)
    anchors = O.generate_anchors(hp)
    rng = np.random.default_rng(1000 * int(cfg[1]))
    gtb, gtl = synthetic.gt_batch(rng, B, G)
    d, l, dbg = check_targets(T, hp, anchors, gtb, gtl, seed=2026, offset=3)
    lab = l.reshape(B, -1)
    assert np.all((lab == 1).sum(-1) <= hp["total_pos_bboxes"])
    assert np.all((lab == 1).sum(-1) + (lab == 0).sum(-1) <= hp["total_pos_bboxes"] + hp["total_neg_bboxes"])


def test_targets_edge_cases(T):
    hp = dict(O.get_hyper_params("vgg16"))
    anchors = O.generate_anchors(hp)
    # all-padding image, G = 1, exact duplicates (ties), tiny GT, one image with 50 valid boxes
    gtb = np.zeros((4, 1, 4), F32); gtl = np.full((4, 1), -1, np.int32)
    gtb[1, 0] = anchors[8648]; gtl[1, 0] = 5
    gtb[2, 0] = [0.4, 0.4, 0.41, 0.41]; gtl[2, 0] = 1
    gtb[3, 0] = [0, 0, 1, 1]; gtl[3, 0] = 1
    check_targets(T, hp, anchors, gtb, gtl, seed=1, offset=0)
    hp2 = dict(hp, total_pos_bboxes=3, total_neg_bboxes=5)
    rng = np.random.default_rng(3)
    from tfrpn import synthetic
    gtb, gtl = synthetic.gt_batch(rng, 3, 50)
    uniq, inv, counts = np.unique(anchors, axis=0, return_inverse=True, return_counts=True)
    dup = np.flatnonzero(counts[inv.reshape(-1)] > 1)
    gtb[0, 0] = anchors[dup[-1]]; gtl[0, 0] = 2     # tie group at IoU 1: lowest index must win
    check_targets(T, hp2, anchors, gtb, gtl, seed=77, offset=12345678901)


def test_targets_sharded_equals_unsharded(T):
    """shard r of R must equal rows [r*B/R, (r+1)*B/R) of the unsharded run bit for bit."""
    from tfrpn import synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    anchors = T.bbox.generate_anchors(hp)
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(9), 8, 20)
    full_d, full_l = T.train.calculate_rpn_actual_outputs(anchors, T.cu(gtb), T.cu(gtl), hp, seed=4, offset=1)
    for r in range(4):
        sl = slice(2 * r, 2 * r + 2)
        d, l = T.train.calculate_rpn_actual_outputs(anchors, T.cu(gtb[sl]), T.cu(gtl[sl]), hp, seed=4, offset=1,
                                                    image_offset=2 * r)
        assert T.torch.equal(d, full_d[sl]) and T.torch.equal(l, full_l[sl])


def test_targets_sampling_changes_with_offset_only(T):
    from tfrpn import synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    anchors = T.bbox.generate_anchors(hp)
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(10), 2, 30)
    a = T.train.calculate_rpn_actual_outputs(anchors, T.cu(gtb), T.cu(gtl), hp, seed=1, offset=5)
    b = T.train.calculate_rpn_actual_outputs(anchors, T.cu(gtb), T.cu(gtl), hp, seed=1, offset=5)
    c = T.train.calculate_rpn_actual_outputs(anchors, T.cu(gtb), T.cu(gtl), hp, seed=1, offset=6)
    assert T.torch.equal(a[1], b[1]) and T.torch.equal(a[0], b[0])
    assert not T.torch.equal(a[1], c[1])


@pytest.mark.parametrize("B,N,quota", [(4, 500, [50]), (4, 8649, [0, 10, 100000, 128]), (2, 70000, [1, 69999])])
def test_select_mask_bit_exact(T, B, N, quota):
    rng = np.random.default_rng(N)
    mask = rng.uniform(size=(B, N)) < 0.3
    got = T.np(T.train.randomly_select_xyz_mask(T.cu(mask), T.cu(np.asarray(quota, np.int32)), seed=9, offset=2,
                                                stream=1, image_offset=3))
    want = O.randomly_select_xyz_mask(mask, quota, seed=9, offset=2, stream=1, image_offset=3)
    assert got.dtype == bool and np.array_equal(got, want)


# ---------------------------------------------------------------- top-k
@pytest.mark.parametrize("B,N,k", [(2, 8649, 10), (3, 8649, 6000), (2, 9216, 9216), (1, 37, 1), (2, 100000, 6000),
                                   (1, 500000, 6000), (2, 5000, 4999)])
def test_topk_bit_exact(T, B, N, k):
    rng = np.random.default_rng(N + k)
    scores = rng.uniform(0, 1, size=(B, N)).astype(F32)
    boxes = rand_boxes(rng, B, N)
    v, i, g = T.bbox.top_k_boxes(T.cu(scores), k, T.cu(boxes))
    ov, oi = O.top_k(scores, k)
    assert np.array_equal(T.np(i), oi) and bits_equal(T.np(v), ov)
    assert bits_equal(T.np(g), np.take_along_axis(boxes, oi[..., None].astype(np.int64), axis=1))


@pytest.mark.parametrize("N,k", [(60000, 100), (131072, 6000), (200003, 3)])
def test_topk_large_n_prefilter(T, N, k):
    """N >= 40000 and k <= N/4 goes through the multi-CTA prefilter (histogram select + stable compaction):
    image 0 random scores, image 1 heavy ties straddling rank k (lower index first must survive the
    compaction), image 2 ONE tie group larger than the candidate array (falls back to the unfiltered kernel),
    image 3 negative / zero / -0 scores."""
    rng = np.random.default_rng(N + k)
    scores = rng.uniform(0, 1, size=(4, N)).astype(F32)
    scores[1] = np.round(scores[1] * 200) / 200
    scores[2] = F32(0.25)
    scores[2, rng.integers(0, N, size=max(1, k // 2))] = F32(0.75)
    scores[3] = (rng.integers(-50, 50, size=N).astype(F32)) / F32(16)
    scores[3, :7] = -0.0
    boxes = rand_boxes(rng, 4, N)
    v, i, g = T.bbox.top_k_boxes(T.cu(scores), k, T.cu(boxes))
    ov, oi = O.top_k(scores, k)
    assert np.array_equal(T.np(i), oi) and np.array_equal(T.np(v), ov)
    assert bits_equal(T.np(g), np.take_along_axis(boxes, oi[..., None].astype(np.int64), axis=1))
    v2, i2 = T.bbox.top_k_boxes(T.cu(scores), k)             # without the gather
    assert np.array_equal(T.np(i2), oi)


def test_topk_ties_and_negative_scores(T):
    rng = np.random.default_rng(0)
    scores = rng.integers(-3, 4, size=(3, 4000)).astype(F32) / F32(4)      # massive ties, +-0
    scores[0, :10] = -0.0
    v, i = T.bbox.top_k_boxes(T.cu(scores), 1500)
    ov, oi = O.top_k(scores, 1500)
    assert np.array_equal(T.np(i), oi) and np.array_equal(T.np(v), ov)
    s = np.asarray([[.5, .9, .5, .9, .1, .5]], F32)
    v, i = T.bbox.top_k_boxes(T.cu(s), 5)
    assert list(T.np(i)[0]) == [1, 3, 0, 2, 5]


def test_topk_golden_predictor_sequence(T, golden):
    hp = T.train.get_hyper_params("vgg16")
    anchors = T.bbox.generate_anchors(hp)
    reg, cls = golden["pred_reg"], golden["pred_cls"]
    B = reg.shape[0]
    deltas = T.cu(reg).reshape(B, -1, 4) * T.torch.tensor(hp["variances"], device=T.dev)   # predictor.py:55
    boxes = T.bbox.get_bboxes_from_deltas(anchors, deltas)
    _, idx, sel = T.bbox.top_k_boxes(T.cu(cls).reshape(B, -1), 10, boxes)
    assert np.array_equal(T.np(idx), golden["pred_top10_idx"])
    assert close(T.np(sel), golden["pred_top10_boxes"])
    assert close(T.np(boxes)[:, golden["pred_boxes_sample_idx"]], golden["pred_boxes_sample"])


# ---------------------------------------------------------------- NMS
def run_nms(T, boxes, scores, **kw):
    B, K = scores.shape
    r = T.bbox.non_max_suppression(T.cu(boxes.reshape(B, K, 1, 4)), T.cu(scores.reshape(B, K, 1)),
                                   return_indices=True, **kw)
    return [T.np(x) for x in r]


def check_nms(T, boxes, scores, **kw):
    B, K = scores.shape
    got = run_nms(T, boxes, scores, **kw)
    want = O.combined_non_max_suppression(boxes.reshape(B, K, 1, 4), scores.reshape(B, K, 1), return_indices=True,
                                          **kw)
    assert np.array_equal(got[3], want[3])                     # valid_detections
    assert np.array_equal(got[4], want[4])                     # keep list, order included
    assert bits_equal(got[0], want[0]) and bits_equal(got[1], want[1]) and bits_equal(got[2], want[2])
    return got


def test_nms_golden(T, golden):
    got = run_nms(T, golden["nms_in_boxes"], golden["nms_in_scores"], max_output_size_per_class=50,
                  max_total_size=60, iou_threshold=0.3, score_threshold=0.25)
    assert bits_equal(got[0], golden["nms_out_boxes"]) and bits_equal(got[1], golden["nms_out_scores"])
    assert bits_equal(got[2], golden["nms_out_classes"]) and np.array_equal(got[3], golden["nms_out_valid"])


@pytest.mark.parametrize("K,per_class,total,thr,sthr", [
    (300, 50, 60, 0.3, 0.25), (6000, 300, 300, 0.7, float("-inf")), (1000, 1000, 1000, 0.5, float("-inf")),
    (129, 7, 5, 0.1, 0.5), (1, 3, 3, 0.5, float("-inf")), (2500, 300, 300, 0.05, float("-inf")),
    (700, 10, 10, 0.5, 2.0)])
def test_nms_vs_oracle(T, K, per_class, total, thr, sthr):
    from tfrpn import synthetic
    rng = np.random.default_rng(K)
    boxes, scores = synthetic.nms_boxes(rng, 3, K, 0.05, 0.4)
    if K > 10:
        boxes[0, 3] = boxes[0, 3][[2, 3, 0, 1]]
        boxes[0, 4] = [0.5, 0.5, 0.5, 0.7]
        boxes[1, 7] = [-0.2, 0.1, 0.4, 1.3]
    got = check_nms(T, boxes, scores, max_output_size_per_class=per_class, max_total_size=total,
                    iou_threshold=thr, score_threshold=sthr)
    # property: kept boxes are pairwise below the threshold and scores are non-increasing
    for b in range(3):
        n = got[3][b]
        assert np.all(np.diff(got[1][b, :n]) <= 0)


def test_nms_heavy_suppression_and_flags(T):
    rng = np.random.default_rng(1)
    K = 4000
    c = rng.uniform(0.45, 0.55, size=(2, K, 2)); s = rng.uniform(0.3, 0.4, size=(2, K, 2))
    boxes = np.concatenate([c - s / 2, c + s / 2], -1).astype(F32)     # nearly everything overlaps
    scores = (rng.permutation(2 * K).reshape(2, K).astype(F32) + 1) / F32(2 * K + 2)
    check_nms(T, boxes, scores, max_output_size_per_class=300, max_total_size=300, iou_threshold=0.7)
    check_nms(T, boxes, scores, max_output_size_per_class=20, max_total_size=300, iou_threshold=0.7, pad_per_class=True)
    check_nms(T, boxes * 1.5 - 0.2, scores, max_output_size_per_class=30, max_total_size=30, iou_threshold=0.6,
              clip_boxes=False)


def test_nms_kat_one_ulp_threshold(T):
    f = F32
    boxes = np.stack([np.asarray(v, f) for v in ([0, 0, 1, 1], [0, 0, 1, .5], [.2, .2, .2, .8], [1, .5, 0, 0])])[None]
    scores = np.asarray([[.9, .8, .7, .6]], f)
    lo, hi = np.nextafter(f(.5), f(0)), np.nextafter(f(.5), f(1))
    keep = {}
    for thr in (lo, f(.5), hi):
        r = run_nms(T, boxes, scores, max_output_size_per_class=10, max_total_size=10, iou_threshold=float(thr))
        keep[float(thr)] = list(r[4][0, :r[3][0]])
    assert keep[float(lo)] == [0, 2] and keep[float(f(.5))] == [0, 1, 2] and keep[float(hi)] == [0, 1, 2]


def test_nms_equal_scores_lower_index_first(T):
    rng = np.random.default_rng(2)
    from tfrpn import synthetic
    boxes, _ = synthetic.nms_boxes(rng, 2, 500, 0.05, 0.4)
    scores = (rng.integers(0, 5, size=(2, 500)).astype(F32)) / F32(8)
    check_nms(T, boxes, scores, max_output_size_per_class=100, max_total_size=100, iou_threshold=0.4)


@pytest.mark.parametrize("K,k,sthr", [(5000, 600, float("-inf")), (20000, 6000, float("-inf")), (3000, 3000, 0.2),
                                      (70000, 2000, float("-inf")), (4000, 50, 0.6)])
def test_nms_fused_pre_nms_topk(T, K, k, sthr):
    """pre_nms_topn = tf.nn.top_k + tf.gather (predictor.py:58-60) in front of the NMS: same result as the
    oracle's top_k -> gather -> combined NMS, keep indices mapped back to the K inputs."""
    from tfrpn import synthetic
    rng = np.random.default_rng(K + k)
    boxes, scores = synthetic.nms_boxes(rng, 2, K, 0.05, 0.4)
    scores[1, : K // 2] = np.round(scores[1, : K // 2] * 64) / 64          # tie groups straddling the top-k cut
    kw = dict(max_output_size_per_class=300, max_total_size=300, iou_threshold=0.6, score_threshold=sthr)
    got = run_nms(T, boxes, scores, pre_nms_topn=k, **kw)
    v, i = O.top_k(scores, k)
    sel = np.take_along_axis(boxes, i[..., None].astype(np.int64), axis=1)
    want = O.combined_non_max_suppression(sel.reshape(2, k, 1, 4), v.reshape(2, k, 1), return_indices=True, **kw)
    keep = np.where(want[4] >= 0, np.take_along_axis(i, np.maximum(want[4], 0).astype(np.int64), axis=1), -1)
    assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], keep)
    assert bits_equal(got[0], want[0]) and bits_equal(got[1], want[1])


def test_nms_large_k_truncated_candidates(T):
    """NMS over all K = 90000 boxes: the prefilter keeps the top ~6144 ranks; image 0 finishes inside them,
    image 1 (tiny disjoint boxes on a grid: nothing is ever suppressed, max_output 300 reached quickly) too,
    image 2 has only 40 boxes above the score threshold among its candidates' scores but all its boxes are
    near-duplicates, so the candidates run out before 300 are kept and the unfiltered kernel redoes it."""
    from tfrpn import synthetic
    rng = np.random.default_rng(90000)
    K = 90000
    boxes, scores = synthetic.nms_boxes(rng, 3, K, 0.05, 0.4)
    g = np.arange(K)
    y, x = (g // 300) / F32(300), (g % 300) / F32(300)
    boxes[1] = np.stack([y, x, y + F32(0.002), x + F32(0.002)], -1).astype(F32)
    c = rng.uniform(0.49, 0.51, size=(K, 2)); sz = rng.uniform(0.3, 0.31, size=(K, 2))
    boxes[2] = np.concatenate([c - sz / 2, c + sz / 2], -1).astype(F32)
    check_nms(T, boxes, scores, max_output_size_per_class=300, max_total_size=300, iou_threshold=0.7)
    check_nms(T, boxes, scores, max_output_size_per_class=300, max_total_size=300, iou_threshold=0.7,
              score_threshold=0.999)


# ---------------------------------------------------------------- composed proposal stage
def test_proposals_golden(T, golden):
    hp = T.train.get_hyper_params("vgg16")
    anchors = T.bbox.generate_anchors(hp)
    k, post, thr = golden["prop_k_post_thr"]
    b, s, v, keep = T.tfrpn.generate_proposals(T.cu(golden["pred_reg"]), T.cu(golden["pred_cls"]), anchors, hp,
                                               pre_nms_topn=int(k), post_nms_topn=int(post),
                                               nms_iou_threshold=float(thr))
    assert np.array_equal(T.np(v), golden["prop_valid"])
    assert bits_equal(T.np(s), golden["prop_scores"])
    assert close(T.np(b), golden["prop_boxes"])


@pytest.mark.parametrize("cfg,B", [("C1", 1), ("C2", 64), ("C3", 4), ("C4", 2)])
def test_proposals_vs_oracle(T, cfg, B):
    from tfrpn import synthetic
    bb, _, _, over = synthetic.CONFIGS[cfg]
    hp = dict(O.get_hyper_params(bb), **over)
    fh, fw = O._pair(hp["feature_map_shape"])
    anchors = O.generate_anchors(hp)
    rng = np.random.default_rng(7 + int(cfg[1]))
    reg, cls = synthetic.head_outputs(rng, B, fh, fw, 9)
    gb, gs, gv, gk = (T.np(x) for x in T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), T.cu(anchors), hp))
    # (a) chain parity, bit-exact: the oracle's top-k + NMS applied to the boxes the CUDA decode produced
    var = np.asarray(hp["variances"], F32)
    dec = T.np(T.bbox.get_bboxes_from_deltas(T.cu(anchors), T.cu((reg.reshape(B, -1, 4) * var).astype(F32))))
    dec = np.clip(dec, 0, 1)
    assert close(dec, np.clip(O.get_bboxes_from_deltas(anchors, reg.reshape(B, -1, 4) * var), 0, 1))
    sc = cls.reshape(B, -1)
    k = min(6000, sc.shape[1])
    ts, ti = O.top_k(sc, k)
    tb = np.take_along_axis(dec, ti[..., None].astype(np.int64), axis=1)
    nb, ns, _, nv, ni = O.combined_non_max_suppression(tb.reshape(B, k, 1, 4), ts.reshape(B, k, 1), 300, 300,
                                                       iou_threshold=0.7, return_indices=True)
    keep = np.where(ni >= 0, np.take_along_axis(ti, np.maximum(ni, 0).astype(np.int64), axis=1), -1)
    assert np.array_equal(gv, nv) and np.array_equal(gk, keep)
    assert bits_equal(gs, ns) and bits_equal(gb, nb)
    # (b) end to end against the pure oracle: identical keep lists unless a 1-ulp exp difference
    #     lands exactly on the 0.7 threshold (never observed on these seeds)
    ob, os_, ov, ok = O.generate_proposals(reg, cls, anchors, hp)
    assert np.array_equal(gv, ov) and np.array_equal(gk, ok)
    assert close(gb, ob) and bits_equal(gs, os_)


def test_proposals_large_feature_map_prefilter(T):
    """1333x800 at stride 8: N = 100*167*9 = 150300 anchors, pre-NMS top-6000 -> the fused stage runs on the
    prefiltered candidates (gathered deltas + anchors, keep indices mapped back through the remap)."""
    hp = dict(O.get_hyper_params("vgg16"), img_size=(800, 1333), feature_map_shape=(100, 167))
    anchors = O.generate_anchors(hp)
    from tfrpn import synthetic
    rng = np.random.default_rng(99)
    B = 2
    reg, cls = synthetic.head_outputs(rng, B, 100, 167, 9)
    gb, gs, gv, gk = (T.np(x) for x in T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), T.cu(anchors), hp))
    ob, os_, ov, ok = O.generate_proposals(reg, cls, anchors, hp)
    assert np.array_equal(gv, ov) and np.array_equal(gk, ok)
    assert close(gb, ob) and bits_equal(gs, os_)
    k = 50
    boxes, vals, idx = T.tfrpn.predict_top_boxes(T.cu(reg), T.cu(cls), T.cu(anchors), hp, k=k)
    ts, ti = O.top_k(cls.reshape(B, -1), k)
    assert np.array_equal(T.np(idx), ti) and bits_equal(T.np(vals), ts)
    var = np.asarray(hp["variances"], F32)
    dec = O.get_bboxes_from_deltas(anchors, reg.reshape(B, -1, 4) * var)
    assert close(T.np(boxes), np.take_along_axis(dec, ti[..., None].astype(np.int64), axis=1))


def test_proposals_host_buffers_equal_device_path(T):
    from tfrpn import synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    anchors = T.bbox.generate_anchors(hp)
    reg, cls = synthetic.head_outputs(np.random.default_rng(3), 3, 31, 31, 9)
    dev = T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), anchors, hp)
    host = T.tfrpn.generate_proposals(reg, cls, T.np(anchors), hp)
    for a, b in zip(dev, host):
        assert isinstance(b, np.ndarray) and np.array_equal(T.np(a), b)


# ---------------------------------------------------------------- boundary behaviour
def test_rejects_bad_inputs(T):
    hp = T.train.get_hyper_params("vgg16")
    anchors = T.bbox.generate_anchors(hp)
    with pytest.raises(ValueError):
        T.bbox.generate_iou_map(anchors.double(), T.cu(np.zeros((1, 2, 4), F32)))
    with pytest.raises(ValueError):
        T.bbox.get_bboxes_from_deltas(anchors, T.cu(np.zeros((2, 5, 4), F32)))
    with pytest.raises(ValueError):
        T.train.calculate_rpn_actual_outputs(anchors[:100], T.cu(np.zeros((1, 2, 4), F32)),
                                             T.cu(np.zeros((1, 2), np.int32)), hp)
    with pytest.raises(ValueError):    # q must be 1 or the number of classes (TF raises too)
        T.bbox.non_max_suppression(T.cu(np.zeros((1, 4, 2, 4), F32)), T.cu(np.zeros((1, 4, 3), F32)),
                                   max_output_size_per_class=2, max_total_size=2)
    with pytest.raises(ValueError):
        T.bbox.top_k_boxes(T.cu(np.zeros((1, 4), F32)), 5)


def test_c_abi_host_entry_points(T):
    """tfrpn_rpn_targets_host / tfrpn_proposals_host: host pointers in, host results out."""
    import ctypes as C
    from tfrpn import _lib, synthetic
    from tfrpn.proposals import proposal_cfg
    from tfrpn.utils.train_utils import _target_cfg
    hp = dict(O.get_hyper_params("vgg16"))
    anchors_np = O.generate_anchors(hp)
    anchors = T.cu(anchors_np)
    B, G, N = 5, 13, 8649
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(21), B, G)
    d = np.empty((B, N, 4), F32); l = np.empty((B, N), F32)
    lib, h = _lib.load(), _lib.handle(T.dev.index)
    tc = _target_cfg(hp, 5, 6, 0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.tfrpn_rpn_targets_host(h, anchors.data_ptr(), vp(gtb), vp(gtl), B, N, G, C.byref(tc), vp(d), vp(l), None))
    od, ol = O.calculate_rpn_actual_outputs(anchors_np, gtb, gtl, hp, seed=5, offset=6)
    assert bits_equal(l, ol.reshape(B, N)) and close(d, od)
    reg, cls = synthetic.head_outputs(np.random.default_rng(22), B, 31, 31, 9)
    pc = proposal_cfg(hp)
    ob = np.empty((B, 300, 4), F32); os_ = np.empty((B, 300), F32)
    ov = np.empty((B,), np.int32); ok = np.empty((B, 300), np.int32)
    _lib.check(lib.tfrpn_proposals_host(h, vp(reg), vp(cls), anchors.data_ptr(), B, N, C.byref(pc), vp(ob), vp(os_),
                                        vp(ov), vp(ok), None))
    wb, ws, wv, wk = O.generate_proposals(reg, cls, anchors_np, hp)
    assert np.array_equal(ov, wv) and np.array_equal(ok, wk) and bits_equal(os_, ws) and close(ob, wb)


def test_c_abi_fused_host_step_equals_separate_calls(T):
    """tfrpn_rpn_step_host (both halves, two streams, duplex copies) == the two host calls."""
    import ctypes as C
    from tfrpn import _lib, synthetic
    from tfrpn.proposals import proposal_cfg
    from tfrpn.utils.train_utils import _target_cfg
    hp = dict(O.get_hyper_params("vgg16"))
    anchors_np = O.generate_anchors(hp)
    anchors = T.cu(anchors_np)
    B, G, N, P = 6, 17, 8649, 300
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(31), B, G)
    reg, cls = synthetic.head_outputs(np.random.default_rng(32), B, 31, 31, 9)
    lib, h = _lib.load(), _lib.handle(T.dev.index)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    tc, pc = _target_cfg(hp, 8, 9, 0), proposal_cfg(hp)
    outs = []
    for fused in (False, True, True):
        d = np.empty((B, N, 4), F32); l = np.empty((B, N), F32)
        ob = np.empty((B, P, 4), F32); os_ = np.empty((B, P), F32)
        ov = np.empty((B,), np.int32); ok = np.empty((B, P), np.int32)
        if fused:
            _lib.check(lib.tfrpn_rpn_step_host(h, anchors.data_ptr(), vp(gtb), vp(gtl), B, N, G, C.byref(tc), vp(d), vp(l),
                                               vp(reg), vp(cls), C.byref(pc), vp(ob), vp(os_), vp(ov), vp(ok), None))
        else:
            _lib.check(lib.tfrpn_rpn_targets_host(h, anchors.data_ptr(), vp(gtb), vp(gtl), B, N, G, C.byref(tc), vp(d), vp(l), None))
            _lib.check(lib.tfrpn_proposals_host(h, vp(reg), vp(cls), anchors.data_ptr(), B, N, C.byref(pc), vp(ob), vp(os_),
                                                vp(ov), vp(ok), None))
        outs.append((d, l, ob, os_, ov, ok))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)
    od, ol = O.calculate_rpn_actual_outputs(anchors_np, gtb, gtl, hp, seed=8, offset=9)
    assert bits_equal(outs[1][1], ol.reshape(B, N)) and close(outs[1][0], od)


def test_c_abi_pipeline_in_flight_steps_equal_oracle(T):
    """tfrpn_pipeline_*: several host steps in flight (slots reused), pinned and pageable buffers,
    both halves / targets only / proposals only -- every step equals the oracle."""
    import ctypes as C
    from tfrpn import _lib, synthetic
    from tfrpn.proposals import proposal_cfg
    from tfrpn.utils.train_utils import _target_cfg
    hp = dict(O.get_hyper_params("vgg16"))
    anchors_np = O.generate_anchors(hp)
    anchors = T.cu(anchors_np)
    B, G, N, P = 9, 11, 8649, 300
    lib, h = _lib.load(), _lib.handle(T.dev.index)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    pipe = C.c_void_p()
    _lib.check(lib.tfrpn_pipeline_create(h, 2, C.byref(pipe)))
    pinned_ptrs = []

    def buf(shape, dtype, pinned):
        if not pinned:
            return np.empty(shape, dtype)
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _lib.check(lib.tfrpn_host_alloc(C.byref(p), n))
        pinned_ptrs.append(p)
        return np.frombuffer((C.c_char * n).from_address(p.value), dtype=dtype).reshape(shape)

    pc = proposal_cfg(hp)
    steps = []
    try:
        for i in range(5):
            pinned = i % 2 == 0
            mode = ("both", "targets", "proposals", "both", "both")[i]
            gtb0, gtl0 = synthetic.gt_batch(np.random.default_rng(100 + i), B, G)
            reg0, cls0 = synthetic.head_outputs(np.random.default_rng(200 + i), B, 31, 31, 9)
            st = dict(mode=mode, tc=_target_cfg(hp, 3, i, 7 * i))
            for k, (a, dt) in dict(gtb=(gtb0, F32), gtl=(gtl0, np.int32), reg=(reg0, F32), cls=(cls0, F32)).items():
                st[k] = buf(a.shape, dt, pinned)
                st[k][...] = a
            for k, (sh, dt) in dict(d=((B, N, 4), F32), l=((B, N), F32), ob=((B, P, 4), F32), os=((B, P), F32),
                                    ov=((B,), np.int32), ok=((B, P), np.int32)).items():
                st[k] = buf(sh, dt, pinned)
                st[k][...] = 77
            t_on, p_on = mode != "proposals", mode != "targets"
            ticket = C.c_int64(-1)
            _lib.check(lib.tfrpn_pipeline_submit(
                pipe, anchors.data_ptr(), B, N,
                vp(st["gtb"]) if t_on else None, vp(st["gtl"]) if t_on else None, G, C.byref(st["tc"]),
                vp(st["d"]) if t_on else None, vp(st["l"]) if t_on else None,
                vp(st["reg"]) if p_on else None, vp(st["cls"]) if p_on else None, C.byref(pc),
                vp(st["ob"]), vp(st["os"]), vp(st["ov"]), vp(st["ok"]), C.byref(ticket)))
            assert ticket.value == i
            st["ticket"] = ticket.value
            steps.append(st)
        _lib.check(lib.tfrpn_pipeline_wait(pipe, 4))        # out of order: 4 first, then the rest (already retired or not)
        for st in steps:
            _lib.check(lib.tfrpn_pipeline_wait(pipe, st["ticket"]))
        assert lib.tfrpn_pipeline_wait(pipe, 99) == -1
        _lib.check(lib.tfrpn_pipeline_drain(pipe))
        for i, st in enumerate(steps):
            if st["mode"] != "proposals":
                od, ol = O.calculate_rpn_actual_outputs(anchors_np, st["gtb"], st["gtl"], hp, seed=3, offset=i, image_offset=7 * i)
                assert bits_equal(st["l"], ol.reshape(B, N)) and close(st["d"], od)
            else:
                assert np.all(st["l"] == 77)
            if st["mode"] != "targets":
                wb, ws, wv, wk = O.generate_proposals(st["reg"], st["cls"], anchors_np, hp)
                assert np.array_equal(st["ov"], wv) and np.array_equal(st["ok"], wk)
                assert bits_equal(st["os"], ws) and close(st["ob"], wb)
            else:
                assert np.all(st["ov"] == 77)
    finally:
        _lib.check(lib.tfrpn_pipeline_destroy(pipe))
        for p in pinned_ptrs:
            lib.tfrpn_host_free(p)


def test_host_pipeline_acquired_slots_equal_oracle(T):
    """tfrpn.HostPipeline (tfrpn_pipeline_acquire / submit_acquired): inputs written into the slot's pinned
    block, one copy per direction, three steps in flight, changing batch shape -- every step equals the oracle."""
    from tfrpn import HostPipeline, synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    anchors_np = O.generate_anchors(hp)
    pipe = HostPipeline(hp, depth=3)
    N = pipe.N
    pending, checked = [], 0

    def check(item):
        t, v, mode, i, gtb, gtl, reg, cls = item
        pipe.wait(t)
        B = gtb.shape[0]
        if mode != "proposals":
            od, ol = O.calculate_rpn_actual_outputs(anchors_np, gtb, gtl, hp, seed=11, offset=i, image_offset=3 * i)
            assert bits_equal(v.labels, ol) and close(v.deltas, od)
        if mode != "targets":
            wb, ws, wv, wk = O.generate_proposals(reg, cls, anchors_np, hp)
            assert np.array_equal(v.valid, wv) and np.array_equal(v.keep_idx, wk)
            assert bits_equal(v.out_scores, ws) and close(v.out_boxes, wb)

    try:
        for i in range(8):
            B, G = (5, 9) if i < 5 else (7, 21)       # the slot staging regrows at step 5
            mode = ("both", "targets", "proposals", "both")[i % 4]
            gtb, gtl = synthetic.gt_batch(np.random.default_rng(300 + i), B, G)
            reg, cls = synthetic.head_outputs(np.random.default_rng(400 + i), B, 31, 31, 9)
            if len(pending) == 3:
                check(pending.pop(0)); checked += 1
            v = pipe.acquire(B, G)
            assert v.deltas.shape == (B, N, 4) and v.labels.shape == (B, 31, 31, 9)
            v.gt_boxes[...] = gtb; v.gt_labels[...] = gtl; v.rpn_reg[...] = reg; v.rpn_cls[...] = cls
            t = pipe.submit(targets=mode != "proposals", proposals=mode != "targets", seed=11, offset=i, image_offset=3 * i)
            pending.append((t, v, mode, i, gtb, gtl, reg, cls))
        while pending:
            check(pending.pop(0)); checked += 1
        assert checked == 8
    finally:
        pipe.close()


def test_prefetching_rpn_generator_matches_reference_generator(T):
    """train_utils.rpn_generator(prefetch=2) (utils/train_utils.py:67-82 with steps in flight) yields the
    same (img, (deltas, labels)) sequence as the oracle computes batch by batch, and cycles forever."""
    from tfrpn import synthetic
    hp = dict(O.get_hyper_params("vgg16"), seed=5)
    anchors_np = O.generate_anchors(hp)
    anchors = T.cu(anchors_np)
    data = []
    for i in range(3):
        gtb, gtl = synthetic.gt_batch(np.random.default_rng(500 + i), 4, 12)
        data.append(("img%d" % i, gtb, gtl))
    gen = T.train.rpn_generator(data, anchors, hp, prefetch=2)
    for step in range(7):                      # more than two epochs of the 3-batch dataset
        img, (deltas, labels) = next(gen)
        name, gtb, gtl = data[step % 3]
        assert img == name
        od, ol = O.calculate_rpn_actual_outputs(anchors_np, gtb, gtl, hp, seed=5, offset=step)
        assert bits_equal(labels, ol) and close(deltas, od)
    gen.close()


# ---------------------------------------------------------------- losses
# Per-entry terms are float32 in the reference's op order (logf within 2 ulp of Eigen's log); the
# sums are carried in float64 on both sides, so the scalar losses agree to ~1e-6 relative.
LOSS_RTOL = 2e-6


def loss_close(got, want, rtol=LOSS_RTOL):
    got, want = float(got), float(want)
    return (np.isnan(got) and np.isnan(want)) or abs(got - want) <= rtol * max(abs(want), 1e-30)


def test_losses_match_golden(T, golden):
    """cls_loss / reg_loss with the reference's signatures on the vectors its own source produced."""
    reg = T.train.reg_loss(T.cu(golden["loss_reg_true"]), T.cu(golden["loss_reg_pred"]))
    cls = T.train.cls_loss((T.cu(golden["loss_cls_true"]), T.cu(golden["loss_cls_pred"])))    # ((y_true, y_pred),) form
    assert reg.shape == () and cls.shape == ()
    assert loss_close(T.np(reg), golden["loss_reg"]) and loss_close(T.np(cls), golden["loss_cls"])


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_losses_on_real_targets(T, cfg):
    """Losses of synthetic head outputs against the targets the CUDA path itself assigned."""
    from tfrpn import synthetic
    bb, B, G, over = synthetic.CONFIGS[cfg]
    B = min(B, 16)
    hp = dict(O.get_hyper_params(bb), **over)
    rng = np.random.default_rng(77)
    anchors = O.generate_anchors(hp)
    gtb, gtl = synthetic.gt_batch(rng, B, G)
    fm = hp["feature_map_shape"]
    reg, cls = synthetic.head_outputs(rng, B, fm, fm, 9)
    deltas, labels = T.train.calculate_rpn_actual_outputs(T.cu(anchors), T.cu(gtb), T.cu(gtl), hp, seed=3, offset=0)
    r = T.train.rpn_losses(deltas, T.cu(reg), labels, T.cu(cls), with_grads=True)
    d_np, l_np = T.np(deltas), T.np(labels)
    assert loss_close(T.np(r["reg_loss"]), O.reg_loss(d_np, reg))
    assert loss_close(T.np(r["cls_loss"]), O.cls_loss(l_np, cls))
    assert int(r["n_pos"]) == int(np.any(d_np != 0, axis=-1).sum()) and int(r["n_cls"]) == int((l_np != -1).sum())
    gd, gl = O.loss_grads(d_np, reg, l_np, cls)
    assert r["grad_deltas"].shape == reg.shape and r["grad_labels"].shape == cls.shape
    assert close(T.np(r["grad_deltas"]), gd) and close(T.np(r["grad_labels"]), gl)
    # the single-loss entry points give the same numbers
    assert bits_equal(T.np(T.train.reg_loss(deltas, T.cu(reg))), T.np(r["reg_loss"]))
    assert bits_equal(T.np(T.train.cls_loss(labels, T.cu(cls))), T.np(r["cls_loss"]))


def test_losses_edge_cases(T):
    rng = np.random.default_rng(5)
    # no positives, no valid labels: reg = 0 / max(1, 0), cls = mean of nothing = NaN (as TF)
    t = np.zeros((2, 33, 4), F32)
    p = rng.normal(size=(2, 33, 4)).astype(F32)
    lab = np.full((2, 33), -1, F32)
    sc = rng.uniform(size=(2, 33)).astype(F32)
    r = T.train.rpn_losses(T.cu(t), T.cu(p), T.cu(lab), T.cu(sc), with_grads=True)
    assert float(r["reg_loss"]) == 0.0 and np.isnan(float(r["cls_loss"]))
    assert not T.np(r["grad_deltas"]).any() and not T.np(r["grad_labels"]).any()
    # scores outside (eps, 1 - eps) are clipped and get no gradient; Huber kink at |e| == delta
    lab = np.array([[1, 0, 1, 0, -1]], F32)
    sc = np.array([[0.0, 1.0, 1.0, 0.0, 0.5]], F32)
    t = np.zeros((1, 5, 4), F32)
    t[0, 0] = [1, 0, 0, 0]
    p = np.zeros((1, 5, 4), F32)
    p[0, 0] = [2, -1, 1, 5]          # e = [1, -1, 1, 5]
    r = T.train.rpn_losses(T.cu(t), T.cu(p), T.cu(lab), T.cu(sc), with_grads=True)
    assert loss_close(T.np(r["cls_loss"]), O.cls_loss(lab, sc)) and loss_close(T.np(r["reg_loss"]), O.reg_loss(t, p))
    gd, gl = O.loss_grads(t, p, lab, sc)
    assert bits_equal(T.np(r["grad_deltas"]), gd) and bits_equal(T.np(r["grad_labels"]), gl)
    assert np.array_equal(T.np(r["grad_deltas"])[0, 0], np.array([1, -1, 1, 1], F32))
    # NumPy in -> NumPy out, like every other drop-in
    out = T.train.reg_loss(t, p)
    assert isinstance(out, np.ndarray) and loss_close(out, O.reg_loss(t, p))
    with pytest.raises(ValueError):
        T.train.rpn_losses(T.cu(t), T.cu(p[:, :4]))


# ---------------------------------------------------------------- callers either side
@pytest.fixture(scope="module")
def nxt():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return dict(np.load(os.path.join(root, "tests", "golden", "next_vectors.npz")))


def test_predictor_body_matches_golden(T, nxt):
    """tfrpn.predict_top_boxes == predictor.py:52-60 exec'd from the reference file (k = 10, no clip)."""
    hp = dict(O.get_hyper_params("vgg16"))
    anchors = T.bbox.generate_anchors(hp)
    boxes, vals, idx = T.tfrpn.predict_top_boxes(T.cu(nxt["pred_reg"]), T.cu(nxt["pred_cls"]), anchors, hp, k=10)
    assert np.array_equal(T.np(idx), nxt["pred_top_indices"])
    assert close(T.np(boxes), nxt["pred_selected_bboxes"])
    assert bits_equal(T.np(vals), np.take_along_axis(nxt["pred_cls"].reshape(3, -1), nxt["pred_top_indices"].astype(np.int64), axis=1))


@pytest.mark.parametrize("B,k", [(2, 1), (5, 300), (3, 2000), (1, 8649)])
def test_predictor_body_vs_oracle(T, B, k):
    from tfrpn import synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    a_np = O.generate_anchors(hp)
    reg, cls = synthetic.head_outputs(np.random.default_rng(B * 31 + k), B, 31, 31, 9)
    boxes, vals, idx = T.tfrpn.predict_top_boxes(T.cu(reg), T.cu(cls), T.cu(a_np), hp, k=k)
    ob, ov, oi = O.predictor_top_boxes(reg, cls, a_np, hp, k=k)
    assert np.array_equal(T.np(idx), oi) and bits_equal(T.np(vals), ov) and close(T.np(boxes), ob)
    # same boxes as the unfused sequence decode -> top_k_boxes
    var = T.cu(np.asarray(hp["variances"], F32))
    full = T.bbox.get_bboxes_from_deltas(T.cu(a_np), T.cu(reg).reshape(B, -1, 4) * var)
    _, idx2, gat = T.bbox.top_k_boxes(T.cu(cls).reshape(B, -1), k, full)
    assert np.array_equal(T.np(idx2), oi) and bits_equal(T.np(gat), T.np(boxes))
    cb, _, _ = T.tfrpn.predict_top_boxes(T.cu(reg), T.cu(cls), T.cu(a_np), hp, k=k, clip=True)
    assert bits_equal(T.np(cb), np.clip(T.np(boxes), 0, 1))


def test_gt_padding_and_flip(T, nxt):
    from tfrpn.utils import data_utils
    assert bits_equal(T.np(data_utils.flip_horizontally_boxes(T.cu(nxt["flip_in"]))), nxt["flip_out"])
    rng = np.random.default_rng(9)
    counts = [3, 0, 7, 1, 50]
    bl = [rand_boxes(rng, n) for n in counts]
    ll = [rng.integers(0, 20, size=n) for n in counts]
    flip = [True, True, False, True, False]
    for G, add in [(None, 1), (50, 0), (4, 1)]:
        gb, gl = data_utils.pad_gt_batch(bl, ll, max_boxes=G, flip=flip, label_add=add)
        ob, ol = O.pad_gt_batch(bl, ll, max_boxes=G, flip=flip, label_add=add)
        assert bits_equal(T.np(gb), ob) and np.array_equal(T.np(gl), ol) and gl.dtype == T.torch.int32
    # the padded batch feeds target assignment exactly like a host-padded one
    hp = dict(O.get_hyper_params("vgg16"))
    a_np = O.generate_anchors(hp)
    gb, gl = data_utils.pad_gt_batch(bl, ll, flip=flip, label_add=1)
    d, l = T.train.calculate_rpn_actual_outputs(T.cu(a_np), gb, gl, hp, seed=4, offset=2)
    ob, ol = O.pad_gt_batch(bl, ll, flip=flip, label_add=1)
    od, olab = O.calculate_rpn_actual_outputs(a_np, ob, ol, hp, seed=4, offset=2)
    assert bits_equal(T.np(l), olab) and close(T.np(d), od)
    # an all-empty batch still pads to G >= 1
    gb, gl = data_utils.pad_gt_batch([np.zeros((0, 4), F32)] * 2, [np.zeros((0,), np.int32)] * 2)
    assert gb.shape == (2, 1, 4) and not T.np(gb).any() and list(T.np(gl).ravel()) == [-1, -1]


def test_targets_compact_form_equals_dense(T):
    """tfrpn_rpn_targets_compact: same labels, and scattering its (index, row) pairs gives bbox_deltas bit for bit."""
    import ctypes as C
    from tfrpn import _lib, synthetic
    hp = dict(O.get_hyper_params("vgg16"))
    a_np = O.generate_anchors(hp)
    B, G, N, TP = 6, 20, a_np.shape[0], hp["total_pos_bboxes"]
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(21), B, G)
    gtb[2], gtl[2] = 0, -1                                     # an image without ground truth: no rows at all
    anchors, dgtb, dgtl = T.cu(a_np), T.cu(gtb), T.cu(gtl)
    d, l = T.train.calculate_rpn_actual_outputs(anchors, dgtb, dgtl, hp, seed=9, offset=4)
    cfg = T.train._target_cfg(hp, 9, 4, 0)
    labels = T.torch.empty((B, N), device=T.dev)
    idx = T.torch.empty((B, TP), dtype=T.torch.int32, device=T.dev)
    rows = T.torch.empty((B, TP, 4), device=T.dev)
    lib = _lib.load()
    _lib.check(lib.tfrpn_rpn_targets_compact(_lib.handle(0), anchors.data_ptr(), dgtb.data_ptr(), dgtl.data_ptr(), B, N, G,
                                             C.byref(cfg), labels.data_ptr(), idx.data_ptr(), rows.data_ptr(),
                                             T.torch.cuda.current_stream().cuda_stream))
    assert bits_equal(T.np(labels).reshape(T.np(l).shape), T.np(l))
    idx_np, rows_np, d_np = T.np(idx), T.np(rows), T.np(d)
    dense = np.zeros((B, N, 4), F32)
    for b in range(B):
        k = int((idx_np[b] >= 0).sum())
        assert np.all(idx_np[b, :k] >= 0) and np.all(idx_np[b, k:] == -1) and len(set(idx_np[b, :k])) == k
        assert k == int((T.np(l)[b] == 1).sum())
        dense[b, idx_np[b, :k]] = rows_np[b, :k]
    assert bits_equal(dense, d_np) and (idx_np[2] == -1).all()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    host = np.empty((B, N, 4), F32)
    assert lib.tfrpn_expand_targets_host(vp(idx_np), vp(rows_np), B, N, TP, None, 0, vp(host)) == 0
    assert bits_equal(host, d_np)


def test_range_check_free_division_is_exact(T):
    """div_rn_inrange (the IoU kernels' division on 'nice' boxes) == __fdiv_rn, bit for bit, on 2^31 operand
    pairs of its domain, including zero numerators, a == b and extreme mantissas."""
    from tfrpn import _lib
    bad = T.torch.ones((1,), dtype=T.torch.int64, device=T.dev)
    for seed in (1, 2):
        _lib.check(_lib.load().tfrpn_selftest_division(1 << 30, seed, bad.data_ptr(), T.torch.cuda.current_stream().cuda_stream))
        assert int(bad.item()) == 0


def test_iou_map_nice_and_fallback_paths_agree_with_oracle(T):
    """K1 takes the range-check-free division when the tile's boxes are 'nice' and the generic path otherwise
    (a coordinate below 2^-16, a flipped GT box, a zero-area box): both must match the oracle bit for bit."""
    rng = np.random.default_rng(123)
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    from tfrpn import synthetic
    gtb, _ = synthetic.gt_batch(rng, 4, 50)
    assert bits_equal(T.np(T.bbox.generate_iou_map(T.cu(anchors), T.cu(gtb))), O.generate_iou_map(anchors, gtb))
    weird = gtb.copy()
    weird[0, 0] = [1e-7, 0.1, 0.3, 0.4]          # coordinate below 2^-16: not nice
    weird[1, 1] = [0.5, 0.5, 0.2, 0.2]           # flipped
    weird[2, 2] = [0.3, 0.3, 0.3, 0.9]           # degenerate but not all-zero
    assert bits_equal(T.np(T.bbox.generate_iou_map(T.cu(anchors), T.cu(weird))), O.generate_iou_map(anchors, weird))
    boxes = rand_boxes(rng, 4, 600)
    boxes[0, 5] = [0.2, 0.2, 0.2, 0.2]           # zero-area box in the tile: generic path, NaN-free here
    assert bits_equal(T.np(T.bbox.generate_iou_map(T.cu(boxes), T.cu(gtb))), O.generate_iou_map(boxes, gtb))
    gt52, _ = synthetic.gt_batch(rng, 4, 52)     # four columns per thread, nice and not nice
    assert bits_equal(T.np(T.bbox.generate_iou_map(T.cu(anchors), T.cu(gt52))), O.generate_iou_map(anchors, gt52))
    gt52[3, 1] = [0.5, 0.5, 0.2, 0.2]
    gt52[0, 2] = [1e-7, 0.1, 0.3, 0.4]
    assert bits_equal(T.np(T.bbox.generate_iou_map(T.cu(anchors), T.cu(gt52))), O.generate_iou_map(anchors, gt52))


# ---------------------------------------------------------------- anchors regenerated in registers (north star bullet 1)
@pytest.mark.parametrize("hp", [O.get_hyper_params("vgg16"), O.get_hyper_params("mobilenet_v2"), HP_C4],
                         ids=["vgg16", "mobilenet_v2", "c4_1333x800"])
def test_decode_with_regenerated_anchors_equals_anchor_tensor(T, hp):
    """tfrpn_decode_anchor_cfg (generate_anchors fused into the decode) == tfrpn_decode on the anchor tensor, bit for bit"""
    anchors = T.bbox.generate_anchors(hp)
    N = anchors.shape[0]
    rng = np.random.default_rng(17)
    deltas = T.cu(rng.normal(0, 0.5, size=(3, N, 4)).astype(F32))
    for var, clip in ((None, False), (hp["variances"], True)):
        want = T.bbox.get_bboxes_from_deltas(anchors, deltas * T.cu(np.asarray(var, F32)) if var is not None else deltas)
        if clip:
            want = want.clamp(0, 1)
        got = T.bbox.get_bboxes_from_hyper_params(hp, deltas, variances=var, clip=clip)
        assert T.torch.equal(got, want)


@pytest.mark.parametrize("cfg,B", [("C1", 1), ("C2", 64), ("C3", 16), ("C4", 3)])
def test_proposals_with_regenerated_anchors_equal_anchor_tensor(T, cfg, B):
    """tfrpn_proposals_anchor_cfg == tfrpn_proposals with the anchor tensor (and hence the oracle), bit for bit"""
    from tfrpn import synthetic
    bb, _, G, over = synthetic.CONFIGS[cfg]
    hp = dict(O.get_hyper_params(bb), **over)
    fm = hp["feature_map_shape"]
    fm_h, fm_w = (fm, fm) if isinstance(fm, int) else fm
    anchors = T.bbox.generate_anchors(hp)
    reg, cls = synthetic.head_outputs(np.random.default_rng(77), B, fm_h, fm_w, 9)
    a = T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), anchors, hp, pre_nms_topn=6000)
    b = T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), None, hp, pre_nms_topn=6000)
    for x, y in zip(a, b):
        assert T.torch.equal(x, y)


# ---------------------------------------------------------------- multi-class NMS (utils/bbox_utils.py:53-55 allows total_labels > 1)
@pytest.mark.parametrize("K,C,q,per_class,total,pad,sthr", [(200, 3, 1, 20, 30, False, float("-inf")), (300, 4, 4, 10, 100, True, 0.2),
                                                            (64, 2, 1, 5, 7, True, 0.5), (500, 5, 5, 50, 40, False, 0.1)])
def test_nms_multi_class_vs_oracle(T, K, C, q, per_class, total, pad, sthr):
    rng = np.random.default_rng(K + C)
    B = 3
    ctr = rng.uniform(0.1, 0.9, size=(B, K, q, 2)); sz = rng.uniform(0.05, 0.4, size=(B, K, q, 2))
    boxes = np.clip(np.concatenate([ctr - sz / 2, ctr + sz / 2], -1), 0, 1).astype(F32)
    scores = rng.uniform(0, 1, size=(B, K, C)).astype(F32)
    scores[0, :5, :] = 0.75                       # equal scores inside a class and across classes
    kw = dict(max_output_size_per_class=per_class, max_total_size=total, iou_threshold=0.5, score_threshold=sthr,
              pad_per_class=pad)
    got = [T.np(x) for x in T.bbox.non_max_suppression(T.cu(boxes), T.cu(scores), return_indices=True, **kw)]
    want = O.combined_non_max_suppression(boxes, scores, return_indices=True, **kw)
    assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4])
    assert bits_equal(got[0], want[0]) and bits_equal(got[1], want[1]) and bits_equal(got[2], want[2])
