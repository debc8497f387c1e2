"""Known-answer tests derivable from IEEE-exact ops (SURVEY.md 8c list), against the oracle."""
import numpy as np

from oracle import rpn_oracle as O

F32 = np.float32


def test_kat_anchors():
    hp = O.get_hyper_params("vgg16")
    a = O.generate_anchors(hp)
    base = O.generate_base_anchors(hp)
    assert a.shape == (8649, 4)
    assert np.array_equal(a[0], np.asarray([0, 0, 0.14412903785705566, 0.14412903785705566], F32))
    assert np.array_equal(a[4], np.asarray([0, 0, 0.37816768884658813, 0.1971483677625656], F32))
    assert base[0, 2] == F32(0.12800000607967377) and base[0, 0] == -base[0, 2]
    assert np.array_equal(base[1, :2], np.asarray([-0.18101933598518372, -0.09050966799259186], F32))
    assert len(np.unique(a, axis=0)) == 7905
    hp2 = O.get_hyper_params("mobilenet_v2")
    a2 = O.generate_anchors(hp2)
    assert a2.shape == (9216, 4) and len(np.unique(a2, axis=0)) == 8384
    assert a2[0, 2] == F32(0.14362500607967377)
    area = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    assert abs(float(area.min()) - 0.02077318) < 1e-7


def test_kat_nonsquare_reduces_to_square():
    hp = O.get_hyper_params("vgg16")
    hp2 = dict(hp, img_size=(500, 500), feature_map_shape=(31, 31))
    assert np.array_equal(O.generate_anchors(hp), O.generate_anchors(hp2))
    hp3 = dict(hp, img_size=(800, 1333), feature_map_shape=(50, 84))
    assert O.generate_anchors(hp3).shape == (37800, 4)


def test_kat_iou():
    a = np.asarray([[0, 0, .5, .5]], F32)
    g = np.asarray([[[.25, .25, .75, .75], [0, 0, .5, .5], [.6, .6, .9, .9], [0, 0, 0, 0]]], F32)
    iou = O.generate_iou_map(a, g)[0, 0]
    assert iou[0] == F32(0.0625) / F32(0.4375) == F32(0.14285715)
    assert iou[1] == 1.0 and iou[2] == 0.0 and iou[3] == 0.0


def test_kat_encode_decode():
    a = np.asarray([[[0, 0, .5, .5]]], F32)
    g = np.asarray([[[.25, .25, .75, .75]]], F32)
    var = np.asarray([.1, .1, .2, .2], F32)
    d = O.get_deltas_from_bboxes(a, g)
    assert np.array_equal(d[0, 0], np.asarray([.5, .5, 0, 0], F32))
    assert np.array_equal((d / var)[0, 0], np.asarray([5, 5, 0, 0], F32))
    assert np.array_equal(O.get_deltas_from_bboxes(a, np.zeros_like(g))[0, 0], np.zeros(4, F32))
    back = O.get_bboxes_from_deltas(a, np.asarray([[[5, 5, 0, 0]]], F32) * var)
    assert np.array_equal(back[0, 0], g[0, 0])
    assert np.array_equal(O.get_bboxes_from_deltas(a, np.zeros((1, 1, 4), F32)), a)


def test_kat_labels_exact_match_and_ties():
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    n_int = (15 * 31 + 15) * 9  # interior, unclipped scale-128 anchor
    gt = np.zeros((1, 3, 4), F32)
    gl = np.full((1, 3), -1, np.int32)
    gt[0, 0] = anchors[n_int]; gl[0, 0] = 1
    uniq, inv, counts = np.unique(anchors, axis=0, return_inverse=True, return_counts=True)
    dup = np.flatnonzero(counts[inv.reshape(-1)] > 1)  # clipping makes duplicate anchors (fact 0.4)
    gt[0, 1] = anchors[dup[-1]]; gl[0, 1] = 2          # GT == a duplicated box: IoU 1 ties
    d, l, dbg = O.calculate_rpn_actual_outputs(anchors, gt, gl, hp, return_debug=True)
    assert dbg["argmax_col"][0, 0] == n_int and dbg["max_iou"][0, n_int] == 1.0
    iou = O.generate_iou_map(anchors, gt)[0]
    tied = np.flatnonzero(iou[:, 1] == iou[:, 1].max())
    assert tied.size > 1 and dbg["argmax_col"][0, 1] == tied.min()
    assert dbg["pos_pre"][0, n_int] and np.all(dbg["pos_pre"][0][dbg["max_iou"][0] > F32(0.7)])
    # the padded GT (IoU 0 everywhere) does not force anchor 0 positive
    assert not dbg["pos_pre"][0, 0] or dbg["max_iou"][0, 0] > F32(0.7)


def test_kat_dropped_forced_positive_may_become_negative():
    hp = O.get_hyper_params("vgg16", total_pos_bboxes=1, total_neg_bboxes=8000)
    anchors = O.generate_anchors(hp)
    gt = np.zeros((1, 2, 4), F32); gl = np.asarray([[1, 1]], np.int32)
    gt[0, 0] = [0.40, 0.40, 0.41, 0.41]      # tiny: best IoU < 0.3
    gt[0, 1] = anchors[(26 * 31 + 26) * 9]
    found = False
    for seed in range(8):
        d, l, dbg = O.calculate_rpn_actual_outputs(anchors, gt, gl, hp, seed=seed, return_debug=True)
        n_tiny = dbg["argmax_col"][0, 0]
        assert dbg["max_iou"][0, n_tiny] < F32(0.3) and dbg["pos_pre"][0, n_tiny]
        assert dbg["pos_count"][0] == 1
        if not dbg["pos"][0, n_tiny]:
            assert dbg["neg_pre"][0, n_tiny]
            found = True
    assert found


def test_kat_nms_threshold_and_degenerate():
    f = F32
    b0 = np.asarray([0, 0, 1, 1], f)
    b1 = np.asarray([0, 0, 1, .5], f)          # IoU(b0,b1) = 0.5 exactly
    zero = np.asarray([.2, .2, .2, .8], f)     # zero area: never suppressed, never suppresses
    flip = np.asarray([1, .5, 0, 0], f)        # b1 with flipped corners
    boxes = np.stack([b0, b1, zero, flip])[None, :, None, :]
    scores = np.asarray([.9, .8, .7, .6], f)[None, :, None]
    lo, hi = np.nextafter(f(.5), f(0)), np.nextafter(f(.5), f(1))
    keep_at = {}
    for thr in (lo, f(.5), hi):
        _, _, _, nv, ni = O.combined_non_max_suppression(boxes, scores, 10, 10, iou_threshold=thr,
                                                         return_indices=True)
        keep_at[float(thr)] = list(ni[0, :nv[0]])
    assert keep_at[float(lo)] == [0, 2]            # 0.5 > lo: b1 and its flipped twin suppressed
    assert keep_at[float(f(.5))] == [0, 1, 2]      # strict '>': b1 kept, twin (IoU 1 with b1) dropped
    assert keep_at[float(hi)] == [0, 1, 2]


def test_kat_topk_ties_ascending_index():
    s = np.asarray([[.5, .9, .5, .9, .1, .5]], F32)
    v, i = O.top_k(s, 5)
    assert list(i[0]) == [1, 3, 0, 2, 5] and list(v[0]) == [F32(.9), F32(.9), F32(.5), F32(.5), F32(.5)]


def test_sampler_counts_and_uniformity():
    rng = np.random.default_rng(0)
    mask = rng.uniform(size=(4, 500)) < 0.3
    sel = O.randomly_select_xyz_mask(mask, [50], seed=3)
    assert np.all(sel <= mask) and np.array_equal(sel.sum(-1), np.minimum(mask.sum(-1), 50))
    sel2 = O.randomly_select_xyz_mask(mask, np.asarray([0, 10, 1000, 7]), seed=3, stream=1)
    assert list(sel2.sum(-1)) == [0, 10, int(mask[2].sum()), 7]
    # chi-square over offsets: each True entry is kept with equal probability
    m = np.ones((1, 64), bool)
    hits = np.zeros(64)
    T = 600
    for off in range(T):
        hits += O.randomly_select_xyz_mask(m, [16], seed=11, offset=off)[0]
    exp = T * 16 / 64
    chi2 = ((hits - exp) ** 2 / (exp * (1 - 16 / 64))).sum()
    assert chi2 < 110, chi2      # 63 dof: p(chi2 > 110) ~ 2e-4


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = O.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = O.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
