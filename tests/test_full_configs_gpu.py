"""BASELINE.json configs 2-4 at their FULL batch sizes, and every template variant of the IoU/argmax kernel,
against the C restatement of the reference (oracle/rpn_oracle.c, itself pinned to the NumPy oracle and the
golden vectors by tests/test_oracle_cross.py).  The NumPy oracle needs ~1 GB temporaries at these sizes.

These are the exact kernels bench.py launches for C2 (B=64), C3 (B=128: rpn_iou_argmax_kernel<1>) and C4
(B=32, G=200: rpn_iou_argmax_kernel<4,packed>), plus the forced variants <2>, <8> and the scalar forms.

Float bars: dy/dx (IEEE-exact ops only) bit-exact; dh/dw (one logf) within 4 ulp; decoded boxes within 4 ulp
at the magnitude of the operands of their final add/sub (exp, then a cancelling subtraction).
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import rpn_oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32
EPS = float(np.finfo(F32).eps)   # 2^-23 = 1 ulp at 1.0


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, F32), np.ascontiguousarray(b, F32)
    return a.shape == b.shape and bool(np.all(a.view(np.uint32) == b.view(np.uint32)))


def ulp_distance(a, b):
    """distance in units in the last place between two float32 arrays (same sign assumed where it matters)"""
    def ordered(x):
        u = np.ascontiguousarray(x, F32).view(np.int32).astype(np.int64)
        return np.where(u < 0, -(u & 0x7FFFFFFF), u)
    return np.abs(ordered(a) - ordered(b))


def within_ulps(a, b, ulps=4, scale=None):
    """|a - b| <= ulps units in the last place of max(|b|, scale) (scale: magnitude of the operands of the
    final add / sub when the result is a cancelling difference)"""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    mag = np.abs(b64) if scale is None else np.maximum(np.abs(b64), np.asarray(scale, np.float64))
    return a64.shape == b64.shape and bool(np.all(np.abs(a64 - b64) <= ulps * EPS * np.maximum(mag, 1e-30)))


@pytest.fixture(scope="module")
def T(cuda_device):
    import torch
    import tfrpn
    from tfrpn import _lib, synthetic
    from tfrpn.utils import bbox_utils, train_utils
    if not CO.available():
        pytest.skip("oracle/_build/librpn_oracle.so not built (python -c 'import __graft_entry__ as g; g.build()')")

    class Ns:
        pass
    ns = Ns()
    ns.torch, ns.bbox, ns.train, ns.tfrpn, ns.dev, ns.lib, ns.syn = torch, bbox_utils, train_utils, tfrpn, cuda_device, _lib, synthetic
    ns.cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    ns.np = lambda t: t.detach().cpu().numpy()
    return ns


def config(T, name):
    bb, B, G, over = T.syn.CONFIGS[name]
    hp = dict(O.get_hyper_params(bb), **over)
    return hp, O.generate_anchors(hp), B, G


def compare_targets(d, l, dbg, od, ol, odbg):
    assert bits_equal(l.reshape(ol.shape), ol)                                   # labels {1,0,-1}
    assert np.array_equal(d != 0, od != 0)
    assert bits_equal(d[..., :2], od[..., :2])                                   # dy, dx
    assert within_ulps(d[..., 2:], od[..., 2:], 4)                               # dh, dw: one logf
    if dbg is not None:
        assert np.array_equal(dbg["argmax_row"], odbg["argmax_row"])
        assert np.array_equal(dbg["argmax_col"], odbg["argmax_col"])
        assert np.array_equal(dbg["max_iou"], odbg["max_iou"])
        assert np.array_equal(dbg["pos_pre"].astype(bool), odbg["pos_pre"].astype(bool))
        assert np.array_equal(dbg["neg_pre"].astype(bool), odbg["neg_pre"].astype(bool))


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_targets_full_batch_vs_c_oracle(T, name):
    hp, anchors, B, G = config(T, name)
    rng = np.random.default_rng(4000 + int(name[1]))
    gtb, gtl = T.syn.gt_batch(rng, B, G)
    if name == "C4":   # half of the images carry all 200 boxes (the config's stress case)
        for b in range(0, B, 2):
            c, sz = rng.uniform(0.1, 0.9, size=(G, 2)), rng.uniform(0.05, 0.6, size=(G, 2))
            gtb[b] = np.clip(np.concatenate([c - sz / 2, c + sz / 2], axis=1), 0, 1).astype(F32)
            gtl[b] = rng.integers(1, 21, size=G)
    d, l, dbg = T.train.calculate_rpn_actual_outputs(T.cu(anchors), T.cu(gtb), T.cu(gtl), hp, seed=99, offset=7,
                                                     image_offset=3, return_debug=True)
    od, ol, odbg = CO.rpn_targets(anchors, gtb, gtl, hp, seed=99, offset=7, image_offset=3, debug=True)
    compare_targets(T.np(d), T.np(l), {k: T.np(v) for k, v in dbg.items()}, od, ol, odbg)


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_proposals_full_batch_vs_c_oracle(T, name):
    hp, anchors, B, G = config(T, name)
    fm = hp["feature_map_shape"]
    fm_h, fm_w = (fm, fm) if isinstance(fm, int) else fm
    rng = np.random.default_rng(5000 + int(name[1]))
    reg, cls = T.syn.head_outputs(rng, B, fm_h, fm_w, 9)
    pb, ps, pv, pk = T.tfrpn.generate_proposals(T.cu(reg), T.cu(cls), T.cu(anchors), hp, pre_nms_topn=6000)
    ob, os_, ov, ok = CO.proposals(reg.reshape(B, -1, 4), cls.reshape(B, -1), anchors, hp, 6000)
    assert np.array_equal(T.np(pv), ov)
    assert np.array_equal(T.np(pk), ok)                                          # keep lists bit-exact, in order
    assert bits_equal(T.np(ps), os_)
    # decoded + clipped boxes: exp, then centre -+ half size (a cancelling difference of values <= ~1)
    assert within_ulps(T.np(pb), ob, 4, scale=1.0)


def fresh_handle(T, **env):
    """a library handle created with A/B switches set (they are read once, by tfrpn_create)"""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        h = C.c_void_p()
        T.lib.check(T.lib.load().tfrpn_create(C.byref(h), T.dev.index or 0))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return h


@pytest.mark.parametrize("apt", [1, 2, 4, 8])
@pytest.mark.parametrize("scalar", [False, True])
@pytest.mark.parametrize("shape", ["C2x8", "C4x2", "odd"])
def test_iou_argmax_kernel_variants(T, apt, scalar, shape):
    """rpn_iou_argmax_kernel<APT, packed> for every APT the launcher can pick, packed FP32 and scalar."""
    if shape == "C2x8":
        hp, anchors, _, G = config(T, "C2"); B = 8
    elif shape == "C4x2":
        hp, anchors, _, G = config(T, "C4"); B = 2
    else:   # N not a multiple of 32 * APT, odd G, one image of padding only
        hp = dict(O.get_hyper_params("vgg16"), feature_map_shape=7, anchor_scales=[64, 300], anchor_ratios=[1., 3., 0.25, 0.7, 1.3])
        hp["anchor_count"] = 10
        anchors, B, G = O.generate_anchors(hp), 5, 7
    N = anchors.shape[0]
    rng = np.random.default_rng(apt * 10 + int(scalar))
    gtb, gtl = T.syn.gt_batch(rng, B, G)
    gtb[-1] = 0; gtl[-1] = -1
    env = {"TFRPN_K2_APT": apt}
    if scalar:
        env["TFRPN_K2_SCALAR"] = 1
    h = fresh_handle(T, **env)
    lib = T.lib.load()
    try:
        torch = T.torch
        d = torch.empty((B, N, 4), device=T.dev); l = torch.empty((B, N), device=T.dev)
        dbg_t = dict(argmax_row=torch.empty((B, N), dtype=torch.int32, device=T.dev),
                     argmax_col=torch.empty((B, G), dtype=torch.int32, device=T.dev),
                     max_iou=torch.empty((B, N), device=T.dev),
                     pos_pre=torch.empty((B, N), dtype=torch.uint8, device=T.dev),
                     neg_pre=torch.empty((B, N), dtype=torch.uint8, device=T.dev),
                     pos_count=torch.empty((B,), dtype=torch.int32, device=T.dev),
                     neg_count=torch.empty((B,), dtype=torch.int32, device=T.dev))
        dbg = T.lib.TargetDebug(*[dbg_t[k].data_ptr() for k in ("argmax_row", "argmax_col", "max_iou", "pos_pre", "neg_pre",
                                                                "pos_count", "neg_count")])
        cfg = T.train._target_cfg(hp, 5, 11, 0)
        a_t, g_t, gl_t = T.cu(anchors), T.cu(gtb), T.cu(gtl)
        T.lib.check(lib.tfrpn_rpn_targets(h, a_t.data_ptr(), g_t.data_ptr(), gl_t.data_ptr(), B, N, G, C.byref(cfg),
                                          d.data_ptr(), l.data_ptr(), C.byref(dbg),
                                          torch.cuda.current_stream(T.dev).cuda_stream))
        torch.cuda.synchronize()
        od, ol, odbg = CO.rpn_targets(anchors, gtb, gtl, hp, seed=5, offset=11, debug=True)
        compare_targets(T.np(d), T.np(l), {k: T.np(v) for k, v in dbg_t.items()}, od, ol, odbg)
    finally:
        lib.tfrpn_destroy(h)


def test_decode_and_encode_ulp_bounds(T):
    """the 1e-6 bar of the other parity tests is absolute below 1; here the same ops get a ulp bound"""
    hp, anchors, _, _ = config(T, "C2")
    rng = np.random.default_rng(8)
    B, N = 16, anchors.shape[0]
    deltas = rng.normal(0, 0.5, size=(B, N, 4)).astype(F32)
    got = T.np(T.bbox.get_bboxes_from_deltas(T.cu(anchors), T.cu(deltas)))
    want = O.get_bboxes_from_deltas(anchors, deltas)
    h = want[..., 2] - want[..., 0]; w = want[..., 3] - want[..., 1]
    scale = np.stack([np.abs(want[..., 0]) + h, np.abs(want[..., 1]) + w, np.abs(want[..., 0]) + h, np.abs(want[..., 1]) + w], -1)
    assert within_ulps(got, want, 4, scale=scale)
    pts = np.sort(rng.uniform(0, 1, size=(B, N, 2, 2)), axis=-2).astype(F32)   # [..., 0, :] = (y1, x1) <= [..., 1, :] = (y2, x2)
    gt = np.ascontiguousarray(pts.reshape(B, N, 4))
    got = T.np(T.bbox.get_deltas_from_bboxes(T.cu(anchors), T.cu(gt)))
    want = O.get_deltas_from_bboxes(anchors, gt)
    assert bits_equal(got[..., :2], want[..., :2])
    assert within_ulps(got[..., 2:], want[..., 2:], 4)


M = {"TFRPN_NMS_PATH": "matrix"}


@pytest.mark.parametrize("env", [{}, M, dict(M, TFRPN_NMS_ROWS=64), dict(M, TFRPN_NMS_ROWS=256), dict(M, TFRPN_NMS_ROWS=1024),
                                 {"TFRPN_PROP_CLUSTER": 0}, {"TFRPN_PROP_CLUSTER": 12}, {"TFRPN_PROP_CLUSTER": 18},
                                 {"TFRPN_PROP_CLUSTER": 14}, {"TFRPN_PROP_CLUSTER": 1}, dict(M, TFRPN_PROP_CLUSTER=18)],
                         ids=["lazy", "matrix", "matrix64_redo", "matrix256_redo", "matrix1024", "lazy_one_cta", "lazy_2x512",
                              "lazy_8x256", "lazy_4x256", "lazy_1x1024", "matrix_rank_8x256"])
def test_nms_paths_agree_with_oracle(T, env):
    """the lazy NMS kernels (default) and the matrix NMS (rank launch + mask + sweep), also with too few rows (every
    image redone by the lazy kernel): the same keep lists as the oracle"""
    hp, anchors, _, _ = config(T, "C2")
    B, N, P = 9, anchors.shape[0], 300
    rng = np.random.default_rng(31)
    reg, cls = T.syn.head_outputs(rng, B, 31, 31, 9)
    h = fresh_handle(T, **env)
    lib = T.lib.load()
    torch = T.torch
    try:
        from tfrpn.proposals import proposal_cfg
        pc = proposal_cfg(hp, pre_nms_topn=6000)
        ob = torch.empty((B, P, 4), device=T.dev); os_ = torch.empty((B, P), device=T.dev)
        ov = torch.empty((B,), dtype=torch.int32, device=T.dev); ok = torch.empty((B, P), dtype=torch.int32, device=T.dev)
        r_t, c_t, a_t = T.cu(reg.reshape(B, -1, 4)), T.cu(cls.reshape(B, -1)), T.cu(anchors)
        st = torch.cuda.current_stream(T.dev).cuda_stream
        T.lib.check(lib.tfrpn_proposals(h, r_t.data_ptr(), c_t.data_ptr(), a_t.data_ptr(), B, N, C.byref(pc), ob.data_ptr(),
                                        os_.data_ptr(), ov.data_ptr(), ok.data_ptr(), st))
        torch.cuda.synchronize()
        wb, ws, wv, wk = CO.proposals(reg.reshape(B, -1, 4), cls.reshape(B, -1), anchors, hp, 6000)
        assert np.array_equal(T.np(ov), wv) and np.array_equal(T.np(ok), wk)
        assert bits_equal(T.np(os_), ws) and within_ulps(T.np(ob), wb, 4, scale=1.0)
        # plain NMS over given boxes with a score threshold and fewer candidates than rows
        K = 700
        boxes, scores = T.syn.nms_boxes(rng, 3, K)
        scores[1, 100:] = 0.001
        nc = T.lib.NmsCfg(300, 300, 0.5, 0.01, 0, 1, 0)
        nb = torch.empty((3, 300, 4), device=T.dev); ns = torch.empty((3, 300), device=T.dev); ncl = torch.empty((3, 300), device=T.dev)
        nv = torch.empty((3,), dtype=torch.int32, device=T.dev); nk = torch.empty((3, 300), dtype=torch.int32, device=T.dev)
        b_t, s_t = T.cu(boxes), T.cu(scores)
        T.lib.check(lib.tfrpn_nms(h, b_t.data_ptr(), s_t.data_ptr(), 3, K, C.byref(nc), nb.data_ptr(), ns.data_ptr(),
                                  ncl.data_ptr(), nv.data_ptr(), nk.data_ptr(), st))
        torch.cuda.synchronize()
        wb, ws, wc, wv, wk = CO.nms(boxes, scores, 300, 300, 0.5, 0.01)
        assert np.array_equal(T.np(nv), wv) and np.array_equal(T.np(nk), wk)
        assert bits_equal(T.np(ns), ws) and bits_equal(T.np(nb), wb)
    finally:
        lib.tfrpn_destroy(h)
