"""ctypes wrapper of oracle/rpn_oracle.c (built by `make -C oracle`).  TEST INFRASTRUCTURE ONLY:
the independent C restatement used to cross-check oracle/rpn_oracle.py and timed as bench.py's CPU
baseline.  Never imported by the product package."""
import ctypes as C
import os

import numpy as np

F32 = np.float32
_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "librpn_oracle.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("build the C oracle first: make -C oracle")
        _lib = C.CDLL(_PATH)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=F32)


def iou_map(boxes, gt):
    boxes, gt = _f(boxes), _f(gt)
    B, G = gt.shape[:2]
    N = boxes.shape[-2]
    out = np.empty((B, N, G), F32)
    lib().oracle_iou_map(_p(boxes), C.c_int(boxes.ndim == 3), _p(gt), B, N, G, _p(out))
    return out


def decode(anchors, deltas):
    anchors, deltas = _f(anchors), _f(deltas)
    B, N = deltas.shape[:2]
    out = np.empty((B, N, 4), F32)
    lib().oracle_decode(_p(anchors), C.c_int(anchors.ndim == 3), _p(deltas), B, N, _p(out))
    return out


def encode(boxes, gt):
    boxes, gt = _f(boxes), _f(gt)
    B, N = gt.shape[:2]
    out = np.empty((B, N, 4), F32)
    lib().oracle_encode(_p(boxes), C.c_int(boxes.ndim == 3), _p(gt), B, N, _p(out))
    return out


def select_mask(mask, select, seed=0, offset=0, stream=0, image_offset=0):
    mask = np.ascontiguousarray(mask, np.uint8)
    B, N = mask.shape
    sel = np.ascontiguousarray(np.asarray(select).reshape(-1), np.int32)
    out = np.empty((B, N), np.uint8)
    lib().oracle_select_mask(_p(mask), _p(sel), sel.size, B, N, C.c_uint64(seed), C.c_uint64(offset), stream,
                             image_offset, _p(out))
    return out.astype(bool)


def rpn_targets(anchors, gt_boxes, gt_labels, hp, seed=0, offset=0, image_offset=0, threads=0, debug=False):
    anchors, gt_boxes = _f(anchors), _f(gt_boxes)
    gt_labels = np.ascontiguousarray(gt_labels, np.int32)
    B, G = gt_labels.shape
    N = anchors.shape[0]
    var = _f(hp["variances"])
    deltas = np.empty((B, N, 4), F32)
    labels = np.empty((B, N), F32)
    dbg = {}
    if debug:
        dbg = dict(argmax_row=np.empty((B, N), np.int32), argmax_col=np.empty((B, G), np.int32),
                   max_iou=np.empty((B, N), F32), pos_pre=np.empty((B, N), np.uint8),
                   neg_pre=np.empty((B, N), np.uint8))
    lib().oracle_rpn_targets(_p(anchors), _p(gt_boxes), _p(gt_labels), B, N, G,
                             C.c_float(hp.get("pos_iou_threshold", 0.7)), C.c_float(hp.get("neg_iou_threshold", 0.3)),
                             int(hp["total_pos_bboxes"]), int(hp["total_neg_bboxes"]), _p(var), C.c_uint64(seed),
                             C.c_uint64(offset), image_offset, threads, _p(deltas), _p(labels),
                             _p(dbg.get("argmax_row")), _p(dbg.get("argmax_col")), _p(dbg.get("max_iou")),
                             _p(dbg.get("pos_pre")), _p(dbg.get("neg_pre")))
    if debug:
        return deltas, labels, dbg
    return deltas, labels


def top_k(scores, k):
    scores = _f(scores)
    B, N = scores.shape
    v = np.empty((B, k), F32)
    i = np.empty((B, k), np.int32)
    lib().oracle_topk(_p(scores), B, N, k, _p(v), _p(i))
    return v, i


def nms(boxes, scores, per_class, total, iou_threshold=0.5, score_threshold=float("-inf"), pad_per_class=False,
        clip_boxes=True):
    boxes, scores = _f(boxes), _f(scores)
    B, K = scores.shape
    rows = min(total, per_class) if pad_per_class else total
    ob = np.empty((B, rows, 4), F32); os_ = np.empty((B, rows), F32); oc = np.empty((B, rows), F32)
    ov = np.empty((B,), np.int32); ok = np.empty((B, rows), np.int32)
    lib().oracle_nms(_p(boxes), _p(scores), B, K, per_class, total, C.c_float(iou_threshold),
                     C.c_float(score_threshold), int(pad_per_class), int(clip_boxes), _p(ob), _p(os_), _p(oc), _p(ov),
                     _p(ok))
    return ob, os_, oc, ov, ok


def proposals(rpn_reg, rpn_cls, anchors, hp, pre_nms_topn=6000, post_nms_topn=None, thr=0.7, clip=True, threads=0):
    rpn_reg, rpn_cls, anchors = _f(rpn_reg), _f(rpn_cls), _f(anchors)
    B = rpn_reg.shape[0]
    N = anchors.shape[0]
    post = int(post_nms_topn if post_nms_topn is not None else hp["test_nms_topn"])
    var = _f(hp["variances"])
    ob = np.empty((B, post, 4), F32); os_ = np.empty((B, post), F32)
    ov = np.empty((B,), np.int32); ok = np.empty((B, post), np.int32)
    lib().oracle_proposals(_p(rpn_reg), _p(rpn_cls), _p(anchors), B, N, _p(var), int(pre_nms_topn), post,
                           C.c_float(thr), int(clip), threads, _p(ob), _p(os_), _p(ov), _p(ok))
    return ob, os_, ov, ok
