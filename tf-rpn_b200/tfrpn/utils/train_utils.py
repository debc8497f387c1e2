"""Drop-in for the hot-path half of the reference's utils/train_utils.py on libtfrpn_cuda.so."""
import ctypes as C
import itertools
import math

import torch

from .. import _lib
from .._tensor import Origin, from_device, ptr, stream_ptr, to_device
from . import bbox_utils

F32 = torch.float32

RPN = {
    "vgg16": {
        "img_size": 500,
        "feature_map_shape": 31,
        "anchor_ratios": [1., 2., 1. / 2.],
        "anchor_scales": [128, 256, 512],
    },
    "mobilenet_v2": {
        "img_size": 500,
        "feature_map_shape": 32,
        "anchor_ratios": [1., 2., 1. / 2.],
        "anchor_scales": [128, 256, 512],
    },
}


def get_hyper_params(backbone, **kwargs):
    """utils/train_utils.py:20-38, including its quirks: kwargs override only keys that already
    exist and only with truthy values, and the module-level RPN[backbone] dict is the object
    returned (the reference mutates it too)."""
    hyper_params = RPN[backbone]
    hyper_params["test_nms_topn"] = 300
    hyper_params["total_pos_bboxes"] = 128
    hyper_params["total_neg_bboxes"] = 128
    hyper_params["variances"] = [0.1, 0.1, 0.2, 0.2]
    for key, value in kwargs.items():
        if key in hyper_params and value:
            hyper_params[key] = value
    hyper_params["anchor_count"] = len(hyper_params["anchor_ratios"]) * len(hyper_params["anchor_scales"])
    return hyper_params


def get_step_size(total_items, batch_size):
    """utils/train_utils.py:40-48."""
    return math.ceil(total_items / batch_size)


# The reference draws fresh tf.random numbers on every call; the counter RNG advances an offset
# instead.  Pass seed/offset explicitly (or hyper_params["seed"]) for a reproducible / resumed run.
_auto_offset = itertools.count()


def randomly_select_xyz_mask(mask, select_xyz, seed=0, offset=None, stream=0, image_offset=0):
    """utils/train_utils.py:50-65: keep min(#True, select) True entries per row, chosen by a
    counter RNG (Philox4x32-10 keyed by seed, counter (n, image, offset)) instead of tf.random."""
    o = Origin()
    m = to_device(mask, torch.uint8, o, "mask")
    sel = to_device(select_xyz, torch.int32, o, "select_xyz").reshape(-1)
    if m.dim() != 2:
        raise ValueError("mask must be (batch_size, m)")
    B, N = m.shape
    if sel.numel() not in (1, B):
        raise ValueError("select_xyz must have 1 or batch_size entries")
    off = next(_auto_offset) if offset is None else int(offset)
    out = torch.empty_like(m)
    dev = m.device
    _lib.check(_lib.load().tfrpn_select_mask(_lib.handle(dev.index), ptr(m), ptr(sel), sel.numel(), B, N,
                                             int(seed), off, int(stream), int(image_offset), ptr(out),
                                             stream_ptr(dev)))
    return from_device(out.to(torch.bool), o)


def rpn_generator(dataset, anchors, hyper_params, prefetch=0):
    """utils/train_utils.py:67-82 (glue kept verbatim in behaviour: the call site of the path).
    ``prefetch`` > 0 (host batches only) keeps that many steps in flight through tfrpn.pipeline."""
    if prefetch:
        from ..pipeline import prefetching_rpn_generator
        yield from prefetching_rpn_generator(dataset, anchors, hyper_params, prefetch=prefetch,
                                             seed=hyper_params.get("seed", 0))
        return
    while True:
        for image_data in dataset:
            img, gt_boxes, gt_labels = image_data
            bbox_deltas, bbox_labels = calculate_rpn_actual_outputs(anchors, gt_boxes, gt_labels, hyper_params)
            yield img, (bbox_deltas, bbox_labels)


def _target_cfg(hyper_params, seed, offset, image_offset):
    cfg = _lib.TargetCfg()
    cfg.pos_iou_threshold = float(hyper_params.get("pos_iou_threshold", 0.7))
    cfg.neg_iou_threshold = float(hyper_params.get("neg_iou_threshold", 0.3))
    cfg.total_pos = int(hyper_params["total_pos_bboxes"])
    cfg.total_neg = int(hyper_params["total_neg_bboxes"])
    for i, v in enumerate(hyper_params["variances"]):
        cfg.variances[i] = float(v)
    cfg.seed = int(hyper_params.get("seed", 0) if seed is None else seed)
    cfg.offset = next(_auto_offset) if offset is None else int(offset)
    cfg.image_offset = int(image_offset)
    return cfg


def calculate_rpn_actual_outputs(anchors, gt_boxes, gt_labels, hyper_params, seed=None, offset=None,
                                 image_offset=0, return_debug=False):
    """utils/train_utils.py:84-144.

    anchors (N,4); gt_boxes (B,G,4) zero padded; gt_labels (B,G) int32, -1 padded.
    Returns bbox_deltas (B,N,4) and bbox_labels (B,fm_h,fm_w,anchor_count) in {1,0,-1}.
    ``image_offset`` = global index of image 0 when the batch is a shard of a larger one, so
    that sharded and unsharded runs sample identically.
    """
    o = Origin()
    gtb = to_device(gt_boxes, F32, o, "gt_boxes")
    gtl = to_device(gt_labels, torch.int32, o, "gt_labels")
    anc = to_device(anchors, F32, o, "anchors")
    if gtb.dim() != 3 or gtb.shape[-1] != 4 or gtl.shape != gtb.shape[:2]:
        raise ValueError("gt_boxes must be (B,G,4) and gt_labels (B,G)")
    if anc.dim() != 2 or anc.shape[-1] != 4:
        raise ValueError("anchors must be (total_anchors, 4)")
    B, G = gtl.shape
    N = anc.shape[0]
    fm_h, fm_w = bbox_utils._pair(hyper_params["feature_map_shape"])
    A = int(hyper_params["anchor_count"])
    if fm_h * fm_w * A != N:
        raise ValueError("anchors (%d) do not match feature_map_shape x anchor_count (%d)" % (N, fm_h * fm_w * A))
    cfg = _target_cfg(hyper_params, seed, offset, image_offset)
    dev = gtb.device
    deltas = torch.empty((B, N, 4), dtype=F32, device=dev)
    labels = torch.empty((B, N), dtype=F32, device=dev)
    dbg_struct, dbg = None, None
    if return_debug:
        i32, u8 = torch.int32, torch.uint8
        dbg = dict(argmax_row=torch.empty((B, N), dtype=i32, device=dev),
                   argmax_col=torch.empty((B, G), dtype=i32, device=dev),
                   max_iou=torch.empty((B, N), dtype=F32, device=dev),
                   pos_pre=torch.empty((B, N), dtype=u8, device=dev),
                   neg_pre=torch.empty((B, N), dtype=u8, device=dev),
                   pos_count=torch.empty((B,), dtype=i32, device=dev),
                   neg_count=torch.empty((B,), dtype=i32, device=dev))
        dbg_struct = _lib.TargetDebug(*(ptr(dbg[k]) for k in ("argmax_row", "argmax_col", "max_iou", "pos_pre",
                                                              "neg_pre", "pos_count", "neg_count")))
    _lib.check(_lib.load().tfrpn_rpn_targets(
        _lib.handle(dev.index), ptr(anc), ptr(gtb), ptr(gtl), B, N, G, C.byref(cfg), ptr(deltas), ptr(labels),
        C.byref(dbg_struct) if dbg_struct is not None else None, stream_ptr(dev)))
    out = (from_device(deltas, o), from_device(labels.reshape(B, fm_h, fm_w, A), o))
    if return_debug:
        return out + ({k: from_device(v, o) for k, v in dbg.items()},)
    return out
