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


def _loss_args(args):
    """The reference accepts (y_true, y_pred) or ((y_true, y_pred),) (utils/train_utils.py:156,174)."""
    y_true, y_pred = args if len(args) == 2 else args[0]
    return y_true, y_pred


def rpn_losses(true_deltas=None, pred_deltas=None, true_labels=None, pred_labels=None, huber_delta=1.0,
               with_grads=False):
    """Both RPN losses (utils/train_utils.py:146-185) in one pass over the target tensors.

    true_deltas (B,N,4) / pred_deltas (B,...) reshaped to (B,N,4) as reg_loss does (:175);
    true_labels / pred_labels any matching shape, e.g. (B,F,F,A).  Either pair may be None.
    Returns a dict with 0-d ``reg_loss`` / ``cls_loss`` tensors, ``n_pos`` / ``n_cls`` counts and,
    with ``with_grads``, ``grad_deltas`` / ``grad_labels`` (d loss / d prediction, prediction-shaped).
    """
    o = Origin()
    td = pd = tl = pl = None
    B = N = None
    pd_shape = pl_shape = None
    if true_deltas is not None:
        td = to_device(true_deltas, F32, o, "true_deltas")
        pd = to_device(pred_deltas, F32, o, "pred_deltas")
        pd_shape = tuple(pd.shape)
        if td.dim() != 3 or td.shape[-1] != 4 or pd.numel() != td.numel() or pd.shape[0] != td.shape[0]:
            raise ValueError("true_deltas must be (B,N,4) and pred_deltas reshapeable to it")
        B, N = td.shape[:2]
    if true_labels is not None:
        tl = to_device(true_labels, F32, o, "true_labels")
        pl = to_device(pred_labels, F32, o, "pred_labels")
        pl_shape = tuple(pl.shape)
        if tl.shape != pl.shape:
            raise ValueError("true_labels and pred_labels must have the same shape")
        if B is None:
            B, N = 1, tl.numel()      # cls_loss is a flat mean: the batch split does not matter
        elif tl.numel() != B * N:
            raise ValueError("labels (%d entries) do not match deltas (%d rows)" % (tl.numel(), B * N))
    if B is None:
        raise ValueError("rpn_losses needs at least one (true, predicted) pair")
    dev = (td if td is not None else tl).device
    out = torch.empty((4,), dtype=F32, device=dev)
    gd = torch.empty_like(pd) if (with_grads and pd is not None) else None
    gl = torch.empty_like(pl) if (with_grads and pl is not None) else None
    _lib.check(_lib.load().tfrpn_rpn_losses(_lib.handle(dev.index), ptr(td), ptr(pd), ptr(tl), ptr(pl), B, N,
                                            float(huber_delta), ptr(out), ptr(gd), ptr(gl), stream_ptr(dev)))
    counts = out.view(torch.int32)
    res = {"n_pos": from_device(counts[2], o), "n_cls": from_device(counts[3], o)}
    if td is not None:
        res["reg_loss"] = from_device(out[0], o)
    if tl is not None:
        res["cls_loss"] = from_device(out[1], o)
    if gd is not None:
        res["grad_deltas"] = from_device(gd.reshape(pd_shape), o)
    if gl is not None:
        res["grad_labels"] = from_device(gl.reshape(pl_shape), o)
    return res


def cls_loss(*args):
    """utils/train_utils.py:146-161: BinaryCrossentropy over the entries whose true label is not -1."""
    y_true, y_pred = _loss_args(args)
    return rpn_losses(true_labels=y_true, pred_labels=y_pred)["cls_loss"]


def reg_loss(*args):
    """utils/train_utils.py:163-185: Huber over the rows with a non-zero true delta / max(1, #rows)."""
    y_true, y_pred = _loss_args(args)
    return rpn_losses(true_deltas=y_true, pred_deltas=y_pred)["reg_loss"]
