"""Drop-in for the two box-side pieces of the reference's utils/data_utils.py that sit just before
target assignment: flip_horizontally's box transform (:54-68) and the padded
batch (:145-157, boxes padded with 0, labels with -1) -- on the GPU, from ragged per-image arrays.
The image pipeline itself (tfds, resize, JPEG) is out of scope."""
import numpy as np
import torch

from .. import _lib
from .._tensor import Origin, default_device, from_device, ptr, stream_ptr, to_device

F32 = torch.float32


def get_padding_values():
    """utils/data_utils.py:152-157 (images, boxes, labels) as plain numbers."""
    return (0.0, 0.0, -1)


def pad_gt_batch(gt_boxes_list, gt_labels_list, max_boxes=None, flip=None, label_add=0):
    """padded_batch of variable-length ground truth (trainer.py:44-45 with data_utils.py:145-157).

    gt_boxes_list[b] (n_b,4) float32, gt_labels_list[b] (n_b,) int.  ``flip`` (B,) bools applies
    flip_horizontally's box transform [y1, 1-x2, y2, 1-x1] (:66-69) to the flagged images;
    ``label_add`` = 1 reproduces preprocessing's ``label + 1`` (:20).  G = ``max_boxes`` or the
    longest list (at least 1).  Returns device tensors gt_boxes (B,G,4), gt_labels (B,G) int32.
    """
    B = len(gt_boxes_list)
    if B != len(gt_labels_list):
        raise ValueError("gt_boxes_list and gt_labels_list differ in length")
    counts = [int(np.shape(b)[0]) for b in gt_boxes_list]
    for n, lab in zip(counts, gt_labels_list):
        if int(np.shape(lab)[0]) != n:
            raise ValueError("boxes and labels of an image differ in length")
    G = int(max_boxes) if max_boxes is not None else max(counts + [1])
    dev = default_device()
    offsets = np.zeros(B + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    M = int(offsets[-1])
    flat_b = np.zeros((max(M, 1), 4), np.float32)
    flat_l = np.zeros((max(M, 1),), np.int32)
    for b in range(B):
        if counts[b]:
            bx = gt_boxes_list[b]
            bx = bx.detach().cpu().numpy() if isinstance(bx, torch.Tensor) else np.asarray(bx)
            if bx.dtype != np.float32:
                raise ValueError("gt boxes must be float32, got %s" % bx.dtype)
            lb = gt_labels_list[b]
            lb = lb.detach().cpu().numpy() if isinstance(lb, torch.Tensor) else np.asarray(lb)
            flat_b[offsets[b]:offsets[b + 1]] = bx
            flat_l[offsets[b]:offsets[b + 1]] = lb
    o = Origin()
    d_b = to_device(flat_b, F32, o, "gt_boxes")
    d_l = to_device(flat_l, torch.int32, o, "gt_labels")
    d_o = to_device(offsets, torch.int32, o, "offsets")
    d_f = to_device(np.asarray(flip, np.uint8), torch.uint8, o, "flip") if flip is not None else None
    if d_f is not None and d_f.numel() != B:
        raise ValueError("flip must have one entry per image")
    out_b = torch.empty((B, G, 4), dtype=F32, device=dev)
    out_l = torch.empty((B, G), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().tfrpn_pad_gt(ptr(d_b), ptr(d_l), ptr(d_o), ptr(d_f), B, G, int(label_add), ptr(out_b),
                                        ptr(out_l), stream_ptr(dev)))
    return out_b, out_l


def flip_horizontally_boxes(gt_boxes):
    """The box half of utils/data_utils.py:54-68: [y1, 1 - x2, y2, 1 - x1] for (...,4) boxes."""
    o = Origin()
    bx = to_device(gt_boxes, F32, o, "gt_boxes")
    flat = bx.reshape(-1, 4)
    M = flat.shape[0]
    dev = bx.device
    off = torch.tensor([0, M], dtype=torch.int32, device=dev)
    lab = torch.zeros((max(M, 1),), dtype=torch.int32, device=dev)
    flag = torch.ones((1,), dtype=torch.uint8, device=dev)
    out_b = torch.empty((1, max(M, 1), 4), dtype=F32, device=dev)
    out_l = torch.empty((1, max(M, 1)), dtype=torch.int32, device=dev)
    if M:
        _lib.check(_lib.load().tfrpn_pad_gt(ptr(flat), ptr(lab), ptr(off), ptr(flag), 1, M, 0, ptr(out_b), ptr(out_l),
                                            stream_ptr(dev)))
    return from_device(out_b[0, :M].reshape(bx.shape), o)
