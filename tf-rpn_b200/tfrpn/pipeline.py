"""Pipelined host steps: the generator call site of the reference (utils/train_utils.py:67-82, driven by
Keras from trainer.py:48-49,64-69) and the predictor loop (predictor.py:48-60), over the C-ABI
``tfrpn_pipeline_*`` entry points.

A dense step would move ~11 MB over PCIe each way at C2 while its kernels take ~0.1 ms, so several steps are
kept in flight and the bytes are cut to ~3 MB each way (csrc/pipeline.cu): only the scores are copied in, the
device ranks them, the library's host threads gather the ~640 candidate rows per image of ``rpn_reg`` and only
those follow; ``bbox_deltas`` is sparse by construction (exactly 0 outside the <= 128 sampled positives per
image), so only those rows come back and the library scatters them into the dense array when the step is
waited for.  ``acquire()`` hands out NumPy views of one slot's page-locked buffers -- the data loader writes the
padded batch (and the head outputs) straight into them and reads the results from them; ``submit_arrays()``
takes the caller's own (pageable) arrays instead.
"""
import collections
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._tensor import Origin, default_device, to_device
from .proposals import proposal_cfg
from .utils import bbox_utils, train_utils

StepViews = collections.namedtuple(
    "StepViews", ["gt_boxes", "gt_labels", "rpn_reg", "rpn_cls", "deltas", "labels", "out_boxes", "out_scores",
                  "valid", "keep_idx"])


def _view(addr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * n).from_address(addr), dtype=dtype).reshape(shape)


class HostPipeline:
    """``depth`` host steps in flight on one GPU.

        pipe = HostPipeline(hyper_params, depth=3)
        v = pipe.acquire(B, G)              # NumPy views of the next slot's pinned buffers
        v.gt_boxes[...] = ...; v.gt_labels[...] = ...; v.rpn_reg[...] = ...; v.rpn_cls[...] = ...
        t = pipe.submit(offset=step)        # enqueue only; returns a ticket
        ...                                 # acquire/fill/submit the next steps
        pipe.wait(t)                        # v.deltas, v.labels, v.out_boxes ... now hold step t's results

    The views of a slot stay valid until that slot is acquired again, ``depth`` steps later.
    """

    def __init__(self, hyper_params, depth=3, device=None, anchors=None, pre_nms_topn=None):
        dev = default_device() if device is None else torch.device(device)
        self.device = dev
        self.hp = hyper_params
        self.depth = int(depth)
        self._pipe = self._h = None
        with torch.cuda.device(dev):
            if anchors is None:
                anchors = bbox_utils.generate_anchors(hyper_params)
            else:   # host arrays / CPU tensors are copied to this GPU; a tensor on another GPU is rejected
                o = Origin()
                o.note(None, dev)
                anchors = to_device(anchors, torch.float32, o, "anchors")
            torch.cuda.synchronize()
        if anchors.dim() != 2 or anchors.shape[1] != 4:
            raise ValueError("anchors must be (total_anchors, 4), got %s" % (tuple(anchors.shape),))
        self.anchors = anchors          # contiguous float32 CUDA tensor on self.device, kept alive here
        self.N = int(self.anchors.shape[0])
        self.pcfg = proposal_cfg(hyper_params, pre_nms_topn=pre_nms_topn)
        self.P = int(self.pcfg.post_nms_topn)
        self._lib = _lib.load()
        # The pipeline's steps run on its own streams while the caller keeps using the drop-in functions
        # (losses, calculate_rpn_actual_outputs, ...) on torch's stream with the thread's shared handle: the
        # pipeline therefore owns a handle (= workspace) of its own.
        h = C.c_void_p()
        _lib.check(self._lib.tfrpn_create(C.byref(h), int(dev.index)))
        self._h = h
        pipe = C.c_void_p()
        _lib.check(self._lib.tfrpn_pipeline_create(self._h, self.depth, C.byref(pipe)))
        self._pipe = pipe
        self._views = {}
        self._keep = {}
        # per-step ctypes objects are reused: building them costs more than the C call they are passed to
        self._sb = _lib.StepBuffers()
        self._sb_ref = C.byref(self._sb)
        self._tcfg = train_utils._target_cfg(hyper_params, 0, 0, 0)
        self._tcfg_ref = C.byref(self._tcfg)
        self._pcfg_ref = C.byref(self.pcfg)
        self._ticket = C.c_int64()
        self._ticket_ref = C.byref(self._ticket)
        self._anchors_ptr = self.anchors.data_ptr()

    def acquire(self, batch, max_gt):
        B, G, N, P = int(batch), int(max_gt), self.N, self.P
        sb = self._sb
        rc = self._lib.tfrpn_pipeline_acquire(self._pipe, B, N, G, P, self._sb_ref)
        if rc:
            _lib.check(rc)
        key = (sb.gt_boxes, B, G)
        v = self._views.get(key)
        if v is None:
            fm_h, fm_w = bbox_utils._pair(self.hp["feature_map_shape"])
            A = int(self.hp["anchor_count"])
            f32, i32 = np.float32, np.int32
            v = StepViews(_view(sb.gt_boxes, (B, G, 4), f32), _view(sb.gt_labels, (B, G), i32),
                          _view(sb.rpn_reg, (B, fm_h, fm_w, 4 * A), f32), _view(sb.rpn_cls, (B, fm_h, fm_w, A), f32),
                          _view(sb.deltas, (B, N, 4), f32), _view(sb.labels, (B, fm_h, fm_w, A), f32),
                          _view(sb.out_boxes, (B, P, 4), f32), _view(sb.out_scores, (B, P), f32),
                          _view(sb.valid, (B,), i32), _view(sb.keep_idx, (B, P), i32))
            if len(self._views) > 4 * self.depth:
                self._views.clear()
            self._views[key] = v
        return v

    def submit(self, targets=True, proposals=True, seed=None, offset=None, image_offset=0):
        if targets:   # only the RNG fields change from step to step (utils/train_utils.py:50-65 has no state to carry)
            tcfg = self._tcfg
            tcfg.seed = int(self.hp.get("seed", 0) if seed is None else seed)
            tcfg.offset = next(train_utils._auto_offset) if offset is None else int(offset)
            tcfg.image_offset = int(image_offset)
        rc = self._lib.tfrpn_pipeline_submit_acquired(self._pipe, self._anchors_ptr, self._tcfg_ref if targets else None,
                                                      self._pcfg_ref if proposals else None, self._ticket_ref)
        if rc:
            _lib.check(rc)
        return self._ticket.value

    def set_stable_outputs(self, on=True):
        """Promise that the ``out`` arrays given to :meth:`submit_arrays` are written by this pipeline only (the usual
        ring of ``depth`` output dicts).  The library then resets just the rows it wrote into an array the last time
        instead of zeroing the whole dense ``bbox_deltas`` every step (TFRPN_PIPE_OPT_STABLE_OUTPUTS)."""
        _lib.check(self._lib.tfrpn_pipeline_set_option(self._pipe, 1, int(bool(on))))

    def submit_arrays(self, gt_boxes=None, gt_labels=None, rpn_reg=None, rpn_cls=None, out=None, seed=None, offset=None,
                      image_offset=0):
        """One step on the caller's own host arrays (``tfrpn_pipeline_submit``): NumPy arrays as the reference's
        generator / predictor hold them (utils/train_utils.py:78-82, predictor.py:50), pageable or page-locked.
        Nothing is copied by the caller: the library stages the small inputs, gathers the candidate rows of
        ``rpn_reg`` in place and writes the results into ``out`` (allocated here unless given).  The arrays must
        stay alive and unchanged until ``wait(ticket)``.  Returns ``(ticket, out)`` with ``out`` a dict of
        bbox_deltas (B,N,4), bbox_labels (B,N), boxes (B,P,4), scores (B,P), valid (B,), keep_idx (B,P)."""
        do_t, do_p = gt_boxes is not None, rpn_reg is not None
        f32, i32 = np.float32, np.int32
        if do_t:
            gt_boxes = np.ascontiguousarray(gt_boxes, f32)
            gt_labels = np.ascontiguousarray(gt_labels, i32)
            B, G = gt_labels.shape
        if do_p:
            rpn_reg = np.ascontiguousarray(rpn_reg, f32)
            rpn_cls = np.ascontiguousarray(rpn_cls, f32)
            B = rpn_reg.shape[0]
        N, P = self.N, self.P
        if out is None:
            out = {}
        want = ([("bbox_deltas", (B, N, 4), f32), ("bbox_labels", (B, N), f32)] if do_t else []) + \
               ([("boxes", (B, P, 4), f32), ("scores", (B, P), f32), ("valid", (B,), i32), ("keep_idx", (B, P), i32)] if do_p else [])
        for name, shape, dt in want:
            a = out.get(name)
            if a is None:
                out[name] = np.empty(shape, dt)
            elif a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("out[%r] must be a C-contiguous %s array of shape %s" % (name, np.dtype(dt).name, shape))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        tcfg = train_utils._target_cfg(self.hp, seed, offset, image_offset) if do_t else _lib.TargetCfg()
        t = C.c_int64()
        _lib.check(self._lib.tfrpn_pipeline_submit(
            self._pipe, self.anchors.data_ptr(), B, N,
            vp(gt_boxes) if do_t else None, vp(gt_labels) if do_t else None, G if do_t else 0, C.byref(tcfg),
            vp(out["bbox_deltas"]) if do_t else None, vp(out["bbox_labels"]) if do_t else None,
            vp(rpn_reg) if do_p else None, vp(rpn_cls) if do_p else None, C.byref(self.pcfg),
            vp(out["boxes"]) if do_p else None, vp(out["scores"]) if do_p else None, vp(out["valid"]) if do_p else None,
            vp(out["keep_idx"]) if do_p else None, C.byref(t)))
        self._keep[t.value % max(self.depth, 1)] = (gt_boxes, gt_labels, rpn_reg, rpn_cls, out)   # alive until the slot is reused
        return t.value, out

    def wait(self, ticket):
        rc = self._lib.tfrpn_pipeline_wait(self._pipe, ticket)
        if rc:
            _lib.check(rc)

    def last_copy_bytes(self):
        """(H2D, D2H) bytes of the last submitted step.  The target tensors cross PCIe in compact form (labels as
        they are, bbox_deltas as its <= total_pos non-zero rows per image) and wait() expands the dense deltas view."""
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self._lib.tfrpn_pipeline_last_copy_bytes(self._pipe, C.byref(a), C.byref(b)))
        return a.value, b.value

    def drain(self):
        _lib.check(self._lib.tfrpn_pipeline_drain(self._pipe))

    def close(self):
        if self._pipe:
            self._lib.tfrpn_pipeline_destroy(self._pipe)
            self._pipe = None
            self._views.clear()
        if self._h:
            self._lib.tfrpn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def prefetching_rpn_generator(dataset, anchors, hyper_params, prefetch=2, seed=None):
    """utils/train_utils.py:67-82 with the target assignment of the next ``prefetch`` batches already in
    flight: yields ``(img, (bbox_deltas, bbox_labels))`` forever, like the reference.
    ``dataset`` yields ``(img, gt_boxes (B,G,4), gt_labels (B,G))`` host arrays.  The yielded target arrays
    are NumPy views of page-locked memory, valid until the next ``next()`` call (copy them to keep them)."""
    pipe = HostPipeline(hyper_params, depth=int(prefetch) + 2, anchors=anchors)
    pending = collections.deque()
    step = 0
    try:
        while True:
            for img, gt_boxes, gt_labels in dataset:
                gt_boxes = np.asarray(gt_boxes, np.float32)
                gt_labels = np.asarray(gt_labels, np.int32)
                B, G = gt_labels.shape
                v = pipe.acquire(B, G)
                v.gt_boxes[...] = gt_boxes
                v.gt_labels[...] = gt_labels
                pending.append((pipe.submit(targets=True, proposals=False, seed=seed, offset=step), img, v))
                step += 1
                if len(pending) > prefetch:
                    t, im, pv = pending.popleft()
                    pipe.wait(t)
                    yield im, (pv.deltas, pv.labels)
            if step == 0:
                return
    finally:
        pipe.close()
