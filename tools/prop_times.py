#!/usr/bin/env python
"""proposal kernel time (library tracing hooks) for a few batch shapes; env TFRPN_PROP_CLUSTER selects the variant."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import _lib, synthetic
from tfrpn.utils import bbox_utils, train_utils
dev = torch.device("cuda:0")
lib = _lib.load(); h = _lib.handle(0)
def read():
    parts, total, launches = [], 0.0, 1
    for kid in (3, 5, 6, 7):   # lazy one-CTA kernel, cluster kernel (lazy or rank launch), mask, sweep
        tot, n = C.c_double(), C.c_int()
        _lib.check(lib.tfrpn_profile_read(h, kid, C.byref(tot), C.byref(n)))
        if n.value:
            parts.append("%s %.1f" % (lib.tfrpn_kernel_name(kid).decode().replace("_kernel", ""), 1e3 * tot.value / n.value))
            total += 1e3 * tot.value
            launches = max(launches, n.value)
    return "%.1f us (%s)" % (total / launches, ", ".join(parts))
out = []
for name, hp_over, fm, B in (("C1", {}, (31, 31), 1), ("C2", {}, (31, 31), 64), ("C2b128", {}, (31, 31), 128),
                             ("C4", {"img_size": (800, 1333), "feature_map_shape": (50, 84)}, (50, 84), 32)):
    hp = dict(train_utils.get_hyper_params("vgg16"), **hp_over)
    anchors = bbox_utils.generate_anchors(hp)
    rng = np.random.default_rng(7)
    sets = []
    for _ in range(4):
        reg, cls = synthetic.head_outputs(rng, B, fm[0], fm[1], 9)
        sets.append((torch.from_numpy(reg).to(dev), torch.from_numpy(cls).to(dev)))
    for i in range(4):
        tfrpn.generate_proposals(sets[i][0], sets[i][1], anchors, hp)
    torch.cuda.synchronize()
    _lib.check(lib.tfrpn_profile_enable(h, 1))
    for i in range(24):
        tfrpn.generate_proposals(sets[i % 4][0], sets[i % 4][1], anchors, hp)
    out.append("%s(B=%d) %s" % (name, B, read()))
    _lib.check(lib.tfrpn_profile_enable(h, 0))
for K in (20000, 200000):
    rng = np.random.default_rng(K)
    boxes, scores = synthetic.nms_boxes(rng, 8, K)
    tb, ts = torch.from_numpy(boxes.reshape(8, K, 1, 4)).to(dev), torch.from_numpy(scores.reshape(8, K, 1)).to(dev)
    for mode, kw in (("topk_nms", dict(pre_nms_topn=6000)), ("nms_all", {})):
        for i in range(3):
            bbox_utils.non_max_suppression(tb, ts, max_output_size_per_class=300, max_total_size=300, iou_threshold=0.7, **kw)
        torch.cuda.synchronize()
        _lib.check(lib.tfrpn_profile_enable(h, 1))
        for i in range(10):
            bbox_utils.non_max_suppression(tb, ts, max_output_size_per_class=300, max_total_size=300, iou_threshold=0.7, **kw)
        out.append("C5 K=%d %s %s" % (K, mode, read()))
        _lib.check(lib.tfrpn_profile_enable(h, 0))
print(" | ".join(out))
