#!/usr/bin/env python
"""Quick per-kernel device times at config C2 (library tracing hooks; serial launches)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import _lib, synthetic
from tfrpn.utils import bbox_utils, train_utils
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0"); hp = dict(train_utils.get_hyper_params("vgg16"))
rng = np.random.default_rng(2000); anchors = bbox_utils.generate_anchors(hp)
sets = []
for _ in range(6):
    gtb, gtl = synthetic.gt_batch(rng, B, 50); reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
    sets.append([torch.from_numpy(a).to(dev) for a in (gtb, gtl, reg, cls)])
lib = _lib.load(); h = _lib.handle(0)
def run(n):
    for i in range(n):
        gtb, gtl, reg, cls = sets[i % 6]
        train_utils.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=1, offset=i)
        tfrpn.generate_proposals(reg, cls, anchors, hp)
run(6); torch.cuda.synchronize()
_lib.check(lib.tfrpn_profile_enable(h, 1)); run(60)
for kid in range(_lib.KERNEL_IDS):
    tot, n = C.c_double(), C.c_int()
    _lib.check(lib.tfrpn_profile_read(h, kid, C.byref(tot), C.byref(n)))
    if n.value: print("%-28s %7.2f us  (n=%d)" % (lib.tfrpn_kernel_name(kid).decode(), 1e3 * tot.value / n.value, n.value))
