#!/usr/bin/env python
"""Cycles per phase of proposal kernels (library built with EXTRA=-DTFRPN_PHASE_TIMING), config C2, image 0."""
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
reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
reg, cls = torch.from_numpy(reg).to(dev), torch.from_numpy(cls).to(dev)
lib = _lib.load()
lib.tfrpn_debug_phase_cycles.argtypes = [C.c_void_p]
names = {1: "phase0", 2: "select", 3: "compact", 4: "sort", 5: "rank", 6: "boxes/stage", 7: "kept-test", 8: "survivors", 9: "triangle",
         10: "resolve", 11: "commit", 12: "tail"}
acc = np.zeros(32)
for it in range(6):
    out = tfrpn.generate_proposals(reg, cls, anchors, hp)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    assert lib.tfrpn_debug_phase_cycles(buf) == 0
    if it >= 2:
        acc += np.array(list(buf), dtype=np.float64)
acc /= 4
tot = acc.sum()
print("valid[0] =", int(out[2][0]), " total cycles %.0f = %.2f us @1.965GHz" % (tot, tot / 1965))
for k, n in names.items():
    print("  %-12s %8.0f cyc  %5.1f%%" % (n, acc[k], 100 * acc[k] / tot))

lib.tfrpn_debug_cta_times.argtypes = [C.c_void_p]
buf = (C.c_longlong * 4096)()
assert lib.tfrpn_debug_cta_times(buf) == 0
a = np.array(list(buf), dtype=np.int64).reshape(1024, 4)
a = a[a[:, 1] > 0]
t0 = a[:, 0].min()
st, en = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3
print("CTAs %d: start us min/med/max %.1f %.1f %.1f | end us min/med/max %.1f %.1f %.1f | duration us min/med/max %.1f %.1f %.1f"
      % (len(a), st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max(), (en - st).min(), np.median(en - st), (en - st).max()))
