#!/usr/bin/env python
"""Device timeline of consecutive HostPipeline steps at C2 (TFRPN_PIPE_TRACE=1): when each step's H2D, target
kernels, proposal kernels and D2H begin and end, us relative to the first traced step.  argv: depth mode."""
import ctypes as C, os, sys
os.environ["TFRPN_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import synthetic, _lib
from tfrpn.utils import train_utils
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
B, G = 64, 50
hp = dict(train_utils.get_hyper_params("vgg16"))
rng = np.random.default_rng(1)
gtb, gtl = synthetic.gt_batch(rng, B, G)
reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
pipe = tfrpn.HostPipeline(hp, depth=DEPTH, pre_nms_topn=6000)
lib = _lib.load()
for i in range(DEPTH):
    v = pipe.acquire(B, G)
    v.gt_boxes[...], v.gt_labels[...], v.rpn_reg[...], v.rpn_cls[...] = gtb, gtl, reg, cls
    pipe.submit(offset=i)
pipe.drain()
rows = []
def trace(t):
    ms = (C.c_float * 10)()
    _lib.check(lib.tfrpn_pipeline_trace(pipe._pipe, t, ms))
    rows.append((t, [1e3 * x for x in ms]))   # us
tk = []
n = 40
for i in range(n):
    if i >= DEPTH - 1:
        pipe.wait(tk[i - (DEPTH - 1)]); trace(tk[i - (DEPTH - 1)])
    pipe.acquire(B, G)
    tk.append(pipe.submit(targets=mode != "proposals", proposals=mode != "targets", offset=i))
for t in tk[n - (DEPTH - 1):]:
    pipe.wait(t); trace(t)
rows = rows[16:32]
t0 = rows[0][1][0]
print("depth %d mode %s; us since step %d's H2D began" % (DEPTH, mode, rows[0][0]))
print("%6s | %15s | %15s | %15s | %15s" % ("step", "H2D", "targets", "proposals", "D2H"))
for t, m in rows:
    m = [x - t0 for x in m[:8]] + m[8:]
    print("%6d | %7.1f %7.1f | %7.1f %7.1f | %7.1f %7.1f | %7.1f %7.1f | gather %6.1f expand %6.1f" % (t, *m))
d = np.array([m for _, m in rows])
print("durations us (median): H2D %.1f targets %.1f proposals %.1f D2H %.1f | step period %.1f"
      % (tuple(np.median(d[:, 2 * k + 1] - d[:, 2 * k]) for k in range(4)) + (np.median(np.diff(d[:, 7])),)))
