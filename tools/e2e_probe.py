#!/usr/bin/env python
"""e2e images/s of the pipelined host step at C2 (tfrpn.HostPipeline, acquired slots) for a given depth
(argv[1]) and mode (argv[2]: both | targets | proposals)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import synthetic
from tfrpn.utils import train_utils
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
B, G = 64, 50
hp = dict(train_utils.get_hyper_params("vgg16"))
pipe = tfrpn.HostPipeline(hp, depth=DEPTH, pre_nms_topn=6000)
rng = np.random.default_rng(1)
for i in range(DEPTH):
    v = pipe.acquire(B, G)
    v.gt_boxes[...], v.gt_labels[...] = synthetic.gt_batch(rng, B, G)
    v.rpn_reg[...], v.rpn_cls[...] = synthetic.head_outputs(rng, B, 31, 31, 9)
    pipe.submit(offset=i)
pipe.drain()
def run(n):
    tk = []
    for i in range(n):
        if i >= DEPTH - 1: pipe.wait(tk[i - (DEPTH - 1)])
        pipe.acquire(B, G)
        tk.append(pipe.submit(targets=mode != "proposals", proposals=mode != "targets", offset=i))
    pipe.drain()
run(10)
t0 = time.perf_counter(); run(300); t = time.perf_counter() - t0
print("depth %d mode %s: %.1f us/step  %.0f images/s" % (DEPTH, mode, 1e6 * t / 300, 64 * 300 / t))
