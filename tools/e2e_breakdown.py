#!/usr/bin/env python
"""Where the e2e step of tfrpn.HostPipeline goes at C2: host time inside acquire / submit / wait (perf_counter
around each call), per mode (both | targets | proposals) and depth.  argv: depths (comma list), default 2,4,6."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import synthetic
from tfrpn.utils import train_utils
depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2,4,6").split(",")]
B, G = 64, 50
hp = dict(train_utils.get_hyper_params("vgg16"))
rng = np.random.default_rng(1)
gtb, gtl = synthetic.gt_batch(rng, B, G)
reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
for DEPTH in depths:
    pipe = tfrpn.HostPipeline(hp, depth=DEPTH, pre_nms_topn=6000)
    for i in range(DEPTH):
        v = pipe.acquire(B, G)
        v.gt_boxes[...], v.gt_labels[...], v.rpn_reg[...], v.rpn_cls[...] = gtb, gtl, reg, cls
        pipe.submit(offset=i)
    pipe.drain()
    for mode in ("both", "targets", "proposals"):
        acc = {"acquire": 0.0, "submit": 0.0, "wait": 0.0}
        def run(n):
            tk = []
            pc = time.perf_counter
            for i in range(n):
                if i >= DEPTH - 1:
                    t0 = pc(); pipe.wait(tk[i - (DEPTH - 1)]); acc["wait"] += pc() - t0
                t0 = pc(); pipe.acquire(B, G); acc["acquire"] += pc() - t0
                t0 = pc(); tk.append(pipe.submit(targets=mode != "proposals", proposals=mode != "targets", offset=i)); acc["submit"] += pc() - t0
            pipe.drain()
        run(20)
        for k in acc: acc[k] = 0.0
        n = 300
        torch.cuda.synchronize()
        t0 = time.perf_counter(); run(n); torch.cuda.synchronize(); t = time.perf_counter() - t0
        print("depth %d %-9s: %.1f us/step %.0f images/s | host us/step: acquire %.1f submit %.1f wait %.1f | copy bytes %s"
              % (DEPTH, mode, 1e6 * t / n, B * n / t, 1e6 * acc["acquire"] / n, 1e6 * acc["submit"] / n, 1e6 * acc["wait"] / n,
                 pipe.last_copy_bytes()), flush=True)
    pipe.close()
