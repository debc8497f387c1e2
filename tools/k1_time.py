#!/usr/bin/env python
"""K1 (materialised IoU map) time at C2 / C3 / C4 shapes: rotating outputs > L2, one CUDA graph of back-to-back launches."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
from tfrpn import _lib, synthetic
from tfrpn.utils import bbox_utils, train_utils
dev = torch.device("cuda:0"); lib = _lib.load()
peak = 6549.4
for name in ("C2", "C3", "C4"):
    bb, B, G, over = synthetic.CONFIGS[name]
    hp = dict(train_utils.get_hyper_params(bb), **over)
    anchors = bbox_utils.generate_anchors(hp); N = anchors.shape[0]
    gtb, _ = synthetic.gt_batch(np.random.default_rng(1), B, G)
    gt = torch.from_numpy(gtb).to(dev)
    n_out = max(3, int(np.ceil(2.1 * 126e6 / (4.0 * B * N * G))))
    outs = [torch.empty((B, N, G), device=dev) for _ in range(n_out)]
    cs = torch.cuda.Stream(dev)
    with torch.cuda.stream(cs):
        fn = lambda r: _lib.check(lib.tfrpn_iou_map(anchors.data_ptr(), 0, gt.data_ptr(), B, N, G, outs[r % n_out].data_ptr(), cs.cuda_stream))
        for r in range(3): fn(r)
        cs.synchronize()
        g = torch.cuda.CUDAGraph()
        reps = 10 * n_out
        with torch.cuda.graph(g, stream=cs):
            for r in range(reps): fn(r)
        g.replay(); cs.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cs)
        for _ in range(3): g.replay()
        b.record(cs); cs.synchronize()
    us = 1e3 * a.elapsed_time(b) / (3 * reps)
    by = 4 * B * N * G + 16 * (N + B * G)
    print("%s B=%d N=%d G=%d: %.2f us, %.0f GB/s, %.3f of %.0f GB/s" % (name, B, N, G, us, by / us / 1e3, by / us / 1e3 / peak, peak))
    del outs
