#!/usr/bin/env python
"""Time the two HBM-bound drop-in kernels alone (iou_map K1, decode K3) at a named config."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
from tfrpn import _lib, synthetic
from tfrpn.utils import bbox_utils, train_utils
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
bb, B, G, over = synthetic.CONFIGS[cfg]
hp = dict(train_utils.get_hyper_params(bb), **over)
dev = torch.device("cuda:0"); lib = _lib.load()
anchors = bbox_utils.generate_anchors(hp); N = anchors.shape[0]
rng = np.random.default_rng(1)
S = 8
gts = [torch.from_numpy(synthetic.gt_batch(rng, B, G)[0]).to(dev) for _ in range(S)]
regs = [torch.randn((B, N, 4), device=dev) * 0.5 for _ in range(S)]
outs = [torch.empty((B, N, 4), device=dev) for _ in range(S)]
iou_out = [torch.empty((B, N, G), device=dev) for _ in range(3)]
st = torch.cuda.current_stream().cuda_stream
var = (C.c_float * 4)(*hp["variances"])
def timed(fn, reps):
    for r in range(3): fn(r)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(reps): fn(r)
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps
us = timed(lambda r: lib.tfrpn_iou_map(anchors.data_ptr(), 0, gts[r % S].data_ptr(), B, N, G, iou_out[r % 3].data_ptr(), st), 30)
by = 4 * B * N * G + 16 * (N + B * G)
print("%s iou_map  %8.2f us  %7.1f GB/s  (%.1f%% of 6449)" % (cfg, us, by / us / 1e3, 100 * by / us / 1e3 / 6449))
us = timed(lambda r: lib.tfrpn_decode(anchors.data_ptr(), 0, regs[r % S].data_ptr(), var, 1, B, N, outs[r % S].data_ptr(), st), 80)
by = 32 * B * N + 16 * N
print("%s decode   %8.2f us  %7.1f GB/s  (%.1f%% of 6449)  PT=%s" % (cfg, us, by / us / 1e3, 100 * by / us / 1e3 / 6449, os.environ.get("TFRPN_DECODE_PT", "2")))
