#!/usr/bin/env python
"""A few generate_proposals calls at C2 (B=64) on rotating inputs: the workload for an ncu capture of the proposal kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch
import tfrpn
from tfrpn import synthetic
from tfrpn.utils import bbox_utils, train_utils
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0"); hp = dict(train_utils.get_hyper_params("vgg16"))
anchors = bbox_utils.generate_anchors(hp)
rng = np.random.default_rng(2000)
sets = []
for _ in range(n):
    reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
    sets.append((torch.from_numpy(reg).to(dev), torch.from_numpy(cls).to(dev)))
for i in range(n):
    out = tfrpn.generate_proposals(sets[i][0], sets[i][1], anchors, hp, pre_nms_topn=6000)
torch.cuda.synchronize()
print("valid", int(out[2].min()), int(out[2].max()))
