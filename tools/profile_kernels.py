#!/usr/bin/env python
"""Launch each hot-path kernel a few times on config C2 inputs -- the short command that ncu wraps:

  ncu --set full --clock-control none --import-source on -k regex:'kernel' -s 5 -c 15 \
      -o gpurun_out/prof python tools/profile_kernels.py
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_kernels.py --steps 20
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import tfrpn  # noqa: E402
from tfrpn import synthetic  # noqa: E402
from tfrpn.utils import bbox_utils, train_utils  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--extras", action="store_true", help="also launch iou_map / decode / encode / topk / nms")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    hp = dict(train_utils.get_hyper_params("vgg16"))
    B = args.batch
    rng = np.random.default_rng(2000)
    anchors = bbox_utils.generate_anchors(hp)
    sets = []
    for _ in range(3):
        gtb, gtl = synthetic.gt_batch(rng, B, 50)
        reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
        sets.append([torch.from_numpy(a).to(dev) for a in (gtb, gtl, reg, cls)])
    big = None
    if args.extras:   # C5-style input for the large-N prefilter: 8 images x 200k boxes
        bx, sc = synthetic.nms_boxes(rng, 8, 200000)
        big = (torch.from_numpy(bx).to(dev).reshape(8, -1, 1, 4), torch.from_numpy(sc).to(dev).reshape(8, -1, 1))
    for i in range(args.steps):
        gtb, gtl, reg, cls = sets[i % 3]
        deltas, labels = train_utils.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=1, offset=i)
        tfrpn.generate_proposals(reg, cls, anchors, hp)
        if args.extras:
            bbox_utils.generate_iou_map(anchors, gtb)
            var = torch.tensor(hp["variances"], device=dev)
            boxes = bbox_utils.get_bboxes_from_deltas(anchors, reg.reshape(B, -1, 4) * var)
            bbox_utils.get_deltas_from_bboxes(anchors, boxes)
            bbox_utils.top_k_boxes(cls.reshape(B, -1), 6000, boxes)
            train_utils.rpn_losses(deltas, reg, labels, cls, with_grads=True)
            tfrpn.predict_top_boxes(reg, cls, anchors, hp, k=10)
            bbox_utils.non_max_suppression(big[0], big[1], max_output_size_per_class=300, max_total_size=300,
                                           iou_threshold=0.7, pre_nms_topn=6000)
    torch.cuda.synchronize()
    print("done", tfrpn._lib.launch_count(), "launches")


if __name__ == "__main__":
    main()
