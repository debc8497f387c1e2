#!/usr/bin/env python
"""Config sweep on one B200: every configuration BASELINE.json names (C1-C4 target assignment +
proposals, C5 NMS / top-k stress at K = 10k..500k), device-resident, next to the CPU restatement on
the host cores.  Each GPU result of the run is also compared with the oracle on a sample.

    python tests/sweep_configs.py [--out profiles/rNN_sweep.json] [--quick]

Timing: CUDA events around `reps` back-to-back steps over rotating input sets (so that inputs do not
sit in L2 between steps); the CPU column is oracle/rpn_oracle.c with OpenMP over images on all host
threads, timed on a bounded sample (kind "port": the reference is TF 2.0 Python, not installable).
The script lives under tests/ (pytest does not collect it) because it links oracle/ -- as the in-run
checker and as the timed CPU baseline, never on the GPU path; nothing outside tests/, bench.py's CPU legs
and __graft_entry__.smoke() touches the oracle.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import tfrpn  # noqa: E402
from oracle import c_oracle, rpn_oracle as O  # noqa: E402
from tfrpn import _lib, synthetic  # noqa: E402
from tfrpn.proposals import proposal_cfg  # noqa: E402
from tfrpn.utils import bbox_utils, train_utils  # noqa: E402

dev = torch.device("cuda:0")
F32 = np.float32


def timed(fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps          # ms per step


def kernel_us(h, lib, fn, n):
    _lib.check(lib.tfrpn_profile_enable(h, 1))
    for i in range(n):
        fn(i)
    out = {}
    for kid in range(5):
        tot, cnt = C.c_double(), C.c_int()
        _lib.check(lib.tfrpn_profile_read(h, kid, C.byref(tot), C.byref(cnt)))
        if cnt.value:
            out[lib.tfrpn_kernel_name(kid).decode()] = round(1e3 * tot.value / cnt.value, 2)
    _lib.check(lib.tfrpn_profile_enable(h, 0))
    return out


def run_config(name, quick):
    bb, B, G, over = synthetic.CONFIGS[name]
    hp = dict(train_utils.get_hyper_params(bb), **over)
    lib, h = _lib.load(), _lib.handle(0)
    anchors = bbox_utils.generate_anchors(hp)
    a_np = O.generate_anchors(hp)
    N = anchors.shape[0]
    fm_h, fm_w = bbox_utils._pair(hp["feature_map_shape"])
    rng = np.random.default_rng(1000 * int(name[1]))
    set_bytes = B * (G * 20 + N * 40)
    S = max(2, min(12, int(300e6 // set_bytes)))
    sets, np_sets = [], []
    for s in range(S):
        if s < 2:
            gtb, gtl = synthetic.gt_batch(rng, B, G)
            reg, cls = synthetic.head_outputs(rng, B, fm_h, fm_w, 9)
            np_sets.append((gtb, gtl, reg, cls))
        else:
            gtb, gtl, reg, cls = (np.ascontiguousarray(np.roll(a, s // 2, axis=0)) for a in np_sets[s % 2])
        sets.append(dict(gtb=torch.from_numpy(gtb).to(dev), gtl=torch.from_numpy(gtl).to(dev),
                         reg=torch.from_numpy(reg).to(dev), cls=torch.from_numpy(cls).to(dev),
                         deltas=torch.empty((B, N, 4), device=dev), labels=torch.empty((B, N), device=dev),
                         pb=torch.empty((B, 300, 4), device=dev), ps=torch.empty((B, 300), device=dev),
                         pv=torch.empty((B,), dtype=torch.int32, device=dev),
                         pk=torch.empty((B, 300), dtype=torch.int32, device=dev)))
    _lib.check(lib.tfrpn_reserve(h, B, N, G, 6000))
    pcfg = proposal_cfg(hp, pre_nms_topn=6000)
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)

    def targets(i, st):
        s = sets[i % S]
        cfg = train_utils._target_cfg(hp, 7, i, 0)
        _lib.check(lib.tfrpn_rpn_targets(h, anchors.data_ptr(), s["gtb"].data_ptr(), s["gtl"].data_ptr(), B, N, G,
                                         C.byref(cfg), s["deltas"].data_ptr(), s["labels"].data_ptr(), None, st.cuda_stream))

    def proposals(i, st):
        s = sets[i % S]
        _lib.check(lib.tfrpn_proposals(h, s["reg"].data_ptr(), s["cls"].data_ptr(), anchors.data_ptr(), B, N, C.byref(pcfg),
                                       s["pb"].data_ptr(), s["ps"].data_ptr(), s["pv"].data_ptr(), s["pk"].data_ptr(),
                                       st.cuda_stream))

    def step(i):
        side.wait_stream(cur)
        targets(i, cur)
        proposals(i, side)
        cur.wait_stream(side)

    reps = 50 if quick else 300
    ms = timed(step, reps)
    ms_t = timed(lambda i: targets(i, cur), reps)
    ms_p = timed(lambda i: proposals(i, cur), reps)
    kern = kernel_us(h, lib, lambda i: (targets(i, cur), proposals(i, cur)), 2 * S)
    # parity of this very run, on the first images of set 0 (labels / keep lists bit-exact)
    nb = min(B, 4)
    step(0)
    torch.cuda.synchronize()
    gtb, gtl, reg, cls = (a[:nb] for a in np_sets[0])
    od, ol = c_oracle.rpn_targets(a_np, gtb, gtl, hp, seed=7, offset=0)
    ob, os_, ov, ok = c_oracle.proposals(reg.reshape(nb, -1, 4), cls.reshape(nb, -1), a_np, hp, 6000)
    s0 = sets[0]
    parity = bool(np.array_equal(s0["labels"][:nb].cpu().numpy().reshape(ol.shape), ol)
                  and np.array_equal(s0["pk"][:nb].cpu().numpy(), ok) and np.array_equal(s0["pv"][:nb].cpu().numpy(), ov)
                  and np.allclose(s0["deltas"][:nb].cpu().numpy(), od, rtol=1e-6, atol=1e-6))
    # CPU restatement, bounded sample
    threads = len(os.sched_getaffinity(0))
    nc = min(B, 64 if name != "C4" else 16)
    gtb, gtl, reg, cls = (a[:nc] for a in np_sets[0])
    t0 = time.perf_counter()
    n_cpu = 0
    while time.perf_counter() - t0 < (1.0 if quick else 4.0):
        c_oracle.rpn_targets(a_np, gtb, gtl, hp, seed=7, offset=0, threads=threads)
        c_oracle.proposals(reg.reshape(nc, -1, 4), cls.reshape(nc, -1), a_np, hp, 6000, threads=threads)
        n_cpu += nc
    cpu_ips = n_cpu / (time.perf_counter() - t0)
    del sets
    torch.cuda.empty_cache()
    return {"config": name, "backbone": bb, "B": B, "N": int(N), "G": G, "input_sets": S,
            "step_ms": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1), "targets_ms": round(ms_t, 4),
            "proposals_ms": round(ms_p, 4), "kernel_us": kern, "parity_vs_oracle": parity,
            "cpu_images_per_s": round(cpu_ips, 1), "cpu_threads": threads, "speedup_vs_cpu": round(B / ms * 1e3 / cpu_ips, 1)}


def run_c5(K, quick):
    """C5: B = 8 images of K boxes.  (a) top-6000-of-K -> NMS 300 @ 0.7; (b) NMS over all K, 300 @ 0.7."""
    B = 8
    lib, h = _lib.load(), _lib.handle(0)
    S = max(2, min(8, int(200e6 // (B * K * 20))))
    rng = np.random.default_rng(5000 + K // 1000)
    np_sets = [synthetic.nms_boxes(rng, B, K) for _ in range(2)]
    sets = [(torch.from_numpy(np_sets[s % 2][0]).to(dev), torch.from_numpy(np_sets[s % 2][1]).to(dev)) for s in range(S)]
    k = min(6000, K)
    cur = torch.cuda.current_stream(dev).cuda_stream
    tv = torch.empty((B, k), device=dev)
    ti = torch.empty((B, k), dtype=torch.int32, device=dev)
    tg = torch.empty((B, k, 4), device=dev)
    ob = torch.empty((B, 300, 4), device=dev)
    os_ = torch.empty((B, 300), device=dev)
    oc = torch.empty((B, 300), device=dev)
    ov = torch.empty((B,), dtype=torch.int32, device=dev)
    ok = torch.empty((B, 300), dtype=torch.int32, device=dev)
    cfg = _lib.NmsCfg(300, 300, 0.7, float("-inf"), 0, 1)

    cfg_a = _lib.NmsCfg(300, 300, 0.7, float("-inf"), 0, 1, k)     # pre_nms_topn = k: top-k fused in front of the NMS

    def a2(i):   # the two separate calls of predictor.py:58-60 + bbox_utils.py:48-70 (full sort of the top k)
        bx, sc = sets[i % S]
        _lib.check(lib.tfrpn_topk(h, sc.data_ptr(), B, K, k, tv.data_ptr(), ti.data_ptr(), bx.data_ptr(), 1, tg.data_ptr(), cur))
        _lib.check(lib.tfrpn_nms(h, tg.data_ptr(), tv.data_ptr(), B, k, C.byref(cfg), ob.data_ptr(), os_.data_ptr(),
                                 oc.data_ptr(), ov.data_ptr(), ok.data_ptr(), cur))

    def a(i):    # one call, candidates consumed lazily
        bx, sc = sets[i % S]
        _lib.check(lib.tfrpn_nms(h, bx.data_ptr(), sc.data_ptr(), B, K, C.byref(cfg_a), ob.data_ptr(), os_.data_ptr(),
                                 oc.data_ptr(), ov.data_ptr(), ok.data_ptr(), cur))

    def b(i):
        bx, sc = sets[i % S]
        _lib.check(lib.tfrpn_nms(h, bx.data_ptr(), sc.data_ptr(), B, K, C.byref(cfg), ob.data_ptr(), os_.data_ptr(),
                                 oc.data_ptr(), ov.data_ptr(), ok.data_ptr(), cur))

    reps = 20 if quick else 100
    ms_a = timed(a, reps)
    ms_a2 = timed(a2, reps)
    ms_b = timed(b, reps)
    # parity: (b) keep list of image 0 of set 0 against the C oracle; (a) keeps the same boxes
    b(0)
    torch.cuda.synchronize()
    keep_b = ok[0].cpu().numpy().copy()
    t0 = time.perf_counter()
    _, _, _, cv, ck = c_oracle.nms(np_sets[0][0][:1], np_sets[0][1][:1], 300, 300, 0.7)
    cpu_b = time.perf_counter() - t0
    parity = bool(np.array_equal(keep_b, ck[0]) and int(ov[0]) == int(cv[0]))
    a(0)
    torch.cuda.synchronize()
    keep_a = ok[0].cpu().numpy().copy()
    parity = parity and bool(np.array_equal(keep_a, keep_b))
    a2(0)
    torch.cuda.synchronize()
    keep_a2 = ti[0].cpu().numpy()[ok[0].cpu().numpy()]
    parity = parity and bool(np.array_equal(keep_a2, keep_b))
    t0 = time.perf_counter()
    v, i = c_oracle.top_k(np_sets[0][1][:1], k)
    c_oracle.nms(np.take_along_axis(np_sets[0][0][:1], i[..., None].astype(np.int64), axis=1), v, 300, 300, 0.7)
    cpu_a = time.perf_counter() - t0
    return {"config": "C5", "K": K, "B": B, "topk6000_nms_ms": round(ms_a, 4), "topk6000_then_nms_two_calls_ms": round(ms_a2, 4), "nms_all_ms": round(ms_b, 4),
            "topk6000_nms_images_per_s": round(B / ms_a * 1e3, 1), "nms_all_images_per_s": round(B / ms_b * 1e3, 1),
            "cpu_ms_per_image_topk_nms": round(cpu_a * 1e3, 3), "cpu_ms_per_image_nms_all": round(cpu_b * 1e3, 3),
            "speedup_vs_cpu_1thread_topk_nms": round(cpu_a * 1e3 / (ms_a / B), 1),
            "speedup_vs_cpu_1thread_nms_all": round(cpu_b * 1e3 / (ms_b / B), 1), "parity_vs_oracle": parity}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    assert c_oracle.available(), "build oracle/ first (python -c 'import __graft_entry__ as g; g.build()')"
    rows = [run_config(c, args.quick) for c in ("C1", "C2", "C3", "C4")]
    rows += [run_c5(K, args.quick) for K in (10000, 20000, 50000, 100000, 200000, 500000)]
    res = {"gpu": torch.cuda.get_device_name(0), "host_threads": len(os.sched_getaffinity(0)),
           "note": "device-resident inputs, CUDA-event timed; CPU = oracle/rpn_oracle.c (port, not TensorFlow)", "rows": rows}
    for r in rows:
        print(json.dumps(r))
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
