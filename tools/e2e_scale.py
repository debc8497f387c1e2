#!/usr/bin/env python
"""e2e images/s of tfrpn.HostPipeline at C2 under torchrun (one rank per GPU), for several host-side policies, in one
process group: who gathers the candidate rows (host threads / device), how many pool threads, dense input.
   python -m torch.distributed.run --nproc-per-node N tools/e2e_scale.py [depth]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
import numpy as np, torch, torch.distributed as dist
import tfrpn
from tfrpn import synthetic
from tfrpn.utils import train_utils
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B, G = 64, 50
hp = dict(train_utils.get_hyper_params("vgg16"))
rng = np.random.default_rng(1 + rank)
gtb, gtl = synthetic.gt_batch(rng, B, G)
reg, cls = synthetic.head_outputs(rng, B, 31, 31, 9)
modes = [("auto", {}),
         ("host expand, device gather", {"TFRPN_PIPE_EXPAND": "host", "TFRPN_PIPE_GATHER": "device"}),
         ("device expand, device gather", {"TFRPN_PIPE_EXPAND": "device", "TFRPN_PIPE_GATHER": "device"}),
         ("device expand, host gather", {"TFRPN_PIPE_EXPAND": "device", "TFRPN_PIPE_GATHER": "host"}),
         ("device expand, device gather, traced", {"TFRPN_PIPE_EXPAND": "device", "TFRPN_PIPE_GATHER": "device", "TFRPN_PIPE_TRACE": "1"}),
         ("targets only, device expand, traced", {"TFRPN_PIPE_EXPAND": "device", "TFRPN_PIPE_TRACE": "1", "_mode": "targets"}),
         ("proposals only, device gather, traced", {"TFRPN_PIPE_GATHER": "device", "TFRPN_PIPE_TRACE": "1", "_mode": "proposals"})]
if len(sys.argv) > 2:
    modes = [m for m in modes if any(k in m[0] for k in sys.argv[2].split(","))]
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for name, env in modes:
    mode = env.pop("_mode", "both")
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    pipe = tfrpn.HostPipeline(hp, depth=DEPTH, device=dev, pre_nms_topn=6000)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    for i in range(DEPTH):
        v = pipe.acquire(B, G)
        v.gt_boxes[...], v.gt_labels[...], v.rpn_reg[...], v.rpn_cls[...] = gtb, gtl, reg, cls
        pipe.submit(offset=i)
    pipe.drain()
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
    run(30)
    barrier()
    for k in acc: acc[k] = 0.0
    t0 = time.perf_counter(); n = 300; run(n); torch.cuda.synchronize(); t = time.perf_counter() - t0
    mine = t
    if world > 1:
        tt = torch.tensor([t], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t = float(tt.item())
    if rank == 0:
        print("%d GPUs, depth %d, %-34s: %.1f us/step per GPU, %.0f images/s in total, copy bytes %s | rank 0: %.1f us/step, host acquire %.1f submit %.1f wait %.1f"
              % (world, DEPTH, name, 1e6 * t / n, world * B * n / t, pipe.last_copy_bytes(), 1e6 * mine / n,
                 1e6 * acc["acquire"] / n, 1e6 * acc["submit"] / n, 1e6 * acc["wait"] / n), flush=True)
    if "traced" in name and rank == 0:
        import ctypes as C
        from tfrpn import _lib
        lib = _lib.load()
        rows, tk = [], []
        def tr(t):
            ms = (C.c_float * 10)()
            _lib.check(lib.tfrpn_pipeline_trace(pipe._pipe, t, ms))
            rows.append([1e3 * x for x in ms])
        for i in range(40):
            if i >= DEPTH - 1:
                pipe.wait(tk[i - (DEPTH - 1)]); tr(tk[i - (DEPTH - 1)])
            pipe.acquire(B, G)
            tk.append(pipe.submit(targets=mode != "proposals", proposals=mode != "targets", offset=i))
        pipe.drain()
        d = np.array(rows[12:])
        print("    trace (us, medians): H2D %.1f | targets %.1f | proposals %.1f | D2H %.1f | gather %.1f | expand %.1f | step period %.1f | H2D begin -> D2H end %.1f"
              % (tuple(np.median(d[:, 2 * k + 1] - d[:, 2 * k]) for k in range(4)) + (np.median(d[:, 8]), np.median(d[:, 9]),
                 np.median(np.diff(d[:, 7])), np.median(d[:, 7] - d[:, 0]))), flush=True)
    elif "traced" in name:
        tk = []
        for i in range(40):
            if i >= DEPTH - 1: pipe.wait(tk[i - (DEPTH - 1)])
            pipe.acquire(B, G)
            tk.append(pipe.submit(targets=mode != "proposals", proposals=mode != "targets", offset=i))
        pipe.drain()
    pipe.close()
if world > 1:
    dist.destroy_process_group()
