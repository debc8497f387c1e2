"""Multi-rank host logic on CPU (gloo, world_size 2): shard bounds, global image offsets and the
optional final gather.  The compute itself is checked with the oracle standing in for the kernels --
what is under test is that shard r + image_offset reproduces rows [lo, hi) of the unsharded run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_batch_exactly():
    sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
    from tfrpn.sharding import shard_bounds
    for batch in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tf-rpn_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rpn_oracle as O
    from tfrpn import synthetic
    from tfrpn.sharding import gather_equal, gather_rows, shard, shard_bounds
    hp = O.get_hyper_params("vgg16")
    anchors = O.generate_anchors(hp)
    B = 5                                        # ragged split: 3 + 2
    gtb, gtl = synthetic.gt_batch(np.random.default_rng(4), B, 9)
    lo, hi = shard_bounds(B, rank, world)
    my_b, off = shard(torch.from_numpy(gtb), rank, world)
    my_l, _ = shard(torch.from_numpy(gtl), rank, world)
    assert off == lo and my_b.shape[0] == hi - lo
    d, l = O.calculate_rpn_actual_outputs(anchors, my_b.numpy(), my_l.numpy(), hp, seed=3, offset=2, image_offset=off)
    full_l = gather_rows(torch.from_numpy(l.reshape(hi - lo, -1)), B)
    full_d = gather_rows(torch.from_numpy(d), B)
    eq = gather_equal(torch.full((2, 3), float(rank)))          # equal blocks: one all_gather_into_tensor
    assert eq.shape == (2 * world, 3) and all(bool((eq[2 * r:2 * r + 2] == r).all()) for r in range(world))
    if rank == 0:
        rd, rl = O.calculate_rpn_actual_outputs(anchors, gtb, gtl, hp, seed=3, offset=2)
        q.put(bool(np.array_equal(full_l.numpy(), rl.reshape(B, -1)) and np.array_equal(full_d.numpy(), rd)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_targets_equal_unsharded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
