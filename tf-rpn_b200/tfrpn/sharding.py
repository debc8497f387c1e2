"""Batch sharding across the GPUs of one box: images are independent units, so rank r of
R takes the contiguous block [r*B/R, (r+1)*B/R) and runs the hot path on it with NO collective.  The
counter RNG is keyed by the GLOBAL image index (``image_offset``), so sharded and unsharded runs
produce bit-identical targets.  ``gather_*`` is the optional final all-gather (NCCL over NVLink when
the process group is NCCL; gloo works for CPU-side tests of the plumbing)."""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous, balanced split: the first (batch % world) ranks get one extra image."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor, rank, world):
    lo, hi = shard_bounds(tensor.shape[0], rank, world)
    return tensor[lo:hi], lo


def sharded_rpn_targets(anchors, gt_boxes, gt_labels, hyper_params, rank, world, seed=0, offset=0):
    """This rank's rows of calculate_rpn_actual_outputs over the global batch (utils/train_utils.py:84-144)."""
    from .utils import train_utils
    gtb, lo = shard(gt_boxes, rank, world)
    gtl, _ = shard(gt_labels, rank, world)
    return train_utils.calculate_rpn_actual_outputs(anchors, gtb, gtl, hyper_params, seed=seed, offset=offset,
                                                    image_offset=lo)


def sharded_proposals(rpn_bbox_deltas, rpn_labels, anchors, hyper_params, rank, world, **kw):
    from .proposals import generate_proposals
    reg, _ = shard(rpn_bbox_deltas, rank, world)
    cls, _ = shard(rpn_labels, rank, world)
    return generate_proposals(reg, cls, anchors, hyper_params, **kw)


def gather_equal(local, out=None, group=None):
    """All-gather of equally sized row blocks into one (world * rows, ...) tensor in rank order: a single
    ``all_gather_into_tensor`` (NCCL: one kernel over NVLink; no padding, no concatenation)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gather_rows(local, batch, group=None):
    """All-gather variable-sized row blocks back into the global batch order (ragged-safe)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
