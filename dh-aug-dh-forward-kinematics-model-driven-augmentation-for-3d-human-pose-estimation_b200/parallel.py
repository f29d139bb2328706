"""Data-parallel plumbing for the augmentation path.

Poses (rows) are independent, so rank r of R simply takes rows [r*N/R, (r+1)*N/R) -- there is no
collective on the FK / projection data path.  The only exchange in a GAN step is the gradient
all-reduce of the small generator / critic MLPs; ``allreduce_grads_flat`` sends each model's
gradients as ONE flat buffer (1-4 MB => latency-bound over NVLink/NVSwitch; one launch instead of
one per parameter).  Works with the nccl backend on GPUs and with gloo on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world_size: int):
    """Contiguous row range [lo, hi) of rank `rank`; sizes differ by at most one row."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_grads_flat(params, group=None, average: bool = True):
    """Sum (or average) the .grad of `params` across ranks with a single all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    views, off = [], 0
    for g in grads:
        k = g.numel()
        views.append(flat[off:off + k].view_as(g))
        off += k
    torch._foreach_copy_(grads, views)      # one multi-tensor launch instead of one copy per parameter
    return flat.numel()


def broadcast_camera_choice(subject_id: int, cam_id: int, device, src: int = 0, group=None):
    """All ranks must project with the same (subject, camera) per iteration, as the reference does per
    batch (model_fk_gan_train.py:344-347): rank `src` draws, everyone receives."""
    t = torch.tensor([subject_id, cam_id], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return int(t[0].item()), int(t[1].item())
