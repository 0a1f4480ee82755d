"""Scene sharding across the GPUs of one box.

Scenes are independent, so the batch dimension is split contiguously: rank r of
W owns scenes [r*B/W, (r+1)*B/W) for every input and output; forward and
backward need NO collective. The only exchange of the path is the gather of the
(small) per-rank loss gradients w.r.t. shared parameters after the local chain
rule -- one `all_gather` (NCCL over NVLink on the GPU box, gloo in CPU tests).
The reference has no distributed code at all; this is the new contract.
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous scene range [lo, hi) owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_inputs(tensors, rank, world):
    """Slice batched tensors to this rank's scenes (1-D empty A/b pass through)."""
    out = []
    for t in tensors:
        if t is None or t.dim() <= 1 and t.numel() == 0:
            out.append(t)
            continue
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return tuple(out)


def gather_loss_gradients(local_grad, group=None):
    """all_gather of one flat per-rank gradient vector -> [world, len]. Summing over
    dim 0 gives the gradient of the global loss w.r.t. shared parameters."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_grad.unsqueeze(0)
    world = dist.get_world_size(group)
    flat = local_grad.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view(world, flat.numel())


def gather_scene_outputs(local, B_total, group=None):
    """Reassemble a per-scene output [B_local, ...] into [B_total, ...] on every rank
    (used by tests; production keeps outputs sharded)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(B_total, r, world) for r in range(world)]
    maxb = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxb,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
