"""Ray-sharded data parallelism.

Rays are independent, the model (<= 50 MiB) is replicated, so the only exchange step of a
training iteration is one sum-allreduce of the gradients: the hash-grid gradient table
(13.0 M fp32 = 52 MB) reduced in place, and the two flat MLP gradient vectors (18 432 floats)
coalesced into one small bucket that is launched first.  Inference shards image rows and
all-gathers the result.  Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests); the
reference itself is single-process (its DDP hooks are unreachable, nerf/utils.py:351-353).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_total, rank, world_size):
    """Contiguous [lo, hi) slice of `n_total` units owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_o, rays_d, rank=None, world_size=None, *extra):
    """Slice [N,3] ray tensors (and any per-ray extras) for this rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(rays_o.shape[-2], rank, world_size)
    out = [rays_o[..., lo:hi, :], rays_d[..., lo:hi, :]]
    out += [e[..., lo:hi, :] if e.dim() == rays_o.dim() else e[..., lo:hi] for e in extra]
    return out


class GradientAllReduce:
    """Sum (or average) the gradients of `params` over all ranks, in as few collectives as possible.

    Parameters with at least `big` elements are reduced in place, one collective each (the
    hash-grid table); everything smaller shares one flat bucket.  Call `reduce()` after
    backward and before the optimizer / GradScaler step so the inf-check sees reduced grads."""

    def __init__(self, params, average=True, big=1 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.average = average
        self.big = [p for p in self.params if p.numel() >= big]
        self.small = [p for p in self.params if p.numel() < big]
        self._bucket = None

    def reduce(self, async_op=False):
        """Launch the collectives.  With async_op the caller overlaps other work and then calls
        `finish(pending)`; otherwise this returns after the gradients hold the reduced values."""
        rank, world_size = world()
        if world_size == 1:
            return []
        scale = 1.0 / world_size if self.average else 1.0
        pending = []
        small = [p for p in self.small if p.grad is not None]
        if small:
            n = sum(p.grad.numel() for p in small)
            dev = small[0].grad.device
            if self._bucket is None or self._bucket.numel() != n or self._bucket.device != dev:
                self._bucket = torch.empty(n, dtype=torch.float32, device=dev)
            o = 0
            for p in small:
                k = p.grad.numel()
                self._bucket[o:o + k].copy_(p.grad.reshape(-1))
                o += k
            pending.append((dist.all_reduce(self._bucket, op=dist.ReduceOp.SUM, async_op=True), "bucket", small, scale))
        for p in self.big:
            if p.grad is not None:
                pending.append((dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True), "inplace", [p], scale))
        if async_op:
            return pending
        self.finish(pending)
        return []

    def finish(self, pending):
        for work, kind, ps, scale in pending:
            work.wait()
            if kind == "bucket":
                o = 0
                for p in ps:
                    k = p.grad.numel()
                    p.grad.copy_(self._bucket[o:o + k].view_as(p.grad))
                    o += k
            if scale != 1.0:
                for p in ps:
                    p.grad.mul_(scale)


def all_gather_rows(local, n_total):
    """Concatenate per-rank row shards [n_r, ...] (made by `shard_bounds`) into [n_total, ...]."""
    rank, world_size = world()
    if world_size == 1:
        return local
    sizes = [shard_bounds(n_total, r, world_size) for r in range(world_size)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
