"""Ray-sharded data parallelism.

Rays are independent, the model (<= 50 MiB) is replicated, so the only exchange step of a
training iteration is the gradient exchange.  Two implementations with one interface
(`begin_step()`, `before_step(scaler)`, `after_step()`):

  AllReduceExchange   sum-allreduce of every gradient (the hash-grid gradient table, 13.0 M fp32 =
                      52 MB, in place; the MLP gradient vectors coalesced into one small bucket
                      launched first), every rank then runs the full optimizer step.
  ShardedExchange     reduce-scatter of the table gradient -> each rank runs Adam on its 1/N slice
                      (FusedAdam, which also writes that slice of the fp16 table) -> all-gather of the
                      fp16 table (what the kernels read under autocast): 45 + 23 MB on the wire instead
                      of 91 MB at N = 8, 1/N of the Adam work, no 1/N scaling pass (folded into the Adam
                      kernel).  The fp32 master copy of a rank is current for its own slice only;
                      `gather_master()` completes it (checkpoints, fp32 evaluation).

Inference shards image rows and all-gathers the result.  Backend-agnostic (`nccl` on GPUs, `gloo`
in the CPU tests); the reference itself is single-process (its DDP hooks are unreachable,
nerf/utils.py:351-353).
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


class NoExchange:
    """no communication (single GPU, or diagnosis of what the exchange costs)"""
    name = "none"

    def __init__(self, model=None, optimizer=None, big=1 << 20):
        pass

    def begin_step(self):
        pass

    def before_step(self, scaler=None):
        pass

    def after_step(self):
        pass

    def gather_master(self):
        pass


class AllReduceExchange:
    """every gradient sum-allreduced (fp32); the optimizer step is replicated.  With FusedAdam the 1/N of the average is applied
    inside the Adam kernel instead of by a separate pass over the 52 MB table gradient."""
    name = "allreduce"

    def __init__(self, model, optimizer=None, big=1 << 20):
        self._fused = optimizer is not None and getattr(optimizer, "honours_enerf_shard", False)
        self.reducer = GradientAllReduce(list(model.parameters()), average=not self._fused, big=big)

    def begin_step(self):
        pass

    def before_step(self, scaler=None):
        rank, world_size = world()
        if world_size == 1:
            return
        self.reducer.reduce()
        if self._fused:
            for p in self.reducer.params:
                if p.grad is not None and p.grad.is_contiguous():
                    p._enerf_shard = (0, p.numel(), p.grad.view(-1), 1.0 / world_size)

    def after_step(self):
        pass

    def gather_master(self):
        pass


def _reduce_scatter_sum(out, full):
    """out <- this rank's slice of the sum over ranks of `full` (gloo has no reduce-scatter: all-reduce a copy and slice)"""
    if dist.get_backend() == "gloo":
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        n = out.numel()
        out.copy_(tmp.view(-1)[dist.get_rank() * n:(dist.get_rank() + 1) * n])
    else:
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM)


def _all_gather_inplace(full, lo, hi):
    """every rank contributes full[lo:hi] (its slice, equal sizes); afterwards `full` is complete everywhere"""
    if dist.get_backend() == "gloo":
        parts = [torch.empty(hi - lo, dtype=full.dtype) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, full[lo:hi].clone())
        full.copy_(torch.cat(parts))
    else:
        dist.all_gather_into_tensor(full, full[lo:hi])


class ShardedExchange:
    """reduce-scatter of the big gradients, optimizer step on the local slice, all-gather of the updated (fp16) table.

    Per step, for the hash table (13.0 M parameters, N ranks):
      before_step  gradient cast fp32 -> fp16 (the wire format; the reference itself accumulates these gradients in fp16,
                   gridencoder.cu:296-302), reduce-scatter (26 MB * (N-1)/N per rank instead of the 52 MB * 2(N-1)/N of an fp32
                   all-reduce), non-finite check of the reduced slice;
      optimizer    FusedAdam on the rank's 1/N slice: un-scales (GradScaler) and divides by N in the kernel, writes the fp32
                   master slice and the fp16 table slice;
      begin_step   (of the NEXT step) all-gather of the fp16 table on a side stream, overlapped with ray set-up and the
                   occupancy-grid march, which do not read the table; the first hash-grid kernel waits for it.
    Requires an optimizer that honours `param._enerf_shard = (lo, hi, reduced_grad_slice, 1/world)` — `enerf_b200.optim.FusedAdam`.
    GradScaler: its inf check runs on every rank's LOCAL gradients, so a non-finite value in the reduced slice of any rank (an
    overflow anywhere, incl. in the fp16 cast) is made visible to all: one 4-byte all-reduce, after which the local gradient's first
    element is NaN on every rank — all ranks skip the same steps and keep the same scale.
    The fp32 master copy of a rank is current for its own slice only; `gather_master()` completes it (checkpoints, EMA, fp32 eval)."""
    name = "fp16 reduce-scatter + sharded Adam + fp16 all-gather (overlapped with the next march)"

    def __init__(self, model, optimizer=None, big=1 << 20, wire_dtype=torch.float16):
        self.rank, self.world = world()
        self.wire_dtype = wire_dtype
        params = [p for p in model.parameters() if p.requires_grad]
        self.big = [p for p in params if p.numel() >= big and p.numel() % (8 * max(self.world, 1)) == 0] if self.world > 1 else []
        ids = {id(p) for p in self.big}
        self.small = GradientAllReduce([p for p in params if id(p) not in ids], average=True, big=1 << 62)
        self._slices, self._wire = {}, {}
        self._dirty = False            # table slices written by the optimizer, not yet gathered
        self._event = None             # completion of the last gather (side stream)
        self._side = None
        self._fused = optimizer is not None and getattr(optimizer, "honours_enerf_shard", False)
        if self.world > 1 and self.big and not self._fused:
            raise RuntimeError("ShardedExchange needs an optimizer that understands sharded parameters (enerf_b200.optim.FusedAdam)")
        for p in self.big:
            p._enerf_wait = self._wait_table          # consulted by gridencoder.grid._half_table before the table is read

    def _slice(self, p):
        n = p.numel() // self.world
        return self.rank * n, (self.rank + 1) * n

    def _gather_tables(self):
        for p in self.big:
            lo, hi = self._slice(p)
            slot = getattr(p, "_enerf_half", None)
            if slot is not None and slot[0].shape == p.shape and slot[0].device == p.device:
                _all_gather_inplace(slot[0].view(-1), lo, hi)          # the fp16 table the kernels read
            else:
                _all_gather_inplace(p.data.view(-1), lo, hi)

    def begin_step(self):
        """start gathering the table slices the previous optimizer step wrote; returns immediately"""
        if self.world == 1 or not self._dirty:
            return
        self._dirty = False
        if self.big and self.big[0].is_cuda:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._gather_tables()
                self._event = torch.cuda.Event()
                self._event.record(self._side)
        else:
            self._gather_tables()

    def _wait_table(self, p=None):
        """the table is about to be read on the current stream"""
        if self.world == 1:
            return
        if self._dirty:
            self.begin_step()
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None

    def _to_wire(self, g, wire, flag):
        """wire <- fp16(g); flag[0] = 1 if a value is non-finite or too large for the fp16 sum over the ranks"""
        limit = 65504.0 / self.world
        if g.is_cuda:
            from . import _lib
            _lib.call("enerf_grad_to_half", _lib.ptr(g), _lib.ptr(wire), g.numel(), limit, _lib.ptr(flag), _lib.stream())
        else:                                                # gloo tests
            wire.copy_(g)
            flag.copy_(torch.maximum(flag, (~(g.abs() <= limit)).any().to(flag.dtype).reshape(1)))

    def before_step(self, scaler=None):
        """Collectives per step: ONE all-reduce of [small gradients | overflow flag] (74 KB) and one reduce-scatter per table.
        Nothing is scaled here: the sums carry a factor N that the optimizer kernel removes (`_enerf_shard[3]`)."""
        if self.world == 1:
            return
        small = [p for p in self.small.params if p.grad is not None]
        big = [p for p in self.big if p.grad is not None]
        if not small and not big:
            return
        dev = (small + big)[0].grad.device
        flag = torch.zeros(1, dtype=torch.float32, device=dev)
        wires = []
        for p in big:
            g = p.grad.contiguous().view(-1)
            key = id(p)
            lo, hi = self._slice(p)
            if key not in self._slices or self._slices[key].numel() != hi - lo or self._slices[key].device != g.device:
                self._slices[key] = torch.empty(hi - lo, dtype=self.wire_dtype, device=g.device)
                self._wire[key] = torch.empty(g.numel(), dtype=self.wire_dtype, device=g.device) if self.wire_dtype != g.dtype else None
            if self._wire[key] is not None:
                self._to_wire(g, self._wire[key], flag)
                g = self._wire[key]
            wires.append(g)
        # small gradients and the overflow flag travel together
        bucket = torch.cat([p.grad.reshape(-1).float() for p in small] + [flag])
        work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True)
        for p, g in zip(big, wires):
            lo, hi = self._slice(p)
            shard = self._slices[id(p)]
            _reduce_scatter_sum(shard, g)
            p._enerf_shard = (lo, hi, shard, 1.0 / self.world)
        work.wait()
        o = 0
        for p in small:
            k = p.grad.numel()
            view = bucket[o:o + k]
            if self._folds(p, view):
                p._enerf_shard = (0, k, view, 1.0 / self.world)      # the summed gradient, 1/N applied inside the Adam kernel
            else:
                p.grad.copy_((view / self.world).view_as(p.grad))
            o += k
        # a non-finite / too-large value on ANY rank -> NaN in the first element of every rank's local gradient: the GradScaler of every
        # rank then finds an inf, skips the step and backs off
        poison = torch.where(bucket[-1:] > 0, torch.full_like(flag, float("nan")), torch.zeros_like(flag))
        for p in (big or small)[:1]:
            p.grad.view(-1)[:1].add_(poison.to(p.grad.dtype))

    def _folds(self, p, view):
        return self._fused and p.is_cuda and view.data_ptr() % 16 == 0

    def after_step(self):
        if self.world > 1 and self.big:
            self._dirty = True

    @torch.no_grad()
    def gather_master(self):
        """complete the fp32 parameters (and the fp16 table) on every rank"""
        if self.world == 1:
            return
        self._wait_table()
        for p in self.big:
            lo, hi = self._slice(p)
            _all_gather_inplace(p.data.view(-1), lo, hi)


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
