"""N>1 host logic on CPU: world_size-2 gloo processes exercise ray sharding, the gradient
allreduce bucket logic and the inference all-gather of enerf_b200.parallel."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from enerf_b200 import parallel


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 65536, 640000):
        for w in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.manual_seed(0)
        N = 101
        rays_o, rays_d = torch.randn(1, N, 3), torch.randn(1, N, 3)
        target = torch.randn(N)
        big = torch.nn.Parameter(torch.randn(1 << 20, 2) * 0.01)          # stands in for the hash table
        w1 = torch.nn.Parameter(torch.randn(3))
        w2 = torch.nn.Parameter(torch.randn(3))

        def loss_of(o, d, tg):
            feat = big[:o.shape[-2] * 2].view(-1, 2).sum() * 0 + (o.reshape(-1, 3) * w1).sum(-1) + (d.reshape(-1, 3) * w2).sum(-1)
            return ((feat - tg) ** 2).sum() + (big[:8] ** 2).sum() * (rank + 1)

        o, d, tg = parallel.shard_rays(rays_o, rays_d, None, None, target.view(1, N))
        loss = loss_of(o, d, tg.reshape(-1)) / N
        loss.backward()
        red = parallel.GradientAllReduce([big, w1, w2], average=False)
        assert len(red.big) == 1 and len(red.small) == 2
        pending = red.reduce(async_op=True)
        red.finish(pending)

        # single-process reference on the full batch
        if rank == 0:
            b2, a1, a2 = (torch.nn.Parameter(p.detach().clone()) for p in (big, w1, w2))
            feat = (rays_o.reshape(-1, 3) * a1).sum(-1) + (rays_d.reshape(-1, 3) * a2).sum(-1)
            full = ((feat - target) ** 2).sum() / N + sum((b2[:8] ** 2).sum() * (r + 1) for r in range(world_size)) / N
            full.backward()
            ok = torch.allclose(w1.grad, a1.grad, atol=1e-5) and torch.allclose(w2.grad, a2.grad, atol=1e-5) \
                and torch.allclose(big.grad, b2.grad, atol=1e-6)
            out.put(("grads", bool(ok)))

        lo, hi = parallel.shard_bounds(N, rank, world_size)
        img = torch.arange(N, dtype=torch.float32).view(N, 1).repeat(1, 3)[lo:hi]
        full_img = parallel.all_gather_rows(img, N)
        if rank == 0:
            out.put(("gather", bool(torch.equal(full_img[:, 0], torch.arange(N, dtype=torch.float32)))))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(out.get(timeout=10) for _ in range(2))
    assert got == {"grads": True, "gather": True}


def _sharded_worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.manual_seed(0)
        big = torch.nn.Parameter(torch.randn(1 << 20, 2) * 0.01)           # stands in for the hash table (same values on every rank)
        small = torch.nn.Parameter(torch.randn(7))
        model = torch.nn.ParameterList([big, small])

        class _SliceSGD:                                                    # an optimizer that honours `_enerf_shard` like FusedAdam does
            honours_enerf_shard = True

            def step(self):
                lo, hi, g, mul = big._enerf_shard
                big.data.view(-1)[lo:hi] -= 0.5 * mul * g.float()
                small.data -= 0.5 * small.grad

        ex = parallel.ShardedExchange(model, _SliceSGD())
        assert len(ex.big) == 1 and "reduce-scatter" in ex.name
        g_local = torch.full_like(big, float(rank + 1))
        g_local.view(-1)[5] = 10.0 * (rank + 1)
        big.grad, small.grad = g_local.clone(), torch.full((7,), float(rank))
        before = big.detach().clone()
        ex.before_step()
        lo, hi, shard, mul = big._enerf_shard
        want_sum = sum(r + 1 for r in range(world_size))
        ok = (hi - lo) * world_size == big.numel() and abs(mul - 1.0 / world_size) < 1e-12
        ok &= bool(torch.all(shard[(1 if rank == 0 else 0):6 if rank == 0 else None] == want_sum)) if rank != 0 else bool(shard[5] == 10.0 * want_sum and shard[0] == want_sum)
        ok &= bool(torch.allclose(small.grad, torch.full((7,), sum(range(world_size)) / world_size)))
        ok &= bool(torch.isfinite(big.grad).all())                         # nothing overflowed: the local gradient is not poisoned
        assert shard.dtype == torch.float16                                # the wire format
        _SliceSGD().step()
        ex.after_step()
        ex.begin_step()                                                    # deferred gather (no fp16 shadow here: the fp32 slices travel)
        expect = before - 0.5 * (want_sum / world_size)
        expect.view(-1)[5] = before.view(-1)[5] - 0.5 * 10.0 * want_sum / world_size
        ok &= bool(torch.allclose(big.detach(), expect, atol=1e-6))
        # overflow on ONE rank is seen by all
        big.grad = torch.ones_like(big)
        if rank == world_size - 1:
            big.grad.view(-1)[-3] = float("inf")
        small.grad = torch.zeros(7)
        ex.before_step()
        ok &= bool(torch.isnan(big.grad.view(-1)[0]))
        ex.gather_master()
        out.put((f"rank{rank}", bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_exchange():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = dict(out.get(timeout=10) for _ in range(2))
    assert got == {"rank0": True, "rank1": True}
