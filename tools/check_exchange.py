#!/usr/bin/env python
"""N-GPU check of the gradient exchanges (enerf_b200/parallel.py) on the real training step:
  torchrun --nproc-per-node N tools/check_exchange.py
Every rank trains the same model on its own rays for a few steps with (a) AllReduceExchange + replicated FusedAdam and
(b) ShardedExchange (fp16 reduce-scatter, sharded FusedAdam, overlapped fp16 all-gather), eager and from a CUDA graph, starting
from identical parameters.  Reports, on rank 0: max |table_a - table_b| relative to max |update|, the MLP weights' difference,
that all ranks hold the same table after (b), and that an injected overflow on one rank makes every rank skip the step."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import parallel, synthetic  # noqa: E402
from enerf_b200.gridencoder.grid import half_shadow  # noqa: E402
from enerf_b200.nerf.network_ff import NeRFNetwork  # noqa: E402
from enerf_b200.optim import FusedAdam  # noqa: E402

BOUND = 2


def train(exchange_cls, steps, rank, dev, graph=False, poison_at=-1):
    torch.manual_seed(0)
    model = NeRFNetwork(bound=BOUND, cuda_ray=True, out_dim_color=1).to(dev).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.2, 0.2)
    grid = synthetic.ball_density_grid(BOUND, model.cascade)
    model.density_grid.copy_(torch.from_numpy(grid))
    model.density_bitfield.copy_(torch.from_numpy(synthetic.packbits_np(grid)))
    init = model.encoder.embeddings.detach().clone()
    opt = FusedAdam(model.get_params(5e-3), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    ex = exchange_cls(model, opt)
    o, d = synthetic.random_rays(1024, BOUND, seed=50 + rank)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    tg = torch.rand(1024, 1, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    bomb = torch.zeros(1, device=dev)

    def step():
        ex.begin_step()
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(o[None], d[None], staged=False, bg_color=1, perturb=False, out_dim_color=1, force_all_rays=not graph)
        loss = F.mse_loss(out["image"].reshape(-1, 1).float(), tg) * (1 + bomb * float("inf")).nan_to_num(posinf=float("inf"), nan=1.0)
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        ex.before_step(scaler)
        scaler.step(opt)
        scaler.update()
        ex.after_step()
        return loss

    run = step
    if graph:
        step()
        model.mean_count = int(model.step_counter[0, 0].item())
        from enerf_b200.graphs import GraphedStep
        g = GraphedStep(lambda: step(), [], warmup=2)
        run = lambda: g()                                   # noqa: E731
    scales = []
    for i in range(steps):
        bomb.fill_(1.0 if (i == poison_at and rank == dist.get_world_size() - 1) else 0.0)
        run()
        scales.append(float(scaler.get_scale()))
    ex.gather_master()
    torch.cuda.synchronize()
    with torch.autocast("cuda", dtype=torch.float16), torch.no_grad():
        model.encoder(o[:128], bound=BOUND)                 # makes sure the fp16 table is complete
    table16 = half_shadow(model.encoder.embeddings)[0].detach().float().clone()
    return dict(init=init, table=model.encoder.embeddings.detach().clone(), table16=table16, ws=model.sigma_net.weights.detach().clone(),
                wc=model.color_net.weights.detach().clone(), scales=scales)


def main():
    rank, world_size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world_size}
    for graph in (False, True):
        a = train(parallel.AllReduceExchange, 6, rank, dev, graph)
        b = train(parallel.ShardedExchange, 6, rank, dev, graph)
        assert torch.equal(a["init"], b["init"])
        upd = float((a["table"] - a["init"]).abs().max())
        tag = "graph" if graph else "eager"
        moved = (a["table"] - a["init"]).abs() > 0
        # Adam (eps 1e-15) turns any non-zero gradient into a full-size step, so entries whose gradient is within fp16 rounding of zero
        # may move differently with the fp16 wire; what must agree is the bulk: relative L2 distance of the UPDATES
        da, db = (a["table"] - a["init"]).double(), (b["table"] - a["init"]).double()
        out[tag] = {"max_update": upd, "entries_moved": int(moved.sum()), "update_rel_l2": float((da - db).norm() / da.norm()),
                    "table_diff_max_over_max_update": float((a["table"] - b["table"]).abs().max()) / upd,
                    "sigma_w_diff": float((a["ws"] - b["ws"]).abs().max()), "color_w_diff": float((a["wc"] - b["wc"]).abs().max())}
        ref = b["table16"].clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1.0 if torch.equal(ref, b["table16"]) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out[tag]["all_ranks_same_table"] = bool(same.item())
    # overflow on the last rank at step 2: every rank must halve its scale at that step
    c = train(parallel.ShardedExchange, 4, rank, dev, False, poison_at=2)
    s = torch.tensor(c["scales"], device=dev)
    lo, hi = s.clone(), s.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["overflow"] = {"scales": c["scales"], "identical_on_all_ranks": bool(torch.equal(lo, hi)), "halved_at_step_2": c["scales"][2] == c["scales"][1] / 2}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
