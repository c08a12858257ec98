#!/usr/bin/env python
"""bench.py — train rays/sec of the E-NeRF volume-rendering hot path on B200 (BASELINE.json).

One "step" = one full training iteration on one batch of synthetic rays through the reference-
facing API (`model.render(...)` of the mirrored NeRFNetwork): occupancy-grid march -> hash grid ->
sigma-net -> SH -> colour-net -> composite, MSE loss, backward through every stage, GradScaler +
Adam step, and every 16th step the occupancy-grid refresh (`update_extra_state`, as
nerf/utils.py:945-947 schedules it); with N>1 GPUs each rank renders its own 4096-ray batch (weak
scaling) and the gradients are exchanged over NCCL before the optimizer step.

  python bench.py [--gpus N --steps K --warmup W]        our arm (default N=1)
  python bench.py --impl reference ...                   reference arm: the reference's pure-
        PyTorch renderer (nerf/network.py + NeRFRenderer.run, no --cuda_ray/--ff) on the host
        CPU cores, restated in oracle/cpu_reference.py (the reference tree is not on the box; the
        port is pinned to the reference's own classes by tests/test_cpu_port_pinned.py).

Prints ONE JSON line.  Headline workload = BASELINE.json configs[1]: shakeCarpet1-shaped scene
(bound 3, cascade 3, hashgrid L=16 T=2^19 F=2, ffmlp 64x2 sigma-net + 64x3 colour-net,
out_dim_color 1, fp16 autocast, cuda_ray), 4096 rays/batch, fwd+bwd.  The same line carries, outside
the headline's timed region, the other BASELINE configs as objects:
  strong_scaling  configs[4]: 65 536 rays per global batch split over the ranks (+ gradient exchange)
  event_step      configs[2]: 8192 event pairs, two renders, normalised event loss (N = 1 only)
  run_variant     the path every shipped config runs: 4096 rays x 512 fixed steps, nerf/network.py topology
  render          configs[3]: 800x800 full-frame inference, rows sharded over the ranks
  gpu_bar         the reference's own CUDA build (oracle/_ref) next to this repo's kernels (N = 1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "train_rays_per_sec", "rays/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload (bytes)
NCU_TRAFFIC_SRC = "profiles/r2_43_ncu_full_step_kernels.md (3.29 M samples/launch)"
NCU_TRAFFIC = {"grid_encode_backward": 292.828672e6 + 6.303488e6, "grid_encode_forward": 62.323968e6 + 167.823360e6,
               "march_rays_train": 0.622848e6 + 46.673664e6, "field_color_backward": 237.623808e6 + 168.468224e6,
               "field_sigma_backward": 448.601344e6 + 182.442240e6, "field_sigma_forward": 250.630144e6 + 180.914176e6,
               "field_color_forward": 211.083008e6 + 12.192256e6}


def canon(name):
    """C-ABI entry point -> the name its roofline is booked under: `_xf` (positions mapped inside the hash-grid kernels) and `_bounded`
    (march limited to the occupied box) are the same kernels called with extra arguments"""
    for suffix in ("_xf", "_bounded"):
        if name.endswith(suffix):
            name = name[:-len(suffix)]
    return name
RAYS = 4096
BOUND = 3
N_BATCHES = 8            # distinct ray batches cycled through the timed region
REFRESH_EVERY = 16       # nerf/utils.py:945-947


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="replay the step from a CUDA graph (auto: fall back to eager launches if capture fails)")
    ap.add_argument("--cpu-rays", type=int, default=256, help="rays per CPU-baseline step (bounded sample)")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"], help="Adam step: this repo's fused kernel or torch.optim.Adam(fused=True)")
    ap.add_argument("--skip", default="", help="comma-separated extra objects to skip: strong_scaling,event_step,run_variant,render,gpu_bar,extra_state")
    ap.add_argument("--no-render", action="store_true", help="same as --skip render")
    ap.add_argument("--nvtx", action="store_true", help="NVTX range around every C-ABI call (profiling; named after the entry point)")
    ap.add_argument("--only", default="", help="profiling aid: run ONE of event_step / run_variant alone and print its object (no headline line)")
    ap.add_argument("--fuse-encoder", default="on", choices=["on", "off"],
                    help="off: the inference render evaluates the field as three kernels (gather, sigma-net, colour-net) instead of one (csrc/field_infer.cu); diagnosis")
    ap.add_argument("--exchange", default="sharded", choices=["sharded", "allreduce", "none"],
                    help="gradient exchange at N > 1 (enerf_b200/parallel.py); none = no exchange at all (diagnosis only: the ranks diverge)")
    return ap.parse_args()


def workload_config(n_rays, extra=None):
    cfg = {"workload": "BASELINE configs[1]: shakeCarpet1-shaped, bound 3 / cascade 3, hashgrid L=16 T=2^19 F=2, ffmlp sigma 32-64-64-16 + "
                       "colour 32-64-64-64-16, out_dim_color 1, fp16 autocast, cuda_ray, max_steps 1024, dt_gamma 0, perturb, "
                       "analytic-ball occupancy (r = 0.5*bound), cameras at 0.6*bound; full train step (fwd+bwd+GradScaler+Adam)",
           "rays_per_gpu": n_rays}
    cfg.update(extra or {})
    return cfg


# --------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_reference
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded sample: a few hundred rays per step so that K+W steps end within minutes
    r = cpu_reference.time_train_steps(n_rays=args.cpu_rays, num_steps=512, bound=BOUND, out_dim_color=1, steps=steps, warmup=warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.cpu_rays, {"note": "reference renderer = fixed 512 steps/ray (cuda_ray=False, ff=False: what every shipped "
                                                              "config runs); each step is a bounded sample of the workload; our arm's `run_variant` "
                                                              "object times the same topology and step count on the GPU"}),
            "cpu_baseline": {"value": r["rays_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                             "pinned_by": "tests/test_cpu_port_pinned.py (image, depth, gradients vs the reference's own classes, 1e-6)"},
            "samples_per_sec": r["rays_per_s"] * 512,
            "e2e": {"value": r["rays_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 20 ms in the background; `stop(t0, t1)` keeps the samples whose timestamp falls inside the
    timed region [t0, t1] (host wall clock, taken right after the synchronisations that bracket it)."""
    QUERY = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 3:          # region shorter than a few sampling periods: fall back to every sample taken under load
            inside, window = rows, "whole run (timed region shorter than 3 samples)"
        if not inside:
            return out
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": max(r[2] for r in inside), "power_w_max": max(r[3] for r in inside),
                "reasons": sorted(reasons), "samples": len(inside), "window": window}


class Dist:
    """the handful of collective helpers the measurements need (no-ops at world 1)"""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, value, dev):
        import torch
        t = torch.tensor([float(value)], device=dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum(self, value, dev):
        import torch
        t = torch.tensor([float(value)], device=dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def all_ok(self, ok, dev):
        import torch
        t = torch.tensor([1 if ok else 0], device=dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t[0]) == 1


FUSE_ENCODER = True      # --fuse-encoder


def make_ff_model(dev, bound):
    import torch
    from enerf_b200 import synthetic
    from enerf_b200.nerf.network_ff import NeRFNetwork
    model = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                        out_dim_color=1).to(dev)
    model.fuse_infer = FUSE_ENCODER
    model.train()
    grid = synthetic.ball_density_grid(bound, model.cascade)
    model.density_grid.copy_(torch.from_numpy(grid))
    model.density_bitfield.copy_(torch.from_numpy(synthetic.packbits_np(grid)))
    return model


def make_optimizer(model, kind):
    import torch
    # E-NeRF's optimizer (main_nerf.py:211-214: Adam, betas (0.9, 0.99), eps 1e-15): this repo's fused step (enerf_b200/optim.py,
    # same update rule, parity-tested against torch's) or torch's own fused Adam with --optimizer torch
    if kind == "fused":
        from enerf_b200.optim import FusedAdam
        return FusedAdam(model.get_params(5e-3), betas=(0.9, 0.99), eps=1e-15)
    return torch.optim.Adam(model.get_params(5e-3), betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True)


class TrainLoop:
    """render -> MSE -> backward -> gradient exchange -> GradScaler + Adam on `batches` (device-resident (o, d, target) triples),
    optionally replayed from a CUDA graph; `refresh()` runs the occupancy-grid refresh for real and then restores the analytic
    grid so that the marched workload stays the one named in `config` (the refresh's cost is paid, its result is discarded)."""

    def __init__(self, model, optimizer, exchange, batches, D, graph_mode, dev):
        import torch
        import torch.nn.functional as F
        self.model, self.optimizer, self.exchange, self.batches, self.D, self.dev = model, optimizer, exchange, batches, D, dev
        self.scaler = torch.amp.GradScaler("cuda", enabled=True)
        self.bg = torch.ones(1, device=dev)
        self.kw = dict(num_steps=512, upsample_steps=0, max_ray_batch=5096, dt_gamma=0, out_dim_color=1)
        self.grid_saved = model.density_grid.clone()
        self.bits_saved = model.density_bitfield.clone()

        def step(o, d, tg):
            exchange.begin_step()
            with torch.autocast("cuda", dtype=torch.float16):
                out = model.render(o.unsqueeze(0), d.unsqueeze(0), staged=False, bg_color=self.bg, perturb=True, **self.kw)
            loss = F.mse_loss(out["image"].reshape(-1, 1).float(), tg)
            optimizer.zero_grad(set_to_none=True)
            self.scaler.scale(loss).backward()
            exchange.before_step(self.scaler)
            self.scaler.step(optimizer)
            self.scaler.update()
            exchange.after_step()
            return loss

        self.eager_step = step
        # every batch once with exact sizing (one D2H read each): the largest sample count fixes M like update_extra_state does
        model.mean_count = 0
        totals = []
        for b in batches:
            model.local_step = 0
            step(*b)
            totals.append(int(model.step_counter[0, 0].item()))
        self.totals = totals
        model.mean_count = max(totals)
        self.samples_per_step = sum(totals) / len(totals)
        self.M = model.mean_count + (128 - model.mean_count % 128)
        self.graphed = None
        self.step = step
        if graph_mode in ("on", "auto"):       # NCCL collectives are capturable: the exchange is part of the graph at N > 1
            try:
                from enerf_b200.graphs import GraphedStep
                self.graphed = GraphedStep(step, list(batches[0]), warmup=3)
                self.step = self.graphed
            except Exception as e:  # noqa: BLE001
                if graph_mode == "on":
                    raise
                print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager launches", file=sys.stderr)
                torch.cuda.synchronize()
                self.graphed, self.step = None, step
            if not D.all_ok(self.graphed is not None, dev):          # all ranks replay, or none does
                self.graphed, self.step = None, step

    def refresh(self):
        import torch
        m = self.model
        keep = m.mean_count
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
        m.density_grid.copy_(self.grid_saved)
        m.density_bitfield.copy_(self.bits_saved)
        m.mean_count = keep

    def run(self, K, refresh_every=0, host=None):
        """K steps, device-timed between two barriers; host: list of pinned (o, d, target) triples -> copied in every step and the
        loss read back every step (the end-to-end variant).  Returns (ms_total, last loss tensor or float)."""
        import torch
        from enerf_b200 import _lib
        nb = len(self.batches)
        self.D.barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        e0.record()
        h0 = time.perf_counter()
        loss = None
        for i in range(K):
            if refresh_every and i % refresh_every == 0:
                self.refresh()
            if host is None:
                loss = self.step(*self.batches[i % nb])
            else:
                o, d, tg = (h.to(self.dev, non_blocking=True) for h in host[i % nb])
                loss = self.step(o, d, tg).item()          # D2H read of the step's result
        enqueue_ms = (time.perf_counter() - h0) * 1e3 / max(K, 1)
        e1.record()
        self.D.barrier()
        wall1 = time.time()
        ms = self.D.max(e0.elapsed_time(e1), self.dev)
        launches = _lib.launch_count() - n0
        if self.graphed is not None:
            launches += K * self.graphed.launches_per_replay
        return dict(ms=ms, loss=loss, launches=int(launches), wall=(wall0, wall1), enqueue_ms=enqueue_ms)

    def release(self):
        if self.graphed is not None:
            self.graphed.graph = None
            self.graphed.static_out = None
            self.graphed = None


def ray_batches(n_rays, n_batches, bound, seed0, dev, pinned=False):
    import numpy as np
    import torch
    from enerf_b200 import synthetic
    dev_b, host_b = [], []
    for i in range(n_batches):
        o_np, d_np = synthetic.random_rays(n_rays, bound, seed=seed0 + i)
        tgt_np = np.random.default_rng(seed0 + i).random((n_rays, 1)).astype(np.float32)
        dev_b.append(tuple(torch.from_numpy(a).to(dev) for a in (o_np, d_np, tgt_np)))
        if pinned:
            host_b.append(tuple(torch.from_numpy(a).pin_memory() for a in (o_np, d_np, tgt_np)))
    return dev_b, host_b


def render_bench(model, dev, D, frames=3, res=800):
    """BASELINE configs[3]: full-frame inference render (eval mode, perturb off, max_steps 1024), image rows sharded over
    the ranks and gathered on every rank inside the timed region.  Msamples/s counts the samples actually shaded."""
    import numpy as np
    import torch
    from enerf_b200 import parallel, synthetic
    pose = synthetic.look_at_poses(1, 0.6 * BOUND, seed=7)[0]
    lo, hi = parallel.shard_bounds(res, D.rank, D.world)
    pix = np.arange(lo * res, hi * res)
    o_np, d_np = synthetic.pinhole_rays(pose, res, res, 50.0, pix)
    o, d = torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)
    was_training = model.training
    model.eval()
    samples, ms, full = 0, [], None
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for f in range(frames + 1):
            D.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = model.render(o.unsqueeze(0), d.unsqueeze(0), staged=False, bg_color=1, perturb=False, dt_gamma=0, max_steps=1024, out_dim_color=1)
            local = torch.cat([out["image"][0], out["depth"][0].unsqueeze(-1)], dim=-1)
            full = parallel.all_gather_rows(local, res * res)          # [res*res, C+1] on every rank
            e1.record()
            torch.cuda.synchronize()
            if f > 0:
                ms.append(e0.elapsed_time(e1))
                samples = model.last_render_stats["samples"]
    model.train(was_training)
    frame_ms = D.max(float(np.median(ms)), dev)
    total_samples = D.sum(samples, dev)
    return {"workload": f"BASELINE configs[3]: {res}x{res} full-frame inference, rows sharded over {D.world} GPU(s), all_gather of the image inside the timed region",
            "frame_ms": frame_ms, "msamples_per_s": total_samples / (frame_ms * 1e-3) / 1e6, "samples_shaded": int(total_samples),
            "iterations": int(model.last_render_stats["iterations"]), "host_syncs": int(model.last_render_stats.get("host_syncs", -1)),
            "image_finite": bool(torch.isfinite(full).all()), "image_rows": int(full.shape[0])}


def strong_scaling_bench(args, dev, D, exchange_cls, K):
    """BASELINE configs[4]: spiral1-shaped training, 65 536 rays per GLOBAL batch split evenly over the ranks, gradients exchanged
    over NCCL every step.  Same network and step as the headline; reported as its own object so that the headline stays configs[1]."""
    import torch
    from enerf_b200 import parallel
    total = 65536
    lo, hi = parallel.shard_bounds(total, D.rank, D.world)
    torch.manual_seed(0)
    model = make_ff_model(dev, BOUND)
    optimizer = make_optimizer(model, args.optimizer)
    exchange = exchange_cls(model, optimizer)
    import numpy as np
    from enerf_b200 import synthetic
    o_np, d_np = synthetic.random_rays(total, BOUND, seed=4242)
    tgt_np = np.random.default_rng(4242).random((total, 1)).astype(np.float32)
    batch = tuple(torch.from_numpy(a[lo:hi]).to(dev) for a in (o_np, d_np, tgt_np))
    loop = TrainLoop(model, optimizer, exchange, [batch], D, args.graph, dev)
    for _ in range(3):
        loop.step(*batch)
    r = loop.run(K)
    ms = r["ms"] / K
    samples = D.sum(loop.samples_per_step, dev)
    out = {"workload": "BASELINE configs[4]: spiral1-shaped (bound 3, ff + cuda_ray, fp16), 65 536 rays per global batch split over the ranks, "
                       f"gradient exchange '{exchange.name}' every step; full train step",
           "global_rays": total, "rays_per_gpu": hi - lo, "n_gpus": D.world, "scaling": "strong", "steps": K, "ms_per_step": ms,
           "rays_per_s": total / (ms * 1e-3), "samples_per_step": int(samples), "samples_per_s": samples / (ms * 1e-3),
           "launch": "cuda-graph replay" if loop.graphed is not None else "eager", "loss_finite": bool(torch.isfinite(r["loss"]).all())}
    loop.release()
    del loop, model, optimizer, exchange
    torch.cuda.empty_cache()
    return out


def event_step_bench(args, dev, K):
    """BASELINE configs[2]: mocapDesk2-shaped event step (event_only + accumulate_evs, C_thres = -1 normalised loss, 8192 event pairs,
    bound 2): device pair sampler (N3) -> event rays + near/far (N2) -> two renders -> fused event loss (N1) -> backward -> GradScaler +
    FusedAdam (N4) — nerf/utils.py:482-573 with nerf/provider.py:1364-1448 in front."""
    import numpy as np
    import torch
    from enerf_b200 import events, synthetic
    from enerf_b200.graphs import GraphedStep
    bound, pairs = 2, 8192
    torch.manual_seed(0)
    model = make_ff_model(dev, bound)
    ev, num_succ, no_succ, poses = synthetic.event_frame(bound=bound)
    sampler = events.EventPairSampler(ev, num_succ, no_succ, acc_max_num_evs=8, poses_evs=poses, device=dev)
    focal = synthetic.EVENT_H / (2 * np.tan(np.radians(50.0) / 2))
    intr = (focal, focal, synthetic.EVENT_W / 2, synthetic.EVENT_H / 2)
    optimizer = make_optimizer(model, args.optimizer)
    scaler = torch.amp.GradScaler("cuda")
    kw = dict(num_steps=512, upsample_steps=0, max_ray_batch=5096, dt_gamma=0, out_dim_color=1)
    bg = torch.rand(1, 1, 1, device=dev)

    def step():
        batch = sampler.sample(pairs, intr)
        with torch.autocast("cuda", dtype=torch.float16):
            out1 = model.render(batch["rays_evs_o1"], batch["rays_evs_d1"], staged=False, bg_color=bg, perturb=True, **kw)
            out2 = model.render(batch["rays_evs_o2"], batch["rays_evs_d2"], staged=False, bg_color=bg, perturb=True, **kw)
        loss, _ = events.event_loss(out1["image"].float(), out2["image"].float(), batch["pols"], use_luma=False, linlog=True, C_thres=-1,
                                    event_only=True)
        optimizer.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        return loss

    step()                                              # sizes the sample buffers
    totals = model.step_counter[:2, 0].tolist()
    model.mean_count = int(1.03 * max(totals))
    run, graphed = step, None
    if args.graph in ("on", "auto"):
        try:
            graphed = GraphedStep(lambda: step(), [], warmup=3)
            run = lambda: graphed()                     # noqa: E731
        except Exception as e:  # noqa: BLE001
            if args.graph == "on":
                raise
            print(f"[bench] event step: graph capture failed ({type(e).__name__}: {e})", file=sys.stderr)
            torch.cuda.synchronize()
    for _ in range(3):
        loss = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out = {"workload": "BASELINE configs[2]: mocapDesk2-shaped event step: event_only + accumulate_evs, normalised loss (C_thres=-1), bound 2, ff + cuda_ray, "
                       "fp16 autocast; sampler (N3) -> event rays (N2) -> 2 renders -> event loss (N1) -> backward -> GradScaler + Adam (N4)",
           "event_pairs_per_step": pairs, "steps": K, "ms_per_step": ms, "event_pairs_per_s": pairs / (ms * 1e-3),
           "rendered_rays_per_s": 2 * pairs / (ms * 1e-3), "samples_per_render": [int(t) for t in totals],
           "samples_per_s": float(sum(totals)) / (ms * 1e-3), "launch": "cuda-graph replay" if graphed is not None else "eager",
           "loss": float(loss.detach()), "loss_finite": bool(torch.isfinite(loss.detach()))}
    if graphed is not None:
        graphed.graph = None
    del model, optimizer, sampler
    torch.cuda.empty_cache()
    return out


def run_variant_bench(args, dev, K, peaks):
    """What every shipped E-NeRF config executes (cuda_ray = False, ff = False, configs/*/*.txt:23-30): NeRFRenderer.run with 512 fixed
    steps per ray through the nerf/network.py topology — here on the tcgen05 kernels (enerf_b200/nerf/network.py), fp16 autocast,
    4096 rays -> 2.10 M samples per step, fwd + bwd + GradScaler + Adam.  The colour mask's row count stays on the device (every kernel of
    the compacted colour batch reads it there), so the step has no host synchronisation and replays from a CUDA graph."""
    import torch
    import torch.nn.functional as F
    from enerf_b200 import _lib
    from enerf_b200.nerf.network import NeRFNetwork
    n_rays, T = 4096, 512
    torch.manual_seed(0)
    model = NeRFNetwork(encoding="hashgrid", bound=BOUND, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                        out_dim_color=1).to(dev).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.3, 0.3)          # "trained-like" table so that densities (and the colour mask) are non-degenerate
    optimizer = make_optimizer(model, args.optimizer)
    scaler = torch.amp.GradScaler("cuda")
    batches, _ = ray_batches(n_rays, 2, BOUND, 900, dev)
    bg = torch.ones(1, device=dev)
    kw = dict(num_steps=T, upsample_steps=0, max_ray_batch=5096, dt_gamma=0, out_dim_color=1)

    def step(o, d, tg):
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(o.unsqueeze(0), d.unsqueeze(0), staged=False, bg_color=bg, perturb=True, **kw)
        loss = F.mse_loss(out["image"].reshape(-1, 1).float(), tg)
        optimizer.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        return loss

    eager, graphed = step, None
    for i in range(3):
        step(*batches[i % 2])
    if args.graph in ("on", "auto"):      # no host synchronisation inside the step (the colour mask's count stays on the device): capturable
        try:
            from enerf_b200.graphs import GraphedStep
            graphed = GraphedStep(eager, list(batches[0]), warmup=2)
            step = graphed
        except Exception as e:  # noqa: BLE001
            if args.graph == "on":
                raise
            print(f"[bench] run_variant: graph capture failed ({type(e).__name__}: {e})", file=sys.stderr)
            torch.cuda.synchronize()
            step = eager
    for i in range(3):
        step(*batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = step(*batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    P = min(K, 5)
    _lib.profile_start()
    for i in range(P):
        eager(*batches[i % 2])
    prof = _lib.profile_stop()
    S = n_rays * T
    work = {"enerf_grid_encode_forward": ("hbm", 588.0 * S), "enerf_grid_encode_backward": ("hbm", 1100.0 * S),
            "enerf_field_density_forward": ("tensor", 2.0 * (32 * 64 + 64 * 16) * S), "enerf_field_density_backward": ("tensor", 4.0 * (32 * 64 + 64 * 16) * S)}
    kernels = {}
    for name, (calls, tot_ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        name = canon(name)
        e = {"ms_per_step": tot_ms / P, "calls_per_step": calls / P}
        if name in work and tot_ms > 0:
            bound_, amount = work[name]
            peak = peaks["hbm_gbs"] if bound_ == "hbm" else peaks["bf16_tflops_sustained"]
            ach = amount / (e["ms_per_step"] * 1e-3) / (1e9 if bound_ == "hbm" else 1e12)
            e.update(bound=bound_, achieved=ach, unit="GB/s" if bound_ == "hbm" else "TFLOP/s", peak=peak, frac=ach / peak)
        kernels[name.replace("enerf_", "")] = e
    out = {"workload": "shipped-config path (configs/*/*.txt: cuda_ray = False, ff = False): NeRFRenderer.run, 512 fixed steps/ray, nerf/network.py "
                       "topology (sigma-net 32-64-16, colour-net 31-64-64-1 on the weights > 1e-4 samples) on tcgen05, hashgrid bound 3, fp16 autocast; "
                       "4096 rays, full train step (fwd+bwd+GradScaler+Adam)",
           "launch": "cuda-graph replay" if graphed is not None else "eager",
           "rays": n_rays, "steps_per_ray": T, "samples_per_step": S, "steps": K, "ms_per_step": ms, "rays_per_s": n_rays / (ms * 1e-3),
           "msamples_per_s": S / (ms * 1e-3) / 1e6, "kernels": kernels, "kernels_sum_ms": sum(v["ms_per_step"] for v in kernels.values()),
           "mlp": "tcgen05 (enerf_field_density_* / enerf_field_color_*)" if "field_density_forward" in kernels else "nn.Linear (cuBLAS) fallback",
           "loss_finite": bool(torch.isfinite(loss))}
    if graphed is not None:
        graphed.graph = None
    del model, optimizer
    torch.cuda.empty_cache()
    return out


def extra_state_bench(loop, dev):
    """device time of one occupancy-grid refresh (update_extra_state, nerf/renderer.py:474-563): the full pass of the first 16 refreshes
    (every cell of every cascade) and the steady-state partial pass"""
    import torch
    m = loop.model
    out = {}
    for name, it in (("full", 0), ("partial", 16)):
        ts = []
        for _ in range(3):
            m.iter_density = it
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loop.refresh()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name + "_ms"] = sorted(ts)[1]
    cells = m.cascade * m.grid_size ** 3
    out.update(cells=cells, density_queries_full=cells, density_queries_partial=cells // 2, every_n_steps=REFRESH_EVERY,
               amortised_ms_per_step=out["partial_ms"] / REFRESH_EVERY, mean_density=float(m.mean_density))
    m.iter_density = 16
    return out


def shutdown(world, loops):
    """Leave without hanging: a captured graph that contains NCCL kernels must be released before the communicator goes away, and a
    communicator teardown that blocks (seen after graph capture) must not keep the job alive — a watchdog ends the process."""
    import threading
    import torch
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    import torch.distributed as dist
    threading.Timer(15.0, lambda: os._exit(0)).start()
    for l in loops:
        l.release()
    torch.cuda.synchronize()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        os._exit(0)


def our_arm(args):
    import torch

    from enerf_b200 import _lib, parallel

    global FUSE_ENCODER
    FUSE_ENCODER = args.fuse_encoder == "on"
    D = Dist()
    world, rank = D.world, D.rank
    if args.nvtx:
        _lib.enable_nvtx(True)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(D.local_rank)
    dev = torch.device("cuda", D.local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_rays, K = args.rays, max(1, args.steps)
    W = max(3, args.warmup)                         # the timing rules ask for >= 3 warm-up steps
    skip = set(x for x in args.skip.split(",") if x)
    if args.no_render:
        skip.add("render")
    exchange_cls = {"sharded": parallel.ShardedExchange, "allreduce": parallel.AllReduceExchange, "none": parallel.NoExchange}[args.exchange]

    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            m = json.load(f)
        peaks = {"hbm_gbs": m["hbm_gbs"], "bf16_tflops_sustained": m.get("bf16_tflops_sustained", m["bf16_tflops"]), "src": "MEASURED_PEAKS.json"}

    if args.only:                                   # profiling aid (ncu captures of one object's kernels); not the contract line
        K1 = max(3, min(K, 10))
        obj = {"event_step": lambda: event_step_bench(args, dev, K1), "run_variant": lambda: run_variant_bench(args, dev, K1, peaks)}[args.only]()
        print(json.dumps({"only": args.only, args.only: obj}), flush=True)
        return

    # ---------------- headline: configs[1], 4096 rays per GPU
    torch.manual_seed(0)
    model = make_ff_model(dev, BOUND)
    optimizer = make_optimizer(model, args.optimizer)
    exchange = exchange_cls(model, optimizer)
    batches, host = ray_batches(n_rays, N_BATCHES, BOUND, 100 + rank * N_BATCHES, dev, pinned=True)
    loop = TrainLoop(model, optimizer, exchange, batches, D, args.graph, dev)
    model.iter_density = 16                        # steady state: partial refreshes (the first 16 refreshes of a run are full passes)
    for i in range(W):
        if i == 0:
            loop.refresh()
        loop.step(*batches[i % N_BATCHES])

    clocks = ClockSampler(D.local_rank) if rank == 0 else None
    for i in range(2):
        loop.step(*batches[i % N_BATCHES])
    r = loop.run(K, refresh_every=REFRESH_EVERY)
    ms = r["ms"]
    clock_info = clocks.stop(*r["wall"]) if clocks else None
    value = world * n_rays * K / (ms * 1e-3)
    S = loop.samples_per_step
    samples_per_sec = D.sum(S, dev) * K / (ms * 1e-3)

    # ---------------- end to end: host buffers, H2D + D2H inside the timed region
    r2 = loop.run(K, refresh_every=REFRESH_EVERY, host=host)
    e2e = {"value": world * n_rays * K / (r2["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host[0])),
           "d2h_bytes_per_step": 4, "loss": r2["loss"]}

    # ---------------- per-kernel device time (CUDA events around every C-ABI call) -> roofline of the dominant one
    P = min(K, 10)
    _lib.profile_start()
    for i in range(P):
        loop.eager_step(*batches[i % N_BATCHES])
    prof = _lib.profile_stop()
    # algorithmic work per STEP of each entry point; both MLPs go through the same entry points
    work = {
        "enerf_grid_encode_forward": ("hbm", 588.0 * S),
        "enerf_grid_encode_backward": ("hbm", 1100.0 * S),
        "enerf_ffmlp_forward": ("tensor", 36864.0 * S),
        "enerf_ffmlp_backward": ("tensor", 73728.0 * S),
        "enerf_field_sigma_forward": ("tensor", 14336.0 * S),
        "enerf_field_color_forward": ("tensor", 22528.0 * S),
        "enerf_field_sigma_backward": ("tensor", 2 * 14336.0 * S),
        "enerf_field_color_backward": ("tensor", 2 * 22528.0 * S),
        "enerf_march_rays_train": ("hbm", 32.0 * S + 44.0 * n_rays),
        "enerf_composite_rays_train_forward": ("hbm", 16.0 * S + 32.0 * n_rays),
        "enerf_composite_rays_train_backward": ("hbm", 24.0 * S + 44.0 * n_rays),
        "enerf_sh_encode_forward": ("hbm", 44.0 * S),
    }
    kernels = {}
    for name, (calls, tot_ms) in prof.items():
        name = canon(name)
        per_step_ms = tot_ms / P
        entry = {"ms_per_step": per_step_ms, "calls_per_step": calls / P}
        if name in work and per_step_ms > 0:
            bound, amount = work[name]
            if bound == "hbm":
                entry.update(bound="hbm", achieved=amount / (per_step_ms * 1e-3) / 1e9, unit="GB/s", peak=peaks["hbm_gbs"])
            else:
                entry.update(bound="tensor", achieved=amount / (per_step_ms * 1e-3) / 1e12, unit="TFLOP/s", peak=peaks["bf16_tflops_sustained"])
            entry["frac"] = entry["achieved"] / entry["peak"]
        kernels[name.replace("enerf_", "")] = entry
    top = max((k for k in kernels if "frac" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    roofline = {"kernel": top, "bound": kernels[top]["bound"], "achieved": kernels[top]["achieved"], "peak": kernels[top]["peak"],
                "unit": kernels[top]["unit"], "frac": kernels[top]["frac"], "traffic": NCU_TRAFFIC.get(top), "traffic_source": NCU_TRAFFIC_SRC,
                "peak_source": peaks["src"],
                "ms_per_launch": kernels[top]["ms_per_step"] / kernels[top]["calls_per_step"],
                "share_of_step": kernels[top]["ms_per_step"] / (ms / K)}

    # ---------------- the other BASELINE configs (each outside the headline's timed region; a failure is reported, not fatal)
    def guarded(name, fn):
        if name in skip:
            return None
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            torch.cuda.synchronize()
            return {"error": f"{type(e).__name__}: {e}"[:400]}

    K2 = max(5, min(K, 30))
    extra_state = guarded("extra_state", lambda: extra_state_bench(loop, dev))
    render = guarded("render", lambda: render_bench(model, dev, D))
    loop.release()
    strong = guarded("strong_scaling", lambda: strong_scaling_bench(args, dev, D, exchange_cls, max(5, min(K, 20))))
    event_step = run_variant = gpu_bar = None
    if world == 1:
        event_step = guarded("event_step", lambda: event_step_bench(args, dev, K2))
        run_variant = guarded("run_variant", lambda: run_variant_bench(args, dev, max(3, min(K, 10)), peaks))
        if "gpu_bar" not in skip:
            from oracle import ref
            if all(ref.available(nm) for nm in ref.NAMES):
                from oracle import gpu_bar as gb
                gpu_bar = guarded("gpu_bar", lambda: gb.measure(n_rays=n_rays, iters=3, bound=BOUND))
            else:
                gpu_bar = {"unavailable": "oracle/_ref/*.so (the reference's own CUDA build) is not on this box"}

    if rank != 0:
        shutdown(world, [loop])
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_reference
        rc = cpu_reference.time_train_steps(n_rays=args.cpu_rays, num_steps=512, bound=BOUND, out_dim_color=1, steps=3, warmup=1)
        cpu_baseline = {"value": rc["rays_per_s"], "unit": UNIT, "cores": rc["cores"], "kind": "port", "sample": rc["sample"],
                        "samples_per_sec": rc["rays_per_s"] * 512,
                        "pinned_by": "tests/test_cpu_port_pinned.py (image, depth, gradients vs the reference's own classes, 1e-6)",
                        "like_for_like": "run_variant (same topology, same 512 steps/ray, on the GPU)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "samples_per_sec": samples_per_sec,
            "config": workload_config(n_rays, {"samples_per_step_per_gpu": S, "sample_buffer_rows": loop.M, "ray_batches_cycled": N_BATCHES,
                                               "occupancy_refresh": f"update_extra_state (partial pass) every {REFRESH_EVERY} steps inside the timed region; "
                                                                    "its result is discarded and the analytic grid restored so the marched workload stays fixed",
                                               "parallelism": f"dp{world} (ray-sharded, NCCL gradient exchange: {exchange.name})",
                                               "launch": "cuda-graph replay" if loop.step is not loop.eager_step else "eager",
                                               "optimizer": "enerf_b200.optim.FusedAdam" if args.optimizer == "fused" else "torch.optim.Adam(fused)",
                                               "l2": "per-step working set (samples x ~1.7 KB of activations + 52 MB grad table) is >> 126 MB L2; no explicit flush"}),
            "e2e": e2e, "gpu_launches": int(r["launches"]), "clocks": clock_info, "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "extra_state": extra_state, "render": render, "strong_scaling": strong, "event_step": event_step,
            "run_variant": run_variant, "gpu_bar": gpu_bar, "final_loss": float(r2["loss"]), "host_enqueue_ms_per_step": r["enqueue_ms"]}
    if W != args.warmup:
        line["warmup_requested"] = args.warmup
    print(json.dumps(line), flush=True)
    shutdown(world, [loop])


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        our_arm(a)
