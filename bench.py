#!/usr/bin/env python
"""bench.py — latent tokens/s of the DiT train step (fwd + bwd + loss + fused AdamW) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload debug-8k|debug-512|B|XL] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the step shards by data-parallel batch (weak
scaling: fixed per-rank batch), parameters/gradients are sharded with our own all-gather / reduce-scatter.
Rank 0 prints ONE JSON line.

The headline workload is BASELINE.json configs[1] (`debug-8k`: the run_debug.sh model on a dataset-shaped
[16,16,64,64] latent, per-rank batch 2).  The same line carries, under "workloads", a short measurement of the other
BASELINE configs: `B` (configs[2], DiT-B on 16-frame 256x256 latents), `XL` (configs[3], DiT-XL on 32-frame 512x512
latents) and — at N = 1 — `sampling` (configs[4]: the sample.py model, batch 8, one denoising step = two forwards).

`--impl reference` times the reference's CPU eager step (the oracle restatement of /root/reference/model.py +
train.py, fp32, all host threads) on the SAME workload (model, batch and latent shape) with few steps.

The step is captured once and replayed as one CUDA graph (train.GraphedTrainStep; at N > 1 the graph contains the
NCCL all-gathers / reduce-scatters on their side stream) or issued kernel by kernel from Python with programmatic
dependent launch.  By default an untimed 2-step probe of each (`issue_mode_probe`) picks the faster; `--graph` /
`--eager` force one.  `value` times K steps with the batch resident in HBM; `e2e` times K steps that each copy the
batch from pinned host memory (vds_b200.data.DevicePrefetcher, copy of step i+1 overlapped with step i) and read the
loss back; `roofline` is the self-attention backward kernel timed by CUDA events around every one of its launches inside
the timed steps (event-record nodes when graphed).
"""
import argparse
import gc
import itertools
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (hidden, depth, heads, per-rank batch, latent T,H,W)   — SURVEY.md §8 shapes
    "debug-8k": (512, 24, 4, 2, (16, 64, 64)),   # S_dbg: run_debug.sh model on a [16,16,64,64] latent, L = 8208
    "debug-512": (512, 24, 4, 2, (4, 32, 32)),   # S_small
    "B": (768, 12, 6, 8, (2, 32, 32)),           # S_B
    "XL": (1152, 28, 9, 2, (4, 64, 64)),         # S_XL
}
SAMPLING = (2048, 24, 16, 8, (16, 64, 64))       # S_smp: sample.py:43-53 model, batch 8 of [16,16,64,64] latents
LC, DC = 512, 4096
METRIC = "latent tokens/s per train step (fwd+bwd+AdamW)"


def model_cfg(hidden, depth, heads):
    return dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=hidden, depth=depth, num_heads=heads,
                mlp_ratio=4.0, cross_attn_input_size=DC, residual_v=True, train_bias_and_rms=False, use_rope=True)


def flops_fwd(h, depth, B, N, Lc=LC, Dc=DC):
    Lr = N + 16
    blk = 28 * B * Lr * h * h + 4 * B * Lr * Lr * h + 4 * B * Lr * Lc * h + 4 * B * Lc * Dc * h + 18 * B * h * h
    return depth * blk + 4 * B * N * 128 * h + 20 * B * h * h


def flops_fwd_bwd(h, depth, B, N, Lc=LC, Dc=DC):
    """Algorithmic FLOPs of one train step (SURVEY.md §8d / BASELINE.md §4)."""
    fwd = flops_fwd(h, depth, B, N, Lc, Dc)
    bwd = 2 * fwd - depth * 4 * B * Lc * Dc * h - 2 * B * N * 128 * h
    return fwd + bwd


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measured_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the last `ncu --set full` capture
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep).  None when this kernel /
    workload has no capture: the number is a profiler reading, it cannot be re-measured inside a timed run."""
    try:
        ent = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{kernel}@{workload}")
        if ent:
            return float(ent["dram_bytes_per_launch"]), ent.get("source", "profiles/ncu_traffic.json")
    except Exception:
        pass
    return None, "no ncu --set full capture of this kernel on this workload in profiles/ncu_traffic.json"


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.index)], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.03)   # nvidia-smi itself takes ~40 ms: ~14 samples per second of timed region

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def workload_config(name, gpus, B=None, thw=None):
    hidden, depth, heads, B0, thw0 = WORKLOADS[name]
    B = B0 if B is None else B
    thw = thw0 if thw is None else thw
    N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)
    tag = name if (B == B0 and tuple(thw) == tuple(thw0)) else f"{name} model, bounded sample"
    return {"workload": f"{tag}: DiT h={hidden} depth={depth} heads={heads}x128, per-rank batch {B} of "
                        f"[16,{thw[0]},{thw[1]},{thw[2]}] latents ({N} tokens + 16 registers / sample), "
                        f"context [B,{LC},{DC}], fwd+bwd+loss+AdamW",
            "global_batch": B * gpus, "tokens_per_sample": N, "parallelism": f"dp{gpus} (own param-shard AG / grad RS)",
            "l2_policy": "256 MiB L2 flush write between timed steps; per-step working set (GBs of activations) >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(workload, steps, warmup, sample_B=None, sample_thw=None):
    """The reference's eager CPU train step (fwd + bwd + AdamW), fp32, all host threads: the workload's model (same
    width / depth / heads) on a [sample_B,16,T,H,W] latent batch (default: the workload's own batch and latent shape).
    Executed through the oracle restatement (the GPU box has no /root/reference)."""
    import torch
    from oracle import dit_oracle as O
    hidden, depth, heads, B0, thw0 = WORKLOADS[workload]
    sample_B = B0 if sample_B is None else sample_B
    sample_thw = thw0 if sample_thw is None else sample_thw
    cfg = model_cfg(hidden, depth, heads)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    torch.manual_seed(0)
    m = DiT(**cfg)
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    sd = O.randomise_zero_init({n: p.detach().clone() for n, p in m.named_parameters()}, seed=1)
    del m
    P = {n: (v * 0.1 if v.dim() == 2 else v).clone().requires_grad_(True) for n, v in sd.items()}
    settings = O.mup_settings(shapes, 2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
    groups = {}
    for n, p in P.items():
        groups.setdefault(settings[n], []).append(p)
    opt = torch.optim.AdamW([{"params": ps, "lr": lr, "weight_decay": wd} for (lr, wd), ps in groups.items()],
                            betas=(0.95, 0.99))
    latent, noise, context, t = [a.float() for a in O.make_inputs(cfg, sample_B, sample_thw, LC, DC, 1234)]
    N = (sample_thw[0] // 2) * (sample_thw[1] // 2) * (sample_thw[2] // 2)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch.manual_seed(i)
        opt.zero_grad()
        loss, _ = O.train_loss(P, cfg, latent, context, t, noise)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    same = sample_B == B0 and tuple(sample_thw) == tuple(thw0)
    return {"value": sample_B * N / mean, "unit": "latent tokens/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed step(s) (+{warmup} warm-up) of the {workload} model (h={hidden}, depth={depth}) on "
                      + ("the workload's own batch: " if same else "a bounded sample: ")
                      + f"[{sample_B},16,{sample_thw[0]},{sample_thw[1]},{sample_thw[2]}] latents "
                      f"({sample_B * N} tokens/step, L = {N + 16}), fp32 eager torch CPU (oracle restatement of the "
                      f"reference step), {mean:.2f} s/step",
            "s_per_step": mean, "B": sample_B, "thw": tuple(sample_thw)}


def run_reference(args):
    """Reference arm: the SAME workload (model, per-rank batch, latent shape) as our arm, on the host cores.  One
    step of debug-8k is ~30 TFLOP in fp32 on the CPU (tens of seconds), so K is clamped to 2 timed steps after 1
    warm-up; `--ref-sample B,T,H,W` bounds the batch further (the config line then says so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sB, sthw = None, None
    if args.ref_sample:
        f = [int(x) for x in args.ref_sample.split(",")]
        sB, sthw = f[0], tuple(f[1:4])
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    base = cpu_reference_step(args.workload, steps=steps, warmup=warm, sample_B=sB, sample_thw=sthw)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"],
            "unit": "latent tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": base["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, 1, base["B"], base["thw"]),
            "same_workload_as_ours": args.ref_sample is None,
            "note": "one host process on rank 0 regardless of --gpus (the reference CPU path is not sharded); value is "
                    "tokens/s of that one process",
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "latent tokens/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
class TrainRun:
    """Model + optimizer + synthetic batch of one training workload on this rank, and the two ways of issuing a step."""

    def __init__(self, name, dev, world, rank, gpu_init=False):
        import torch
        import vds_b200  # noqa: F401
        from vds_b200.model import DiT, apply_fsdp
        from vds_b200.optim import FusedAdamW
        from oracle import dit_oracle as O  # only for the synthetic-input recipe (+ the cpu_baseline leg)
        self.name, self.dev, self.world, self.rank = name, dev, world, rank
        hidden, depth, heads, B, thw = WORKLOADS[name]
        self.hidden, self.depth, self.heads, self.B, self.thw = hidden, depth, heads, B, thw
        self.N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)
        cfg = model_cfg(hidden, depth, heads)
        if not gpu_init:
            torch.manual_seed(0)  # identical init on every rank (the reference leaves this unseeded: SURVEY.md §2.3)
            model = DiT(**cfg)
            with torch.no_grad():
                sd = O.randomise_zero_init({n: p.detach().clone() for n, p in model.named_parameters()}, seed=1)
                for n, p in model.named_parameters():
                    p.copy_(sd[n] * 0.1 if p.dim() == 2 else sd[n])  # train.py:247-251
            model = model.to(dev)
        else:
            # extra workloads: same recipe drawn on the device (seeded, identical on every rank) — CPU init of a
            # 1.1 B-parameter model would dominate the bench's wall time
            torch.manual_seed(0)
            torch.cuda.manual_seed(0)
            with torch.device(dev):
                model = DiT(**cfg)
            with torch.no_grad():
                for n, p in model.named_parameters():
                    if any(z in n for z in O.ZERO_INIT):
                        p.normal_(0.0, 0.02)
                    if p.dim() == 2:
                        p.mul_(0.1)
        self.model = apply_fsdp(model, torch.bfloat16, torch.float32)
        groups, _ = self.model.get_mup_setup(2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
        self.opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=self.model._flat)
        lat, noi, ctx, t = O.make_inputs(cfg, B, thw, LC, DC, 1234 + rank)
        self.latent_h, self.noise_h, self.context_h = lat.pin_memory(), noi.pin_memory(), ctx.pin_memory()
        self.latent, self.noise, self.context, self.t = (self.latent_h.to(dev), self.noise_h.to(dev),
                                                         self.context_h.to(dev), t.to(dev))
        self.stepper = None
        self.graph_error = None

    def try_graph(self, warmup=2):
        from vds_b200 import train
        self.stepper = train.GraphedTrainStep(self.model, self.opt, self.latent.shape, self.context.shape,
                                              device=self.dev, warmup=warmup)

    def step(self, i, lat=None, ctx=None):
        import torch
        from vds_b200 import train
        lat = self.latent if lat is None else lat
        ctx = self.context if ctx is None else ctx
        torch.manual_seed(i)  # RoPE offset draws (model.py:224-226)
        if self.stepper is not None:
            return self.stepper(lat, ctx, self.t, self.noise)
        self.opt.zero_grad()
        loss, _ = train.forward(self.model, lat, ctx, t=self.t, noise=self.noise)
        loss.backward()
        self.opt.step()
        return loss

    def close(self):
        import torch
        if self.stepper is not None:
            self.stepper.close()
        self.stepper = None
        hook = getattr(self.model, "_post_step_hook", None)
        if hook is not None:
            hook.remove()                 # the global optimizer hook of apply_fsdp keeps the flat buffers alive
        self.model._flat = None
        self.model = self.opt = None
        gc.collect()
        torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as dist
    import vds_b200  # noqa: F401
    from vds_b200 import lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn(i), each bracketed by CUDA events on the launch stream, L2 flushed in between; mean ms."""
        evs = []
        barrier()
        for i in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs) / n

    def settle(run, warmup, allow_graph, force_graph, profile):
        """Warm-up + choice of the issue mode (graph replay / Python issue).  Returns the probe record."""
        if profile:
            ops.PROFILE["attn_bwd_self"] = []
        if allow_graph:
            run.try_graph(warmup=2)
        try:
            for i in range(max(warmup, 3)):
                run.step(i)
            barrier()
            if run.stepper is not None and run.stepper.graph is None:
                for i in range(3):
                    run.step(100 + i)
                barrier()
        except Exception as ex:                      # a failed graph capture must not take the bench line down
            if run.stepper is None:
                raise
            run.graph_error = f"{type(ex).__name__}: {ex}"
            sys.stderr.write(f"[{run.name}] CUDA-graph step unavailable ({run.graph_error}); issuing from Python\n")
            run.stepper = None
            ops.PROFILE.pop("attn_bwd_self", None)
            torch.cuda.synchronize()
            for i in range(max(warmup, 3)):
                run.step(i)
            barrier()
        pick = None
        if run.stepper is not None and not force_graph:
            def probe(use_graph, n=2):
                saved = run.stepper
                if not use_graph:
                    run.stepper = None
                try:
                    run.step(5000)
                    barrier()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(n):
                        run.step(5001 + i)
                    b.record()
                    barrier()
                    return a.elapsed_time(b) / n
                finally:
                    run.stepper = saved
            keep = ops.PROFILE.pop("attn_bwd_self", None)     # the eager probe must not append to the graph's events
            ms_g, ms_e = probe(True), probe(False)
            if keep is not None:
                ops.PROFILE["attn_bwd_self"] = keep
            tm = torch.tensor([ms_g, ms_e], device=dev, dtype=torch.float64)
            if world > 1:                                      # every rank must take the same decision
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_g, ms_e = tm.tolist()
            pick = {"graph_ms": ms_g, "eager_ms": ms_e}
            if ms_e < 0.98 * ms_g:
                run.stepper = None
        return pick

    # =========================================================================== headline workload
    run = TrainRun(args.workload, dev, world, rank)
    hidden, depth, B, N = run.hidden, run.depth, run.B, run.N
    pick = settle(run, args.warmup, allow_graph=not args.eager, force_graph=args.graph, profile=True)
    graph_prof = None
    if run.stepper is not None:
        graph_prof = ops.PROFILE.get("attn_bwd_self", [])[-depth:]   # the event nodes recorded during the capture
    ops.PROFILE.pop("attn_bwd_self", None)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM; dominant kernel timed live with events on the launch stream
    if run.stepper is None:
        ops.PROFILE["attn_bwd_self"] = []
    launches0 = lib.launch_count()
    step_ms = timed(lambda i: run.step(args.warmup + i), args.steps)
    launches = lib.launch_count() - launches0
    prof = ops.PROFILE.pop("attn_bwd_self", [])
    if graph_prof is not None:
        prof = graph_prof                    # re-recorded by every replay: these are the last timed step's launches
    kern_ms = sum(a.elapsed_time(b) for a, b in prof) / max(1, len(prof))
    ops.PROFILE.clear()

    # ---- timed region 2 (e2e): host buffers -> H2D every step, loss read back (D2H) every step.  The copies go
    # through the repo's own input pipeline (vds_b200.data.DevicePrefetcher): batch i+1 is copied from pinned memory on
    # a copy stream while step i computes.
    from vds_b200.data import DevicePrefetcher
    barrier()
    host_batches = itertools.repeat({"latent": run.latent_h, "context": run.context_h, "noise": run.noise_h})
    # one untimed pipeline warm-up step (copy stream, staging allocations, first two copies in flight): the timed steps
    # then run in the pipeline's steady state — each issues exactly one H2D batch copy and waits for one
    pf = DevicePrefetcher(host_batches, device=dev, depth=2)
    bt = next(pf)
    run.noise.copy_(bt["noise"], non_blocking=True)
    run.step(args.warmup + args.steps, bt["latent"], bt["context"]).item()
    last = {}

    def e2e_step(i):
        bt = next(pf)
        run.noise.copy_(bt["noise"], non_blocking=True)
        loss = run.step(args.warmup + args.steps + 1 + i, bt["latent"], bt["context"])
        last["loss"] = loss.item()               # D2H read of the step's result inside the timed region

    pipelined = run.stepper is not None and run.stepper.graph is not None

    def stage_next(i):                           # host side of step i: wait for its H2D copy, fill the graph's inputs
        bt = next(pf)
        run.noise.copy_(bt["noise"], non_blocking=True)
        torch.manual_seed(args.warmup + args.steps + 1 + i)   # RoPE offset draws, as TrainRun.step
        run.stepper.stage(bt["latent"], bt["context"], run.t, run.noise)

    def e2e_step_pipelined(i):
        # software pipeline of a training loop on a captured step: launch step i (staged one iteration ago), prepare
        # step i+1 on the host while the GPU computes (its copies queue behind the replay), then read step i's loss
        run.stepper.replay()                     # two graph launches: forward + loss | backward + optimizer
        stage_next(i + 1)
        last["loss"] = run.stepper.loss_value()  # D2H read of THIS step's loss (final once the forward graph is done)

    if pipelined:
        stage_next(0)
        e2e_ms = timed(e2e_step_pipelined, args.steps)
    else:
        e2e_ms = timed(e2e_step, args.steps)
    h2d_per_step = sum(v.numel() * v.element_size() for v in (run.latent_h, run.context_h, run.noise_h))
    sampler.stop_flag = True

    # ---- extra: the other issue mode of the same step (not the headline)
    other = None
    if not args.no_extras and pick is not None:
        other = {"graph_replay_ms": pick["graph_ms"], "python_issue_ms": pick["eager_ms"],
                 "note": "untimed-probe numbers (2 steps each) of the two ways of issuing the same kernels"}

    # ---- cross-rank parity signal (N > 1): after the timed steps every rank must hold the same parameters.  Each rank
    # owns 1/N of the fp32 master; the gathered bf16 compute copy (what the next forward reads) is compared by checksum.
    param_check = None
    if world > 1:
        full = run.model._flat.full16
        torch.cuda.synchronize()
        cs = torch.stack([full.float().sum().double(), full.float().abs().sum().double(),
                          full[:: max(1, full.numel() // 65536)].float().square().sum().double()])
        allc = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allc, cs)
        same = all(torch.equal(allc[0], c) for c in allc)
        finite = bool(torch.isfinite(cs).all().item())
        param_check = {"ranks_hold_identical_params": bool(same), "finite": finite,
                       "how": "sum / abs-sum / strided square-sum of every rank's gathered bf16 parameter buffer "
                              "after the timed steps, all-gathered and compared bit for bit"}

    tm = torch.tensor([step_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = tm.tolist()
    used_graph = run.stepper is not None
    launches_per_step = run.stepper.launches_per_step if used_graph else launches / max(1, args.steps)
    graph_error = run.graph_error
    run.close()
    del run, pf

    # =========================================================================== other BASELINE configs (short)
    pk, pk_src = peaks()
    peak = pk["bf16_tflops_sustained"]
    extras = {}
    if not args.no_extras:
        xsteps = max(3, min(args.steps, 5))
        for name in ("B", "XL"):
            if name == args.workload:
                continue
            try:
                r = TrainRun(name, dev, world, rank, gpu_init=True)
                pk2 = settle(r, 3, allow_graph=not args.eager, force_graph=args.graph, profile=False)
                ms = timed(lambda i: r.step(10 + i), xsteps)
                t2 = torch.tensor([ms], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                ms = t2.item()
                fl = flops_fwd_bwd(r.hidden, r.depth, r.B, r.N)
                extras[name] = {"config": workload_config(name, world)["workload"], "ms_per_step": ms, "steps": xsteps,
                                "value": world * r.B * r.N / (ms * 1e-3), "unit": "latent tokens/s",
                                "step_tflops_per_gpu": fl / (ms * 1e-3) / 1e12,
                                "step_frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / peak,
                                "cuda_graph": r.stepper is not None, "issue_mode_probe": pk2,
                                "graph_error": r.graph_error}
                r.close()
                del r
            except Exception as ex:
                extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
                gc.collect()
                torch.cuda.empty_cache()
        if world == 1:
            try:
                extras["sampling"] = sampling_workload(dev, timed, xsteps, peak)
            except Exception as ex:
                extras["sampling"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        Lr = N + 16
        kern_flops = 8.0 * B * Lr * Lr * hidden  # self-attention backward, algorithmic 2 x forward
        achieved = kern_flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
        step_flops = flops_fwd_bwd(hidden, depth, B, N)
        traffic, traffic_src = measured_traffic("attn_bwd", args.workload)
        line = {
            "metric": METRIC, "value": world * B * N / (step_ms * 1e-3),
            "unit": "latent tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args.workload, world),
            "step_tflops": step_flops / (step_ms * 1e-3) / 1e12,
            "step_frac_of_bf16_peak": step_flops / (step_ms * 1e-3) / 1e12 / peak,
            "roofline": {"kernel": "self-attention backward (vds_attn_bwd: main + tail-balancing launches, L=%d)" % Lr,
                         "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_src,
                         "peak_source": pk_src + " (bf16_tflops_sustained)",
                         "kernel_ms": kern_ms, "launches_timed": len(prof),
                         "kernel_share_of_step": kern_ms * depth / step_ms},
            "e2e": {"value": world * B * N / (e2e_ms * 1e-3), "unit": "latent tokens/s", "ms_per_step": e2e_ms,
                    "pipeline": ("step i+1 staged on the host (H2D copy waited for, inputs / RoPE offsets / optimizer scalars "
                                 "copied into the graph's buffers) behind the replay of step i; the step is two graphs "
                                 "(forward + loss | backward + optimizer) and the loss of step i is read back over a side "
                                 "stream as soon as the first has run, before step i+1 is launched behind the second"
                                 ) if pipelined else "stage, launch, read back, in sequence",
                    "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": 4, "last_loss": last.get("loss")},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "cuda_graph": used_graph, "graph_error": graph_error, "issue_mode_probe": pick, "clocks": sampler.summary(),
            "issue_modes": other, "param_check": param_check, "workloads": extras,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload: ONE sample of the per-rank batch at the full latent shape (same
            # sequence length, same per-token work), one step without warm-up — ~20-40 s of host time
            _, _, _, _, thw = WORKLOADS[args.workload]
            cb = cpu_reference_step(args.workload, steps=1, warmup=0, sample_B=1, sample_thw=thw)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        if os.environ.get("VDS_BENCH_OUT"):
            with open(os.environ["VDS_BENCH_OUT"], "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sampling_workload(dev, timed, steps, peak):
    """BASELINE configs[4]: the sampling/sample.py model (h=2048 x 24, 16 heads, bf16 module incl. bf16 RoPE tables),
    batch 8 of [16,16,64,64] latents, one denoising step = conditional + unconditional forward + CFG + fp32 Euler
    update (sample.py:122-146), replayed as one CUDA graph (sampling.GraphedDenoiser).  Cosmos decoder / T5 excluded."""
    import torch
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    from vds_b200.sampling.sample import GraphedDenoiser
    from oracle import dit_oracle as O
    hidden, depth, heads, B, thw = SAMPLING
    cfg = model_cfg(hidden, depth, heads)
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    with torch.device(dev):
        model = DiT(**cfg)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if any(z in n for z in O.ZERO_INIT):
                p.normal_(0.0, 0.02)
            if p.dim() == 2:
                p.mul_(0.1)
    model = model.to(dev, torch.bfloat16).eval()          # sample.py:63
    N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)
    shape = (B, 16) + tuple(thw)
    g = torch.Generator(device=dev).manual_seed(42)
    prompt = torch.randn((B, LC, DC), device=dev, dtype=torch.bfloat16, generator=g)
    latents = torch.randn(shape, device=dev, dtype=torch.bfloat16, generator=g)
    den = GraphedDenoiser(model, prompt, shape, cfg_scale=6.0, device=dev)
    with torch.no_grad():
        den.run(latents, inference_steps=3)                # eager step + capture + one replay
        n_steps = 50

        def one(i):
            torch.manual_seed(i)
            k = n_steps - i
            t = den_shift(k / n_steps)
            den._refresh(t, t - den_shift((k - 1) / n_steps))
            den.graph.replay()
        ms = timed(one, steps)
    fl = 2 * flops_fwd(hidden, depth, B, N)
    out = {"config": f"sampling: DiT h={hidden} depth={depth} heads={heads}x128 (bf16 module), batch {B} of "
                     f"[16,{thw[0]},{thw[1]},{thw[2]}] latents, one denoising step = cond + uncond forward + CFG + Euler "
                     f"(context_kv of the prompt cached), CUDA graph replay",
           "ms_per_step": ms, "steps": steps, "value": B * N / (ms * 1e-3), "unit": "latent tokens/s per denoising step",
           "step_tflops_per_gpu": fl / (ms * 1e-3) / 1e12, "step_frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / peak,
           "flops_note": "2 forwards incl. the context_kv GEMMs the cache skips (algorithmic, SURVEY §8d)",
           "finite": bool(torch.isfinite(den.acc).all().item())}
    del den, model
    gc.collect()
    torch.cuda.empty_cache()
    return out


def den_shift(t, alpha=8.0):
    return t * alpha / (1 + (alpha - 1) * t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="debug-8k", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="force the CUDA-graph step (default: an untimed 2-step probe "
                    "picks graph replay or Python issue, whichever is faster for the workload)")
    ap.add_argument("--eager", action="store_true", help="issue the kernels from Python instead of replaying a graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the short B / XL / sampling measurements")
    ap.add_argument("--ref-sample", default="", help="reference arm only: B,T,H,W of a bounded sample instead of the "
                    "workload's own batch (the printed config then names the sample)")
    ap.add_argument("--depth", type=int, default=0, help="profiling only: override the model depth (NOT a bench line)")
    args = ap.parse_args()
    if args.depth > 0:
        hidden, depth, heads, B, thw = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (hidden, args.depth, heads, B, thw)
    if args.impl == "reference":
        run_reference(args)
    else:
        assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
        run_ours(args)


if __name__ == "__main__":
    main()
