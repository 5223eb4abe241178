#!/usr/bin/env python
"""bench.py — latent tokens/s of the DiT train step (fwd + bwd + loss + fused AdamW) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload debug-8k|debug-512|B|XL] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the step shards by data-parallel batch (weak
scaling: fixed per-rank batch), parameters/gradients are sharded with our own all-gather / reduce-scatter.
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's CPU eager step (the oracle
restatement of /root/reference/model.py + train.py, fp32, all host threads) on a bounded sample.

At world size 1 the same kernels can be issued two ways: replayed as one CUDA graph (train.GraphedTrainStep) or
launched one by one from Python with programmatic dependent launch.  By default an untimed 2-step probe of each
(`issue_mode_probe` in the JSON line) picks the faster one for the workload; `--graph` / `--eager` force one.  `value` times K steps with the batch resident in
HBM; `e2e` times K steps that each copy the batch from pinned host memory (vds_b200.data.DevicePrefetcher, copy of
step i+1 overlapped with step i) and read the loss back; `roofline` is the self-attention backward kernel timed by
CUDA events around every one of its launches inside the timed steps (event-record nodes when graphed).
"""
import argparse
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
LC, DC = 512, 4096


def model_cfg(hidden, depth, heads):
    return dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=hidden, depth=depth, num_heads=heads,
                mlp_ratio=4.0, cross_attn_input_size=DC, residual_v=True, train_bias_and_rms=False, use_rope=True)


def flops_fwd_bwd(h, depth, B, N, Lc=LC, Dc=DC):
    """Algorithmic FLOPs of one train step (SURVEY.md §8d / BASELINE.md §4)."""
    Lr = N + 16
    blk = 28 * B * Lr * h * h + 4 * B * Lr * Lr * h + 4 * B * Lr * Lc * h + 4 * B * Lc * Dc * h + 18 * B * h * h
    fwd = depth * blk + 4 * B * N * 128 * h + 20 * B * h * h
    bwd = 2 * fwd - depth * 4 * B * Lc * Dc * h - 2 * B * N * 128 * h
    return fwd + bwd


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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
            time.sleep(0.15)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(workload, steps, warmup, sample_B=1, sample_thw=(4, 32, 32)):
    """The reference's eager CPU train step (fwd + bwd + AdamW), fp32, all host threads, on a bounded sample of
    the workload's model: same DiT (width/depth/heads), a [sample_B,16,4,32,32] latent batch (N = 512 tokens
    per sample).  Executed through the oracle restatement (the GPU box has no /root/reference)."""
    import torch
    from oracle import dit_oracle as O
    hidden, depth, heads, _, _ = WORKLOADS[workload]
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
    return {"value": sample_B * N / mean, "unit": "latent tokens/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} step(s) of the {workload} model (h={hidden}, depth={depth}) on a "
                      f"[{sample_B},16,{sample_thw[0]},{sample_thw[1]},{sample_thw[2]}] latent batch "
                      f"({sample_B * N} tokens/step), fp32 eager torch CPU, {mean:.2f} s/step",
            "s_per_step": mean}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hidden, depth, heads, B, thw = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 3))
    base = cpu_reference_step(args.workload, steps=steps, warmup=1 if args.warmup > 0 else 0)
    line = {"impl": "reference", "metric": "latent tokens/s per train step (fwd+bwd+AdamW)", "value": base["value"],
            "unit": "latent tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1 if args.warmup > 0 else 0,
            "ms_per_step": base["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "latent tokens/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, gpus):
    hidden, depth, heads, B, thw = WORKLOADS[name]
    N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)
    return {"workload": f"{name}: DiT h={hidden} depth={depth} heads={heads}x128, per-rank batch {B} of "
                        f"[16,{thw[0]},{thw[1]},{thw[2]}] latents ({N} tokens + 16 registers / sample), "
                        f"context [B,{LC},{DC}], fwd+bwd+loss+AdamW",
            "global_batch": B * gpus, "tokens_per_sample": N, "parallelism": f"dp{gpus} (own param-shard AG / grad RS)",
            "l2_policy": "256 MiB L2 flush write between timed steps; per-step working set (~10 GB activations) >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vds_b200  # noqa: F401
    from vds_b200 import lib, ops, train
    from vds_b200.model import DiT, apply_fsdp
    from vds_b200.optim import FusedAdamW
    from oracle import dit_oracle as O  # only for synthetic-input recipe + cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hidden, depth, heads, B, thw = WORKLOADS[args.workload]
    cfg = model_cfg(hidden, depth, heads)
    N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)

    torch.manual_seed(0)  # identical init on every rank (the reference leaves this unseeded: SURVEY.md §2.3)
    model = DiT(**cfg)
    with torch.no_grad():
        sd = O.randomise_zero_init({n: p.detach().clone() for n, p in model.named_parameters()}, seed=1)
        for n, p in model.named_parameters():
            p.copy_(sd[n] * 0.1 if p.dim() == 2 else sd[n])  # train.py:247-251
    model = model.to(dev)
    model = apply_fsdp(model, torch.bfloat16, torch.float32)
    groups, _ = model.get_mup_setup(2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
    opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)

    latent_h, noise_h, context_h, t_h = O.make_inputs(cfg, B, thw, LC, DC, 1234 + rank)
    latent_h, noise_h, context_h = latent_h.pin_memory(), noise_h.pin_memory(), context_h.pin_memory()
    latent, noise, context, t = latent_h.to(dev), noise_h.to(dev), context_h.to(dev), t_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # World size 1: the step is captured once and replayed as ONE CUDA graph (train.GraphedTrainStep, the repo's public
    # API for a launch-bound step); --eager issues the same ~1200 kernels from Python instead.  Multi-GPU runs are
    # eager (the per-block NCCL collectives are issued from Python).  The dominant kernel is timed live in both modes:
    # its events are event-record nodes inside the graph (ops.attn_bwd).
    stepper = None
    if world == 1 and not args.eager:
        ops.PROFILE["attn_bwd_self"] = []
        stepper = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=dev, warmup=2)

    def step(i, lat, ctx):
        torch.manual_seed(i)  # RoPE offset draws (model.py:224-226)
        if stepper is not None:
            return stepper(lat, ctx, t, noise)
        opt.zero_grad()
        loss, _ = train.forward(model, lat, ctx, t=t, noise=noise)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        for i in range(args.warmup):
            step(i, latent, context)
        barrier()
    except Exception as ex:                      # a failed graph capture must not take the bench line down
        if stepper is None:
            raise
        sys.stderr.write(f"CUDA-graph step unavailable ({type(ex).__name__}: {ex}); issuing the kernels from Python\n")
        stepper, opt.hyper_dev = None, None
        ops.PROFILE.clear()
        torch.cuda.synchronize()
        for i in range(args.warmup):
            step(i, latent, context)
        barrier()
    graph_prof = None
    mode_pick = None
    if stepper is not None:
        if stepper.graph is None:            # fewer warm-up steps than the stepper's own eager warm-up: capture now
            for i in range(3):
                step(args.warmup + 100 + i, latent, context)
            barrier()
        graph_prof = ops.PROFILE.get("attn_bwd_self", [])[-depth:]   # the event nodes recorded during the capture
        if not args.graph:
            # Untimed pick between the two ways of issuing the SAME kernels: graph replay wins when the host cannot keep
            # up (small workloads), Python issue + programmatic dependent launch wins when every kernel is long (XL).
            def probe(use_graph, n=2):
                nonlocal stepper
                saved, hyper = stepper, getattr(opt, "hyper_dev", None)
                if not use_graph:
                    stepper, opt.hyper_dev = None, None
                try:
                    step(5000, latent, context)
                    barrier()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(n):
                        step(5001 + i, latent, context)
                    b.record()
                    barrier()
                    return a.elapsed_time(b) / n
                finally:
                    stepper, opt.hyper_dev = saved, hyper
            ms_g, ms_e = probe(True), probe(False)
            mode_pick = {"graph_ms": ms_g, "eager_ms": ms_e}
            if ms_e < 0.98 * ms_g:
                stepper, graph_prof, opt.hyper_dev = None, None, None
                ops.PROFILE.clear()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM; dominant kernel timed live with events on the launch stream
    if stepper is None:
        ops.PROFILE["attn_bwd_self"] = []
    launches0 = lib.launch_count()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(args.warmup + i, latent, context)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = lib.launch_count() - launches0
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
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
    e2e_evs = []
    import itertools
    host_batches = itertools.repeat({"latent": latent_h, "context": context_h, "noise": noise_h})
    # one untimed pipeline warm-up step (copy stream, staging allocations, first two copies in flight): the timed steps
    # then run in the pipeline's steady state — each issues exactly one H2D batch copy and waits for one
    pf = DevicePrefetcher(host_batches, device=dev, depth=2)
    bt = next(pf)
    noise.copy_(bt["noise"], non_blocking=True)
    step(args.warmup + args.steps, bt["latent"], bt["context"]).item()
    barrier()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tdbg = [time.perf_counter()]
        bt = next(pf)
        tdbg.append(time.perf_counter())
        noise.copy_(bt["noise"], non_blocking=True)
        loss = step(args.warmup + args.steps + i, bt["latent"], bt["context"])
        tdbg.append(time.perf_counter())
        loss_host = loss.item()
        tdbg.append(time.perf_counter())
        if os.environ.get("VDS_BENCH_DEBUG"):
            sys.stderr.write("e2e host ms: prefetch %.2f step-issue %.2f loss.item %.2f\n" % tuple(
                1e3 * (tdbg[k + 1] - tdbg[k]) for k in range(3)))
        e1.record()
        e2e_evs.append((e0, e1))
    barrier()
    h2d_per_step = sum(v.numel() * v.element_size() for v in (latent_h, context_h, noise_h))   # one batch per step
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs) / args.steps
    sampler.stop_flag = True

    # ---- extra (world size 1, eager runs only): the same step replayed from a CUDA graph (train.GraphedTrainStep)
    eager_ms = None
    if stepper is not None and not args.no_graph_extra and args.graph:
        ev2 = []
        stepper_saved, stepper = stepper, None
        hyper_saved, opt.hyper_dev = getattr(opt, "hyper_dev", None), None   # eager steps pass lr / wd by value
        try:
            for i in range(2):
                step(3000 + i, latent, context)
            barrier()
            for i in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step(3100 + i, latent, context)
                e1.record()
                ev2.append((e0, e1))
            barrier()
            eager_ms = sum(a.elapsed_time(b) for a, b in ev2) / args.steps
        finally:
            stepper = stepper_saved
            opt.hyper_dev = hyper_saved
    graph_ms = None
    if world == 1 and stepper is None and not args.no_graph_extra and args.eager:
        try:
            gstep = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=dev, warmup=1)
            for i in range(3):
                torch.manual_seed(1000 + i)
                gstep(latent, context, t, noise)
            barrier()
            gev = []
            for i in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.manual_seed(2000 + i)
                gstep(latent, context, t, noise)
                e1.record()
                gev.append((e0, e1))
            barrier()
            graph_ms = sum(a.elapsed_time(b) for a, b in gev) / args.steps
        except Exception as ex:  # the extra must never take the bench line down
            graph_ms = None
            sys.stderr.write(f"graph extra skipped: {ex}\n")

    tm = torch.tensor([step_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = tm.tolist()

    if rank == 0:
        pk, pk_src = peaks()
        Lr = N + 16
        kern_flops = 8.0 * B * Lr * Lr * hidden  # self-attention backward, algorithmic 2 x forward
        achieved = kern_flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        step_flops = flops_fwd_bwd(hidden, depth, B, N)
        h2d = h2d_per_step
        line = {
            "metric": "latent tokens/s per train step (fwd+bwd+AdamW)", "value": world * B * N / (step_ms * 1e-3),
            "unit": "latent tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args.workload, world),
            "step_tflops": step_flops / (step_ms * 1e-3) / 1e12,
            "step_frac_of_bf16_peak": step_flops / (step_ms * 1e-3) / 1e12 / peak,
            "roofline": {"kernel": "attn_bwd_kernel (self-attention backward, L=%d)" % Lr, "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": 127.4e6 if args.workload == "debug-8k" else None,
                         "traffic_note": "dram__bytes_read+write per launch, ncu --set full (profiles/r1_ncu_attention_full.md)",
                         "peak_source": pk_src + " (bf16_tflops_sustained)",
                         "kernel_ms": kern_ms, "launches_timed": len(prof),
                         "kernel_share_of_step": kern_ms * depth / step_ms},
            "e2e": {"value": world * B * N / (e2e_ms * 1e-3), "unit": "latent tokens/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": loss_host},
            "gpu_launches": launches if stepper is None else stepper.launches_per_step * args.steps,
            "cuda_graph": stepper is not None, "issue_mode_probe": mode_pick, "clocks": sampler.summary(),
            "eager_issue": None if eager_ms is None else {
                "ms_per_step": eager_ms, "value": world * B * N / (eager_ms * 1e-3), "unit": "latent tokens/s",
                "note": "same step with its kernels issued one by one from Python (no CUDA graph); extra, not the headline"},
            "graph_replay": None if graph_ms is None else {
                "ms_per_step": graph_ms, "value": world * B * N / (graph_ms * 1e-3), "unit": "latent tokens/s",
                "note": "same step captured once and replayed as ONE CUDA graph (train.GraphedTrainStep); extra, not the headline"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_step(args.workload, steps=2, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        if os.environ.get("VDS_BENCH_OUT"):
            with open(os.environ["VDS_BENCH_OUT"], "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="debug-8k", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="world size 1: force the CUDA-graph step (default: an untimed "
                    "2-step probe picks graph replay or Python issue, whichever is faster for the workload)")
    ap.add_argument("--eager", action="store_true", help="issue the kernels from Python instead of replaying a graph")
    ap.add_argument("--no-graph-extra", action="store_true", help="skip the extra graph-replay measurement")
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
