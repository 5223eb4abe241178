#!/usr/bin/env python
"""Context numbers (NOT the target, NOT a bench line): the reference's own way of running this step on one B200 —
PyTorch eager bf16 on library kernels (cuBLAS, cuDNN / flash SDPA, ATen element-wise, torch's fused AdamW), and the
same under torch.compile (train.py:327-329) — timed with the same inputs and shapes as bench.py's workloads.

/root/reference does not exist on the GPU box, so the model is the oracle's op-for-op restatement of model.py
(oracle/dit_oracle.py: the same F.linear / F.scaled_dot_product_attention / F.conv3d / F.gelu calls in the same order),
run with bf16 parameter copies of fp32 masters and fp32 gradients for AdamW, i.e. FSDP2's MixedPrecisionPolicy
(param_dtype=bf16, reduce_dtype=fp32) at world size 1 (model.py:516-519; the reference's own apply_fsdp raises NameError
at world size 1, model.py:489).

  python scripts/ref_gpu_baseline.py [--workloads debug-8k,B,XL] [--steps 5] [--compile] [--out file.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402


def run(name, steps, compile_it, dev):
    hidden, depth, heads, B, thw = bench.WORKLOADS[name]
    cfg = bench.model_cfg(hidden, depth, heads)
    N = (thw[0] // 2) * (thw[1] // 2) * (thw[2] // 2)
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    with torch.device(dev):
        m = DiT(**cfg)
    master = {}
    with torch.no_grad():
        for n, p in m.named_parameters():
            if any(z in n for z in O.ZERO_INIT):
                p.normal_(0.0, 0.02)
            if p.dim() == 2:
                p.mul_(0.1)
            master[n] = p.detach().clone().float().requires_grad_(True)
    del m
    shapes = {n: tuple(p.shape) for n, p in master.items()}
    settings = O.mup_settings(shapes, 2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
    groups = {}
    for n, p in master.items():
        groups.setdefault(settings[n], []).append(p)
    opt = torch.optim.AdamW([{"params": ps, "lr": lr, "weight_decay": wd} for (lr, wd), ps in groups.items()],
                            betas=(0.95, 0.99), fused=True)
    names = list(master)
    P16 = {n: master[n].detach().bfloat16().requires_grad_(True) for n in names}
    latent, noise, context, t = [a.to(dev) for a in O.make_inputs(cfg, B, thw, bench.LC, bench.DC, 1234)]
    thw_p = tuple(d // 2 for d in thw)

    def loss_fn(P, latent, context, t, noise, starts):
        return O.train_loss(P, cfg, latent, context, t, noise, rope_starts=starts)[0]

    fn = torch.compile(loss_fn) if compile_it else loss_fn

    def step(i):
        torch.manual_seed(i)
        starts = O.draw_rope_starts(thw_p)
        with torch.no_grad():
            torch._foreach_copy_([P16[n] for n in names], [master[n] for n in names])   # the bf16 "all-gather" cast
        for n in names:
            P16[n].grad = None
        loss = fn(P16, latent, context, t, noise, starts)
        loss.backward()
        with torch.no_grad():
            for n in names:
                g = P16[n].grad
                master[n].grad = None if g is None else g.float()                     # fp32 "reduce-scatter"
        opt.step()
        return loss

    t0 = time.time()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    warm_s = time.time() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    evs = []
    for i in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step(10 + i)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    fl = bench.flops_fwd_bwd(hidden, depth, B, N)
    pk, _ = bench.peaks()
    return {"impl": "torch-eager-bf16 (reference ops via the oracle restatement)" + (" + torch.compile" if compile_it else ""),
            "workload": name, "ms_per_step": ms, "tokens_per_s": B * N / (ms * 1e-3),
            "step_tflops": fl / (ms * 1e-3) / 1e12, "frac_of_sustained_bf16": fl / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
            "steps": steps, "warmup_s": warm_s, "loss": float(loss), "torch": torch.__version__,
            "sdpa_backends": {"flash": torch.backends.cuda.flash_sdp_enabled(), "cudnn": torch.backends.cuda.cudnn_sdp_enabled(),
                              "mem_efficient": torch.backends.cuda.mem_efficient_sdp_enabled()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="debug-8k,B,XL")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True          # train.py:23-24
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda:0")
    for name in a.workloads.split(","):
        try:
            r = run(name, a.steps, a.compile, dev)
        except Exception as ex:  # noqa: BLE001
            r = {"workload": name, "compile": a.compile, "error": f"{type(ex).__name__}: {str(ex)[:400]}"}
        print(json.dumps(r), flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(r) + "\n")
        import gc
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
