"""Denoising-step time: eager loop vs CUDA-graph replay vs graph with batched cond+uncond.
usage: python scripts/sample_bench.py [hidden depth heads B T H W steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vds_b200
from vds_b200.model import DiT
from vds_b200.sampling.sample import denoise, GraphedDenoiser

a = [int(v) for v in sys.argv[1:]]
hidden, depth, heads, B, T, H, W, steps = (a + [512, 24, 4, 1, 16, 64, 64, 8][len(a):])[:8]
dev = torch.device("cuda")
torch.manual_seed(0)
model = DiT(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=hidden, depth=depth, num_heads=heads,
            cross_attn_input_size=4096, residual_v=True, train_bias_and_rms=False, use_rope=True)
with torch.no_grad():
    for n, p in model.named_parameters():
        if p.abs().max() == 0:
            p.normal_(0, 0.02)
model = model.to(dev, torch.bfloat16).eval()
ctx = torch.randn((B, 512, 4096), device=dev).bfloat16()
lat = torch.randn((B, 16, T, H, W), device=dev).bfloat16()


def timed(fn):
    fn()                       # warm-up (and graph capture)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


ms_eager = timed(lambda: denoise(model, ctx, inference_steps=steps, latents=lat, device=dev))
g = GraphedDenoiser(model, ctx, tuple(lat.shape), device=dev)
ms_graph = timed(lambda: g.run(lat, inference_steps=steps))
gb = GraphedDenoiser(model, ctx, tuple(lat.shape), device=dev, batch_cfg=True)
ms_batched = timed(lambda: gb.run(lat, inference_steps=steps))
N = (T // 2) * (H // 2) * (W // 2)
print(f"h={hidden} depth={depth} B={B} N={N}: per denoising step (cond+uncond, CFG, Euler): eager {ms_eager:.2f} ms | "
      f"graph {ms_graph:.2f} ms ({g.launches_per_step} kernels/replay) | graph + batched CFG {ms_batched:.2f} ms")
