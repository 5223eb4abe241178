"""How long does the HOST take to issue one train step (no sync) vs the GPU to run it?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, vds_b200
from vds_b200 import train, lib
from vds_b200.model import DiT, apply_fsdp
from vds_b200.optim import FusedAdamW
from oracle import dit_oracle as O
hidden, depth, heads, B, thw = bench.WORKLOADS["debug-8k"]
cfg = bench.model_cfg(hidden, depth, heads)
dev = torch.device("cuda")
torch.manual_seed(0)
model = DiT(**cfg)
with torch.no_grad():
    sd = O.randomise_zero_init({n: p.detach().clone() for n, p in model.named_parameters()}, seed=1)
    for n, p in model.named_parameters():
        p.copy_(sd[n] * 0.1 if p.dim() == 2 else sd[n])
model = apply_fsdp(model.to(dev), torch.bfloat16, torch.float32)
groups, _ = model.get_mup_setup(2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
latent, noise, context, t = [a.to(dev) for a in O.make_inputs(cfg, B, thw, 512, 4096, 1234)]
def step(i):
    torch.manual_seed(i); opt.zero_grad()
    loss, _ = train.forward(model, latent, context, t=t, noise=noise); loss.backward(); opt.step()
for i in range(3): step(i)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); step(10 + i); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host issue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(20); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
