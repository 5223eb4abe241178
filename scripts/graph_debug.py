import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, vds_b200
from vds_b200 import train, lib, ops
from vds_b200.model import DiT, apply_fsdp
from vds_b200.optim import FusedAdamW
from oracle import dit_oracle as O
hidden, depth, heads, B, thw = bench.WORKLOADS["debug-512"]
cfg = bench.model_cfg(hidden, 4, heads)
dev = torch.device("cuda")
torch.manual_seed(0)
model = apply_fsdp(DiT(**cfg).to(dev), torch.bfloat16, torch.float32)
groups, _ = model.get_mup_setup(2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
latent, noise, context, t = [a.to(dev) for a in O.make_inputs(cfg, B, thw, 512, 4096, 1234)]
mode = sys.argv[1] if len(sys.argv) > 1 else "eager_first"
if mode in ("v1", "v2", "v3"):
    sd = torch.zeros(3, device=dev, dtype=torch.int32)
    for i in range(2):
        if mode == "v1":
            opt.zero_grad(); loss, _ = train.forward(model, latent, context, t=t, noise=noise, rope_starts_dev=sd); loss.backward(); opt.step()
        elif mode == "v2":
            opt.zero_grad(); loss, _ = train.forward(model, latent, context, t=t, noise=noise); loss.backward()
        else:
            with torch.no_grad():
                loss, _ = train.forward(model, latent, context, t=t, noise=noise)
    torch.cuda.synchronize()
if mode == "eager_first":
    for i in range(2):
        opt.zero_grad(); loss, _ = train.forward(model, latent, context, t=t, noise=noise); loss.backward(); opt.step()
    torch.cuda.synchronize()
    if len(sys.argv) > 2:
        ops.PROFILE["attn_bwd_self"] = []
        opt.zero_grad(); loss, _ = train.forward(model, latent, context, t=t, noise=noise); loss.backward(); opt.step()
        ops.PROFILE.clear(); torch.cuda.synchronize()
try:
    g = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=dev, warmup=1)
    for i in range(3):
        l = g(latent, context, t, noise)
    torch.cuda.synchronize()
    print("OK", mode, l.item())
except Exception:
    traceback.print_exc()
