"""Shared by CPU and GPU tests: rebuild the golden cases (parameters + inputs) deterministically."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def cos_sim(a, b):
    a, b = a.detach().float().flatten().cpu(), b.detach().float().flatten().cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def build_model(cfg, seed_model, seed_zero, device="cpu", dtype=torch.float32):
    """The mirror DiT draws its parameters exactly like the reference constructor (same RNG order, verified
    bit-exactly by oracle/gen_golden.py's param_checksum), then the zero-init tensors are re-drawn."""
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    from oracle import dit_oracle as O
    torch.manual_seed(seed_model)
    m = DiT(**cfg)
    sd = {k: v.clone() for k, v in m.state_dict().items() if "freqs_hwt" not in k}
    sd = O.randomise_zero_init(sd, seed=seed_zero)
    m.load_state_dict(sd, strict=False)
    return m.to(device=device) if dtype == torch.float32 else m.to(device=device, dtype=dtype)


def golden_case(name):
    from oracle import dit_oracle as O
    fx = load_golden(name)
    cfg = fx["cfg"]
    model = build_model(cfg, fx["seeds"]["model"], fx["seeds"]["zero"])
    latent, noise, context, t = O.make_inputs(cfg, fx["B"], fx["latent_thw"], fx["Lc"], cfg["cross_attn_input_size"],
                                              fx["seeds"]["data"])
    return fx, cfg, model, (latent, noise, context, t)


def params_of(model, device=None, dtype=None, requires_grad=False):
    out = {}
    for n, p in model.named_parameters():
        t = p.detach().clone()
        if device is not None or dtype is not None:
            t = t.to(device=device or t.device, dtype=dtype or t.dtype)
        out[n] = t.requires_grad_(requires_grad)
    return out
