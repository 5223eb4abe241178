"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/model.py, imported
read-only) on CPU in fp32 at fixed seeds.  Run in the build container only (the GPU box has no
/root/reference):   python oracle/gen_golden.py

Fixture contents are small: full model output + loss, and for every parameter gradient / AdamW-updated
parameter its L2 norm and 256 entries at fixed pseudo-random positions (full tensors are compared
against the oracle restatement here, at generation time, and the result is recorded in the fixture).
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dit_oracle as O  # noqa: E402

REF = "/root/reference/model.py"


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_model", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sample_idx(numel, k=256, seed=7):
    g = torch.Generator().manual_seed(seed + numel % 1000)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def summarise(t):
    f = t.detach().float().flatten()
    idx = sample_idx(f.numel())
    return {"norm": f.norm().item(), "idx": idx, "val": f[idx].clone(), "shape": tuple(t.shape)}


CASES = {
    "tiny_nobias": dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=256, depth=2, num_heads=2,
                        mlp_ratio=4.0, cross_attn_input_size=64, residual_v=True, train_bias_and_rms=False,
                        use_rope=True),
    "tiny_bias": dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=256, depth=3, num_heads=2,
                      mlp_ratio=4.0, cross_attn_input_size=64, residual_v=True, train_bias_and_rms=True,
                      use_rope=True),
}
LATENT_THW = (4, 8, 8)
LC, B = 24, 2
SEED_MODEL, SEED_ZERO, SEED_DATA, SEED_ROPE = 0, 1, 2, 123


def run_case(ref, name, cfg):
    torch.manual_seed(SEED_MODEL)
    model = ref.DiT(**cfg)
    sd = model.state_dict()
    rnd = O.randomise_zero_init({k: v.clone() for k, v in sd.items() if "freqs_hwt" not in k}, seed=SEED_ZERO)
    model.load_state_dict(rnd, strict=False)
    latent, noise, context, t = O.make_inputs(cfg, B, LATENT_THW, LC, cfg["cross_attn_input_size"], SEED_DATA)
    latent, noise, context, t = latent.float(), noise.float(), context.float(), t.float()

    # replay the rope draws to record them (model.py:224-226)
    torch.manual_seed(SEED_ROPE)
    thw = (LATENT_THW[0] // 2, LATENT_THW[1] // 2, LATENT_THW[2] // 2)
    starts = O.draw_rope_starts(thw)

    # reference train step (train.py:114-125 restated around the imported reference model)
    torch.manual_seed(SEED_ROPE)
    tr = t.reshape(B, 1, 1, 1, 1)
    z_t = latent * (1 - tr) + noise * tr
    v_obj = latent - noise
    out = model(z_t, context, t)
    loss = (v_obj.float() - out.float()).pow(2).mean(dim=(1, 2, 3, 4)).mean()
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}

    # oracle restatement on the same inputs: full-tensor comparison, recorded in the fixture
    P = {k: v.clone().requires_grad_(True) for k, v in rnd.items()}
    o_loss, o_out = O.train_loss(P, cfg, latent, context, t, noise, rope_starts=starts)
    o_loss.backward()
    worst = 0.0
    for n, g in grads.items():
        if g is None:
            assert P[n].grad is None or P[n].grad.abs().max() == 0, n
            continue
        d = (P[n].grad - g).abs().max().item() / (g.abs().max().item() + 1e-30)
        worst = max(worst, d)
    out_err = (o_out - out).abs().max().item()
    print(f"[{name}] loss {loss.item():.6f} oracle-vs-reference: out max abs err {out_err:.3e}, "
          f"worst grad rel err {worst:.3e}, rope starts {starts}")
    assert out_err < 1e-5 and worst < 1e-4

    # optimizer: reference grouping + torch AdamW, two steps with the same gradient
    groups, settings = model.get_mup_setup(2 ** -7, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
    opt = torch.optim.AdamW(groups, betas=(0.95, 0.99))
    opt.step()
    opt.step()
    after = {n: p.detach() for n, p in model.named_parameters()}

    fx = {
        "cfg": cfg, "seeds": dict(model=SEED_MODEL, zero=SEED_ZERO, data=SEED_DATA, rope=SEED_ROPE),
        "latent_thw": LATENT_THW, "Lc": LC, "B": B, "rope_starts": starts,
        "out": out.detach().clone(), "loss": loss.item(),
        "grads": {n: (summarise(g) if g is not None else None) for n, g in grads.items()},
        "param_checksum": {n: (v.double().sum().item(), v.double().abs().sum().item()) for n, v in rnd.items()},
        "mup": {n: (s["lr"], s["wd"]) for n, s in settings.items()},
        "n_groups": len(groups),
        "adamw_after_2_steps": {n: summarise(v) for n, v in after.items()},
        "oracle_vs_reference": dict(out_max_abs_err=out_err, worst_grad_rel_err=worst),
        "torch": torch.__version__,
    }
    torch.save(fx, os.path.join(ROOT, "tests", "golden", f"{name}.pt"))


def index_maps(ref):
    """Bit-exact index contracts from the reference's own ops on integer-coded tensors."""
    from einops import rearrange
    B, C, T, H, W = 2, 16, 4, 6, 8
    x = (torch.arange(B * C * T * H * W) % 253).float().view(B, C, T, H, W)
    pe = ref.PatchEmbed(2, C, 128, 2)
    with torch.no_grad():
        pe.patch_proj.weight.zero_()
        pe.patch_proj.bias.zero_()
        pe.patch_proj.weight.view(128, -1)[torch.arange(128), torch.arange(128)] = 1.0
        patches = pe(x)  # [B, N, 128]: row = token "(h w t)", col k = Conv3d weight position k
    y = (torch.arange(B * 24 * 128) % 251).float().view(B, 24, 128)
    un = rearrange(y, "b (h w t) (p1 p2 p3 c) -> b c (t p3) (h p1) (w p2)", t=T // 2, h=H // 2, w=W // 2, p1=2,
                   p2=2, p3=2)
    # rope rows from the reference module (dim 8 keeps the table tiny)
    rope = ref.ThreeDimRotary(8, h=128, w=128, t=128)
    torch.manual_seed(5)
    cos, sin = rope(None, time_height_width=(2, 3, 4), extend_with_register_tokens=16)
    torch.manual_seed(5)
    starts = O.draw_rope_starts((2, 3, 4))
    torch.save({"x_shape": (B, C, T, H, W), "patches": patches.to(torch.int16), "y_shape": tuple(y.shape),
                "unpatch": un.to(torch.int16), "rope_thw": (2, 3, 4), "rope_starts": starts, "rope_dim": 8,
                "rope_cos": cos[0, 0].clone(), "rope_sin": sin[0, 0].clone()},
               os.path.join(ROOT, "tests", "golden", "index_maps.pt"))
    print("[index_maps] rope starts", starts)


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = load_ref()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    index_maps(ref)
    for name, cfg in CASES.items():
        run_case(ref, name, cfg)
