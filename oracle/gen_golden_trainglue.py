"""Pins the oracle's train-step glue (oracle/dit_oracle.py: sample_timesteps, train_loss) against the UNMODIFIED
reference ``train.py:forward`` (lines 51-145).  Build container only:   python oracle/gen_golden_trainglue.py

``train.py`` imports ``utils`` (T5 encoders, HF dataset) at module level; that module is replaced by a stub whose
``encode_prompt_with_t5`` returns a synthetic caption embedding — everything else (``click``, ``wandb``,
``transformers`` schedulers, ``model``) is the real thing.  ``forward()`` then runs its own arithmetic on the CPU:
bf16 cast of the latents, the 1 % caption zero-out draw (global RNG), t = shift(sigmoid(N(0,1))) and the noise from
the passed generator, z_t / v-target, the reference DiT in bf16, per-sample MSE and batch mean.
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dit_oracle as O  # noqa: E402

REF_DIR = "/root/reference"
CFG = dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=128, depth=2, num_heads=4, mlp_ratio=4.0,
           cross_attn_input_size=32, residual_v=True, train_bias_and_rms=True, use_rope=True)
B, THW, LC = 3, (4, 8, 8), 10
SEED_MODEL, SEED_ZERO, SEED_DATA, SEED_GEN, SEED_GLOBAL = 0, 1, 5, 77, 4321


def main():
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(SEED_DATA)
    latent = torch.randn((B, 16) + THW, generator=g)                       # dataset rows are fp32/bf16 latents
    caption = torch.randn((B, LC, CFG["cross_attn_input_size"]), generator=g)
    ut = types.ModuleType("utils")
    ut.avg_scalar_across_ranks = lambda x: x
    ut.create_dataloader = lambda *a, **k: None
    ut.load_encoders = lambda *a, **k: (None, None)
    ut.encode_prompt_with_t5 = lambda text_encoder, tokenizer, prompt=None, device=None, return_index=-1: caption.clone()
    sys.modules["utils"] = ut
    sys.path.insert(0, REF_DIR)
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF_DIR, "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    import model as ref_model
    torch.manual_seed(SEED_MODEL)
    m = ref_model.DiT(**CFG)
    sd = O.randomise_zero_init({k: v.clone() for k, v in m.state_dict().items() if "freqs_hwt" not in k}, seed=SEED_ZERO)
    m.load_state_dict(sd, strict=False)
    m = m.to(torch.bfloat16)          # what FSDP2's MixedPrecisionPolicy(param_dtype=bf16) computes with (model.py:516-519)

    gen = torch.Generator().manual_seed(SEED_GEN)
    torch.manual_seed(SEED_GLOBAL)    # caption zero-out draw (train.py:86) + the 3 RoPE draws of the model call
    total, diff = ref_train.forward(m, {"latent": latent, "prompt": ["p"] * B}, None, None, "cpu", 1, False, generator=gen)
    ref_loss = total.item()

    # oracle replay
    gen = torch.Generator().manual_seed(SEED_GEN)
    torch.manual_seed(SEED_GLOBAL)
    lat16 = latent.to(torch.bfloat16)
    cap16 = caption.to(torch.bfloat16)
    zero = torch.rand(B) < 0.01                                             # train.py:86-87
    cap16[zero] = 0
    t = O.sample_timesteps(B, "cpu", torch.bfloat16, gen)                   # train.py:89-96
    noise = torch.randn(lat16.shape, dtype=torch.bfloat16, generator=gen)   # train.py:103-105
    starts = O.draw_rope_starts(tuple(d // 2 for d in THW))
    P16 = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    loss, _ = O.train_loss(P16, CFG, lat16, cap16, t, noise, rope_starts=starts, table_dtype=torch.bfloat16)
    err = abs(loss.item() - ref_loss)
    print(f"reference train.forward loss {ref_loss:.8f} vs oracle {loss.item():.8f} (abs err {err:.3e}); zeroed captions: {int(zero.sum())}")
    fx = {"cfg": CFG, "latent": latent, "caption": caption, "seed_model": SEED_MODEL, "seed_zero": SEED_ZERO,
          "seed_gen": SEED_GEN, "seed_global": SEED_GLOBAL, "loss": ref_loss, "t": t.clone(),
          "param_norms": {k: v.float().norm().item() for k, v in sd.items()}, "oracle_abs_err_at_generation": err}
    out = os.path.join(ROOT, "tests", "golden", "trainglue_tiny.pt")
    torch.save(fx, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
