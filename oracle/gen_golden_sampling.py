"""Pins the oracle's sampling loop (oracle/dit_oracle.py:sample_loop) against the UNMODIFIED reference
``sampling/sample.py:generate_image`` (lines 77-159).  Build container only:   python oracle/gen_golden_sampling.py

``sample.py`` cannot be imported as it stands (streamlit, the Cosmos decoder and the T5 helpers are imported at module
level and none of them is part of the hot path), so the three modules are replaced by stubs *before* the import:
  streamlit  -> cache_resource passthrough, progress() no-op
  decoder    -> get_decoder / save_tensor_to_mp4 that just captures the final latents
  utils      -> encode_prompt_with_t5 returning the synthetic prompt embedding (the prompt "" call returns anything:
                the reference zeroes the negative embedding itself, sample.py:104)
``model`` is the real /root/reference/model.py.  generate_image then runs its own loop — shifted-time Euler, CFG,
fp32 accumulator, two model calls per step each drawing RoPE offsets from the global CPU RNG — on a tiny fp32 DiT on
the CPU.  The fixture stores inputs, the final latents and the RNG seed; tests/test_oracle_cpu.py replays the oracle.
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
           cross_attn_input_size=32, residual_v=True, train_bias_and_rms=False, use_rope=True)
STEPS, CFG_SCALE, SEED_LAT, SEED_RNG, SEED_MODEL, SEED_ZERO, SEED_CTX = 5, 6.0, 42, 1234, 0, 1, 9
HEIGHT = WIDTH = 32          # generate_image: latents (1, 16, 16, 2*(H//16), 2*(W//16)) = [1,16,16,4,4] -> 8*2*2 = 32 tokens
LC = 12


def load_reference_sampler(prompt_embeds, captured):
    st = types.ModuleType("streamlit")
    st.cache_resource = lambda f=None, **k: (f if f is not None else (lambda g: g))

    class _Bar:
        def progress(self, *_a, **_k):
            return None
    st.progress = lambda *_a, **_k: _Bar()
    sys.modules["streamlit"] = st
    dec = types.ModuleType("decoder")
    dec.get_decoder = lambda *a, **k: None

    def save_tensor_to_mp4(latents, vae, out_dir, name):
        captured["latents"] = latents.detach().clone()
    dec.save_tensor_to_mp4 = save_tensor_to_mp4
    sys.modules["decoder"] = dec
    ut = types.ModuleType("utils")
    ut.load_encoders = lambda *a, **k: (None, None)
    ut.encode_prompt_with_t5 = lambda text_encoder, tokenizer, prompt=None, device=None, return_index=-1: prompt_embeds.clone()
    sys.modules["utils"] = ut
    sys.path.insert(0, REF_DIR)                      # `from model import DiT, timestep_embedding`
    spec = importlib.util.spec_from_file_location("ref_sample", os.path.join(REF_DIR, "sampling", "sample.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Vae(torch.nn.Module):                         # generate_image only asks for the dtype of its first parameter
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1, dtype=torch.float32))


def main():
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(SEED_CTX)
    prompt_embeds = torch.randn((1, LC, CFG["cross_attn_input_size"]), generator=g)
    captured = {}
    ref = load_reference_sampler(prompt_embeds, captured)
    import model as ref_model                        # the unmodified reference model.py
    torch.manual_seed(SEED_MODEL)
    m = ref_model.DiT(**CFG)
    sd = O.randomise_zero_init({k: v.clone() for k, v in m.state_dict().items() if "freqs_hwt" not in k}, seed=SEED_ZERO)
    m.load_state_dict(sd, strict=False)
    m.eval()
    torch.manual_seed(SEED_RNG)                      # RoPE offset draws of the 2 x STEPS model calls
    ref.generate_image("a prompt", m, _Vae(), None, None, device="cpu", dtype=torch.float32, inference_steps=STEPS,
                       cfg_scale=CFG_SCALE, height=HEIGHT, width=WIDTH, seed=SEED_LAT)
    final = captured["latents"]                      # acc_latents.squeeze(0): [16, 16, 4, 4] fp32
    # the same start latents generate_image drew (sample.py:108-114)
    gen = torch.Generator(device="cpu").manual_seed(SEED_LAT)
    lat0 = torch.randn((1, 16, 16, 2 * (HEIGHT // 16), 2 * (WIDTH // 16)), dtype=torch.float32, generator=gen)
    # oracle replay, checked here at generation time
    torch.manual_seed(SEED_RNG)
    got = O.sample_loop(sd, CFG, prompt_embeds, lat0, STEPS, cfg_scale=CFG_SCALE, table_dtype=torch.float32,
                        model_dtype=torch.float32)
    err = (got.squeeze(0) - final).abs().max().item()
    print(f"reference generate_image vs oracle sample_loop: max abs err {err:.3e} (|latents| max {final.abs().max().item():.3f})")
    fx = {"cfg": CFG, "steps": STEPS, "cfg_scale": CFG_SCALE, "seed_rng": SEED_RNG, "prompt_embeds": prompt_embeds,
          "lat0": lat0, "final": final, "seed_model": SEED_MODEL, "seed_zero": SEED_ZERO,
          # weights are not stored: DiT(**cfg) under torch.manual_seed(seed_model) + randomise_zero_init(seed_zero)
          # reproduces them (same constructor RNG consumption); the norms below guard that
          "param_norms": {k: v.float().norm().item() for k, v in sd.items()},
          "oracle_max_abs_err_at_generation": err}
    out = os.path.join(ROOT, "tests", "golden", "sampling_tiny.pt")
    torch.save(fx, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
