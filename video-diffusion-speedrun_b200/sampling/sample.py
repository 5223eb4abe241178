"""Mirror of the denoising loop of the reference ``sampling/sample.py`` (generate_image, lines 77-159) on the
CUDA DiT: shifted-time Euler steps on the flow, classifier-free guidance with a zeroed negative embedding, fp32
latent accumulator, bf16 model input.  T5 encoding, the Streamlit UI and the Cosmos decoder / mp4 writer are out of
scope: the caller passes ``prompt_embeds`` and gets the final fp32 latents back.
"""
import torch


def shift_time(t, alpha=8.0):
    """sample.py:131-134."""
    return t * alpha / (1 + (alpha - 1) * t)


@torch.no_grad()
def denoise(model, prompt_embeds, latent_shape=(1, 16, 16, 64, 64), inference_steps=50, cfg_scale=6.0, seed=42,
            device="cuda", dtype=torch.bfloat16, latents=None, on_step=None, cache_context_kv=True):
    """sample.py:92-146.  Returns acc_latents (fp32).  Two separate model calls per step (cond / uncond), each drawing
    its own RoPE offsets from the global CPU generator, exactly like the reference."""
    prompt_embeds = prompt_embeds.to(device=device, dtype=dtype)
    negative_embeds = torch.zeros_like(prompt_embeds)                       # sample.py:104
    if latents is None:
        generator = torch.Generator(device=device).manual_seed(seed)         # sample.py:108
        latents = torch.randn(latent_shape, device=device, dtype=dtype, generator=generator)
    acc_latents = latents.to(dtype=torch.float32)
    if cache_context_kv and hasattr(model, "context_kv_cache"):
        model.context_kv_cache(True)     # prompt / negative embeddings are constant over the loop
    for i in range(inference_steps, 0, -1):
        t = shift_time(i / inference_steps)
        t_next = shift_time((i - 1) / inference_steps)
        dt = t - t_next
        tt = torch.tensor([t] * latents.shape[0]).to(device, dtype)
        model_output = model(latents, prompt_embeds, tt)
        if cfg_scale > 1:
            uncond_output = model(latents, negative_embeds, tt)
            model_output = uncond_output + cfg_scale * (model_output - uncond_output)
        acc_latents = acc_latents + dt * model_output.to(dtype=torch.float32)
        latents = acc_latents.to(dtype=dtype)
        if on_step is not None:
            on_step(i, acc_latents)
    if cache_context_kv and hasattr(model, "context_kv_cache"):
        model.context_kv_cache(False)
    return acc_latents
