"""Mirror of the train-step arithmetic of the reference ``train.py`` (/root/reference/train.py:89-125,
429-435) on the fused CUDA path: timestep sampling, noise, z_t / v-target, model, MSE loss, backward.

``forward(dit_model, latent, caption_encoded, ...) -> (total_loss, diffusion_loss)`` keeps the reference's
name and return convention (train.py:51-145); the T5 encoder / dataloader (out of scope) are
replaced by the caller handing in ``caption_encoded`` directly.
"""
import torch

from . import engine, ops


def shift_time(t, alpha=8.0):
    """train.py:93-96."""
    return t * alpha / (1 + (alpha - 1) * t)


class TrainStepFunction(torch.autograd.Function):
    """loss = mse(v, DiT(z_t)) as ONE autograd node: z_t is formed inside the patch gather, the loss and
    its gradient come from one fused reduction kernel, the model backward is the hand-written engine."""

    @staticmethod
    def forward(ctx, model, need, latent, noise, context, t, *params):
        P = model._param_view()
        out, c = engine.forward(model, P, latent, context, t, save=need, noise=noise)
        loss, _, lb = ops.loss_fwd_bwd(latent, noise, out, want_grad=False, want_batch=True)
        ctx.model, ctx.P, ctx.c = model, P, c
        ctx.saved = (latent, noise, out)
        ctx.n_params = len(params)
        ctx.dtypes = [p.dtype for p in params]
        ctx.mark_non_differentiable(lb)
        return loss.view(()), lb

    @staticmethod
    def backward(ctx, gloss, _glb):
        latent, noise, out = ctx.saved
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        _, d_out, _ = ops.loss_fwd_bwd(latent, noise, out, want_grad=True, want_loss=False, grad_scale_dev=g)
        grads = engine.run_backward(ctx.model, ctx.P, ctx.c, d_out, ctx.dtypes)
        ctx.c = None
        return (None, None, None, None, None, None) + grads


def forward(dit_model, latent, caption_encoded, generator=None, t=None, noise=None):
    """train.py:51-145 on the fused path.  latent [B,16,T,H,W], caption_encoded [B,512,4096] (bf16, CUDA).
    `t` / `noise` may be supplied (parity tests); otherwise drawn exactly like train.py:89-105."""
    device = latent.device
    vae_latent = latent.to(torch.bfloat16).contiguous()  # train.py:73
    batch_size = vae_latent.size(0)
    if t is None:
        z = torch.randn(batch_size, device=device, dtype=torch.bfloat16, generator=generator)
        t = shift_time(torch.sigmoid(z))
    if noise is None:
        noise = torch.randn(vae_latent.shape, device=device, dtype=torch.bfloat16, generator=generator)
    params = [p for _, p in dit_model.named_parameters()]
    need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    loss, loss_batchwise = TrainStepFunction.apply(dit_model, need, vae_latent, noise.contiguous(),
                                                   caption_encoded.to(torch.bfloat16), t.to(torch.bfloat16), *params)
    dit_model.last_loss_batchwise = loss_batchwise
    return loss, loss
