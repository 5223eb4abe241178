"""ORACLE — test infrastructure only (never imported by the product path).

A plain restatement of the reference's DiT train step (fal-ai-community/video-diffusion-speedrun,
/root/reference/model.py + train.py) as *functional* PyTorch over a ``state_dict``-keyed parameter
dict, plus numpy restatements of the integer index maps.  Every function cites the reference lines it
follows.  It runs on CPU (fp32: the checker for small cases and the ``cpu_baseline`` of bench.py) or,
inside ``-m gpu`` tests, on the GPU with stock torch ops in fp32 / bf16 as the comparison target for
the hand-written CUDA path.  Gradients come from torch autograd over this restatement.

Parity pin: the reference has no tests / golden vectors of its own, so the pin is
``tests/golden/*.pt`` — outputs of the *imported reference itself*, generated in the build container
by ``oracle/gen_golden.py`` (committed) — against which ``tests/test_oracle_cpu.py`` checks this file.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

N_REG = 16  # model.py:316, 362, 366


# ------------------------------------------------------------------------------------------------
# integer index maps (numpy) — bit-exact contracts
# ------------------------------------------------------------------------------------------------
def patch_token_index(Tp, Hp, Wp):
    """token id of patch (t', h', w'):  rearrange "b c t h w -> b (h w t) c"  (model.py:185)."""
    t, h, w = np.meshgrid(np.arange(Tp), np.arange(Hp), np.arange(Wp), indexing="ij")
    return (h * Wp + w) * Tp + t  # [Tp, Hp, Wp]


def patch_feature_index(C, pt, p):
    """K index of the Conv3d-as-GEMM weight [h, C, pt, p, p] flattened (model.py:173-178)."""
    c, dt, dh, dw = np.meshgrid(np.arange(C), np.arange(pt), np.arange(p), np.arange(p), indexing="ij")
    return ((c * pt + dt) * p + dh) * p + dw


def unpatch_feature_index(C, pt, p):
    """feature id of (p1<->H, p2<->W, p3<->T, c) in "(p1 p2 p3 c)" (model.py:392-401)."""
    p1, p2, p3, c = np.meshgrid(np.arange(p), np.arange(p), np.arange(pt), np.arange(C), indexing="ij")
    return ((p1 * p + p2) * pt + p3) * C + c


def rope_row_position(n, Tp, Hp, Wp):
    """table position (ti, hi, wi) of patch-token row n: the slice [t,h,w,d] is flattened row-major
    "(t h w)" (model.py:239) although tokens are ordered "(h w t)" — a reference quirk kept on purpose."""
    wi = n % Wp
    hi = (n // Wp) % Hp
    ti = n // (Wp * Hp)
    return ti, hi, wi


def patchify_np(x, p, pt):
    """[B,C,T,H,W] -> [B, N, C*pt*p*p] with the orders above (numpy, any dtype)."""
    B, C, T, H, W = x.shape
    Tp, Hp, Wp = T // pt, H // p, W // p
    x = x.reshape(B, C, Tp, pt, Hp, p, Wp, p)
    x = x.transpose(0, 4, 6, 2, 1, 3, 5, 7)  # b h w t c pt ph pw
    return x.reshape(B, Hp * Wp * Tp, C * pt * p * p)


def unpatchify_np(y, C, T, H, W, p, pt):
    """[B, N, p*p*pt*C] -> [B,C,T,H,W] (model.py:392-401)."""
    B = y.shape[0]
    Tp, Hp, Wp = T // pt, H // p, W // p
    y = y.reshape(B, Hp, Wp, Tp, p, p, pt, C)  # b h w t p1 p2 p3 c
    y = y.transpose(0, 7, 3, 6, 1, 4, 2, 5)  # b c t p3 h p1 w p2
    return y.reshape(B, C, T, H, W)


# ------------------------------------------------------------------------------------------------
# model pieces
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim, max_period=10000):
    """model.py:12-22."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(
        device=t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rmsnorm(x, weight=None, eps=1e-6):
    """model.py:34-41."""
    x_dtype = x.dtype
    x = x.float()
    norm = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    if weight is not None:
        return (x * norm * weight).to(dtype=x_dtype)
    return (x * norm).to(dtype=x_dtype)


def rope_tables_rows(dim, thw, starts, device, base=100, table_dtype=torch.float32):
    """Rows [N_REG + N, dim] of cos / sin that ThreeDimRotary.forward returns (model.py:192-263), computed
    lazily for the requested slice instead of materialising the [128,128,128,dim] buffers.
    dim = hidden // (2*heads); layout [t: dim/2 | h: dim/4 | w: dim/4]; rows flattened "(t h w)"."""
    Tp, Hp, Wp = thw
    st, sh, sw = starts
    inv_freq_space = 1.0 / (base ** (torch.arange(0, dim, 4).float() / dim))  # model.py:192
    inv_freq_time = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))  # model.py:193
    t_t = torch.arange(st, st + Tp).float()
    t_h = torch.arange(sh, sh + Hp).float()
    t_w = torch.arange(sw, sw + Wp).float()
    f_t = torch.outer(t_t, inv_freq_time).reshape(Tp, 1, 1, dim // 2).repeat(1, Hp, Wp, 1)
    f_h = torch.outer(t_h, inv_freq_space).reshape(1, Hp, 1, dim // 4).repeat(Tp, 1, Wp, 1)
    f_w = torch.outer(t_w, inv_freq_space).reshape(1, 1, Wp, dim // 4).repeat(Tp, Hp, 1, 1)
    f = torch.cat([f_t, f_h, f_w], 3)  # model.py:214
    cos = f.cos().to(table_dtype).reshape(Tp * Hp * Wp, -1)  # model.py:216, 239
    sin = f.sin().to(table_dtype).reshape(Tp * Hp * Wp, -1)
    cos = torch.cat([torch.ones(N_REG, dim), cos.float()], 0)  # model.py:243-261 (fp32 ones/zeros promote)
    sin = torch.cat([torch.zeros(N_REG, dim), sin.float()], 0)
    return cos.to(device)[None, None], sin.to(device)[None, None]


def draw_rope_starts(thw, hmax=128, wmax=128, tmax=128):
    """model.py:224-226: three draws on the global CPU generator, order h, w, t."""
    Tp, Hp, Wp = thw
    start_h = torch.randint(0, hmax - Hp + 1, (1,)).item()
    start_w = torch.randint(0, wmax - Wp + 1, (1,)).item()
    start_t = torch.randint(0, tmax - Tp + 1, (1,)).item()
    return start_t, start_h, start_w


def apply_rotary_emb(x, cos, sin):
    """model.py:266-275."""
    orig_dtype = x.dtype
    x = x.to(dtype=torch.float32)
    d = x.shape[3] // 2
    x1, x2 = x[..., :d], x[..., d:]
    y1 = x1 * cos + x2 * sin
    y2 = x1 * (-sin) + x2 * cos
    return torch.cat([y1, y2], 3).to(dtype=orig_dtype)


def _heads(x, k, nh):
    """rearrange "b l (k h d) -> k b h l d" (model.py:126, 149-154)."""
    b, l, _ = x.shape
    return x.view(b, l, k, nh, -1).permute(2, 0, 3, 1, 4)


def _merge(x):
    """rearrange "b h l d -> b l (h d)" (model.py:137, 158)."""
    b, nh, l, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, l, nh * d)


def _lin(x, P, name):
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def block_forward(P, pre, x, context, c, v_0, rope, nh, residual_v):
    """DiTBlock.forward, model.py:96-167."""
    mod = F.linear(F.silu(c), P[pre + "adaLN_modulation.1.weight"], P[pre + "adaLN_modulation.1.bias"])
    (shift_sa, scale_sa, gate_sa, shift_ca, scale_ca, gate_ca, shift_mlp, scale_mlp, gate_mlp) = [
        m[:, None, :] for m in mod.chunk(9, dim=1)]
    norm_x = rmsnorm(x, P.get(pre + "norm1.weight"))
    norm_x = norm_x * (1 + scale_sa) + shift_sa
    q, k, v = _heads(_lin(norm_x, P, pre + "qkv"), 3, nh).unbind(0)
    if residual_v and v_0 is not None:
        lam = P[pre + "lambda_param"]
        v = lam * v + (1 - lam) * v_0  # model.py:130
    q = apply_rotary_emb(q, rope[0], rope[1])
    k = apply_rotary_emb(k, rope[0], rope[1])
    attn = _merge(F.scaled_dot_product_attention(q, k, v))
    x = x + _lin(attn, P, pre + "attn_proj") * gate_sa
    if (pre + "context_kv.weight") in P:
        norm_x = rmsnorm(x, P.get(pre + "norm2.weight"))
        norm_x = norm_x * (1 + scale_ca) + shift_ca
        qc = _heads(_lin(norm_x, P, pre + "q_cross"), 1, nh)[0]
        ck, cv = _heads(_lin(context, P, pre + "context_kv"), 2, nh).unbind(0)
        cross = _merge(F.scaled_dot_product_attention(qc, ck, cv))
        x = x + _lin(cross, P, pre + "cross_proj") * gate_ca
    norm_x = rmsnorm(x, P.get(pre + "norm3.weight"))
    norm_x = norm_x * (1 + scale_mlp) + shift_mlp
    hmid = F.gelu(_lin(norm_x, P, pre + "mlp.0"))  # nn.GELU() = exact erf (model.py:85)
    x = x + _lin(hmid, P, pre + "mlp.2") * gate_mlp
    return x, v


def dit_forward(P, cfg, x, context, timesteps, rope_starts=None, table_dtype=torch.float32):
    """DiT.forward, model.py:358-402.  P: state_dict-keyed tensors (any dtype/device), cfg: dict with
    patch_size, time_patch_size, hidden_size, depth, num_heads, residual_v."""
    p, pt, h, nh = cfg["patch_size"], cfg["time_patch_size"], cfg["hidden_size"], cfg["num_heads"]
    b, c, t, hh, w = x.shape
    x = F.conv3d(x, P["patch_embed.patch_proj.weight"], P["patch_embed.patch_proj.bias"], stride=(pt, p, p))
    x = x.permute(0, 3, 4, 2, 1).reshape(b, -1, h)  # "b c t h w -> b (h w t) c"  model.py:185
    x = torch.cat([P["register_tokens"].repeat(b, 1, 1), x], 1)  # model.py:362
    thw = (t // pt, hh // p, w // p)
    if rope_starts is None:
        rope_starts = draw_rope_starts(thw)
    cos, sin = rope_tables_rows(h // (2 * nh), thw, rope_starts, x.device, table_dtype=table_dtype)
    t_emb = timestep_embedding(timesteps, h).to(x.device, dtype=x.dtype)  # model.py:374-376
    t_emb = F.linear(F.silu(_lin(t_emb, P, "time_embed.0")), P["time_embed.2.weight"], P["time_embed.2.bias"])
    v_0 = None
    for i in range(cfg["depth"]):
        x, v = block_forward(P, f"blocks.{i}.", x, context, t_emb, v_0, (cos, sin), nh, cfg["residual_v"])
        if v_0 is None:
            v_0 = v
    x = x[:, N_REG:, :]
    f_shift, f_scale = F.linear(F.silu(t_emb), P["final_modulation.1.weight"],
                                P["final_modulation.1.bias"]).chunk(2, dim=1)
    x = rmsnorm(x, P.get("final_norm.weight"))
    x = x * (1 + f_scale[:, None, :]) + f_shift[:, None, :]
    x = _lin(x, P, "final_proj")
    Tp, Hp, Wp = thw
    x = x.view(b, Hp, Wp, Tp, p, p, pt, c).permute(0, 7, 3, 6, 1, 4, 2, 5).reshape(b, c, t, hh, w)  # 392-401
    return x


# ------------------------------------------------------------------------------------------------
# train-step glue (train.py:89-125) and optimizer (train.py:335-344, model.py:404-465)
# ------------------------------------------------------------------------------------------------
def shift_time(t, alpha=8.0):
    """train.py:93-96 / sample.py:131-134."""
    return t * alpha / (1 + (alpha - 1) * t)


def sample_timesteps(batch_size, device, dtype=torch.bfloat16, generator=None):
    """train.py:89-96."""
    z = torch.randn(batch_size, device=device, dtype=dtype, generator=generator)
    return shift_time(torch.sigmoid(z))


def train_loss(P, cfg, latent, context, t, noise, rope_starts=None, table_dtype=torch.float32):
    """train.py:114-125: z_t / v-target formation, model call, per-sample MSE, batch mean."""
    b = latent.shape[0]
    tr = t.reshape(b, 1, 1, 1, 1)
    z_t = latent * (1 - tr) + noise * tr
    v_objective = latent - noise
    output = dit_forward(P, cfg, z_t, context, t, rope_starts=rope_starts, table_dtype=table_dtype)
    loss_b = (v_objective.float() - output.float()).pow(2).mean(dim=(1, 2, 3, 4))
    return loss_b.mean(), output


def mup_settings(shapes, learning_rate, weight_decay, constant_param_classes):
    """DiT.get_mup_setup, model.py:404-465: name -> (lr, wd) from the FULL parameter shapes."""
    out = {}
    for n, shape in shapes.items():
        if any(k in n for k in ("bias", "norm", "lambda")):
            lr, wd = learning_rate * 0.01, 0.0
        else:
            hidden_dim = shape[-1]
            lr, wd = learning_rate * (32 / hidden_dim), weight_decay * hidden_dim / 1024
        if any(cls in n for cls in constant_param_classes):
            lr, wd = learning_rate * 0.01, 0.0
        if "time" in n:
            lr = learning_rate * 0.1
        if "modulation" in n:
            lr = learning_rate * 0.1
        out[n] = (lr, wd)
    return out


def adamw_step(p, g, m, v, step, lr, wd, beta1=0.95, beta2=0.99, eps=1e-8):
    """torch.optim.AdamW single-tensor math (train.py:340-344; betas from there, eps = torch default)."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# ------------------------------------------------------------------------------------------------
# helpers shared by tests / bench
# ------------------------------------------------------------------------------------------------
ZERO_INIT = ("adaLN_modulation.1.", "final_modulation.1.", "final_proj.")


def randomise_zero_init(state, seed=0, std=0.02):
    """The reference zero-inits the adaLN heads and final_proj (model.py:93-94,347-350) which makes output
    and 296/300 grads exactly zero at init: re-draw them N(0, std) so every kernel works."""
    g = torch.Generator().manual_seed(seed)
    for n in sorted(state):
        if any(z in n for z in ZERO_INIT):
            state[n] = (torch.randn(state[n].shape, generator=g) * std).to(state[n].dtype)
    return state


def make_inputs(cfg, B, thw_latent, Lc, Dc, seed):
    """Synthetic batch: latent/noise ~ N(0,1) bf16, context ~ N(0,1) bf16, t per train.py:90-96."""
    g = torch.Generator().manual_seed(seed)
    T, H, W = thw_latent
    latent = torch.randn((B, cfg["in_channels"], T, H, W), generator=g).bfloat16()
    noise = torch.randn(latent.shape, generator=g).bfloat16()
    context = torch.randn((B, Lc, Dc), generator=g).bfloat16()
    t = shift_time(torch.sigmoid(torch.randn((B,), generator=g).bfloat16()))
    return latent, noise, context, t


def sample_loop(P, cfg, prompt_embeds, latents, inference_steps, cfg_scale=6.0, table_dtype=torch.bfloat16,
                model_dtype=torch.bfloat16):
    """sampling/sample.py:107-146 restated: shifted-time Euler, CFG with zero negative embeds, fp32 accumulator.
    `P` / inputs are used in `model_dtype` like the reference's ``model.to(device, bf16)`` (sample.py:63)."""
    Pm = {k: v.to(model_dtype) for k, v in P.items()}
    prompt_embeds = prompt_embeds.to(model_dtype)
    negative = torch.zeros_like(prompt_embeds)
    latents = latents.to(model_dtype)
    acc = latents.to(torch.float32)
    for i in range(inference_steps, 0, -1):
        t = shift_time(i / inference_steps)
        t_next = shift_time((i - 1) / inference_steps)
        dt = t - t_next
        tt = torch.tensor([t] * latents.shape[0]).to(latents.device, model_dtype)
        out = dit_forward(Pm, cfg, latents, prompt_embeds, tt, table_dtype=table_dtype)
        if cfg_scale > 1:
            un = dit_forward(Pm, cfg, latents, negative, tt, table_dtype=table_dtype)
            out = un + cfg_scale * (out - un)
        acc = acc + dt * out.to(torch.float32)
        latents = acc.to(model_dtype)
    return acc
