"""Drop-in mirror of the reference ``model.py`` (/root/reference/model.py) for the B200-native path.

Same public names, constructor kwargs, parameter / buffer names and shapes (so ``state_dict`` round-trips
with the reference and ``load_state_dict(strict=True)`` of ``sampling/sample.py:61`` works), same
``forward(x, context, timesteps)`` contract, same ``get_mup_setup``.  The arithmetic, however, never
touches torch ops: ``DiT.forward`` is one autograd node (``engine.DiTFunction``) that runs the
hand-written sm_100a kernels forward and backward.  There is no CPU path: on a CPU tensor
``forward`` raises.
"""
import math
from collections import defaultdict
from functools import reduce

import torch
from torch import nn

from . import engine, ops


def timestep_embedding(t, dim, max_period=10000):
    """model.py:12-22.  CUDA tensors go through the fused kernel (bf16 out, like the reference's
    ``.to(dtype=x.dtype)`` at model.py:374-376 after a bf16 cast); CPU tensors use the reference formula
    (host-side helper used by sample.py)."""
    if t.is_cuda:
        return ops.timestep_embedding(t.to(torch.bfloat16).contiguous(), dim, float(max_period))
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(
        device=t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class RMSNorm(nn.Module):
    """model.py:25-41 (parameter container; the math lives in the fused rmsnorm+modulate kernel)."""

    def __init__(self, dim, eps=1e-6, trainable=False):
        super().__init__()
        self.eps = eps
        self.dim = dim
        if trainable:
            self.weight = nn.Parameter(torch.ones(dim))
        else:
            self.weight = None

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("vds_b200.RMSNorm has no CPU path")
        shp = x.shape
        x2 = x.to(torch.bfloat16).contiguous().view(-1, shp[-1])
        w = self.weight.to(torch.bfloat16) if self.weight is not None else None
        y, _ = ops.rmsnorm_mod_fwd(x2, 1, x2.shape[0], shp[-1], weight=w, eps=self.eps, want_rstd=False)
        return y.view(shp).to(x.dtype)


class DiTBlock(nn.Module):
    """model.py:44-94: identical parameters and init (incl. zero-init adaLN head)."""

    def __init__(self, hidden_size, cross_attn_input_size, num_heads, mlp_ratio=4.0, qkv_bias=True,
                 residual_v=False):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_heads = num_heads
        self.head_dim = hidden_size // num_heads
        self.scale = self.head_dim ** -0.5
        self.residual_v = residual_v

        self.norm1 = RMSNorm(hidden_size, trainable=qkv_bias)
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_bias)
        self.attn_proj = nn.Linear(hidden_size, hidden_size, bias=False)
        if residual_v:
            self.lambda_param = nn.Parameter(torch.tensor(0.5).reshape(1))
        if cross_attn_input_size is not None:
            self.norm2 = RMSNorm(hidden_size, trainable=qkv_bias)
            self.q_cross = nn.Linear(hidden_size, hidden_size, bias=qkv_bias)
            self.context_kv = nn.Linear(cross_attn_input_size, hidden_size * 2, bias=qkv_bias)
            self.cross_proj = nn.Linear(hidden_size, hidden_size, bias=False)
        else:
            self.norm2 = None
            self.q_cross = None
            self.context_kv = None
            self.cross_proj = None
        self.norm3 = RMSNorm(hidden_size, trainable=qkv_bias)
        mlp_hidden = int(hidden_size * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(hidden_size, mlp_hidden), nn.GELU(), nn.Linear(mlp_hidden, hidden_size))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 9 * hidden_size, bias=True))
        self.adaLN_modulation[-1].weight.data.zero_()
        self.adaLN_modulation[-1].bias.data.zero_()

    def forward(self, *a, **k):
        raise RuntimeError("vds_b200.DiTBlock is executed by DiT.forward (fused engine), not called on its own")


class PatchEmbed(nn.Module):
    """model.py:170-186: Conv3d container (weight [h, C, pt, p, p]); executed as patch-gather + tcgen05 GEMM."""

    def __init__(self, patch_size=16, in_channels=3, embed_dim=768, time_patch_size=16):
        super().__init__()
        self.patch_proj = nn.Conv3d(in_channels, embed_dim, kernel_size=(time_patch_size, patch_size, patch_size),
                                    stride=(time_patch_size, patch_size, patch_size))
        self.patch_size = patch_size
        self.time_patch_size = time_patch_size

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("vds_b200.PatchEmbed has no CPU path")
        B = x.shape[0]
        w = self.patch_proj.weight.to(torch.bfloat16)
        A = ops.patchify(x.to(torch.bfloat16).contiguous(), self.patch_size, self.time_patch_size)
        y = ops.gemm(A, w.view(w.shape[0], -1).contiguous(), bias=self.patch_proj.bias.to(torch.bfloat16))
        return y.view(B, -1, w.shape[0])


class ThreeDimRotary(nn.Module):
    """model.py:189-263: same persistent tables (they are part of the state_dict, model.py:216-217)."""

    def __init__(self, dim, base=100, h=128, w=128, t=128):
        super().__init__()
        self.inv_freq_space = 1.0 / (base ** (torch.arange(0, dim, 4).float() / dim))
        self.inv_freq_time = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.h, self.w, self.t = h, w, t
        t_h = torch.arange(h).type_as(self.inv_freq_space)
        t_w = torch.arange(w).type_as(self.inv_freq_space)
        t_t = torch.arange(t).type_as(self.inv_freq_time)
        freqs_h = torch.outer(t_h, self.inv_freq_space).reshape(1, h, 1, dim // 4).repeat(t, 1, w, 1)
        freqs_w = torch.outer(t_w, self.inv_freq_space).reshape(1, 1, w, dim // 4).repeat(t, h, 1, 1)
        freqs_t = torch.outer(t_t, self.inv_freq_time).reshape(t, 1, 1, dim // 2).repeat(1, h, w, 1)
        freqs_hwt = torch.cat([freqs_t, freqs_h, freqs_w], 3)
        self.register_buffer("freqs_hwt_cos", freqs_hwt.cos())
        self.register_buffer("freqs_hwt_sin", freqs_hwt.sin())

    def forward(self, x, time_height_width=None, extend_with_register_tokens=0):
        starts = engine.draw_rope_starts(self, time_height_width)
        cos, sin = ops.rope_rows(self.freqs_hwt_cos, self.freqs_hwt_sin, time_height_width, starts,
                                 extend_with_register_tokens)
        return cos[None, None], sin[None, None]


def apply_rotary_emb(x, cos, sin):
    """model.py:266-275 for a [B, nh, L, hd] tensor (standalone helper; the engine fuses this into the
    QKV post-processing kernel)."""
    assert x.ndim == 4
    B, nh, Lr, hd = x.shape
    buf = torch.zeros((B * Lr, 3 * nh * hd), device=x.device, dtype=torch.bfloat16)
    buf[:, :nh * hd] = x.permute(0, 2, 1, 3).reshape(B * Lr, nh * hd)
    ops.qkv_post_fwd(buf, B, Lr, nh * hd, nh, cos=cos.reshape(Lr, -1).float().contiguous(),
                     sin=sin.reshape(Lr, -1).float().contiguous())
    return buf[:, :nh * hd].reshape(B, Lr, nh, hd).permute(0, 2, 1, 3).to(x.dtype)


class DiT(nn.Module):
    """model.py:278-402."""

    def __init__(self, in_channels=4, patch_size=2, time_patch_size=2, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, cross_attn_input_size=128, residual_v=False, train_bias_and_rms=True, use_rope=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.patch_size = patch_size
        self.time_patch_size = time_patch_size
        self.hidden_size = hidden_size
        self.num_heads = num_heads
        self.depth = depth
        self.mlp_ratio = mlp_ratio
        self.use_rope = use_rope

        self.patch_embed = PatchEmbed(patch_size, in_channels, hidden_size, time_patch_size)
        if self.use_rope:
            self.rope = ThreeDimRotary(hidden_size // (2 * num_heads), h=128, w=128, t=128)
        else:
            # the reference allocates this but its forward still calls self.rope (model.py:313-314,364):
            # only use_rope=True is a working configuration there, and here.
            self.positional_embedding = nn.Parameter(torch.zeros(1, 2048, hidden_size))
        self.register_tokens = nn.Parameter(torch.randn(1, 16, hidden_size))
        self.time_embed = nn.Sequential(nn.Linear(hidden_size, 4 * hidden_size), nn.SiLU(),
                                        nn.Linear(4 * hidden_size, hidden_size))
        self.blocks = nn.ModuleList([
            DiTBlock(hidden_size=hidden_size, num_heads=num_heads, mlp_ratio=mlp_ratio,
                     cross_attn_input_size=cross_attn_input_size, residual_v=residual_v, qkv_bias=train_bias_and_rms)
            for _ in range(depth)
        ])
        self.depth = depth
        self.final_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))
        self.final_norm = RMSNorm(hidden_size, trainable=train_bias_and_rms)
        self.final_proj = nn.Linear(hidden_size, patch_size * patch_size * time_patch_size * self.out_channels)
        nn.init.zeros_(self.final_modulation[-1].weight)
        nn.init.zeros_(self.final_modulation[-1].bias)
        nn.init.zeros_(self.final_proj.weight)
        nn.init.zeros_(self.final_proj.bias)
        self.paramstatus = {}
        for n, p in self.named_parameters():
            self.paramstatus[n] = {"shape": p.shape, "requires_grad": p.requires_grad}
        self._ckv_cache = None
        self._flat = None  # set by apply_fsdp (shard.FlatShards): flat master / gathered bf16 / gradient buffers

    def context_kv_cache(self, enabled=True):
        """Forward-only runs (sampling): keep every block's ``context_kv(context)`` ([B*Lc, 2h] bf16) keyed by the context
        tensor, so the K = 4096 GEMM runs once per prompt instead of once per denoising step and block.  Results are
        bit-identical (same kernel, same inputs).  Call with ``False`` (or change the weights) to drop the cache."""
        self._ckv_cache = {} if enabled else None
        return self

    def _param_view(self):
        """bf16 compute parameters: views into the gathered flat buffers when sharded, else (cast) copies."""
        if self._flat is not None:
            return engine.ParamView(self._flat.compute_params(), self.depth, self._flat)
        return engine.ParamView(engine.bf16_params(self), self.depth)

    # ------------------------------------------------------------------------------------------
    def forward(self, x, context, timesteps):
        if not x.is_cuda:
            raise RuntimeError(
                "vds_b200.DiT runs only on a CUDA (sm_100a) device: there is no CPU / eager fallback")
        if not self.use_rope:
            raise RuntimeError("use_rope=False is not a working configuration (reference model.py:364)")
        assert self.hidden_size // self.num_heads == 128, "attention kernels are built for head_dim 128"
        params = [p for _, p in self.named_parameters()]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        out = engine.DiTFunction.apply(self, need, x, context, timesteps, *params)
        return out.to(x.dtype) if x.dtype != out.dtype else out

    # ------------------------------------------------------------------------------------------
    def get_mup_setup(self, learning_rate, weight_decay, constant_param_classes):
        """model.py:404-465 (same grouping rules, evaluated on the FULL shapes in ``paramstatus``)."""
        no_decay_name_list = ["bias", "norm", "lambda"]
        custom_lr_multipliers = {"bias": 0.01, "norm": 0.01, "lambda": 0.01}
        final_optimizer_settings = {}
        param_groups = defaultdict(lambda: {"params": [], "weight_decay": None, "lr": None})
        for n, p in self.named_parameters():
            n = n.replace("_fsdp_wrapped_module.", "")
            status = self.paramstatus[n]
            if status["requires_grad"]:
                if any(ndnl in n for ndnl in no_decay_name_list):
                    for ndnl in no_decay_name_list:
                        if ndnl in n:
                            lr_value = learning_rate * custom_lr_multipliers[ndnl]
                            break
                    per_layer_weight_decay_value = 0.0
                else:
                    hidden_dim = status["shape"][-1]
                    lr_value = learning_rate * (32 / hidden_dim)
                    per_layer_weight_decay_value = weight_decay * hidden_dim / 1024
                if any(cls in n for cls in constant_param_classes):
                    lr_value = learning_rate * 0.01
                    per_layer_weight_decay_value = 0.0
                if "time" in n:
                    lr_value = learning_rate * 0.1
                if "modulation" in n:
                    lr_value = learning_rate * 0.1
                group_key = (lr_value, per_layer_weight_decay_value)
                param_groups[group_key]["params"].append(p)
                param_groups[group_key]["weight_decay"] = per_layer_weight_decay_value
                param_groups[group_key]["lr"] = lr_value
                final_optimizer_settings[n] = {"lr": lr_value, "wd": per_layer_weight_decay_value,
                                               "shape": status["shape"]}
        return [v for v in param_groups.values()], final_optimizer_settings


from .shard import apply_fsdp, get_device_mesh  # noqa: E402,F401  (model.py:475-542 surface)


def get_module(module, access_string):
    return reduce(getattr, access_string.split(sep="."), module)


def set_module(module, access_string, value):
    names = access_string.split(sep=".")
    setattr(reduce(getattr, names[:-1], module), names[-1], value)
