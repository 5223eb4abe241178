"""Hand-derived forward / backward of the whole DiT on top of the C-ABI kernels.

Mirrors, op for op, the arithmetic of the reference ``DiT.forward`` (/root/reference/model.py:358-402)
and ``DiTBlock.forward`` (model.py:96-167), including its bf16 rounding points, the "(h w t)" token
order, the "(t h w)" RoPE row order and the three global-CPU-RNG draws (model.py:224-226).  The
backward is written by hand (no autograd inside): every activation that the reference's autograd
would save is stashed here explicitly, and parameter gradients are accumulated in fp32 by the wgrad
GEMM epilogue (split-K ``red.global.add``).
"""
import torch

from . import lib as L
from . import ops

import os

N_REG = 16  # register tokens (model.py:316,362)
# RoPE + value residual inside the QKV GEMM epilogue (VDS_EPI_QKV_ROPE); VDS_FUSE_QKV_ROPE=0: plain GEMM + qkv_post_fwd pass
FUSE_QKV_ROPE = os.environ.get("VDS_FUSE_QKV_ROPE", "1") != "0"


class ParamView:
    """bf16 compute views of the parameters, keyed by the reference state_dict names."""

    def __init__(self, named, depth, flat=None):
        self.p = named
        self.depth = depth
        self.flat = flat  # shard.FlatShards: per-group "parameters gathered" events to wait on

    def wait_group(self, g):
        if self.flat is not None:
            self.flat.wait_group(g)

    def get(self, name):
        return self.p.get(name)

    def __getitem__(self, name):
        return self.p[name]


def bf16_params(model):
    """name -> contiguous bf16 tensor (cast when the module holds fp32 master weights, the FSDP2
    MixedPrecisionPolicy(param_dtype=bf16) behaviour of model.py:516-519)."""
    out = {}
    for n, p in model.named_parameters():
        t = p.detach()
        if t.dtype != torch.bfloat16:
            t = t.to(torch.bfloat16)
        out[n] = t.contiguous()
    return out


def _chunks(mod, h):
    return [mod[:, i * h:(i + 1) * h] for i in range(mod.shape[1] // h)]


def draw_rope_starts(rope, thw):
    """Three draws on the global CPU generator in the reference's order h, w, t (model.py:224-226)."""
    this_t, this_h, this_w = thw
    start_h = torch.randint(0, rope.h - this_h + 1, (1,)).item()
    start_w = torch.randint(0, rope.w - this_w + 1, (1,)).item()
    start_t = torch.randint(0, rope.t - this_t + 1, (1,)).item()
    return start_t, start_h, start_w


class Ctx:
    pass


def forward(model, P, x, context, timesteps, save=True, rope_starts=None, noise=None, rope_starts_dev=None):
    with ops.pinned_stream():
        return _forward(model, P, x, context, timesteps, save, rope_starts, noise, rope_starts_dev)


def _forward(model, P, x, context, timesteps, save=True, rope_starts=None, noise=None, rope_starts_dev=None):
    """Returns (out [B,C,T,H,W] bf16, ctx or None).  With `noise`, `x` is the clean latent and
    z_t = x*(1-t) + noise*t (train.py:115-116) is formed inside the patch gather."""
    dev = x.device
    B, C, T, H, W = x.shape
    p, pt = model.patch_size, model.time_patch_size
    h, nh, depth = model.hidden_size, model.num_heads, model.depth
    hd = h // nh
    Tp, Hp, Wp = T // pt, H // p, W // p
    N = Tp * Hp * Wp
    Lr = N + N_REG
    has_cross = model.blocks[0].context_kv is not None
    residual_v = model.blocks[0].residual_v
    x = x.to(torch.bfloat16).contiguous()
    t_bf = timesteps.to(device=dev, dtype=torch.bfloat16).contiguous()
    if has_cross:
        ctx2d = context.to(torch.bfloat16).contiguous().view(-1, context.shape[-1])
        ctx_is_callers = ctx2d.data_ptr() == context.data_ptr()   # no temporary copy was made
        Lc = context.shape[1]

    c = Ctx()
    c.shape = (B, C, T, H, W)
    c.dims = (N, Lr, h, nh, hd)

    P.wait_group(depth)  # root group (patch embed, time embed, final head)
    # ---- patch embed (+ register tokens) : a2, a3
    A = ops.patchify(x, p, pt, noise=noise, t=t_bf if noise is not None else None)
    X = torch.empty((B * Lr, h), device=dev, dtype=torch.bfloat16)
    X.view(B, Lr, h)[:, :N_REG] = P["register_tokens"]
    ops.gemm(A, P["patch_embed.patch_proj.weight"].view(h, -1), bias=P["patch_embed.patch_proj.bias"], out=X,
             remap=(N, Lr, N_REG))

    # ---- RoPE rows : a4
    if rope_starts_dev is not None:
        rope_starts = (0, 0, 0)          # graph capture: the offsets are read from device memory at replay time
    elif rope_starts is None:
        rope_starts = draw_rope_starts(model.rope, (Tp, Hp, Wp))
    cos, sin = ops.rope_rows(model.rope.freqs_hwt_cos, model.rope.freqs_hwt_sin, (Tp, Hp, Wp), rope_starts, N_REG,
                             starts_dev=rope_starts_dev)
    rope_tab = ops.rope_pack(cos, sin) if FUSE_QKV_ROPE else None   # what the fused QKV epilogue reads by TMA

    # ---- time embedding : a5
    temb0 = ops.timestep_embedding(t_bf, h)
    te_h = ops.gemm(temb0, P["time_embed.0.weight"], bias=P["time_embed.0.bias"])
    te_a = ops.silu(te_h)
    cvec = ops.gemm(te_a, P["time_embed.2.weight"], bias=P["time_embed.2.bias"])
    sc = ops.silu(cvec)  # SiLU(c) is shared by every adaLN head (model.py:90,340)

    blocks = []
    v0 = None
    # Training with flat (sharded) parameters: context_kv(context) of G consecutive blocks is ONE GEMM on the stacked
    # [G*2h, Dc] weight (shard.Layout: ckv groups) — every block applies its own Linear to the SAME caption embedding
    # (model.py:149-155), so the grouping is result-identical and turns depth small-N GEMMs into a few chip-filling ones.
    lay = P.flat.layout if (P.flat is not None and has_cross) else None
    grouped_ckv = lay is not None and lay.ckv_group > 0 and getattr(model, "_ckv_cache", None) is None
    ckv_g = None
    for i in range(depth):
        pre = f"blocks.{i}."
        s = Ctx()
        if grouped_ckv:
            gi, gj, gnb = lay.ckv_of_block(i)
            if gj == 0:
                P.wait_group(depth + 1 + gi)
                Wg, bg, _, _ = P.flat.ckv_views(gi)
                ckv_g = ops.gemm(ctx2d, Wg, bias=bg)        # [B*Lc, nb*2h]
        P.wait_group(i)
        mod = ops.gemm(sc, P[pre + "adaLN_modulation.1.weight"], bias=P[pre + "adaLN_modulation.1.bias"])
        shift_sa, scale_sa, gate_sa, shift_ca, scale_ca, gate_ca, shift_mlp, scale_mlp, gate_mlp = _chunks(mod, h)
        # self attention
        n1, rstd1 = ops.rmsnorm_mod_fwd(X, B, Lr, h, scale=scale_sa, shift=shift_sa, weight=P.get(pre + "norm1.weight"),
                                        want_rstd=save)
        use_mix = residual_v and v0 is not None
        # RoPE + value residual in the QKV GEMM's epilogue (2-CTA tile path); small shapes: plain GEMM + in-place pass
        fused = None
        if rope_tab is not None:
            fused = ops.gemm_qkv_rope(n1, P[pre + "qkv.weight"], P.get(pre + "qkv.bias"), rope_tab, Lr,
                                      v0=v0 if use_mix else None, v0_ld=v0.stride(0) if use_mix else 0,
                                      lam=P.get(pre + "lambda_param") if use_mix else None)
        if fused is not None:
            qkv, vmix = fused
        else:
            qkv = ops.gemm(n1, P[pre + "qkv.weight"], bias=P.get(pre + "qkv.bias"))
            vmix = ops.qkv_post_fwd(qkv, B, Lr, h, nh, cos=cos, sin=sin, v0=v0 if use_mix else None,
                                    v0_ld=v0.stride(0) if use_mix else 0,
                                    lam=P.get(pre + "lambda_param") if use_mix else None)
        V = vmix if use_mix else qkv[:, 2 * h:]
        if v0 is None:
            v0 = V
        a, lse = ops.attn_fwd(qkv[:, :h], qkv[:, h:2 * h], V, B, nh, Lr, Lr, want_lse=save)
        o1, X1 = ops.gemm(a, P[pre + "attn_proj.weight"], epilogue=L.EPI_GATE_RES, aux=X, gate=gate_sa,
                          rows_per_batch=Lr)
        # cross attention
        if has_cross:
            n2, rstd2 = ops.rmsnorm_mod_fwd(X1, B, Lr, h, scale=scale_ca, shift=shift_ca,
                                            weight=P.get(pre + "norm2.weight"), want_rstd=save)
            qc = ops.gemm(n2, P[pre + "q_cross.weight"], bias=P.get(pre + "q_cross.bias"))
            ckv = None
            # inference: context_kv(context) depends only on the prompt, not on the step (§8f n2).  The entry keeps the
            # context tensor itself alive and a hit needs the SAME tensor object at the same version: a freed temporary
            # whose address the allocator hands to the next prompt can never alias an entry.  When the engine had to
            # make its own bf16 / contiguous copy of `context` the copy dies with this call, so the cache is bypassed.
            cache = getattr(model, "_ckv_cache", None) if (not save and ctx_is_callers) else None
            if cache is not None:
                ent = cache.get((i, id(context)))   # the entry holds `context`, so its id cannot be recycled
                if ent is not None and ent[0] is context and ent[1] == context._version:
                    ckv = ent[2]
            if ckv is None and grouped_ckv:
                ckv = ckv_g[:, gj * 2 * h:(gj + 1) * 2 * h]      # this block's columns of the grouped output
            if ckv is None:
                ckv = ops.gemm(ctx2d, P[pre + "context_kv.weight"], bias=P.get(pre + "context_kv.bias"))
                if cache is not None:
                    cache[(i, id(context))] = (context, context._version, ckv)
            ca, lse2 = ops.attn_fwd(qc, ckv[:, :h], ckv[:, h:], B, nh, Lr, Lc, want_lse=save)
            o2, X2 = ops.gemm(ca, P[pre + "cross_proj.weight"], epilogue=L.EPI_GATE_RES, aux=X1, gate=gate_ca,
                              rows_per_batch=Lr)
        else:
            X2 = X1
        # MLP
        n3, rstd3 = ops.rmsnorm_mod_fwd(X2, B, Lr, h, scale=scale_mlp, shift=shift_mlp,
                                        weight=P.get(pre + "norm3.weight"), want_rstd=save)
        h1, g = ops.gemm(n3, P[pre + "mlp.0.weight"], bias=P[pre + "mlp.0.bias"], epilogue=L.EPI_BIAS_GELU)
        o3, X3 = ops.gemm(g, P[pre + "mlp.2.weight"], bias=P[pre + "mlp.2.bias"], epilogue=L.EPI_GATE_RES, aux=X2,
                          gate=gate_mlp, rows_per_batch=Lr)
        if save:
            s.mod, s.X0, s.rstd1, s.n1, s.qkv, s.V, s.use_mix, s.a, s.lse, s.o1 = mod, X, rstd1, n1, qkv, V, use_mix, a, lse, o1
            s.X1 = X1
            if has_cross:
                s.rstd2, s.n2, s.qc, s.ckv, s.ca, s.lse2, s.o2 = rstd2, n2, qc, ckv, ca, lse2, o2
            s.X2, s.rstd3, s.n3, s.h1, s.g, s.o3 = X2, rstd3, n3, h1, g, o3
            blocks.append(s)
        X = X3

    # ---- final head : a15, a16
    fmod = ops.gemm(sc, P["final_modulation.1.weight"], bias=P["final_modulation.1.bias"])
    f_shift, f_scale = fmod[:, :h], fmod[:, h:]
    nf, rstdf = ops.rmsnorm_mod_fwd(X, B, N, h, scale=f_scale, shift=f_shift, weight=P.get("final_norm.weight"),
                                    in_batch_stride=Lr, in_row_offset=N_REG, want_rstd=save)
    y = ops.gemm(nf, P["final_proj.weight"], bias=P["final_proj.bias"])
    out = ops.unpatchify(y, B, C, T, H, W, p, pt)
    if not save:
        return out, None
    c.A, c.cos, c.sin, c.temb0, c.te_h, c.te_a, c.cvec, c.sc = A, cos, sin, temb0, te_h, te_a, cvec, sc
    c.blocks, c.v0, c.Xf, c.fmod, c.nf, c.rstdf = blocks, v0, X, fmod, nf, rstdf
    c.ctx2d = ctx2d if has_cross else None
    c.Lc = Lc if has_cross else 0
    c.has_cross, c.residual_v = has_cross, residual_v
    c.grouped_ckv = grouped_ckv
    return out, c


def wgrad_splits(n_out, n_in, rows, sms=148):
    """Split-K factor of a wgrad GEMM dW[n_out, n_in] += dy[rows, n_out]^T x[rows, n_in] (contraction over the tokens).

    Host-side cost model of the two tile paths of vds_gemm, fitted to scripts/wgrad_splits_bench.py on a B200: a work item
    (one output tile x one k-range) costs 0.25 us per 64-row k-block — a 128 x 128 tile on one SM and a 256 x 256 tile on
    an SM pair take the same time per k-block — plus a fixed 5.3 us (1-CTA) / 6.3 us (2-CTA: fp32 reduction of a 4 x larger
    tile); the 2-CTA path is taken when tiles x splits fill half the SM pairs (gemm.cu: vds_gemm).  Best is usually the
    factor that makes ~148 items: one round of 1-CTA tiles or two rounds of pair tiles."""
    kb = max(1, (rows + 63) // 64)
    t128 = ((n_out + 127) // 128) * ((n_in + 127) // 128)
    pt = ((n_out + 255) // 256) * ((n_in + 255) // 256)
    two_cta_ok = n_in % 64 == 0 and n_in >= 256
    best, best_cost = 1, None
    for s in range(1, min(24, kb) + 1):
        per = -(-kb // s)
        if two_cta_ok and pt * s >= sms // 2:
            cost = -(-pt * s // (sms // 2)) * (per * 0.25 + 6.3)
        else:
            cost = -(-t128 * s // sms) * (per * 0.25 + 5.3)
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = s, cost
    return best


class GradSink:
    """Where parameter gradients go: name -> fp32 tensor of the parameter's shape (accumulated into)."""

    def __init__(self, model, dev, buffers=None):
        self.g = buffers if buffers is not None else {}
        self.model = model
        self.dev = dev
        self._shapes = getattr(model, "_full_shapes", None) or {n: p.shape for n, p in model.named_parameters()}
        self.on_block_done = None
        self.on_ckv_done = None

    def buf(self, name):
        t = self.g.get(name)
        if t is None:
            t = torch.zeros(self._shapes[name], device=self.dev, dtype=torch.float32)
            self.g[name] = t
        return t

    def wgrad(self, name, dy, x, rows=None):
        """dW[out,in] += dy[rows,out]^T x[rows,in]"""
        w = self.buf(name)
        w2 = w.view(w.shape[0], -1)
        K = dy.shape[0] if rows is None else rows
        splits = wgrad_splits(w2.shape[0], w2.shape[1], K)
        ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=w2, splits=splits, K=K)

    def bgrad(self, name, dy):
        if name in self._shapes:
            ops.colsum(dy, self.buf(name))


def _attn_q_splits(n_kv_tiles, B, nh, n_q_tiles, sms=148):
    """Query-range splits for an attention backward with fewer (kv tile, head, batch) items than SMs.

    Cost model in units of one 64-row query sub-tile iteration: waves x (sub-tiles per CTA + fixed prologue /
    epilogue of about 5 iterations: K/V fill, TMEM alloc, fp32 red of dK/dV)."""
    ctas = n_kv_tiles * B * nh
    if ctas >= sms:
        return 1
    n_sub = 2 * n_q_tiles
    best, best_cost = 1, None
    for qs in range(1, min(n_sub, 64) + 1):
        waves = (ctas * qs + sms - 1) // sms
        cost = waves * ((n_sub + qs - 1) // qs + 5)
        if best_cost is None or cost < best_cost:
            best, best_cost = qs, cost
    return best


class ZeroArena:
    """The zero-initialised fp32 accumulation targets of ONE block's backward (dmod, the two delta vectors, the two
    fp32 dq buffers, the split-query dk/dv buffer) carved out of one allocation, so a block costs one fill launch
    instead of six ``torch.zeros`` (0.7 ms / step of FillFunctor launches at debug-8k).  The arena is reused by every
    block of a backward pass: all work is stream-ordered, block i's readers are enqueued before block i-1's fill."""

    def __init__(self, dev, sizes):
        self.off, n = {}, 0
        for name, numel in sizes.items():
            self.off[name] = (n, numel)
            n += (numel + 63) // 64 * 64          # 256-byte aligned slices (TMA reduce target needs 16)
        self.buf = torch.empty(max(n, 64), device=dev, dtype=torch.float32)

    def reset(self):
        self.buf.zero_()

    def get(self, name, shape):
        o, numel = self.off[name]
        return self.buf[o:o + numel].view(shape)


def backward(model, P, c, dout, sink):
    """Accumulates every parameter gradient into `sink`; returns nothing (x / context / t get no grad)."""
    B, C, T, H, W = c.shape
    N, Lr, h, nh, hd = c.dims
    p, pt = model.patch_size, model.time_patch_size
    dev = dout.device
    depth = model.depth
    dout = dout.to(torch.bfloat16).contiguous()
    f32 = dict(device=dev, dtype=torch.float32)

    dy = ops.unpatchify(dout, B, C, T, H, W, p, pt, to_tokens=True)
    sink.bgrad("final_proj.bias", dy)
    sink.wgrad("final_proj.weight", dy, c.nf)
    dnf = ops.gemm(dy, P["final_proj.weight"], b_mn=True)
    dfmod = torch.zeros((B, 2 * h), **f32)
    dX = torch.zeros((B * Lr, h), device=dev, dtype=torch.bfloat16)
    has_fw = "final_norm.weight" in P.p
    ops.rmsnorm_mod_bwd(dnf, c.Xf, c.rstdf, B, N, h, scale=c.fmod[:, h:], weight=P.get("final_norm.weight"), dx=dX,
                        dscale=dfmod[:, h:], dshift=dfmod[:, :h],
                        dweight=sink.buf("final_norm.weight") if has_fw else None, in_batch_stride=Lr,
                        in_row_offset=N_REG, dx_full_rows=True)
    dsc_acc = torch.zeros((B, h), **f32)
    dfmod_b = ops.cast_f32_bf16(dfmod)
    sink.wgrad("final_modulation.1.weight", dfmod_b, c.sc)
    sink.bgrad("final_modulation.1.bias", dfmod_b)
    ops.gemm(dfmod_b, P["final_modulation.1.weight"], b_mn=True, epilogue=L.EPI_ACCUM_F32, out=dsc_acc)

    dv0_acc = torch.zeros((B * Lr, h), **f32) if (c.residual_v and depth > 1) else None
    sizes = {"dmod": B * 9 * h, "delta1": B * nh * Lr, "dq_self": B * Lr * h}
    qs = 1
    if c.has_cross:
        qs = _attn_q_splits((c.Lc + 127) // 128, B, nh, (Lr + 127) // 128)
        sizes.update({"delta2": B * nh * Lr, "dq_cross": B * Lr * h})
        if qs > 1:
            sizes["dckv_f"] = B * c.Lc * 2 * h
    arena = ZeroArena(dev, sizes)
    lay = P.flat.layout if (P.flat is not None and c.has_cross) else None
    dckv_g = None
    for i in reversed(range(depth)):
        pre = f"blocks.{i}."
        s = c.blocks[i]
        shift_sa, scale_sa, gate_sa, shift_ca, scale_ca, gate_ca, shift_mlp, scale_mlp, gate_mlp = _chunks(s.mod, h)
        arena.reset()
        dmod = arena.get("dmod", (B, 9 * h))
        dm = _chunks(dmod, h)
        # ---- MLP branch
        do3 = ops.gate_bwd(dX, s.o3, gate_mlp, dm[8], B, Lr, h)
        sink.bgrad(pre + "mlp.2.bias", do3)
        sink.wgrad(pre + "mlp.2.weight", do3, s.g)
        dh1 = ops.gemm(do3, P[pre + "mlp.2.weight"], b_mn=True, epilogue=L.EPI_DGELU, aux=s.h1)
        sink.bgrad(pre + "mlp.0.bias", dh1)
        sink.wgrad(pre + "mlp.0.weight", dh1, s.n3)
        dn3 = ops.gemm(dh1, P[pre + "mlp.0.weight"], b_mn=True)
        w3 = P.get(pre + "norm3.weight")
        dX2 = ops.rmsnorm_mod_bwd(dn3, s.X2, s.rstd3, B, Lr, h, scale=scale_mlp, weight=w3, dx_res=dX, dscale=dm[7],
                                  dshift=dm[6], dweight=sink.buf(pre + "norm3.weight") if w3 is not None else None)
        # ---- cross-attention branch
        if c.has_cross:
            Lc = c.Lc
            do2 = ops.gate_bwd(dX2, s.o2, gate_ca, dm[5], B, Lr, h)
            sink.wgrad(pre + "cross_proj.weight", do2, s.ca)
            delta2 = arena.get("delta2", (B, nh, Lr))
            dca = ops.gemm_dgrad_rowdot(do2, P[pre + "cross_proj.weight"], s.ca, delta2, Lr)
            if dca is None:
                dca, delta2 = ops.gemm(do2, P[pre + "cross_proj.weight"], b_mn=True), None
            dq_acc = arena.get("dq_cross", (B * Lr, h))
            if c.grouped_ckv:   # this block's column slice of the group's [B*Lc, nb*2h] gradient (one wgrad per group)
                gi, gj, gnb = lay.ckv_of_block(i)
                if dckv_g is None:
                    dckv_g = torch.empty((B * Lc, gnb * 2 * h), device=dev, dtype=torch.bfloat16)
                dckv = dckv_g[:, gj * 2 * h:(gj + 1) * 2 * h]
            else:
                dckv = torch.empty((B * Lc, 2 * h), device=dev, dtype=torch.bfloat16)
            if qs > 1:
                dckv_f = arena.get("dckv_f", (B * Lc, 2 * h))
                ops.attn_bwd(s.qc, s.ckv[:, :h], s.ckv[:, h:], s.ca, dca, s.lse2, B, nh, Lr, Lc, dq_acc,
                             dk_acc=dckv_f[:, :h], dv_acc=dckv_f[:, h:], q_splits=qs, delta=delta2)
                ops.cast_f32_bf16_2d(dckv_f, dckv)
            else:
                ops.attn_bwd(s.qc, s.ckv[:, :h], s.ckv[:, h:], s.ca, dca, s.lse2, B, nh, Lr, Lc, dq_acc,
                             dk=dckv[:, :h], dv=dckv[:, h:], delta=delta2)
            dqc = ops.cast_f32_bf16(dq_acc)
            if c.grouped_ckv:
                if gj == 0:     # first block of the group = last one in backward order: the group's gradient is complete
                    _, _, gWg, gbg = P.flat.ckv_views(gi)
                    Kc = B * Lc
                    ops.gemm(dckv_g, c.ctx2d, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gWg,
                             splits=max(1, min(16, Kc // 2048)), K=Kc)
                    if gbg is not None:
                        ops.colsum(dckv_g, gbg)
                    dckv_g = None
                    if sink.on_ckv_done is not None:
                        sink.on_ckv_done(gi)
            else:
                sink.wgrad(pre + "context_kv.weight", dckv, c.ctx2d)
                sink.bgrad(pre + "context_kv.bias", dckv)
            sink.wgrad(pre + "q_cross.weight", dqc, s.n2)
            sink.bgrad(pre + "q_cross.bias", dqc)
            dn2 = ops.gemm(dqc, P[pre + "q_cross.weight"], b_mn=True)
            w2 = P.get(pre + "norm2.weight")
            dX1 = ops.rmsnorm_mod_bwd(dn2, s.X1, s.rstd2, B, Lr, h, scale=scale_ca, weight=w2, dx_res=dX2,
                                      dscale=dm[4], dshift=dm[3],
                                      dweight=sink.buf(pre + "norm2.weight") if w2 is not None else None)
        else:
            dX1 = dX2
        # ---- self-attention branch
        do1 = ops.gate_bwd(dX1, s.o1, gate_sa, dm[2], B, Lr, h)
        sink.wgrad(pre + "attn_proj.weight", do1, s.a)
        delta1 = arena.get("delta1", (B, nh, Lr))
        da = ops.gemm_dgrad_rowdot(do1, P[pre + "attn_proj.weight"], s.a, delta1, Lr)   # dO and delta = rowsum(dO * O)
        if da is None:
            da, delta1 = ops.gemm(do1, P[pre + "attn_proj.weight"], b_mn=True), None
        dqkv = torch.empty((B * Lr, 3 * h), device=dev, dtype=torch.bfloat16)
        dq_acc = arena.get("dq_self", (B * Lr, h))
        ops.attn_bwd(s.qkv[:, :h], s.qkv[:, h:2 * h], s.V, s.a, da, s.lse, B, nh, Lr, Lr, dq_acc, dk=dqkv[:, h:2 * h],
                     dv=dqkv[:, 2 * h:], delta=delta1)
        if s.use_mix:
            mode = 1
        elif i == 0 and dv0_acc is not None:
            mode = 2
        else:
            mode = 0
        ops.qkv_post_bwd(dqkv, B, Lr, h, nh, dq_acc=dq_acc, cos=c.cos, sin=c.sin, qkv_pre=s.qkv,
                         v0=c.v0 if mode == 1 else None, v0_ld=c.v0.stride(0) if mode == 1 else 0,
                         lam=P.get(pre + "lambda_param") if mode == 1 else None,
                         dlambda=sink.buf(pre + "lambda_param") if mode == 1 else None, dv0_acc=dv0_acc, mode=mode)
        sink.wgrad(pre + "qkv.weight", dqkv, s.n1)
        sink.bgrad(pre + "qkv.bias", dqkv)
        dn1 = ops.gemm(dqkv, P[pre + "qkv.weight"], b_mn=True)
        w1 = P.get(pre + "norm1.weight")
        dX = ops.rmsnorm_mod_bwd(dn1, s.X0, s.rstd1, B, Lr, h, scale=scale_sa, weight=w1, dx_res=dX1, dscale=dm[1],
                                 dshift=dm[0], dweight=sink.buf(pre + "norm1.weight") if w1 is not None else None)
        # ---- adaLN modulation head
        dmod_b = ops.cast_f32_bf16(dmod)
        sink.wgrad(pre + "adaLN_modulation.1.weight", dmod_b, c.sc)
        sink.bgrad(pre + "adaLN_modulation.1.bias", dmod_b)
        ops.gemm(dmod_b, P[pre + "adaLN_modulation.1.weight"], b_mn=True, epilogue=L.EPI_ACCUM_F32, out=dsc_acc)
        c.blocks[i] = None  # free this block's activations
        if sink.on_block_done is not None:
            sink.on_block_done(i)

    # ---- register tokens + patch embed
    ops.batch_rowsum(dX, sink.buf("register_tokens").view(N_REG, h), B, Lr * h, N_REG, h)
    wname = "patch_embed.patch_proj.weight"
    gw = sink.buf(wname).view(h, -1)
    dXv = dX.view(B, Lr, h)
    for b in range(B):
        dxb = dXv[b, N_REG:]
        ops.gemm(dxb, c.A[b * N:(b + 1) * N], a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw,
                 splits=max(1, min(16, N // 2048)))
        ops.colsum(dxb, sink.buf("patch_embed.patch_proj.bias"))
    # ---- time embedding MLP
    dsc = ops.cast_f32_bf16(dsc_acc)
    dc = ops.silu_bwd(c.cvec, dsc)
    sink.wgrad("time_embed.2.weight", dc, c.te_a)
    sink.bgrad("time_embed.2.bias", dc)
    da1 = ops.gemm(dc, P["time_embed.2.weight"], b_mn=True)
    dh = ops.silu_bwd(c.te_h, da1)
    sink.wgrad("time_embed.0.weight", dh, c.temb0)
    sink.bgrad("time_embed.0.bias", dh)


def run_backward(model, P, c, dout, dtypes):
    """Shared by DiTFunction / TrainStepFunction: returns the per-parameter grads tuple for autograd (or Nones
    when the model owns flat gradient buffers, see shard.py)."""
    names = [n for n, _ in model.named_parameters()]
    flat = getattr(model, "_flat", None)
    if flat is not None:
        flat.begin_backward()
        sink = GradSink(model, dout.device, flat.grad_views)
        sink.on_block_done = flat.block_backward_done
        sink.on_ckv_done = flat.ckv_backward_done
        with ops.pinned_stream():
            backward(model, P, c, dout, sink)
        flat.end_backward()
        return tuple(None for _ in names)
    sink = GradSink(model, dout.device)
    with ops.pinned_stream():
        backward(model, P, c, dout, sink)
    grads = []
    for n, dt in zip(names, dtypes):
        g = sink.g.get(n)
        grads.append(g.to(dt) if g is not None else None)
    return tuple(grads)


class DiTFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward + hand-written backward."""

    @staticmethod
    def forward(ctx, model, need, x, context, timesteps, *params):
        P = model._param_view()
        out, c = forward(model, P, x, context, timesteps, save=need)
        ctx.model, ctx.P, ctx.c = model, P, c
        ctx.dtypes = [p.dtype for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = run_backward(ctx.model, ctx.P, ctx.c, dout, ctx.dtypes)
        ctx.c = None
        return (None, None, None, None, None) + grads
