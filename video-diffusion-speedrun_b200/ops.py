"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers + stream."""
import ctypes

import torch

from . import lib as L


# bench.py hook: PROFILE["attn_bwd_self"] = [] makes attn_bwd record (start, end) CUDA events on the launch
# stream around every self-attention backward launch (the dominant kernel of the step).
PROFILE = {}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# The engine pins the launch stream for the duration of one forward / backward (saves a
# torch.cuda.current_stream() lookup per launch: ~1100 of them per train step).
_PINNED_STREAM = None


class pinned_stream:
    def __enter__(self):
        global _PINNED_STREAM
        self.prev = _PINNED_STREAM
        _PINNED_STREAM = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        global _PINNED_STREAM
        _PINNED_STREAM = self.prev
        return False


def _stream():
    return ctypes.c_void_p(_PINNED_STREAM if _PINNED_STREAM is not None else torch.cuda.current_stream().cuda_stream)


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.bfloat16, (t.device, t.dtype)


def gemm(a, b, *, a_mn=False, b_mn=False, epilogue=L.EPI_STORE, out=None, out2=None, bias=None,
         aux=None, gate=None, rows_per_batch=0, splits=1, remap=None, M=None, N=None, K=None, tile_n=0, cluster=0):
    """D[M,N] = A * B^T on tcgen05 (see include/vds_b200.h: vds_gemm).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True).
    2-D views with unit inner stride are accepted (leading dim = stride(0)).
    """
    _chk_bf16(a, b, bias, aux, gate)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    assert (b.shape[0] if b_mn else b.shape[1]) == K, (a.shape, b.shape, a_mn, b_mn)
    f32_out = epilogue in (L.EPI_ACCUM_F32, L.EPI_STORE_F32)
    if out is None and epilogue not in (L.EPI_ACCUM_F32,):
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    if epilogue in (L.EPI_BIAS_GELU, L.EPI_GATE_RES) and out2 is None:
        out2 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    args = L.GemmArgs()
    args.A, args.B = a.data_ptr(), b.data_ptr()
    args.lda, args.ldb = a.stride(0), b.stride(0)
    args.M, args.N, args.K = M, N, K
    args.a_mn, args.b_mn = int(a_mn), int(b_mn)
    args.epilogue, args.splits = epilogue, splits
    if out is not None:
        assert out.stride(-1) == 1 and out.dtype == (torch.float32 if f32_out else torch.bfloat16)
        args.C, args.ldc = out.data_ptr(), out.stride(-2)
    if out2 is not None:
        assert out2.stride(-1) == 1
        args.C2, args.ldc2 = out2.data_ptr(), out2.stride(-2)
    if bias is not None:
        args.bias = bias.data_ptr()
    if aux is not None:
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(-2)
    if gate is not None:
        args.gate, args.gate_stride = gate.data_ptr(), gate.stride(0)
    args.rows_per_batch = rows_per_batch
    args.tile_n = tile_n
    args.cluster = cluster
    if remap is not None:
        args.remap_rows, args.remap_stride, args.remap_offset = remap
    rc = L.lib().vds_gemm(ctypes.byref(args), _stream())
    if rc == L.ERR_UNSUPPORTED and epilogue == L.EPI_STORE_ROWDOT:
        return None                      # caller falls back to the plain GEMM + separate row-dot kernel
    L.check(rc, "vds_gemm")
    if epilogue == L.EPI_STORE_ROWDOT:
        return out
    return (out, out2) if out2 is not None else out


def rope_pack(cos, sin):
    """cos / sin rows [L, 64] fp32 -> the packed [L + 32, 4, 32] table the fused QKV epilogue reads by TMA (vds_rope_pack)."""
    Lr = cos.shape[0]
    assert cos.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous() and cos.shape == (Lr, 64) == sin.shape
    tab = torch.empty((Lr + 32, 4, 32), device=cos.device, dtype=torch.float32)
    L.check(L.lib().vds_rope_pack(_p(cos), _p(sin), _p(tab), Lr, _s()), "vds_rope_pack")
    return tab


def gemm_qkv_rope(x, w, bias, tab, rows_per_batch, v0=None, v0_ld=0, lam=None):
    """qkv = x @ W^T (+ bias) with the block's RoPE (q, k heads) and value residual (v heads) applied in the GEMM epilogue
    (model.py:124-134; VDS_EPI_QKV_ROPE).  Returns (qkv, vmix) — qkv[:, 2h:] holds v_pre, vmix is None without v0 — or None
    when the shape has no 2-CTA tile path (the caller then runs the plain GEMM + qkv_post_fwd)."""
    _chk_bf16(x, w, bias, v0, lam)
    M, K = x.shape
    N = w.shape[0]
    assert tab.dtype == torch.float32 and tab.is_contiguous() and tab.shape == (rows_per_batch + 32, 4, 32)
    qkv = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    vmix = torch.empty((M, N // 3), device=x.device, dtype=torch.bfloat16) if v0 is not None else None
    args = L.GemmArgs()
    args.A, args.B, args.lda, args.ldb = x.data_ptr(), w.data_ptr(), x.stride(0), w.stride(0)
    args.M, args.N, args.K = M, N, K
    args.epilogue, args.splits = L.EPI_QKV_ROPE, 1
    args.C, args.ldc = qkv.data_ptr(), qkv.stride(0)
    if vmix is not None:
        args.C2, args.ldc2 = vmix.data_ptr(), vmix.stride(0)
        args.v0, args.ldv0, args.lambda_ = v0.data_ptr(), v0_ld, lam.data_ptr()
    if bias is not None:
        args.bias = bias.data_ptr()
    args.rope_tab = tab.data_ptr()
    args.rows_per_batch = rows_per_batch
    rc = L.lib().vds_gemm(ctypes.byref(args), _stream())
    if rc == L.ERR_UNSUPPORTED:
        return None
    L.check(rc, "vds_gemm")
    return qkv, vmix


def gemm_dgrad_rowdot(dy, w, o, rowdot, rows_per_batch):
    """dX = dy @ W (dgrad, W [N_out, N_in] read MN-major) and, in the same epilogue, rowdot[b, head, r] += <dX_head, o_head>
    per 128-column head — the `delta` of the attention backward.  `rowdot`: zero-initialised fp32 [B, N_in/128,
    rows_per_batch].  Returns dX, or None when the shape has no 2-CTA tile path (the caller then runs the plain
    GEMM and lets vds_attn_bwd compute delta itself)."""
    assert rowdot.dtype == torch.float32 and rowdot.is_contiguous()
    return gemm(dy, w, b_mn=True, epilogue=L.EPI_STORE_ROWDOT, aux=o, out2=rowdot, rows_per_batch=rows_per_batch)


def _p(t):
    return t.data_ptr() if t is not None else None


def _s():
    return _PINNED_STREAM if _PINNED_STREAM is not None else torch.cuda.current_stream().cuda_stream


def patchify(x, p, pt, noise=None, t=None, out=None):
    """[B,C,T,H,W] -> [B*N, C*pt*p*p] in the reference's token / Conv3d-weight order (model.py:173-185);
    with `noise`/`t` also forms z_t = x*(1-t) + noise*t in bf16 (train.py:115-116)."""
    _chk_bf16(x, noise, t)
    B, C, T, H, W = x.shape
    assert x.is_contiguous()
    n = (T // pt) * (H // p) * (W // p)
    if out is None:
        out = torch.empty((B * n, C * pt * p * p), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().vds_patchify(_p(x), _p(noise), _p(t), _p(out), B, C, T, H, W, p, pt, _s()), "vds_patchify")
    return out


def unpatchify(y, B, C, T, H, W, p, pt, to_tokens=False, out=None):
    """tokens [B*N, p*p*pt*C] -> [B,C,T,H,W] (model.py:392-401); to_tokens=True is the inverse gather."""
    _chk_bf16(y)
    n = (T // pt) * (H // p) * (W // p)
    if out is None:
        shape = (B * n, C * pt * p * p) if to_tokens else (B, C, T, H, W)
        out = torch.empty(shape, device=y.device, dtype=torch.bfloat16)
    assert y.is_contiguous() and out.is_contiguous()
    L.check(L.lib().vds_unpatchify(_p(y), _p(out), B, C, T, H, W, p, pt, int(to_tokens), _s()), "vds_unpatchify")
    return out


def rope_rows(tcos, tsin, thw, starts, n_reg, starts_dev=None):
    """Gather cos/sin rows [L, D] (fp32) from the persistent tables at (start_t, start_h, start_w)."""
    Tp, Hp, Wp = thw
    st, sh, sw = starts
    D = tcos.shape[-1]
    Lr = n_reg + Tp * Hp * Wp
    assert tcos.dtype in (torch.float32, torch.bfloat16) and tcos.is_contiguous() and tsin.is_contiguous()
    ocos = torch.empty((Lr, D), device=tcos.device, dtype=torch.float32)
    osin = torch.empty_like(ocos)
    L.check(L.lib().vds_rope_rows(_p(tcos), _p(tsin), int(tcos.dtype == torch.bfloat16), _p(ocos), _p(osin), Lr, D,
                                  n_reg, Tp, Hp, Wp, st, sh, sw, tcos.shape[1], tcos.shape[2], _p(starts_dev), _s()),
            "vds_rope_rows")
    return ocos, osin


def timestep_embedding(t, dim, max_period=10000.0):
    _chk_bf16(t)
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.bfloat16)
    L.check(L.lib().vds_timestep_embedding(_p(t), _p(out), t.shape[0], dim, float(max_period), _s()),
            "vds_timestep_embedding")
    return out


def silu(x):
    _chk_bf16(x)
    y = torch.empty_like(x)
    L.check(L.lib().vds_silu(_p(x), _p(y), x.numel(), _s()), "vds_silu")
    return y


def silu_bwd(x, dy):
    _chk_bf16(x, dy)
    dx = torch.empty_like(x)
    L.check(L.lib().vds_silu_bwd(_p(x), _p(dy), _p(dx), x.numel(), _s()), "vds_silu_bwd")
    return dx


def rmsnorm_mod_fwd(x, B, rows_out, h, scale=None, shift=None, weight=None, in_batch_stride=None, in_row_offset=0,
                    eps=1e-6, want_rstd=True):
    """y = rmsnorm(x)[*w] * (1 + scale[b]) + shift[b] with the reference's bf16 rounding points."""
    _chk_bf16(x, scale, shift, weight)
    if in_batch_stride is None:
        in_batch_stride = rows_out
    y = torch.empty((B * rows_out, h), device=x.device, dtype=torch.bfloat16)
    rstd = torch.empty((B * rows_out,), device=x.device, dtype=torch.float32) if want_rstd else None
    ms = scale.stride(0) if scale is not None else 0
    L.check(L.lib().vds_rmsnorm_mod_fwd(_p(x), _p(y), _p(rstd), _p(weight), _p(scale), _p(shift), ms, B, rows_out,
                                        in_batch_stride, in_row_offset, h, eps, _s()), "vds_rmsnorm_mod_fwd")
    return y, rstd


def rmsnorm_mod_bwd(dy, x, rstd, B, rows_out, h, scale=None, weight=None, dx_res=None, dx=None, dscale=None,
                    dshift=None, dweight=None, in_batch_stride=None, in_row_offset=0, dx_full_rows=False):
    _chk_bf16(dy, x, scale, weight, dx_res)
    if in_batch_stride is None:
        in_batch_stride = rows_out
    if dx is None:
        dx = torch.empty((B * (in_batch_stride if dx_full_rows else rows_out), h), device=x.device,
                         dtype=torch.bfloat16)
    ms = scale.stride(0) if scale is not None else 0
    dms = dscale.stride(0) if dscale is not None else 0
    L.check(L.lib().vds_rmsnorm_mod_bwd(_p(dy), _p(x), _p(rstd), _p(weight), _p(scale), _p(dx_res), _p(dx),
                                        _p(dscale), _p(dshift), _p(dweight), ms, dms, B, rows_out, in_batch_stride,
                                        in_row_offset, int(dx_full_rows), h, _s()), "vds_rmsnorm_mod_bwd")
    return dx


def gate_bwd(dx, o, gate, dgate, B, rows, h):
    """do = dx*gate[b]; dgate[b] += sum_rows dx*o."""
    _chk_bf16(dx, o, gate)
    d_o = torch.empty((B * rows, h), device=dx.device, dtype=torch.bfloat16)
    L.check(L.lib().vds_gate_bwd(_p(dx), _p(o), _p(gate), _p(d_o), _p(dgate), gate.stride(0), dgate.stride(0), B, rows,
                                 h, _s()), "vds_gate_bwd")
    return d_o


def qkv_post_fwd(qkv, B, Lr, h, nh, cos=None, sin=None, v0=None, v0_ld=0, lam=None):
    _chk_bf16(qkv, v0, lam)
    vmix = torch.empty((B * Lr, h), device=qkv.device, dtype=torch.bfloat16) if v0 is not None else None
    L.check(L.lib().vds_qkv_post_fwd(_p(qkv), _p(cos), _p(sin), _p(v0), v0_ld, _p(vmix), _p(lam), B, Lr, h, nh, _s()),
            "vds_qkv_post_fwd")
    return vmix


def qkv_post_bwd(dqkv, B, Lr, h, nh, dq_acc=None, cos=None, sin=None, qkv_pre=None, v0=None, v0_ld=0, lam=None,
                 dlambda=None, dv0_acc=None, mode=0):
    L.check(L.lib().vds_qkv_post_bwd(_p(dqkv), _p(dq_acc), _p(cos), _p(sin), _p(qkv_pre), _p(v0), v0_ld, _p(lam),
                                     _p(dlambda), _p(dv0_acc), mode, B, Lr, h, nh, _s()), "vds_qkv_post_bwd")


def colsum(x, out, rows=None, n=None):
    """out[n] (fp32) += sum_rows x[row, n]"""
    _chk_bf16(x)
    rows = x.shape[0] if rows is None else rows
    n = x.shape[1] if n is None else n
    L.check(L.lib().vds_colsum(_p(x), _p(out), rows, n, x.stride(0), _s()), "vds_colsum")


def batch_rowsum(x, out, B, batch_stride, rows, h):
    L.check(L.lib().vds_batch_rowsum(_p(x), _p(out), B, batch_stride, rows, h, _s()), "vds_batch_rowsum")


def cast_f32_bf16(x, out=None, scale=1.0):
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().vds_cast_f32_bf16(_p(x), _p(out), x.numel(), scale, _s()), "vds_cast_f32_bf16")
    return out


def cast_f32_bf16_2d(x, out, scale=1.0):
    """x fp32 [rows, cols] -> out bf16 [rows, cols]; both may be column slices of wider buffers (unit inner stride)."""
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and x.shape == out.shape and x.dim() == 2
    assert x.stride(1) == 1 and out.stride(1) == 1
    L.check(L.lib().vds_cast_f32_bf16_2d(_p(x), x.stride(0), _p(out), out.stride(0), x.shape[0], x.shape[1], scale, _s()),
            "vds_cast_f32_bf16_2d")
    return out


def accum_bf16_f32(x, out, accumulate=True):
    _chk_bf16(x)
    assert out.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    L.check(L.lib().vds_accum_bf16_f32(_p(x), _p(out), x.numel(), int(accumulate), _s()), "vds_accum_bf16_f32")


def attn_fwd(q, k, v, B, nh, Lq, Lk, out=None, want_lse=True, hd=128):
    """q/k/v: 2-D token-major views [B*L, >= nh*hd] (unit inner stride; may be column slices of a wider buffer)."""
    _chk_bf16(q, k, v)
    if out is None:
        out = torch.empty((B * Lq, nh * hd), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((B, nh, Lq), device=q.device, dtype=torch.float32) if want_lse else None
    L.check(L.lib().vds_attn_fwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                                 _p(lse), B, nh, Lq, Lk, hd, float(hd) ** -0.5, _s()), "vds_attn_fwd")
    return out, lse


_TAIL_WS = {}


def _tail_ws(device, B, nh, Lk):
    """Tail-balancing workspace of vds_attn_bwd, one per device, grown (never shrunk) to the largest request so a
    later, larger problem cannot silently lose tail balancing.  Allocated zero-filled: the C side requires that once and
    hands the workspace back zero-filled after every call (include/vds_b200.h)."""
    need = int(L.lib().vds_attn_bwd_tail_ws_bytes(B, nh, Lk))
    ws = _TAIL_WS.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, device=device, dtype=torch.uint8)
        _TAIL_WS[device] = ws
    return ws


def attn_bwd(q, k, v, o, d_o, lse, B, nh, Lq, Lk, dq_acc, dk=None, dv=None, dk_acc=None, dv_acc=None, q_splits=1,
             hd=128, tail_balance=True, delta=None):
    """dq_acc: zeroed fp32 [B*Lq, nh*hd]; dk/dv: bf16 2-D views (q_splits == 1) or fp32 accumulators.
    delta: precomputed rowsum(dO * O) fp32 [B, nh, Lq] (gemm_dgrad_rowdot); None -> computed here from o and d_o."""
    _chk_bf16(q, k, v, o, d_o)
    have_delta = delta is not None
    if not have_delta:
        delta = torch.empty((B, nh, Lq), device=q.device, dtype=torch.float32)
    else:
        assert delta.dtype == torch.float32 and delta.is_contiguous() and delta.numel() == B * nh * Lq
    ws = _tail_ws(q.device, B, nh, Lk) if (tail_balance and q_splits == 1) else None
    prof = PROFILE.get("attn_bwd_self") if Lq == Lk else None
    if prof is not None:
        # inside a CUDA-graph capture the events become event-record NODES (external=True): every replay re-records
        # them, so the kernel can be timed live inside a graphed step as well
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
    L.check(L.lib().vds_attn_bwd(
        _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), None if have_delta else _p(o), o.stride(0), _p(d_o),
        d_o.stride(0),
        _p(lse), _p(delta), _p(dq_acc), dq_acc.stride(0), _p(dk), dk.stride(0) if dk is not None else 0, _p(dv),
        dv.stride(0) if dv is not None else 0, _p(dk_acc), _p(dv_acc),
        dk_acc.stride(0) if dk_acc is not None else 0, q_splits, B, nh, Lq, Lk, hd, float(hd) ** -0.5,
        _p(ws), ws.numel() if ws is not None else 0, _s()),
        "vds_attn_bwd")
    if prof is not None:
        e1.record()
        prof.append((e0, e1))
    return delta


def loss_fwd_bwd(x, noise, out, want_grad=True, grad_scale=1.0, want_batch=False, want_loss=True,
                 grad_scale_dev=None):
    """Fused v = x - noise, MSE against `out`, and d loss / d out (train.py:117-125)."""
    _chk_bf16(x, noise, out)
    B = x.shape[0]
    per = x.numel() // B
    d_out = torch.empty_like(out) if want_grad else None
    loss = torch.zeros((1,), device=x.device, dtype=torch.float32) if want_loss else None
    lb = torch.zeros((B,), device=x.device, dtype=torch.float32) if want_batch else None
    if grad_scale_dev is not None:
        assert grad_scale_dev.dtype == torch.float32 and grad_scale_dev.is_cuda
    L.check(L.lib().vds_loss_fwd_bwd(_p(x), _p(noise), _p(out), _p(d_out), _p(loss), _p(lb), B, per, grad_scale,
                                     _p(grad_scale_dev), _s()), "vds_loss_fwd_bwd")
    return loss, d_out, lb
