"""Each memory-bound kernel and the attention kernels against plain torch references on the GPU."""
import math

import pytest
import torch
import torch.nn.functional as F
from einops import rearrange

pytestmark = pytest.mark.gpu


def _r(shape, dev, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(dev)


def _close(got, ref, tol=2e-2, name=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"{name}: max err {err} vs scale {den}"


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def test_patchify_bit_exact(cuda_dev):
    from vds_b200 import ops
    B, C, T, H, W = 2, 16, 4, 8, 6
    x = torch.arange(B * C * T * H * W, dtype=torch.float32).remainder(251).view(B, C, T, H, W).bfloat16().to(cuda_dev)
    got = ops.patchify(x, 2, 2)
    ref = rearrange(x, "b c (t pt) (h ph) (w pw) -> (b h w t) (c pt ph pw)", pt=2, ph=2, pw=2)
    assert torch.equal(got, ref)


def test_patchify_zt(cuda_dev):
    from vds_b200 import ops
    B, C, T, H, W = 2, 16, 4, 8, 8
    x, n = _r((B, C, T, H, W), cuda_dev, 1), _r((B, C, T, H, W), cuda_dev, 2)
    t = torch.tensor([0.3, 0.9], device=cuda_dev).bfloat16()
    tr = t.view(B, 1, 1, 1, 1)
    zt = x * (1 - tr) + n * tr
    got = ops.patchify(x, 2, 2, noise=n, t=t)
    ref = rearrange(zt, "b c (t pt) (h ph) (w pw) -> (b h w t) (c pt ph pw)", pt=2, ph=2, pw=2)
    assert torch.equal(got, ref)


def test_unpatchify_bit_exact(cuda_dev):
    from vds_b200 import ops
    B, C, T, H, W = 2, 16, 4, 8, 6
    n = (T // 2) * (H // 2) * (W // 2)
    y = torch.arange(B * n * 128, dtype=torch.float32).remainder(241).view(B * n, 128).bfloat16().to(cuda_dev)
    got = ops.unpatchify(y, B, C, T, H, W, 2, 2)
    ref = rearrange(y.view(B, n, 128), "b (h w t) (p1 p2 p3 c) -> b c (t p3) (h p1) (w p2)", t=T // 2, h=H // 2,
                    w=W // 2, p1=2, p2=2, p3=2)
    assert torch.equal(got, ref)
    back = ops.unpatchify(got, B, C, T, H, W, 2, 2, to_tokens=True)
    assert torch.equal(back, y)


def test_rope_rows(cuda_dev):
    from vds_b200 import ops
    D = 8
    tab = torch.arange(10 * 12 * 14 * D, dtype=torch.float32).view(10, 12, 14, D).to(cuda_dev)
    Tp, Hp, Wp = 2, 3, 4
    st, sh, sw = 5, 2, 7
    c, s = ops.rope_rows(tab, -tab, (Tp, Hp, Wp), (st, sh, sw), 16)
    ref = tab[st:st + Tp, sh:sh + Hp, sw:sw + Wp].clone().reshape(Tp * Hp * Wp, -1)
    assert torch.equal(c[16:], ref) and torch.equal(s[16:], -ref)
    assert torch.equal(c[:16], torch.ones_like(c[:16])) and torch.equal(s[:16], torch.zeros_like(s[:16]))


def test_timestep_embedding_and_silu(cuda_dev):
    from vds_b200 import ops
    t = torch.tensor([0.1, 0.77, 0.5], device=cuda_dev).bfloat16()
    dim = 512
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(cuda_dev)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1).bfloat16()
    _close(ops.timestep_embedding(t, dim), ref, tol=1e-2)
    x = _r((3, 2048), cuda_dev, 3)
    _close(ops.silu(x), F.silu(x), tol=1e-2)
    dy = _r((3, 2048), cuda_dev, 4)
    xx = x.float().requires_grad_(True)
    F.silu(xx).backward(dy.float())
    _close(ops.silu_bwd(x, dy), xx.grad, tol=1e-2)


@pytest.mark.parametrize("h,with_w", [(512, False), (768, True), (1152, False), (2048, False)])
def test_rmsnorm_mod(cuda_dev, h, with_w):
    from vds_b200 import ops
    B, Lr = 2, 300
    x = _r((B * Lr, h), cuda_dev, 5, 2.0)
    mod = _r((B, 9 * h), cuda_dev, 6, 0.3)
    scale, shift = mod[:, h:2 * h], mod[:, 0:h]
    w = (1 + 0.1 * _r((h,), cuda_dev, 7).float()).bfloat16() if with_w else None

    def ref_fn(xf, scf, shf, wf):
        n = xf.float() * torch.rsqrt(xf.float().pow(2).mean(-1, keepdim=True) + 1e-6)
        if wf is not None:
            n = n * wf
        return n.view(B, Lr, h) * (1 + scf[:, None, :]) + shf[:, None, :]

    y, rstd = ops.rmsnorm_mod_fwd(x, B, Lr, h, scale=scale, shift=shift, weight=w)
    ref = ref_fn(x, scale.float(), shift.float(), w.float() if w is not None else None)
    _close(y.view(B, Lr, h), ref, tol=1.5e-2)
    # backward vs autograd of the fp32 reference
    dy = _r((B * Lr, h), cuda_dev, 8)
    res = _r((B * Lr, h), cuda_dev, 9)
    xf = x.float().requires_grad_(True)
    scf, shf = scale.float().clone().requires_grad_(True), shift.float().clone().requires_grad_(True)
    wf = w.float().clone().requires_grad_(True) if w is not None else None
    ref_fn(xf, scf, shf, wf).backward(dy.float().view(B, Lr, h))
    dmod = torch.zeros((B, 9 * h), device=cuda_dev, dtype=torch.float32)
    dw = torch.zeros((h,), device=cuda_dev, dtype=torch.float32) if w is not None else None
    dx = ops.rmsnorm_mod_bwd(dy, x, rstd, B, Lr, h, scale=scale, weight=w, dx_res=res, dscale=dmod[:, h:2 * h],
                             dshift=dmod[:, 0:h], dweight=dw)
    assert _cos(dx.float() - res.float(), xf.grad) > 0.999
    assert _cos(dmod[:, h:2 * h], scf.grad) > 0.9995
    assert _cos(dmod[:, 0:h], shf.grad) > 0.9999
    if w is not None:
        assert _cos(dw, wf.grad) > 0.9995


def test_rmsnorm_final_rowmap(cuda_dev):
    """final norm: reads rows 16.. of every sample, writes compact rows (model.py:386-389)."""
    from vds_b200 import ops
    B, Lr, h = 2, 80, 512
    x = _r((B, Lr, h), cuda_dev, 10)
    mod = _r((B, 2 * h), cuda_dev, 11, 0.3)
    y, rstd = ops.rmsnorm_mod_fwd(x.view(-1, h), B, Lr - 16, h, scale=mod[:, h:], shift=mod[:, :h], in_batch_stride=Lr,
                                  in_row_offset=16)
    xs = x[:, 16:].float()
    ref = xs * torch.rsqrt(xs.pow(2).mean(-1, keepdim=True) + 1e-6) * (1 + mod[:, None, h:].float()) + mod[:, None, :h].float()
    _close(y.view(B, Lr - 16, h), ref, tol=1.5e-2)
    dy = _r((B * (Lr - 16), h), cuda_dev, 12)
    dx = torch.zeros((B * Lr, h), device=cuda_dev, dtype=torch.bfloat16)
    dmod = torch.zeros((B, 2 * h), device=cuda_dev, dtype=torch.float32)
    ops.rmsnorm_mod_bwd(dy, x.view(-1, h), rstd, B, Lr - 16, h, scale=mod[:, h:], dx=dx, dscale=dmod[:, h:],
                        dshift=dmod[:, :h], in_batch_stride=Lr, in_row_offset=16, dx_full_rows=True)
    xf = x.float().requires_grad_(True)
    xs = xf[:, 16:]
    (xs * torch.rsqrt(xs.pow(2).mean(-1, keepdim=True) + 1e-6) * (1 + mod[:, None, h:].float()) +
     mod[:, None, :h].float()).backward(dy.float().view(B, Lr - 16, h))
    assert _cos(dx.view(B, Lr, h)[:, 16:], xf.grad[:, 16:]) > 0.999
    assert dx.view(B, Lr, h)[:, :16].abs().max().item() == 0


def test_gate_bwd(cuda_dev):
    from vds_b200 import ops
    B, Lr, h = 2, 300, 512
    dx, o = _r((B * Lr, h), cuda_dev, 13), _r((B * Lr, h), cuda_dev, 14)
    mod = _r((B, 9 * h), cuda_dev, 15)
    g = mod[:, 2 * h:3 * h]
    dmod = torch.zeros((B, 9 * h), device=cuda_dev, dtype=torch.float32)
    d_o = ops.gate_bwd(dx, o, g, dmod[:, 2 * h:3 * h], B, Lr, h)
    _close(d_o.view(B, Lr, h), dx.view(B, Lr, h).float() * g[:, None, :].float(), tol=1e-2)
    ref = (dx.view(B, Lr, h).float() * o.view(B, Lr, h).float()).sum(1)
    assert _cos(dmod[:, 2 * h:3 * h], ref) > 0.9999
    _close(dmod[:, 2 * h:3 * h], ref, tol=1e-3)


def _ref_rope(x, cos, sin):
    d = x.shape[-1] // 2
    x = x.float()
    x1, x2 = x[..., :d], x[..., d:]
    return torch.cat([x1 * cos + x2 * sin, x1 * (-sin) + x2 * cos], -1)


def test_qkv_post_fwd_bwd(cuda_dev):
    from vds_b200 import ops
    B, Lr, nh, hd = 2, 80, 4, 128
    h = nh * hd
    qkv0 = _r((B * Lr, 3 * h), cuda_dev, 16)
    v0buf = _r((B * Lr, 3 * h), cuda_dev, 17)
    v0 = v0buf[:, 2 * h:]
    ang = _r((Lr, hd // 2), cuda_dev, 18, 3.0, torch.float32)
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    lam = torch.tensor([0.4], device=cuda_dev).bfloat16()
    qkv = qkv0.clone()
    vmix = ops.qkv_post_fwd(qkv, B, Lr, h, nh, cos=cos, sin=sin, v0=v0, v0_ld=v0.stride(0), lam=lam)
    q = rearrange(qkv0[:, :h].view(B, Lr, h), "b l (h d) -> b h l d", h=nh)
    rq = rearrange(_ref_rope(q, cos[None, None], sin[None, None]), "b h l d -> (b l) (h d)")
    _close(qkv[:, :h], rq, tol=1e-2)
    k = rearrange(qkv0[:, h:2 * h].view(B, Lr, h), "b l (h d) -> b h l d", h=nh)
    rk = rearrange(_ref_rope(k, cos[None, None], sin[None, None]), "b h l d -> (b l) (h d)")
    _close(qkv[:, h:2 * h], rk, tol=1e-2)
    assert torch.equal(qkv[:, 2 * h:], qkv0[:, 2 * h:])
    ref_v = lam * qkv0[:, 2 * h:] + (1 - lam) * v0
    assert torch.equal(vmix, ref_v)
    # backward
    dqkv0 = _r((B * Lr, 3 * h), cuda_dev, 19)
    dq_acc = _r((B * Lr, h), cuda_dev, 20, 1.0, torch.float32)
    dqkv = dqkv0.clone()
    dlam = torch.zeros((1,), device=cuda_dev, dtype=torch.float32)
    dv0 = torch.zeros((B * Lr, h), device=cuda_dev, dtype=torch.float32)
    ops.qkv_post_bwd(dqkv, B, Lr, h, nh, dq_acc=dq_acc, cos=cos, sin=sin, qkv_pre=qkv0, v0=v0, v0_ld=v0.stride(0),
                     lam=lam, dlambda=dlam, dv0_acc=dv0, mode=1)
    qf = q.float().requires_grad_(True)
    _ref_rope(qf, cos[None, None], sin[None, None]).backward(
        rearrange(dq_acc.view(B, Lr, h), "b l (h d) -> b h l d", h=nh))
    _close(dqkv[:, :h], rearrange(qf.grad, "b h l d -> (b l) (h d)"), tol=1e-2)
    kf = k.float().requires_grad_(True)
    _ref_rope(kf, cos[None, None], sin[None, None]).backward(
        rearrange(dqkv0[:, h:2 * h].float().view(B, Lr, h), "b l (h d) -> b h l d", h=nh))
    _close(dqkv[:, h:2 * h], rearrange(kf.grad, "b h l d -> (b l) (h d)"), tol=1e-2)
    dvm = dqkv0[:, 2 * h:].float()
    _close(dqkv[:, 2 * h:], lam.float() * dvm, tol=1e-2)
    _close(dv0, (1 - lam).float() * dvm, tol=1e-2)
    ref_dl = (dvm * (qkv0[:, 2 * h:].float() - v0.float())).sum()
    assert abs(dlam.item() - ref_dl.item()) < 2e-3 * (abs(ref_dl.item()) + 10)
    # block-0 mode: dv_pre = dv_mix + dv0_acc
    dqkv2 = dqkv0.clone()
    ops.qkv_post_bwd(dqkv2, B, Lr, h, nh, dq_acc=dq_acc, cos=cos, sin=sin, dv0_acc=dv0, mode=2)
    _close(dqkv2[:, 2 * h:], dvm + dv0, tol=1e-2)


def test_colsum_rowsum_cast(cuda_dev):
    from vds_b200 import ops
    x = _r((1000, 2048), cuda_dev, 21)
    out = torch.ones((2048,), device=cuda_dev, dtype=torch.float32)
    ops.colsum(x, out)
    _close(out, 1 + x.float().sum(0), tol=1e-3)
    xb = _r((3, 40, 512), cuda_dev, 22)
    o2 = torch.zeros((16, 512), device=cuda_dev, dtype=torch.float32)
    ops.batch_rowsum(xb, o2, 3, 40 * 512, 16, 512)
    _close(o2, xb[:, :16].float().sum(0), tol=1e-3)
    f = _r((1001,), cuda_dev, 23, 1.0, torch.float32)
    assert torch.equal(ops.cast_f32_bf16(f), f.bfloat16())
    acc = torch.ones((1000, 2048), device=cuda_dev, dtype=torch.float32)
    ops.accum_bf16_f32(x, acc)
    assert torch.equal(acc, 1 + x.float())


def _ref_attn(q, k, v):
    return F.scaled_dot_product_attention(q.float(), k.float(), v.float())


@pytest.fixture
def pair_mode():
    """Selects the attention-backward kernel (vds_debug_attn_pair_mode) for one test and restores the default."""
    from vds_b200 import lib as L

    def set_mode(m):
        L.check(L.lib().vds_debug_attn_pair_mode(m), "vds_debug_attn_pair_mode")
    yield set_mode
    set_mode(-1)


@pytest.mark.parametrize("pair", [0, 2])    # 0: 1-CTA kernel only; 2: every kv-tile pair on the 2-CTA cluster kernel
@pytest.mark.parametrize("B,nh,Lq,Lk", [(1, 1, 128, 128), (2, 4, 272, 272), (1, 2, 528, 512), (2, 4, 2064, 2064),
                                        (3, 4, 2192, 2192),    # 216 items on 148 SMs -> tail balancing path
                                        (1, 2, 1040, 1040),    # 9 kv tiles: 4 pairs + an unpaired ragged tile per head
                                        (1, 3, 200, 256)])     # one pair per head, ragged query range
def test_attn_fwd_bwd(cuda_dev, pair_mode, pair, B, nh, Lq, Lk):
    from vds_b200 import ops
    pair_mode(pair)
    hd, h = 128, nh * 128
    self_attn = Lq == Lk
    if self_attn:
        qkv = _r((B * Lq, 3 * h), cuda_dev, 30)
        q2, k2, v2 = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
    else:
        q2 = _r((B * Lq, h), cuda_dev, 31)
        kv = _r((B * Lk, 2 * h), cuda_dev, 32)
        k2, v2 = kv[:, :h], kv[:, h:]
    out, lse = ops.attn_fwd(q2, k2, v2, B, nh, Lq, Lk)
    q = rearrange(q2.reshape(B, Lq, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    k = rearrange(k2.reshape(B, Lk, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    v = rearrange(v2.reshape(B, Lk, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    ref = _ref_attn(q, k, v)
    ref2 = rearrange(ref, "b h l d -> (b l) (h d)")
    _close(out, ref2, tol=2e-2, name="attn out")
    # lse (log2 domain)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    ref_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    _close(lse, ref_lse.detach(), tol=1e-3, name="lse")
    d_o = _r((B * Lq, h), cuda_dev, 33)
    ref.backward(rearrange(d_o.float().view(B, Lq, nh, hd), "b l h d -> b h l d"))
    for q_splits in ([1] if self_attn else [1, 3]):
        dq_acc = torch.zeros((B * Lq, h), device=cuda_dev, dtype=torch.float32)
        if q_splits == 1:
            dk = torch.zeros((B * Lk, h), device=cuda_dev, dtype=torch.bfloat16)
            dv = torch.zeros_like(dk)
            ops.attn_bwd(q2, k2, v2, out, d_o, lse, B, nh, Lq, Lk, dq_acc, dk=dk, dv=dv)
        else:
            dkv = torch.zeros((B * Lk, 2 * h), device=cuda_dev, dtype=torch.float32)
            dk, dv = dkv[:, :h], dkv[:, h:]
            ops.attn_bwd(q2, k2, v2, out, d_o, lse, B, nh, Lq, Lk, dq_acc, dk_acc=dk, dv_acc=dv, q_splits=q_splits)
        rdq = rearrange(q.grad, "b h l d -> (b l) (h d)")
        rdk = rearrange(k.grad, "b h l d -> (b l) (h d)")
        rdv = rearrange(v.grad, "b h l d -> (b l) (h d)")
        assert _cos(dq_acc, rdq) > 0.999, ("dq", _cos(dq_acc, rdq))
        assert _cos(dk, rdk) > 0.999, ("dk", _cos(dk, rdk))
        assert _cos(dv, rdv) > 0.999, ("dv", _cos(dv, rdv))
        _close(dq_acc, rdq, tol=3e-2, name="dq")
        _close(dk, rdk, tol=3e-2, name="dk")
        _close(dv, rdv, tol=3e-2, name="dv")


def test_attn_growing_max_takes_the_rescale_path(cuda_dev):
    """The forward exponentiates a tile against the running max of the previous tiles and redoes it exactly when some
    row's max grew by more than 2^8.  Unit-scale random inputs never trigger that, so this case makes the scores of
    later key tiles much larger (keys scaled up with position) and checks forward, LSE and backward against fp32 SDPA."""
    from vds_b200 import ops
    B, nh, L, hd = 1, 2, 1040, 128          # 9 key tiles, last one ragged
    h = nh * hd
    g = torch.Generator(device="cpu").manual_seed(40)
    q = torch.randn((B * L, h), generator=g)
    k = torch.randn((B * L, h), generator=g)
    ramp = torch.linspace(0.5, 6.0, L).repeat(B).unsqueeze(1)     # |scores| grow ~12x from the first to the last tile
    k = k * ramp
    v = torch.randn((B * L, h), generator=g)
    q2, k2, v2 = (t.bfloat16().to(cuda_dev) for t in (q, k, v))
    out, lse = ops.attn_fwd(q2, k2, v2, B, nh, L, L)
    qf = rearrange(q2.reshape(B, L, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    kf = rearrange(k2.reshape(B, L, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    vf = rearrange(v2.reshape(B, L, nh, hd), "b l h d -> b h l d").float().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * hd ** -0.5
    # the premise of the test: row maxima really jump by more than 8 (log2 domain) between early and late tiles
    m_first = (s[..., :128].max(-1).values * 1.4426950408889634)
    m_all = (s.max(-1).values * 1.4426950408889634)
    assert ((m_all - m_first) > 8.0).float().mean().item() > 0.5
    ref = _ref_attn(qf, kf, vf)
    _close(out, rearrange(ref, "b h l d -> (b l) (h d)"), tol=2e-2, name="attn out (rescale path)")
    _close(lse, (torch.logsumexp(s, -1) * 1.4426950408889634).detach(), tol=1e-3, name="lse (rescale path)")
    d_o = _r((B * L, h), cuda_dev, 41)
    ref.backward(rearrange(d_o.float().view(B, L, nh, hd), "b l h d -> b h l d"))
    dq_acc = torch.zeros((B * L, h), device=cuda_dev, dtype=torch.float32)
    dk = torch.zeros((B * L, h), device=cuda_dev, dtype=torch.bfloat16)
    dv = torch.zeros_like(dk)
    ops.attn_bwd(q2, k2, v2, out, d_o, lse, B, nh, L, L, dq_acc, dk=dk, dv=dv)
    for got, want, name in ((dq_acc, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        want = rearrange(want, "b h l d -> (b l) (h d)")
        assert _cos(got, want) > 0.999, (name, _cos(got, want))


def test_attn_full_size_properties(cuda_dev):
    """S_dbg size (B=2, 4 heads, L=8208, what bench.py runs): size-independent properties instead of a dense reference.
    (i) rows of softmax sum to one: with V = ones the output is exactly 1 and dV = column sums of P = sum of dO weights;
    (ii) linearity in V and dO; (iii) a spot check of 64 query rows per (b, head) against fp32 SDPA on those rows."""
    from vds_b200 import ops
    B, nh, L, hd = 2, 4, 8208, 128
    h = nh * hd
    qkv = _r((B * L, 3 * h), cuda_dev, 50)
    q2, k2, v2 = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
    ones = torch.ones((B * L, h), device=cuda_dev, dtype=torch.bfloat16)
    out1, lse = ops.attn_fwd(q2, k2, ones, B, nh, L, L)
    assert (out1.float() - 1.0).abs().max().item() < 1e-2
    out, lse2 = ops.attn_fwd(q2, k2, v2, B, nh, L, L)
    assert torch.equal(lse, lse2)                                   # LSE does not depend on V
    out_2v, _ = ops.attn_fwd(q2, k2, (v2.float() * 2).bfloat16(), B, nh, L, L)
    _close(out_2v, out.float() * 2, tol=1e-2, name="linearity in V")
    # spot rows against fp32 SDPA
    rows = torch.arange(0, L, L // 64, device=cuda_dev)[:64]
    qf = rearrange(q2.reshape(B, L, nh, hd), "b l h d -> b h l d").float()[:, :, rows]
    kf = rearrange(k2.reshape(B, L, nh, hd), "b l h d -> b h l d").float()
    vf = rearrange(v2.reshape(B, L, nh, hd), "b l h d -> b h l d").float()
    ref = F.scaled_dot_product_attention(qf, kf, vf)
    got = rearrange(out.view(B, L, nh, hd), "b l h d -> b h l d")[:, :, rows]
    _close(got, ref, tol=2e-2, name="spot rows")
    # backward: linear in dO; dq of a constant-V problem is zero (softmax rows sum to one => dP - delta = 0)
    d_o = _r((B * L, h), cuda_dev, 51)

    def bwd(vv, oo, dd):
        dq = torch.zeros((B * L, h), device=cuda_dev, dtype=torch.float32)
        dk = torch.zeros((B * L, h), device=cuda_dev, dtype=torch.bfloat16)
        dv = torch.zeros_like(dk)
        ops.attn_bwd(q2, k2, vv, oo, dd, lse, B, nh, L, L, dq, dk=dk, dv=dv)
        return dq, dk.float(), dv.float()
    dq, dk, dv = bwd(v2, out, d_o)
    dq2, dk2, dv2 = bwd(v2, out, (d_o.float() * 2).bfloat16())
    for a, b_, name in ((dq2, dq * 2, "dq"), (dk2, dk * 2, "dk"), (dv2, dv * 2, "dv")):
        assert _cos(a, b_) > 0.9999, (name, _cos(a, b_))
        _close(a, b_, tol=2e-2, name=name + " linearity in dO")
    dq_c, dk_c, dv_c = bwd(ones, out1, d_o)
    scale = dq.abs().max().item()
    assert dq_c.abs().max().item() < 2e-2 * scale and dk_c.abs().max().item() < 2e-2 * dk.abs().max().item()
    # dV = P^T dO: summed over keys it equals the sum of dO over queries (columns of P^T sum ... rows of P sum to one)
    tot_dv = rearrange(dv.view(B, L, nh, hd), "b l h d -> b h l d").sum(2)
    tot_do = rearrange(d_o.float().view(B, L, nh, hd), "b l h d -> b h l d").sum(2)
    _close(tot_dv, tot_do, tol=2e-2, name="sum_k dV == sum_q dO")


def test_loss_fwd_bwd(cuda_dev):
    from vds_b200 import ops
    B, shape = 2, (2, 16, 4, 8, 8)
    x, n, o = _r(shape, cuda_dev, 40), _r(shape, cuda_dev, 41), _r(shape, cuda_dev, 42)
    loss, d_out, lb = ops.loss_fwd_bwd(x, n, o, want_batch=True)
    of = o.float().requires_grad_(True)
    v = x - n
    ref_b = (v.float() - of).pow(2).mean(dim=(1, 2, 3, 4))
    ref = ref_b.mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-7
    _close(lb, ref_b.detach(), tol=1e-5)
    _close(d_out, of.grad, tol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("M_B_L_nh_K,use_bias,mix", [((2, 8208, 4, 512), False, True), ((2, 8208, 4, 512), True, False),
                                                     ((8, 272, 6, 768), False, True), ((2, 2064, 9, 1152), True, True)])
def test_gemm_qkv_rope_epilogue(cuda_dev, M_B_L_nh_K, use_bias, mix):
    """VDS_EPI_QKV_ROPE (RoPE + value residual inside the QKV GEMM epilogue, model.py:124-134) against (i) the plain GEMM +
    the in-place qkv_post_fwd pass it replaces — same arithmetic, so bit for bit — and (ii) the fp32 reference rotation."""
    from vds_b200 import ops
    B, Lr, nh, K = M_B_L_nh_K
    h = nh * 128
    x = _r((B * Lr, K), cuda_dev, 31, 1.0)
    w = _r((3 * h, K), cuda_dev, 32, K ** -0.5)
    bias = _r((3 * h,), cuda_dev, 33, 0.5) if use_bias else None
    ang = _r((Lr, 64), cuda_dev, 34, 3.0, torch.float32)
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    v0buf = _r((B * Lr, 3 * h), cuda_dev, 35)
    v0 = v0buf[:, 2 * h:] if mix else None
    lam = torch.tensor([0.4], device=cuda_dev).bfloat16() if mix else None
    fused = ops.gemm_qkv_rope(x, w, bias, ops.rope_pack(cos, sin), Lr, v0=v0, v0_ld=v0.stride(0) if mix else 0, lam=lam)
    assert fused is not None, "bench-sized QKV projections must take the 2-CTA fused path"
    qkv_f, vmix_f = fused
    qkv0 = ops.gemm(x, w, bias=bias)
    qkv = qkv0.clone()
    vmix = ops.qkv_post_fwd(qkv, B, Lr, h, nh, cos=cos, sin=sin, v0=v0, v0_ld=v0.stride(0) if mix else 0, lam=lam)
    assert torch.equal(qkv_f[:, 2 * h:], qkv[:, 2 * h:])          # v_pre
    if mix:
        assert torch.equal(vmix_f, vmix)
    else:
        assert vmix_f is None
    d = (qkv_f[:, :2 * h].float() - qkv[:, :2 * h].float()).abs()
    assert d.max().item() <= 2 ** -6 * qkv[:, :2 * h].float().abs().max().item()    # at most an fma-contraction ulp apart
    assert (d > 0).float().mean().item() < 1e-3
    q = rearrange(qkv0[:, :h].view(B, Lr, h), "b l (h d) -> b h l d", h=nh)
    rq = rearrange(_ref_rope(q, cos[None, None], sin[None, None]), "b h l d -> (b l) (h d)")
    _close(qkv_f[:, :h], rq, tol=1e-2)
    k = rearrange(qkv0[:, h:2 * h].view(B, Lr, h), "b l (h d) -> b h l d", h=nh)
    rk = rearrange(_ref_rope(k, cos[None, None], sin[None, None]), "b h l d -> (b l) (h d)")
    _close(qkv_f[:, h:2 * h], rk, tol=1e-2)


def test_rope_pack_layout(cuda_dev):
    """vds_rope_pack: [L, 64] cos / sin rows -> [L + 32, 4, 32] = per 16-pair step [16 cos | 16 sin], the last 32 rows wrapping
    around to rows 0..31 (a 32-token TMA box of the fused QKV epilogue may cross a sample boundary).  Bit-exact."""
    from vds_b200 import ops
    for Lr in (80, 272, 8208):
        ang = _r((Lr, 64), cuda_dev, 50 + Lr % 7, 3.0, torch.float32)
        cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
        tab = ops.rope_pack(cos, sin)
        ref = torch.cat([cos.view(Lr, 4, 16), sin.view(Lr, 4, 16)], dim=2)
        assert tab.shape == (Lr + 32, 4, 32)
        assert torch.equal(tab[:Lr], ref)
        assert torch.equal(tab[Lr:], ref[:32])
