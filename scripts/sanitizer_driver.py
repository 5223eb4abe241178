"""Only OUR kernels, small shapes, no torch reference math (torch's own SDPA backward floods racecheck): the launch
sequence compute-sanitizer memcheck / racecheck runs over (scripts/sanitizer.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib as L, ops

dev = "cuda"
bf = lambda *s: (torch.randn(s, device=dev) * 0.5).bfloat16()   # noqa: E731
for pair in (0, 2):
    L.check(L.lib().vds_debug_attn_pair_mode(pair))
    for (B, nh, Lq, Lk) in [(1, 1, 128, 128), (1, 2, 272, 272), (1, 1, 200, 256), (1, 2, 1040, 1040)]:
        h = nh * 128
        q, k, v, d_o = bf(B * Lq, h), bf(B * Lk, h), bf(B * Lk, h), bf(B * Lq, h)
        out, lse = ops.attn_fwd(q, k, v, B, nh, Lq, Lk)
        dq = torch.zeros((B * Lq, h), device=dev, dtype=torch.float32)
        dk = torch.zeros((B * Lk, h), device=dev).bfloat16()
        dv = torch.zeros_like(dk)
        ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk=dk, dv=dv)
        dkf = torch.zeros((B * Lk, 2 * h), device=dev, dtype=torch.float32)
        dq.zero_()
        ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk_acc=dkf[:, :h], dv_acc=dkf[:, h:], q_splits=3)
L.check(L.lib().vds_debug_attn_pair_mode(-1))
M, h = 4200, 512
x, w1, b1, w2, res = bf(M, h), bf(4 * h, h), bf(4 * h), bf(h, 4 * h), bf(M, h)
mod = bf(2, 9 * h)
pre, act = ops.gemm(x, w1, bias=b1, epilogue=L.EPI_BIAS_GELU)
lin, xo = ops.gemm(act, w2, epilogue=L.EPI_GATE_RES, aux=res, gate=mod[:, 8 * h:], rows_per_batch=M // 2)
dh = ops.gemm(x, w2, b_mn=True, epilogue=L.EPI_DGELU, aux=pre)
gw = torch.zeros((4 * h, h), device=dev, dtype=torch.float32)
ops.gemm(dh, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw, splits=2)
big = bf(40000, h)      # large enough for the 2-CTA path
ops.gemm(big, w1, bias=b1)
rowdot = torch.zeros((2, 4, 20000), device=dev, dtype=torch.float32)
ops.gemm_dgrad_rowdot(big, bf(h, h), big, rowdot, 20000)
# QKV projection with the fused RoPE / value-residual epilogue (2-CTA path), incl. a bias and a ragged last m-tile
Lq, nhq = 5003, 4
xq, wq, bq, v0q = bf(2 * Lq, h), bf(3 * nhq * 128, h), bf(3 * nhq * 128), bf(2 * Lq, nhq * 128)
angq = torch.randn((Lq, 64), device=dev)
tabq = ops.rope_pack(angq.cos().contiguous(), angq.sin().contiguous())
assert ops.gemm_qkv_rope(xq, wq, bq, tabq, Lq, v0=v0q, v0_ld=nhq * 128, lam=torch.tensor([0.3], device=dev).bfloat16()) is not None
assert ops.gemm_qkv_rope(xq, wq, None, tabq, Lq) is not None
y, rstd = ops.rmsnorm_mod_fwd(x, 2, M // 2, h, scale=mod[:, h:2 * h], shift=mod[:, :h])
dmod = torch.zeros((2, 9 * h), device=dev, dtype=torch.float32)
ops.rmsnorm_mod_bwd(x, x, rstd, 2, M // 2, h, scale=mod[:, h:2 * h], dx_res=res, dscale=dmod[:, h:2 * h], dshift=dmod[:, :h])
ops.gate_bwd(x, res, mod[:, 2 * h:3 * h], dmod[:, 2 * h:3 * h], 2, M // 2, h)
ops.colsum(pre, torch.zeros(4 * h, device=dev))
torch.cuda.synchronize()
print("sanitizer driver: done")
