"""One launch of every kernel family of the step at the debug-8k shapes (B=2, L=8208, h=512), after one warm-up pass,
for `ncu --set full` (scripts/ncu_kernels.sh).  The second pass is the one profiled (-s skips the first)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib, ops

dev = "cuda"
B, Lr, h, nh = 2, 8208, 512, 4
M = B * Lr
bf = lambda *s: torch.randn(s, device=dev).bfloat16()   # noqa: E731
x, dy, res = bf(M, h), bf(M, h), bf(M, h)
mod = bf(B, 9 * h)
dmod = torch.zeros((B, 9 * h), device=dev, dtype=torch.float32)
w_qkv, w_p, w1, b1, w2, b2 = bf(3 * h, h), bf(h, h), bf(4 * h, h), bf(4 * h), bf(h, 4 * h), bf(h)
big = bf(M, 4 * h)
n_par = 248_000_000
master = torch.randn(n_par, device=dev)
grad = torch.randn(n_par, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def one_pass():
    flush.zero_()
    y, rstd = ops.rmsnorm_mod_fwd(x, B, Lr, h, scale=mod[:, h:2 * h], shift=mod[:, :h])
    flush.zero_()
    ops.rmsnorm_mod_bwd(dy, x, rstd, B, Lr, h, scale=mod[:, h:2 * h], dx_res=res, dscale=dmod[:, h:2 * h], dshift=dmod[:, :h])
    flush.zero_()
    ops.gate_bwd(dy, x, mod[:, 2 * h:3 * h], dmod[:, 2 * h:3 * h], B, Lr, h)
    flush.zero_()
    ops.colsum(big, torch.zeros(4 * h, device=dev))
    flush.zero_()
    qkv = ops.gemm(x, w_qkv)                                                   # plain store, N = 1536
    flush.zero_()
    h1, g = ops.gemm(x, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU)              # fused GELU, N = 2048
    flush.zero_()
    o3, x3 = ops.gemm(g, w2, bias=b2, epilogue=lib.EPI_GATE_RES, aux=x, gate=mod[:, 8 * h:], rows_per_batch=Lr)
    flush.zero_()
    dh1 = ops.gemm(dy, w2, b_mn=True, epilogue=lib.EPI_DGELU, aux=h1)         # dGELU dgrad
    flush.zero_()
    gw = torch.zeros((4 * h, h), device=dev, dtype=torch.float32)
    ops.gemm(dh1, x, a_mn=True, b_mn=True, epilogue=lib.EPI_ACCUM_F32, out=gw, splits=8)   # wgrad split-K
    torch.cuda.synchronize()


one_pass()
one_pass()
