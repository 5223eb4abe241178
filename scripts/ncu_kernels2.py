"""Round-2 additions at the debug-8k shapes for `ncu --set full` (scripts/ncu_kernels2.sh): the QKV projection with the fused
RoPE + value-residual epilogue (VDS_EPI_QKV_ROPE), the plain QKV GEMM + in-place pass it replaces, and qkv_post_bwd.
The second pass is the one profiled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import ops

dev = "cuda"
B, Lr, h, nh = 2, 8208, 512, 4
M = B * Lr
bf = lambda *s: torch.randn(s, device=dev).bfloat16()   # noqa: E731
x, w_qkv, v0 = bf(M, h), bf(3 * h, h) * 0.05, bf(M, h)
ang = torch.randn((Lr, 64), device=dev) * 3
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lam = torch.tensor([0.4], device=dev).bfloat16()
tab = ops.rope_pack(cos, sin)
dq_acc = torch.randn((M, h), device=dev)
dv0 = torch.zeros((M, h), device=dev)
dlam = torch.zeros((1,), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def one_pass():
    flush.zero_()
    qkv_f, vmix = ops.gemm_qkv_rope(x, w_qkv, None, tab, Lr, v0=v0, v0_ld=h, lam=lam)
    flush.zero_()
    qkv = ops.gemm(x, w_qkv)
    flush.zero_()
    ops.qkv_post_fwd(qkv, B, Lr, h, nh, cos=cos, sin=sin, v0=v0, v0_ld=h, lam=lam)
    dqkv = bf(M, 3 * h)
    flush.zero_()
    ops.qkv_post_bwd(dqkv, B, Lr, h, nh, dq_acc=dq_acc, cos=cos, sin=sin, qkv_pre=qkv, v0=v0, v0_ld=h, lam=lam,
                     dlambda=dlam, dv0_acc=dv0, mode=1)
    torch.cuda.synchronize()


one_pass()
one_pass()
