"""End-of-round-2 additions for `ncu --set full` (scripts/ncu_kernels3.sh): wgrad GEMMs with the TMA reduce-add epilogue
(2-CTA and 1-CTA kernel, split factors from engine.wgrad_splits), cross-attention forward (TMA-stored O) and backward
(1-CTA kernel, 4 query splits, fp32 dK / dV added by TMA) at the debug-8k shapes.  The second pass is the one profiled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib as L, ops
from vds_b200.engine import wgrad_splits, _attn_q_splits

dev = "cuda"
B, Lr, h, nh, Lc = 2, 8208, 512, 4, 512
M = B * Lr
bf = lambda *s: torch.randn(s, device=dev).bfloat16()   # noqa: E731
dy4, x, dy1 = bf(M, 4 * h), bf(M, h), bf(M, h)
gw4 = torch.zeros((4 * h, h), device=dev, dtype=torch.float32)
gw1 = torch.zeros((h, h), device=dev, dtype=torch.float32)
q, kv = bf(M, h), bf(B * Lc, 2 * h)
d_o = bf(M, h)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
qs = _attn_q_splits((Lc + 127) // 128, B, nh, (Lr + 127) // 128)


def one_pass():
    flush.zero_()
    ops.gemm(dy4, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw4, splits=wgrad_splits(4 * h, h, M), K=M)
    flush.zero_()
    ops.gemm(dy1, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw1, splits=wgrad_splits(h, h, M), K=M)
    flush.zero_()
    out, lse = ops.attn_fwd(q, kv[:, :h], kv[:, h:], B, nh, Lr, Lc)
    dq = torch.zeros((M, h), device=dev, dtype=torch.float32)
    acc = torch.zeros((B * Lc, 2 * h), device=dev, dtype=torch.float32)
    flush.zero_()
    ops.attn_bwd(q, kv[:, :h], kv[:, h:], out, d_o, lse, B, nh, Lr, Lc, dq, dk_acc=acc[:, :h], dv_acc=acc[:, h:], q_splits=qs)
    torch.cuda.synchronize()


one_pass()
one_pass()
