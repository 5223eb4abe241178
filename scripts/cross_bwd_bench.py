"""1-CTA attention backward at the shapes that use it: cross-attention of debug-8k (Lk = 512, 4 query splits, fp32 dk / dv
accumulators) and the self-attention of DiT-XL (L = 2064) and DiT-B (L = 272)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import ops
from vds_b200.engine import _attn_q_splits
from attn_bench import timeit

for (B, nh, Lq, Lk) in [(2, 4, 8208, 512), (2, 9, 2064, 2064), (8, 6, 272, 272), (2, 9, 2064, 512)]:
    h = nh * 128
    q = torch.randn((B * Lq, h), device="cuda").bfloat16()
    kv = torch.randn((B * Lk, 2 * h), device="cuda").bfloat16()
    k, v = kv[:, :h], kv[:, h:]
    out, lse = ops.attn_fwd(q, k, v, B, nh, Lq, Lk)
    d_o = torch.randn((B * Lq, h), device="cuda").bfloat16()
    dq = torch.zeros((B * Lq, h), device="cuda", dtype=torch.float32)
    qs = _attn_q_splits((Lk + 127) // 128, B, nh, (Lq + 127) // 128) if Lq != Lk else 1
    if qs > 1:
        acc = torch.zeros((B * Lk, 2 * h), device="cuda", dtype=torch.float32)
        fn = lambda: ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk_acc=acc[:, :h], dv_acc=acc[:, h:], q_splits=qs)
    else:
        dkv = torch.zeros((B * Lk, 2 * h), device="cuda").bfloat16()
        fn = lambda: ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk=dkv[:, :h], dv=dkv[:, h:])
    t, _ = timeit(fn, n=4, reps=8)
    print(f"bwd B={B} nh={nh} Lq={Lq} Lk={Lk} q_splits={qs}: {t*1e3:.1f} us", flush=True)
