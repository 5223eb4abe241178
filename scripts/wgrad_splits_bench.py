"""wgrad GEMMs of the debug-8k block (contraction over M = 16416 tokens) under different split-K factors: time per launch
(CUDA graph of 8 launches, L2 flushed between replays) — what engine.GradSink._wgrad_splits is tuned against."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib as L, ops
from attn_bench import timeit

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16416
for (N, K) in [(1536, 512), (2048, 512), (512, 2048), (512, 512), (1024, 4096)]:
    Mc = M if (N, K) != (1024, 4096) else 1024
    dy = torch.randn((Mc, N), device="cuda").bfloat16()
    x = torch.randn((Mc, K), device="cuda").bfloat16()
    gw = torch.zeros((N, K), device="cuda", dtype=torch.float32)
    res = []
    for s in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16):
        if s > max(1, Mc // 64):
            continue
        t, _ = timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=gw, splits=s, K=Mc), n=3, reps=8)
        res.append((s, t * 1e3))
    best = min(res, key=lambda r: r[1])
    print(f"dW[{N},{K}] over {Mc} rows: " + "  ".join(f"s={s}:{t:.1f}" for s, t in res) + f"   best s={best[0]} ({best[1]:.1f} us)", flush=True)
