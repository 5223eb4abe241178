"""QKV projection at the bench shapes: plain GEMM + in-place qkv_post_fwd pass against the fused VDS_EPI_QKV_ROPE epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import ops
from attn_bench import timeit

for (B, Lr, nh, K) in [(2, 8208, 4, 512), (8, 272, 6, 768), (2, 2064, 9, 1152)]:
    h = nh * 128
    x = torch.randn((B * Lr, K), device="cuda").bfloat16()
    w = (torch.randn((3 * h, K), device="cuda") * K ** -0.5).bfloat16()
    ang = torch.randn((Lr, 64), device="cuda") * 3
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    v0 = torch.randn((B * Lr, h), device="cuda").bfloat16()
    lam = torch.tensor([0.4], device="cuda").bfloat16()
    tab = ops.rope_pack(cos, sin)

    def unfused():
        qkv = ops.gemm(x, w)
        ops.qkv_post_fwd(qkv, B, Lr, h, nh, cos=cos, sin=sin, v0=v0, v0_ld=h, lam=lam)

    t_plain, _ = timeit(lambda: ops.gemm(x, w), n=5, reps=8)
    t_un, _ = timeit(unfused, n=5, reps=8)
    t_f, _ = timeit(lambda: ops.gemm_qkv_rope(x, w, None, tab, Lr, v0=v0, v0_ld=h, lam=lam), n=5, reps=8)
    print(f"B={B} L={Lr} h={h}: plain GEMM {t_plain*1e3:.1f} us, GEMM + qkv_post_fwd {t_un*1e3:.1f} us, fused epilogue {t_f*1e3:.1f} us", flush=True)
