"""CTA-pair attention backward (attention_bwd2.cu) against the 1-CTA kernel at the bench shape: same inputs, results
compared tensor by tensor, both timed (CUDA graph of 4 calls, L2 flushed before each replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib as L, ops
from attn_bench import timeit


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def main():
    dev = "cuda"
    shapes = [(2, 4, 8208, 8208), (2, 9, 2064, 2064), (1, 16, 8208, 8208)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for (B, nh, Lq, Lk) in shapes:
        h = nh * 128
        torch.manual_seed(0)
        qkv = torch.randn((B * Lq, 3 * h), device=dev).bfloat16()
        q, k, v = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
        out, lse = ops.attn_fwd(q, k, v, B, nh, Lq, Lk)
        d_o = torch.randn((B * Lq, h), device=dev).bfloat16()
        res = {}
        for mode in (0, 1):
            L.check(L.lib().vds_debug_attn_pair_mode(mode))
            dq = torch.zeros((B * Lq, h), device=dev, dtype=torch.float32)
            dk = torch.zeros((B * Lk, h), device=dev).bfloat16()
            dv = torch.zeros_like(dk)
            ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk=dk, dv=dv)
            torch.cuda.synchronize()
            res[mode] = (dq.clone(), dk.clone(), dv.clone())
            fl = 8.0 * B * nh * Lq * Lk * 128
            mn, av = timeit(lambda: ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk=dk, dv=dv), n=5, reps=4)
            print(f"bwd B={B} nh={nh} L={Lq} pair_mode={mode}: {mn*1e3:8.1f} us  {fl/mn/1e9:7.1f} TFLOP/s (algorithmic)", flush=True)
        for name, a, b in zip(("dq", "dk", "dv"), res[0], res[1]):
            err = (a.float() - b.float()).abs().max().item() / (a.float().abs().max().item() + 1e-30)
            print(f"   {name}: cosine(pair, 1-CTA) = {cos(a, b):.6f}, max rel err {err:.3e}, finite {bool(torch.isfinite(b.float()).all())}")
    L.check(L.lib().vds_debug_attn_pair_mode(-1))


if __name__ == "__main__":
    main()
