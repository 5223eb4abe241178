"""Micro-benchmark of the attention kernels at the debug-8k shapes (CUDA events, L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200
from vds_b200 import ops

def timeit(fn, n=5, reps=8):
    """GPU time per call: `reps` calls captured in a CUDA graph (no host launch gaps), L2 flushed before each replay.
    (Within a replay, calls 2..reps may find inputs in L2 when the working set is < 126 MB.)"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return min(ts), sum(ts) / len(ts)

def main():
    dev = "cuda"
    for (B, nh, Lq, Lk) in [(2, 4, 8208, 8208), (2, 4, 8208, 512), (8, 6, 272, 272), (2, 9, 2064, 2064)]:
        h = nh * 128
        qkv = torch.randn((B * Lq, 3 * h), device=dev).bfloat16()
        if Lq == Lk:
            q, k, v = qkv[:, :h], qkv[:, h:2*h], qkv[:, 2*h:]
        else:
            q = qkv[:, :h]
            kv = torch.randn((B * Lk, 2 * h), device=dev).bfloat16()
            k, v = kv[:, :h], kv[:, h:]
        out, lse = ops.attn_fwd(q, k, v, B, nh, Lq, Lk)
        d_o = torch.randn((B * Lq, h), device=dev).bfloat16()
        fl = 4.0 * B * nh * Lq * Lk * 128
        mn, av = timeit(lambda: ops.attn_fwd(q, k, v, B, nh, Lq, Lk))
        print(f"fwd B={B} nh={nh} Lq={Lq} Lk={Lk}: {mn*1e3:8.1f} us  {fl/mn/1e9:7.1f} TFLOP/s")
        dq = torch.zeros((B * Lq, h), device=dev, dtype=torch.float32)
        dk = torch.zeros((B * Lk, h), device=dev).bfloat16(); dv = torch.zeros_like(dk)
        mn, av = timeit(lambda: ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lk, dq, dk=dk, dv=dv))
        print(f"bwd B={B} nh={nh} Lq={Lq} Lk={Lk}: {mn*1e3:8.1f} us  {2*fl/mn/1e9:7.1f} TFLOP/s (algorithmic 2x fwd)")
    # GEMM shapes of the debug model
    from vds_b200 import lib
    for (M, N, K, tag) in [(16416, 1536, 512, "qkv"), (16416, 2048, 512, "mlp1"), (16416, 512, 2048, "mlp2"),
                           (16416, 512, 512, "proj"), (1024, 1024, 4096, "ctx_kv")]:
        a = torch.randn((M, K), device=dev).bfloat16(); b = torch.randn((N, K), device=dev).bfloat16()
        mn, _ = timeit(lambda: ops.gemm(a, b))
        print(f"gemm {tag} {M}x{N}x{K}: {mn*1e3:8.1f} us {2.0*M*N*K/mn/1e9:7.1f} TFLOP/s")
        dy = torch.randn((M, N), device=dev).bfloat16()
        w = torch.zeros((N, K), device=dev, dtype=torch.float32)
        for sp in (1, 4, 8):
            mn, _ = timeit(lambda: ops.gemm(dy, a, a_mn=True, b_mn=True, epilogue=lib.EPI_ACCUM_F32, out=w, splits=sp))
            print(f"  wgrad {tag} splits={sp}: {mn*1e3:8.1f} us {2.0*M*N*K/mn/1e9:7.1f} TFLOP/s")
        mn, _ = timeit(lambda: ops.gemm(dy, b, b_mn=True))
        print(f"  dgrad {tag}: {mn*1e3:8.1f} us {2.0*M*N*K/mn/1e9:7.1f} TFLOP/s")

if __name__ == "__main__":
    main()


def misc():
    dev = "cuda"
    B, Lr, h = 2, 8208, 512
    x = torch.randn((B * Lr, h), device=dev).bfloat16(); dy = torch.randn_like(x); res = torch.randn_like(x)
    mod = torch.randn((B, 9 * h), device=dev).bfloat16()
    y, rstd = ops.rmsnorm_mod_fwd(x, B, Lr, h, scale=mod[:, h:2*h], shift=mod[:, :h])
    dmod = torch.zeros((B, 9 * h), device=dev, dtype=torch.float32)
    mn, _ = timeit(lambda: ops.rmsnorm_mod_fwd(x, B, Lr, h, scale=mod[:, h:2*h], shift=mod[:, :h]))
    print(f"rmsnorm_fwd: {mn*1e3:.1f} us ({(2*x.numel()*2)/mn/1e6:.0f} GB/s)")
    mn, _ = timeit(lambda: ops.rmsnorm_mod_bwd(dy, x, rstd, B, Lr, h, scale=mod[:, h:2*h], dx_res=res, dscale=dmod[:, h:2*h], dshift=dmod[:, :h]))
    print(f"rmsnorm_bwd: {mn*1e3:.1f} us ({(4*x.numel()*2)/mn/1e6:.0f} GB/s)")
    mn, _ = timeit(lambda: ops.gate_bwd(dy, x, mod[:, 2*h:3*h], dmod[:, 2*h:3*h], B, Lr, h))
    print(f"gate_bwd: {mn*1e3:.1f} us ({(3*x.numel()*2)/mn/1e6:.0f} GB/s)")
    big = torch.randn((B * Lr, 4 * h), device=dev).bfloat16(); o = torch.zeros(4 * h, device=dev)
    mn, _ = timeit(lambda: ops.colsum(big, o))
    print(f"colsum 4h: {mn*1e3:.1f} us ({big.numel()*2/mn/1e6:.0f} GB/s)")


if __name__ == "__main__":
    misc()
