"""Per-sub-tile period of the two attention-backward kernels on a shape that fills whole waves exactly:
B=1, 37 heads, L=8192 -> 2368 kv tiles = 16 waves of 148 CTAs = 16 waves of 74 CTA pairs, 128 query sub-tiles each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vds_b200  # noqa: F401
from vds_b200 import lib as L, ops
from attn_bench import timeit

B, nh, Lq = 1, 37, 8192
h = nh * 128
torch.manual_seed(0)
qkv = torch.randn((B * Lq, 3 * h), device="cuda").bfloat16()
q, k, v = qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:]
out, lse = ops.attn_fwd(q, k, v, B, nh, Lq, Lq)
d_o = torch.randn((B * Lq, h), device="cuda").bfloat16()
dq = torch.zeros((B * Lq, h), device="cuda", dtype=torch.float32)
dk = torch.zeros((B * Lq, h), device="cuda").bfloat16()
dv = torch.zeros_like(dk)
import subprocess
for mode in (0, 1):
    L.check(L.lib().vds_debug_attn_pair_mode(mode))
    mn, av = timeit(lambda: ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, Lq, Lq, dq, dk=dk, dv=dv), n=4, reps=2)
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    fl = 8.0 * B * nh * Lq * Lq * 128
    print(f"pair_mode={mode}: {mn*1e3:8.1f} us = {mn*1e3/16/128*1e3:7.1f} ns per sub-tile per wave; {fl/mn/1e9:7.1f} TFLOP/s algorithmic (SM clock after: {clk} MHz)")
L.check(L.lib().vds_debug_attn_pair_mode(-1))
