import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, vds_b200
from vds_b200 import ops
from attn_bench import timeit
for (B, nh, Lq, Lk) in [(2, 4, 8208, 8208), (2, 4, 8208, 512), (2, 9, 2064, 2064), (8, 16, 8208, 8208)]:
    h = nh * 128
    q = torch.randn((B * Lq, h), device="cuda").bfloat16(); k = torch.randn((B * Lk, h), device="cuda").bfloat16(); v = torch.randn((B * Lk, h), device="cuda").bfloat16()
    fl = 4.0 * B * nh * Lq * Lk * 128
    mn, av = timeit(lambda: ops.attn_fwd(q, k, v, B, nh, Lq, Lk), n=4, reps=4)
    print(f"fwd B={B} nh={nh} Lq={Lq} Lk={Lk}: {mn*1e3:8.1f} us  {fl/mn/1e9:7.1f} TFLOP/s", flush=True)
