import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, vds_b200
from vds_b200 import ops, lib
from attn_bench import timeit
M, h = 16416, 512
dy = torch.randn((M, h), device="cuda").bfloat16(); w2 = (torch.randn((h, 4*h), device="cuda")*0.05).bfloat16()
pre = torch.randn((M, 4*h), device="cuda").bfloat16()
a = torch.randn((M, h), device="cuda").bfloat16(); w1 = (torch.randn((4*h, h), device="cuda")*0.05).bfloat16(); b1 = torch.randn((4*h,), device="cuda").bfloat16()
for cl in (0, 1):
    mn, _ = timeit(lambda: ops.gemm(dy, w2, b_mn=True, epilogue=lib.EPI_DGELU, aux=pre, cluster=cl)); print("dgelu cluster", cl, mn*1e3, "us")
    mn, _ = timeit(lambda: ops.gemm(dy, w2, b_mn=True, cluster=cl)); print("dgrad plain cluster", cl, mn*1e3, "us")
    mn, _ = timeit(lambda: ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU, cluster=cl)); print("bias_gelu cluster", cl, mn*1e3, "us")
