import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, vds_b200
from vds_b200 import ops, lib
from attn_bench import timeit
M, h = 16416, 512
a = torch.randn((M, h), device="cuda").bfloat16(); w1 = (torch.randn((4*h, h), device="cuda")*0.05).bfloat16(); b1 = torch.randn((4*h,), device="cuda").bfloat16()
mn, _ = timeit(lambda: ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU)); print("bias_gelu", mn*1e3, "us")
pre, act = ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU)
dy = torch.randn((M, h), device="cuda").bfloat16(); w2 = (torch.randn((h, 4*h), device="cuda")*0.05).bfloat16()
mn, _ = timeit(lambda: ops.gemm(dy, w2, b_mn=True, epilogue=lib.EPI_DGELU, aux=pre)); print("dgelu", mn*1e3, "us")
mn, _ = timeit(lambda: ops.gemm(a, w1, bias=b1)); print("plain", mn*1e3, "us")
x = torch.randn((M, h), device="cuda").bfloat16(); gate = torch.randn((2, 9*h), device="cuda").bfloat16()
w2t = (torch.randn((h, 4*h), device="cuda")*0.05).bfloat16()
mn, _ = timeit(lambda: ops.gemm(act, w2t, epilogue=lib.EPI_GATE_RES, aux=x, gate=gate[:, :h], rows_per_batch=8208)); print("gate_res K=2048", mn*1e3, "us")
x4 = torch.randn((M, 4*h), device="cuda").bfloat16(); gate4 = torch.randn((2, 4*h), device="cuda").bfloat16()
mn, _ = timeit(lambda: ops.gemm(a, w1, epilogue=lib.EPI_GATE_RES, aux=x4, gate=gate4, rows_per_batch=8208)); print("gate_res N=2048 K=512 (2 outputs + aux read)", mn*1e3, "us")
o1 = torch.empty((M, 4*h), device="cuda").bfloat16(); o2 = torch.empty_like(o1)
mn, _ = timeit(lambda: (ops.gemm(a, w1, bias=b1, out=o1), ops.gemm(a, w1, bias=b1, out=o2))); print("two plain GEMMs", mn*1e3, "us")
mn, _ = timeit(lambda: ops.gemm(a, w1, bias=b1, tile_n=128, cluster=1)); print("plain 1-CTA BN=128", mn*1e3, "us")
mn, _ = timeit(lambda: ops.gemm(a, w1, bias=b1, cluster=1)); print("plain 1-CTA BN=256", mn*1e3, "us")
# dgrad of attn_proj (N = K = 512) with / without the fused delta = rowsum(dO * O)
wp = (torch.randn((h, h), device="cuda")*0.05).bfloat16(); oo = torch.randn((M, h), device="cuda").bfloat16()
delta = torch.zeros((2, 4, 8208), device="cuda", dtype=torch.float32)
mn, _ = timeit(lambda: ops.gemm(dy, wp, b_mn=True)); print("dgrad proj plain", mn*1e3, "us")
mn, _ = timeit(lambda: ops.gemm_dgrad_rowdot(dy, wp, oo, delta, 8208)); print("dgrad proj + rowdot epilogue", mn*1e3, "us")
