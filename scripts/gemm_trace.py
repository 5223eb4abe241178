import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vds_b200
from vds_b200 import ops, lib
M, h = 16416, 512
a = torch.randn((M, h), device="cuda").bfloat16(); w1 = (torch.randn((4*h, h), device="cuda")*0.05).bfloat16(); b1 = torch.randn((4*h,), device="cuda").bfloat16()
x4 = torch.randn((M, 4*h), device="cuda").bfloat16(); gate4 = torch.randn((2, 4*h), device="cuda").bfloat16()
a4 = torch.randn((M, 4*h), device="cuda").bfloat16(); w2 = (torch.randn((h, 4*h), device="cuda")*0.05).bfloat16()
tr = torch.zeros(16, device="cuda", dtype=torch.int64)
def run(name, fn):
    fn(); torch.cuda.synchronize()
    tr.zero_(); lib.lib().vds_debug_gemm2_trace(tr.data_ptr()); fn(); torch.cuda.synchronize(); lib.lib().vds_debug_gemm2_trace(None)
    t = tr.tolist()
    if t[14] or t[12]:
        n = max(1, t[3])
        print(f"{name} [fused fast path] MMA {t[0]/n:.0f}/tile (wait tmem-empty {t[1]/n:.0f}, smem-full {t[2]/n:.0f}); epilogue warp per tile: wait-acc {t[8]/n:.0f} tmem-ld {t[9]/n:.0f} aux {t[10]/n:.0f} math {t[11]/n:.0f} wait-read {t[12]/n:.0f} sts+fence {t[13]/n:.0f} store-issue {t[14]/n:.0f}")
        return
    print(f"{name}: MMA thread total {t[0]} cyc over {t[3]} tiles = {t[0]/max(1,t[3]):.0f}/tile; waiting tmem-empty {t[1]/max(1,t[3]):.0f}/tile, smem-full {t[2]/max(1,t[3]):.0f}/tile; epilogue warp: wait {t[4]/max(1,t[3]):.0f}/tile busy {t[5]/max(1,t[3]):.0f}/tile (tmem ld {t[6]/max(1,t[3]):.0f}, group body {t[7]/max(1,t[3]):.0f}); per tile: compute+STS {t[8]/max(1,t[3]):.0f} syncwarp {t[9]/max(1,t[3]):.0f} stores {t[10]/max(1,t[3]):.0f} tail-sync {t[11]/max(1,t[3]):.0f}")
run("plain N=2048 K=512", lambda: ops.gemm(a, w1, bias=b1))
run("bias_gelu", lambda: ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU))
run("gate_res N=2048", lambda: ops.gemm(a, w1, epilogue=lib.EPI_GATE_RES, aux=x4, gate=gate4, rows_per_batch=8208))
run("plain N=512 K=2048", lambda: ops.gemm(a4, w2))
dy = torch.randn((M, h), device="cuda").bfloat16(); hpre = torch.randn((M, 4*h), device="cuda").bfloat16()
run("dgelu", lambda: ops.gemm(dy, w2, b_mn=True, epilogue=lib.EPI_DGELU, aux=hpre))
