#!/bin/bash
cat > /tmp/gemm_one.py <<PY
import sys; sys.path.insert(0, ".")
import torch, vds_b200
from vds_b200 import ops, lib
M, h = 16416, 512
a = torch.randn((M, h), device="cuda").bfloat16(); w1 = (torch.randn((4*h, h), device="cuda")*0.05).bfloat16(); b1 = torch.randn((4*h,), device="cuda").bfloat16()
x4 = torch.randn((M, 4*h), device="cuda").bfloat16(); gate4 = torch.randn((2, 4*h), device="cuda").bfloat16()
for _ in range(2):
    ops.gemm(a, w1, bias=b1)
    ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU)
    ops.gemm(a, w1, epilogue=lib.EPI_GATE_RES, aux=x4, gate=gate4, rows_per_batch=8208)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 3 -o gpurun_out/prof_gemm_$1 python /tmp/gemm_one.py > gpurun_out/ncu_gemm_$1.log 2>&1
