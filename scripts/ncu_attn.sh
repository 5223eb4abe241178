#!/bin/bash
# ncu --set full on the self-attention kernels alone (micro-benchmark shapes of the debug-8k workload)
TAG=${1:-x}
cat > /tmp/attn_one.py <<PY
import sys; sys.path.insert(0, ".")
import torch, vds_b200
from vds_b200 import ops
B, nh, L = 2, 4, 8208; h = nh * 128
qkv = torch.randn((B * L, 3 * h), device="cuda").bfloat16()
q, k, v = qkv[:, :h], qkv[:, h:2*h], qkv[:, 2*h:]
d_o = torch.randn((B * L, h), device="cuda").bfloat16()
for _ in range(2):
    out, lse = ops.attn_fwd(q, k, v, B, nh, L, L)
    dq = torch.zeros((B * L, h), device="cuda", dtype=torch.float32)
    dk = torch.zeros((B * L, h), device="cuda").bfloat16(); dv = torch.zeros_like(dk)
    ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
PY
# second iteration: attn_fwd, attn_bwd2 (main, unsplit), attn_bwd2 (remainder, split), fix-up
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 5 -c 5 -o gpurun_out/prof_attn_$TAG python /tmp/attn_one.py > gpurun_out/ncu_attn_$TAG.log 2>&1
ls -la gpurun_out/prof_attn_$TAG.ncu-rep
