#!/bin/bash
# ncu --set full: wgrad GEMMs (TMA reduce-add), cross-attention forward / backward at debug-8k shapes -> gpurun_out/prof_kernels3_$TAG.ncu-rep
TAG=${1:-r2}
ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|gemm_kernel|attn_fwd_kernel|attn_bwd_kernel" -s 4 -c 4 \
    -o gpurun_out/prof_kernels3_$TAG python scripts/ncu_kernels3.py > gpurun_out/ncu_kernels3_$TAG.log 2>&1
ls -la gpurun_out/prof_kernels3_$TAG.ncu-rep
