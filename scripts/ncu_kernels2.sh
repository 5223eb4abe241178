#!/bin/bash
# ncu --set full: fused QKV epilogue GEMM, plain QKV GEMM, qkv_post_fwd / bwd at debug-8k shapes -> gpurun_out/prof_kernels2_$TAG.ncu-rep
TAG=${1:-r2}
ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|qkv_post" -s 4 -c 4 \
    -o gpurun_out/prof_kernels2_$TAG python scripts/ncu_kernels2.py > gpurun_out/ncu_kernels2_$TAG.log 2>&1
ls -la gpurun_out/prof_kernels2_$TAG.ncu-rep
