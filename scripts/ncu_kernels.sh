#!/bin/bash
# ncu --set full of one launch of each memory-bound kernel and each GEMM flavour at debug-8k shapes -> gpurun_out/prof_kernels_$TAG.ncu-rep
TAG=${1:-r2}
ncu --set full --clock-control none --import-source on -k regex:"rmsnorm|gate_bwd|colsum|gemm2_kernel|gemm_kernel" -s 9 -c 9 \
    -o gpurun_out/prof_kernels_$TAG python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_$TAG.log 2>&1
ls -la gpurun_out/prof_kernels_$TAG.ncu-rep
