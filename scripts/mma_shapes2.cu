// Tuning aid: cycles per tcgen05.mma (kind::f16, bf16, K = 16) in cta_group::1 vs cta_group::2, by shape and operand
// source, with a tight (compile-time descriptor) issue loop like scripts/mma_shapes.cu.  Every SM (pair) runs the series.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-diffusion-speedrun_b200/csrc scripts/mma_shapes2.cu -o scripts/_build/mma_shapes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vds;

__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n}\n" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, q;\n}\n" ::"r"(d), "r"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// MODE 0: SS, A and B K-major SW128.  1: TS (A in TMEM), B MN-major SW128.  2: SS, A MN-major SW128, B MN-major (CG 2: 64-byte
// swizzle, 32 of N per CTA; CG 1: SW128).  M = total M of the instruction.
template <int CG, int M, int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int R) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t bar = base, slot = base + 16, tiles = base + 1024;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(gen + 1024)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    if (CG == 1) tmem_alloc(slot, 512);
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16);
  if (threadIdx.x < 32 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(M, N, MODE == 2, MODE >= 1);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < R; i += 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t acc = (i + kk) > 0;
            if (MODE == 0) {
              const uint64_t a = umma_smem_desc(tiles + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
              const uint64_t b = umma_smem_desc(tiles + 65536 + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
              if (CG == 1) umma_bf16(tmem, a, b, idesc, acc); else mma2(tmem, a, b, idesc, acc);
            } else if (MODE == 1) {
              const uint64_t b = umma_smem_desc(tiles + 65536 + (kk & 3) * 2048, 8192, 1024);
              if (CG == 1) umma_bf16_ts(tmem, tmem + 256 + (kk & 3) * 8, b, idesc, acc);
              else mma2_ts(tmem, tmem + 256 + (kk & 3) * 8, b, idesc, acc);
            } else {
              const uint64_t a = umma_smem_desc(tiles + kk * 2048, 16384, 1024);
              if (CG == 1) umma_bf16(tmem, a, umma_smem_desc(tiles + 65536 + kk * 2048, 16384, 1024), idesc, acc);
              else mma2(tmem, a, desc_sw64(tiles + 65536 + kk * 1024, 2048, 512), idesc, acc);
            }
          }
        }
        if (CG == 1) umma_commit(bar);
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, rep & 1);
      t2 = clock64();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (CG == 2 && rank == 1 && threadIdx.x < 32) {
    for (int rep = 0; rep < 3; ++rep) mbar_wait(bar, rep & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (threadIdx.x < 32) {
    if (CG == 1) tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, int M, int N, int MODE>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  const int R = 64, smem = 170 * 1024;
  cudaFuncSetAttribute(k<CG, M, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<CG, M, N, MODE>, d, R);
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaGetLastError();
  const int floor_c = (M / CG > 128 ? M / CG : 128) * N / 256;
  printf("cta_group::%d M%3d N%3d %-34s issue %6.1f  complete %6.1f cyc/MMA  (floor %3d; MACs/clk/SM %6.0f) %s\n", CG, M, N, name,
         h[0] / (double)R, h[1] / (double)R, floor_c, (double)M / CG * N * 16 / (h[1] / (double)R), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<1, 128, 64, 0>("SS K-major");
  run<2, 256, 64, 0>("SS K-major (B 32 rows / CTA)");
  run<1, 128, 128, 0>("SS K-major");
  run<2, 256, 128, 0>("SS K-major");
  run<1, 128, 256, 0>("SS K-major");
  run<2, 256, 256, 0>("SS K-major");
  run<1, 128, 128, 1>("TS, B MN-major");
  run<2, 256, 128, 1>("TS, B MN-major (64 of N / CTA)");
  run<1, 128, 64, 2>("SS MN-major A and B");
  run<2, 128, 64, 2>("SS MN-major, M 64 / CTA, B sw64");
  run<2, 256, 64, 2>("SS MN-major, B sw64");
  return 0;
}
