// Tuning aid: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, M=128, K=16) by N, operand source and major-ness.
// One CTA per SM issues R back-to-back MMAs into one accumulator and waits for the commit; prints cycles / MMA of CTA 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-diffusion-speedrun_b200/csrc scripts/mma_shapes.cu -o gpurun_out/mma_shapes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vds;

// mode: 0 SS (A K-major) | 1 TS (A in TMEM) | 2 SS (A MN-major, B MN-major) | 3 alternate SS N / TS 128 (mix)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int R, int distinct) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t bar = base, slot = base + 16, tiles = base + 1024;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(gen + 1024)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16);
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MODE == 2, MODE == 2);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < R; ++i) {
          const int kk = distinct ? (i & 7) : 0;
          const uint32_t a = tiles + (kk >> 2) * 16384 + (kk & 3) * 32, b = tiles + 32768 + (kk >> 2) * 16384 + (kk & 3) * 32;
          if (MODE == 0) umma_bf16(tmem, umma_smem_desc(a, 16, 1024), umma_smem_desc(b, 16, 1024), idesc, i > 0);
          if (MODE == 1) umma_bf16_ts(tmem, tmem + 256 + kk * 8, umma_smem_desc(b, 16, 1024), idesc, i > 0);
          if (MODE == 2)
            umma_bf16(tmem, umma_smem_desc(tiles + kk * 2048, 16384, 1024), umma_smem_desc(tiles + 32768 + kk * 2048, 16384, 1024),
                      idesc, i > 0);
        }
        umma_commit(bar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, rep & 1);
      t2 = clock64();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(const char* name, int distinct) {
  long long* d; cudaMalloc(&d, 16);
  const int R = 64, smem = 162 * 1024;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, MODE><<<148, 128, smem>>>(d, R, distinct);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-28s N=%3d distinct=%d  issue %6.1f cyc/MMA   complete %6.1f cyc/MMA  (floor %d) %s\n", name, N, distinct, h[0] / (double)R,
         h[1] / (double)R, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int distinct = 0; distinct < 2; ++distinct) {
    run<64, 0>("SS K-major", distinct);
    run<128, 0>("SS K-major", distinct);
    run<256, 0>("SS K-major", distinct);
    run<64, 1>("TS (A in TMEM)", distinct);
    run<128, 1>("TS (A in TMEM)", distinct);
    run<256, 1>("TS (A in TMEM)", distinct);
    run<64, 2>("SS MN-major A and B", distinct);
    run<128, 2>("SS MN-major A and B", distinct);
  }
  return 0;
}
