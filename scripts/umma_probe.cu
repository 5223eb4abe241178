// Tuning / discovery aid (not part of the product library): runs a short series of tcgen05.mma instructions whose
// operands, descriptors and instruction descriptor all come from the host, in a 1-CTA kernel (cta_group::1) or in a
// 2-CTA cluster (cta_group::2), and dumps the accumulator as it sits in tensor memory (every lane x column of each
// CTA).  scripts/umma_probe.py builds operand images for a hypothesis about a shared-memory / TMEM layout, runs it and
// compares the dump with the expected product — the layouts the attention backward relies on are pinned this way
// instead of being taken from memory.  Also reports the cycles the series took (issue -> commit arrival).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -I video-diffusion-speedrun_b200/csrc \
//             scripts/umma_probe.cu -o scripts/_build/libumma_probe.so
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace vds;

struct ProbeArgs {
  const uint8_t* a_img;      // [ncta][a_bytes] shared-memory image of the A region of each CTA
  const uint8_t* b_img;      // [ncta][b_bytes]
  const uint32_t* a_tmem;    // TS mode: [ncta][128][a_tmem_cols] words stored to TMEM columns [384, 384 + a_tmem_cols)
  float* out;                // [ncta][128][d_cols]
  long long* cycles;         // [2]: issue, issue -> complete
  unsigned long long a_desc; // descriptor without the start-address field (address of the A region is added in-kernel)
  unsigned long long b_desc;
  int a_bytes, b_bytes, a_tmem_cols;
  int a_step, b_step;        // bytes added to the operand address per k-step (TS: A advances a_step COLUMNS)
  unsigned idesc;
  int ksteps, d_cols, reps;
  int remote_b;              // 1: each CTA's B image is written by its PEER through distributed shared memory
  int ts;                    // 1: A operand from TMEM
  int d_alt;                 // timing: rotate over this many accumulators (128 columns apart) instead of chaining on one
};

constexpr int A_REGION = 1024, B_REGION = 1024 + 96 * 1024, SMEM_BYTES = 200 * 1024;

template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeArgs p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t bar = base, slot = base + 16;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  // operand images
  for (int i = threadIdx.x; i < p.a_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(gen + A_REGION)[i] = reinterpret_cast<const uint4*>(p.a_img + (size_t)rank * p.a_bytes)[i];
  if (CG == 2 && p.remote_b) {
    __syncthreads();
    cluster_sync_all();   // peer's smem is live
    const int peer = rank ^ 1;
    for (int i = threadIdx.x; i < p.b_bytes / 16; i += 128) {
      const uint4 v = reinterpret_cast<const uint4*>(p.b_img + (size_t)peer * p.b_bytes)[i];
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + B_REGION + i * 16), "r"(peer));
      asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (incl. the remote ones) -> async proxy
  } else {
    for (int i = threadIdx.x; i < p.b_bytes / 16; i += 128)
      reinterpret_cast<uint4*>(gen + B_REGION)[i] = reinterpret_cast<const uint4*>(p.b_img + (size_t)rank * p.b_bytes)[i];
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    if (CG == 1) {
      tmem_alloc(slot, 512);
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16);
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  if (p.ts) {   // A operand image into TMEM columns 384.. (thread == lane)
    for (int c = 0; c < p.a_tmem_cols; ++c) {
      uint32_t v = p.a_tmem[((size_t)rank * 128 + threadIdx.x) * p.a_tmem_cols + c];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_off + 384 + c), "r"(v) : "memory");
    }
    tmem_st_wait();
  }
  {  // clear the accumulator columns so untouched lanes / columns read back as a recognisable value
    for (int c = 0; c < p.d_cols; ++c) {
      uint32_t v = 0x7fc00000u;   // NaN
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_off + c), "r"(v) : "memory");
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  if (warp == 0 && rank == 0) {
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < p.reps; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int k = 0; k < p.ksteps; ++k) {
          const uint32_t a_addr = base + A_REGION + k * p.a_step, b_addr = base + B_REGION + k * p.b_step;
          const uint64_t ad = p.a_desc | (uint64_t)((a_addr & 0x3FFFF) >> 4);
          const uint64_t bd = p.b_desc | (uint64_t)((b_addr & 0x3FFFF) >> 4);
          const int alt = p.d_alt > 1 ? p.d_alt : 1;
          const uint32_t acc = k >= alt ? 1u : 0u;
          const uint32_t dt = tmem + (k % alt) * 128;
          if (CG == 1) {
            if (p.ts) umma_bf16_ts(dt, tmem + 384 + k * p.a_step, bd, p.idesc, acc);
            else umma_bf16(dt, ad, bd, p.idesc, acc);
          } else {
            if (p.ts) {
              asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, q;\n}\n" ::"r"(dt),
                           "r"(tmem + 384 + k * p.a_step), "l"(bd), "r"(p.idesc), "r"(acc)
                           : "memory");
            } else {
              asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n}\n" ::"r"(dt),
                           "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc)
                           : "memory");
            }
          }
        }
        if (CG == 1) umma_commit(bar);
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                       "h"((uint16_t)3)
                       : "memory");
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, rep & 1);
      t2 = clock64();
    }
    if (blockIdx.x == 0 && lane == 0 && p.cycles) { p.cycles[0] = t1 - t0; p.cycles[1] = t2 - t0; }
  } else {
    for (int rep = 0; rep < p.reps; ++rep) mbar_wait(bar, rep & 1);
  }
  tc_fence_after();
  for (int c = 0; c < p.d_cols; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + lane_off + c) : "memory");
    tmem_ld_wait();
    p.out[((size_t)rank * 128 + threadIdx.x) * p.d_cols + c] = __uint_as_float(v);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

extern "C" int umma_probe(const ProbeArgs* a, int cg, int nclusters) {
  cudaError_t e;
  if (cg == 1) {
    cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    probe_kernel<1><<<nclusters, 128, SMEM_BYTES>>>(*a);
  } else {
    cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, probe_kernel<2>, *a);
    if (e != cudaSuccess) { printf("probe launch: %s\n", cudaGetErrorString(e)); return -1; }
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe: %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}
