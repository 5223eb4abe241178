// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-B swizzle, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (one thread issues tcgen05.mma 128 x BN x 16; accumulators in TMEM,
//                               double-buffered so the epilogue of tile i overlaps the mainloop of i+1)
//   warps 2..5  epilogue       (tcgen05.ld: thread == output row; fused bias / erf-GELU / gate+residual /
//                               GELU' / fp32 split-K reduction)
// Both operands may be K-major or MN-major (runtime-free: template flags), which covers fprop
// (K,K), dgrad (K,MN) and wgrad (MN,MN) without any transpose copies.
//
// Replaces: every nn.Linear on the reference hot path (model.py:62-90,125-165,318-350,374-390),
// Conv3d patch-embed as a GEMM (model.py:173-185) and their autograd (train.py:432).
#include <cuda.h>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace vds {

constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane quadrant)

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES = 8 * 4096;  // per epilogue warp: 32 rows x 128 B transpose buffer
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const GemmDev p, int tma_c) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);

  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES;   // behind the staging tiles (those are 1024-byte aligned: TMA swizzle atom)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);   // every CTA of the cluster releases the slot (its B half is multicast)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  pdl_wait();      // everything above is launch-independent set-up; global inputs may come from the previous kernel
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;

  // CL == 2: the two CTAs of a cluster work on vertically adjacent tiles (same n-block, m-blocks 2p and 2p+1) in
  // lockstep; each loads its own A tile and HALF of the shared B tile, multicast into both CTAs' shared memory,
  // which cuts the L2 -> SM operand traffic per CTA from A+B to A+B/2 (the K<=2048 GEMMs here are L2-bound).
  const int m_tiles_real = (p.M + BM - 1) / BM;
  const int m_tiles = (CL > 1) ? ((m_tiles_real + CL - 1) / CL) : m_tiles_real;   // in units of cluster rows
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_iters = (p.K + BK - 1) / BK;
  const int total_tiles = m_tiles * n_tiles * p.splits;
  const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int mt = (tile % m_tiles) * CL + crank;
        const int rest = tile / m_tiles;
        const int nt = rest % n_tiles;
        const int sp = rest / n_tiles;
        const int kb0 = sp * p.k_per_split;
        const int kb1 = min(k_iters, kb0 + p.k_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + Cfg::A_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, full_bar(stage), kb * BK, mt * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d(a_dst + i * (BK * 128), &tmA, full_bar(stage), mt * BM + i * 64, kb * BK);
          }
          if constexpr (CL == 1) {
            if constexpr (!B_MN) {
              tma_load_2d(b_dst, &tmB, full_bar(stage), kb * BK, nt * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(b_dst + i * (BK * 128), &tmB, full_bar(stage), nt * BN + i * 64, kb * BK);
            }
          } else {
            constexpr uint16_t kMask = (1u << CL) - 1;
            if constexpr (!B_MN) {   // this CTA fetches rows [crank*BN/2, +BN/2) of the B tile for both CTAs
              tma_load_2d_mc(b_dst + crank * (BN / 2) * 128, &tmB, full_bar(stage), kb * BK,
                             nt * BN + crank * (BN / 2), kMask);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) {
                const int bi = crank * (BN / 128) + i;
                tma_load_2d_mc(b_dst + bi * (BK * 128), &tmB, full_bar(stage), nt * BN + bi * 64, kb * BK, kMask);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // warp-uniform control flow (all lanes wait on the barriers), one elected lane issues MMAs + commits
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int rest = tile / m_tiles;
      const int sp = rest / n_tiles;
      const int kb0 = sp * p.k_per_split;
      const int kb1 = min(k_iters, kb0 + p.k_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? umma_smem_desc(a_addr + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(b_addr + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CL == 1) umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs have read it
          else umma_commit_mc(empty_bar(stage), (1u << CL) - 1);   // ... in BOTH CTAs (the peer multicasts into it)
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int mt = (tile % m_tiles) * CL + crank;
      const int nt = (tile / m_tiles) % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const uint32_t t_base = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      constexpr bool kF32 = (EPI == VDS_EPI_ACCUM_F32 || EPI == VDS_EPI_STORE_F32);
      if constexpr (kF32) {
        uint8_t* stg = smem_gen + STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
#pragma unroll 1
        for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
          uint32_t v[32];
          tmem_ld32(t_base + c * 32, v);
          tmem_ld_wait();
          if (EPI == VDS_EPI_ACCUM_F32 && tma_c) accum_f32_tma(&tmC, stg, mt * BM + q * 32, nt * BN + c * 32, v, lane);
          else epilogue_row<EPI>(p, row, nt * BN + c * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      } else {
        uint8_t* stg = smem_gen + STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
        constexpr int GROUPS = BN / 128;   // 64-column groups per warp (its half of the tile)
#pragma unroll 1
        for (int gi = 0; gi < GROUPS; ++gi) {
          const int cg = chalf * GROUPS + gi;
          uint32_t r0[32], r1[32];
          tmem_ld32(t_base + cg * 64, r0);
          tmem_ld32(t_base + cg * 64 + 32, r1);
          tmem_ld_wait();
          if (gi == GROUPS - 1) {          // accumulator fully read: hand it back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          epilogue_group64<EPI>(p, stg, mt * BM + q * 32, nt * BN + cg * 64, r0, r1, lane);
        }
      }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait_group0();   // TMA reductions of the fp32 epilogue complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // nobody exits while the peer may still multicast / arrive here
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CL = 1>
static int launch_gemm(const vds_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!A_MN) { dims[0] = a.K; dims[1] = a.M; box[0] = BK; box[1] = BM; }
    else       { dims[0] = a.M; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a.lda * 2;
    int r = encode_tmap_bf16(&tmA, a.A, 2, dims, strides, box);
    if (r) return r;
    if (!B_MN) { dims[0] = a.K; dims[1] = a.N; box[0] = BK; box[1] = BN / CL; }
    else       { dims[0] = a.N; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a.ldb * 2;
    r = encode_tmap_bf16(&tmB, a.B, 2, dims, strides, box);
    if (r) return r;
  }
  CUtensorMap tmC = tmA;
  int tma_c = 0;
  if (EPI == VDS_EPI_ACCUM_F32) {
    int r = make_tmap_accum_f32(&tmC, a.C, a.ldc, a.M, a.N, &tma_c);
    if (r) return r;
  }
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  const int k_iters = (a.K + BK - 1) / BK;
  int splits = (EPI == VDS_EPI_ACCUM_F32) ? (a.splits < 1 ? 1 : a.splits) : 1;
  if (splits > k_iters) splits = k_iters;
  p.k_per_split = (k_iters + splits - 1) / splits;
  p.splits = (k_iters + p.k_per_split - 1) / p.k_per_split;  // every split non-empty
  p.C = a.C; p.ldc = a.ldc; p.C2 = a.C2; p.ldc2 = a.ldc2;
  p.bias = reinterpret_cast<const bf16*>(a.bias);
  p.aux = reinterpret_cast<const bf16*>(a.aux); p.ldaux = a.ldaux;
  p.gate = reinterpret_cast<const bf16*>(a.gate); p.gate_stride = a.gate_stride;
  p.rows_per_batch = a.rows_per_batch > 0 ? a.rows_per_batch : 1;
  p.remap_rows = a.remap_rows; p.remap_stride = a.remap_stride; p.remap_offset = a.remap_offset;
  p.v0 = nullptr; p.ldv0 = 0; p.lambda = nullptr;   // (2-CTA QKV_ROPE epilogue only)
  p.dbg = nullptr;

  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return VDS_ERR_CUDA;
    }
    attr_set = true;
  }
  const int m_tiles = ((a.M + BM - 1) / BM + CL - 1) / CL, n_tiles = (a.N + BN - 1) / BN;
  const long long total = (long long)m_tiles * n_tiles * p.splits;   // cluster-tiles
  const int max_clusters = num_sms() / CL;
  const int grid = (int)(total < max_clusters ? total : max_clusters) * CL;
  if constexpr (CL == 1) {
    launch_k(kern, grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, p, tma_c);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see common.h: launch_k
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p, tma_c);
    if (e != cudaSuccess) {
      set_error("gemm: cluster launch failed: %s", cudaGetErrorString(e));
      return VDS_ERR_CUDA;
    }
  }
  VDS_CHECK_LAUNCH("gemm");
  return VDS_OK;
}

// 128 x 256 tiles double the flops per operand byte (the K = 512 GEMMs of the debug model are L2-bandwidth
// bound with 128 x 128 tiles); used when N is a multiple of 256 and there are enough tiles to fill the chip.
template <bool A_MN, bool B_MN, int EPI>
static int launch_bn(const vds_gemm_args& a, cudaStream_t s) {
  const long long tiles256 = (long long)((a.M + BM - 1) / BM) * (a.N / 256);
  const bool cluster_ok = (a.M + BM - 1) / BM >= 2 && a.cluster == 2;   // opt-in: measured no gain on B200 (TMA multicast does not dedup L2 reads at cluster size 2)
  if (a.N % 256 == 0 && tiles256 >= num_sms() && a.tile_n != 128)
    return cluster_ok ? launch_gemm<256, A_MN, B_MN, EPI, 2>(a, s) : launch_gemm<256, A_MN, B_MN, EPI, 1>(a, s);
  return cluster_ok ? launch_gemm<128, A_MN, B_MN, EPI, 2>(a, s) : launch_gemm<128, A_MN, B_MN, EPI, 1>(a, s);
}

template <bool A_MN, bool B_MN>
static int dispatch_epi(const vds_gemm_args& a, cudaStream_t s) {
  switch (a.epilogue) {
    case VDS_EPI_STORE: return launch_bn<A_MN, B_MN, VDS_EPI_STORE>(a, s);
    case VDS_EPI_ACCUM_F32:
      return ((a.M + BM - 1) / BM >= 2 && a.cluster == 2) ? launch_gemm<128, A_MN, B_MN, VDS_EPI_ACCUM_F32, 2>(a, s)
                                                          : launch_gemm<128, A_MN, B_MN, VDS_EPI_ACCUM_F32, 1>(a, s);
    case VDS_EPI_STORE_F32: return launch_gemm<128, A_MN, B_MN, VDS_EPI_STORE_F32, 1>(a, s);
    default: break;
  }
  if constexpr (!A_MN) {
    switch (a.epilogue) {
      case VDS_EPI_BIAS_GELU: if constexpr (!B_MN) return launch_bn<A_MN, B_MN, VDS_EPI_BIAS_GELU>(a, s); break;
      case VDS_EPI_GATE_RES: if constexpr (!B_MN) return launch_bn<A_MN, B_MN, VDS_EPI_GATE_RES>(a, s); break;
      case VDS_EPI_DGELU: if constexpr (B_MN) return launch_bn<A_MN, B_MN, VDS_EPI_DGELU>(a, s); break;
      default: break;
    }
  }
  set_error("gemm: epilogue %d not available for a_mn=%d b_mn=%d", a.epilogue, (int)A_MN, (int)B_MN);
  return VDS_ERR_UNSUPPORTED;
}

int gemm2_dispatch(const vds_gemm_args& a, cudaStream_t s);   // gemm2.cu
void gemm2_set_trace(long long* p);

}  // namespace vds

extern "C" int vds_debug_gemm2_trace(void* buf) {
  vds::gemm2_set_trace((long long*)buf);
  return VDS_OK;
}

extern "C" int vds_gemm(const vds_gemm_args* args, void* stream) {
  using namespace vds;
  VDS_CHECK_ARG(args != nullptr, "gemm: null args");
  const vds_gemm_args& a = *args;
  VDS_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  VDS_CHECK_ARG(a.N % 8 == 0, "gemm: N=%d must be a multiple of 8", a.N);
  VDS_CHECK_ARG(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm: lda=%lld ldb=%lld must be multiples of 8",
                (long long)a.lda, (long long)a.ldb);
  VDS_CHECK_ARG(((uintptr_t)a.A & 15) == 0 && ((uintptr_t)a.B & 15) == 0, "gemm: A/B must be 16-byte aligned");
  VDS_CHECK_ARG(a.C != nullptr || a.C2 != nullptr, "gemm: no output");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // 2-CTA (cta_group::2) 256 x 256 tiles when they can keep every SM pair busy
  if (a.cluster != 1 && a.tile_n != 128 && a.N % 64 == 0 && a.N >= 256) {
    const int k_iters = (a.K + BK - 1) / BK;
    int splits = (a.epilogue == VDS_EPI_ACCUM_F32) ? (a.splits < 1 ? 1 : a.splits) : 1;
    if (splits > k_iters) splits = k_iters;
    const long long pair_tiles = (long long)((a.M + 255) / 256) * ((a.N + 255) / 256) * splits;
    if (pair_tiles >= num_sms() / 2) {
      const int r = gemm2_dispatch(a, s);
      if (r != VDS_ERR_UNSUPPORTED) return r;
    }
  }
  if (!a.a_mn && !a.b_mn) return dispatch_epi<false, false>(a, s);
  if (!a.a_mn && a.b_mn) return dispatch_epi<false, true>(a, s);
  if (a.a_mn && a.b_mn) return dispatch_epi<true, true>(a, s);
  return dispatch_epi<true, false>(a, s);
}
