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
#include "ptx.cuh"

namespace vds {

constexpr int BM = 128;
constexpr int BK = 64;
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

struct GemmDev {
  int M, N, K;
  int splits, k_per_split;
  void* C;
  long long ldc;
  void* C2;
  long long ldc2;
  const bf16* bias;
  const bf16* aux;
  long long ldaux;
  const bf16* gate;
  long long gate_stride;
  int rows_per_batch;
  int remap_rows, remap_stride, remap_offset;
};

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 rounding of the outputs): the libm
// erff costs ~4x more instructions and made the GELU epilogues, not the MMAs, the bottleneck of the MLP GEMMs.
// Returns erf(x/sqrt2) and exp(-x^2/2) (shared by GELU and its derivative).
__device__ __forceinline__ float erf_as(float x, float& gauss) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  gauss = exp2f(-0.72134752044448170368f * x * x);   // exp(-x^2/2) = exp(-z^2)
  const float e = fmaf(-poly * t, gauss, 1.0f);
  return copysignf(e, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float g;
  return 0.5f * x * (1.0f + erf_as(x, g));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float g;
  const float cdf = 0.5f * (1.0f + erf_as(x, g));
  return fmaf(x * 0.39894228040143267794f, g, cdf);
}

__device__ __forceinline__ void ld8_bf16(const bf16* p, float (&o)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}
__device__ __forceinline__ void st8_bf16(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ---- coalesced bf16 epilogue -------------------------------------------------------------------------------
// TMEM hands each thread one output ROW; storing rows straight from registers makes every warp store touch 32
// different lines (32 partial-sector transactions), which caps the K = 512 GEMMs well below the MMA rate.  Each
// epilogue warp therefore transposes 32 rows x 64 columns through a private 4 KiB shared-memory buffer
// (16-byte chunks XOR-swizzled by row) so that one warp instruction moves 4 full 128-byte row segments.
__device__ __forceinline__ uint32_t stg_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

struct RowMap {   // global row of local row rr (0..31) for loads / stores, -1 if out of range
  int row0, M, remap_rows, remap_stride, remap_offset;
  __device__ __forceinline__ long long out_row(int rr) const {
    const int r = row0 + rr;
    if (r >= M) return -1;
    if (remap_rows > 0) return (long long)(r / remap_rows) * remap_stride + remap_offset + r % remap_rows;
    return r;
  }
  __device__ __forceinline__ long long in_row(int rr) const { return row0 + rr < M ? row0 + rr : -1; }
};

// staging -> global: 8 instructions, lane = (row within group of 4, 16-byte chunk)
__device__ __forceinline__ void stg_store(const uint8_t* stg, bf16* C, long long ldc, const RowMap& rm, int col0,
                                          int N, int lane, bool remap) {
  const int ch = lane & 7, col = col0 + ch * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int rr = k * 4 + (lane >> 3);
    const long long row = remap ? rm.out_row(rr) : rm.in_row(rr);
    const uint4 v = *reinterpret_cast<const uint4*>(stg + stg_off(rr, ch));
    if (row >= 0 && col < N) *reinterpret_cast<uint4*>(C + row * ldc + col) = v;
  }
}
// global -> staging
__device__ __forceinline__ void stg_load(uint8_t* stg, const bf16* A, long long lda, const RowMap& rm, int col0, int N,
                                         int lane) {
  const int ch = lane & 7, col = col0 + ch * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int rr = k * 4 + (lane >> 3);
    const long long row = rm.in_row(rr);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row >= 0 && col < N) v = *reinterpret_cast<const uint4*>(A + row * lda + col);
    *reinterpret_cast<uint4*>(stg + stg_off(rr, ch)) = v;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(uint4 u, float (&o)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}

// One warp, 32 rows x 64 columns [col0, col0 + 64): acc = 64 fp32 per thread (its own row).
template <int EPI>
__device__ __forceinline__ void epilogue_group64(const GemmDev& p, uint8_t* stg, int row0, int col0,
                                                 const uint32_t (&r0)[32], const uint32_t (&r1)[32], int lane) {
  RowMap rm{row0, p.M, p.remap_rows, p.remap_stride, p.remap_offset};
  const int my_row = row0 + lane;
  if constexpr (EPI == VDS_EPI_GATE_RES || EPI == VDS_EPI_DGELU) {
    stg_load(stg, p.aux, p.ldaux, rm, col0, p.N, lane);
    __syncwarp();
  }
  uint4 keep[8];  // first output (bf16 Linear result) kept packed while the second goes through the buffer
  const int b = (EPI == VDS_EPI_GATE_RES) ? min(my_row, p.M - 1) / p.rows_per_batch : 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int col = col0 + g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __uint_as_float(g < 4 ? r0[g * 8 + j] : r1[(g - 4) * 8 + j]);
    const bool col_ok = col < p.N;
    if constexpr (EPI != VDS_EPI_DGELU) {
      if (p.bias != nullptr && col_ok) {
        float bb[8];
        ld8_bf16(p.bias + col, bb);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += bb[j];
      }
    }
    uint8_t* slot = stg + stg_off(lane, g);
    if constexpr (EPI == VDS_EPI_STORE) {
      *reinterpret_cast<uint4*>(slot) = pack8(acc);
    } else if constexpr (EPI == VDS_EPI_BIAS_GELU) {
      float act[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = bf16_round(acc[j]);
        act[j] = gelu_erf(acc[j]);
      }
      keep[g] = pack8(acc);
      *reinterpret_cast<uint4*>(slot) = pack8(act);
    } else if constexpr (EPI == VDS_EPI_GATE_RES) {
      float g8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, x8[8], o8[8];
      if (col_ok) ld8_bf16(p.gate + (long long)b * p.gate_stride + col, g8);
      unpack8(*reinterpret_cast<const uint4*>(slot), x8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = bf16_round(acc[j]);                      // Linear output is bf16 in the reference
        o8[j] = x8[j] + bf16_round(acc[j] * g8[j]);      // x + (out * gate), each op rounded to bf16
      }
      keep[g] = pack8(acc);
      *reinterpret_cast<uint4*>(slot) = pack8(o8);
    } else if constexpr (EPI == VDS_EPI_DGELU) {
      float h8[8];
      unpack8(*reinterpret_cast<const uint4*>(slot), h8);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= dgelu_erf(h8[j]);
      *reinterpret_cast<uint4*>(slot) = pack8(acc);
    }
  }
  __syncwarp();
  if constexpr (EPI == VDS_EPI_STORE || EPI == VDS_EPI_DGELU) {
    stg_store(stg, reinterpret_cast<bf16*>(p.C), p.ldc, rm, col0, p.N, lane, EPI == VDS_EPI_STORE);
  } else {
    stg_store(stg, reinterpret_cast<bf16*>(p.C2), p.ldc2, rm, col0, p.N, lane, false);
    if (p.C != nullptr) {
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(stg + stg_off(lane, g)) = keep[g];
      __syncwarp();
      stg_store(stg, reinterpret_cast<bf16*>(p.C), p.ldc, rm, col0, p.N, lane, false);
    }
  }
  __syncwarp();
}

// One thread handles 32 consecutive columns [col0, col0+32) of output row `row`.
template <int EPI>
__device__ __forceinline__ void epilogue_row(const GemmDev& p, int row, int col0, const uint32_t (&raw)[32]) {
  if (row >= p.M) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col >= p.N) break;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __uint_as_float(raw[g * 8 + j]);

    if constexpr (EPI == VDS_EPI_ACCUM_F32) {
      float* c = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col;
      red_add_v4(c, acc[0], acc[1], acc[2], acc[3]);
      red_add_v4(c + 4, acc[4], acc[5], acc[6], acc[7]);
    } else {
      if constexpr (EPI != VDS_EPI_DGELU) {
        if (p.bias != nullptr) {
          float b[8];
          ld8_bf16(p.bias + col, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += b[j];
        }
      }
      if constexpr (EPI == VDS_EPI_STORE_F32) {
        float* c = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col;
        *reinterpret_cast<float4*>(c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(c + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else if constexpr (EPI == VDS_EPI_STORE) {
        long long orow = row;
        if (p.remap_rows > 0)
          orow = (long long)(row / p.remap_rows) * p.remap_stride + p.remap_offset + row % p.remap_rows;
        st8_bf16(reinterpret_cast<bf16*>(p.C) + orow * p.ldc + col, acc);
      } else if constexpr (EPI == VDS_EPI_BIAS_GELU) {
        float act[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = bf16_round(acc[j]);
          act[j] = gelu_erf(acc[j]);
        }
        if (p.C != nullptr) st8_bf16(reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc + col, acc);
        st8_bf16(reinterpret_cast<bf16*>(p.C2) + (long long)row * p.ldc2 + col, act);
      } else if constexpr (EPI == VDS_EPI_GATE_RES) {
        const int b = row / p.rows_per_batch;
        float g8[8], x8[8], o8[8];
        ld8_bf16(p.gate + (long long)b * p.gate_stride + col, g8);
        ld8_bf16(p.aux + (long long)row * p.ldaux + col, x8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = bf16_round(acc[j]);                      // Linear output is bf16 in the reference
          o8[j] = x8[j] + bf16_round(acc[j] * g8[j]);      // x + (out * gate), each op rounded to bf16
        }
        if (p.C != nullptr) st8_bf16(reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc + col, acc);
        st8_bf16(reinterpret_cast<bf16*>(p.C2) + (long long)row * p.ldc2 + col, o8);
      } else if constexpr (EPI == VDS_EPI_DGELU) {
        float h8[8];
        ld8_bf16(p.aux + (long long)row * p.ldaux + col, h8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= dgelu_erf(h8[j]);
        st8_bf16(reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc + col, acc);
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);

  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

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
#pragma unroll 1
        for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
          uint32_t v[32];
          tmem_ld32(t_base + c * 32, v);
          tmem_ld_wait();
          epilogue_row<EPI>(p, row, nt * BN + c * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      } else {
        uint8_t* stg = smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + (warp - 2) * 4096;
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
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
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

}  // namespace vds

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
  if (!a.a_mn && !a.b_mn) return dispatch_epi<false, false>(a, s);
  if (!a.a_mn && a.b_mn) return dispatch_epi<false, true>(a, s);
  if (a.a_mn && a.b_mn) return dispatch_epi<true, true>(a, s);
  return dispatch_epi<true, false>(a, s);
}
