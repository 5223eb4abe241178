// Flash-style attention forward / backward on tcgen05 + TMEM + TMA (sm_100a), head_dim = 128,
// non-causal, no mask (model.py:136 self-attention, model.py:157 cross-attention; autograd of both).
//
// Layout contract: q/k/v live inside row-major token buffers ([B, L, ld] with the head at column
// head*128), so the "(k h d)" split of the QKV GEMM output (model.py:126) is consumed in place by
// 4-D TMA tensor maps {d, l, head, b}; the output is written token-major [B, L, nh*128] so the
// "b h l d -> b l (h d)" rearrange (model.py:137) disappears.
//
// Forward CTA = two 128-row query tiles of one (b, head), ping-pong (see attn_fwd_kernel).
// Backward, 1-CTA kernel (attn_bwd_kernel, below): one 128-row K/V tile of one (b, head) looping over 64-row query
// sub-tiles; serves cross-attention, query-range splits and the remainder of the CTA-pair kernel.
// Backward, CTA-pair kernel (attention_bwd2.cu): two adjacent K/V tiles per 2-CTA cluster (tcgen05 cta_group::2); serves
// whole waves of a self-attention-sized problem.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "ptx.cuh"
#include "attn_common.cuh"

namespace vds {

struct AttnFwdParams {
  bf16* out; long long ldo;       // [B, Lq, ldo], head at column head*128
  float* lse;                      // [B, nh, Lq], log2 domain: m + log2(l)
  int Lq, Lk, nh;
  float scale_log2;                // head_dim^-0.5 * log2(e)
};

// Forward v1: one CTA = TWO 128-row query tiles of one (b, head) ("ping-pong"): while the softmax
// warpgroup of one tile works, the tensor core runs the other tile's GEMMs.
//   warp 0      TMA producer: Q0,Q1 once; K and V rings (2 stages each, released separately)
//   warp 1      MMA issuer:   S_t = Q_t K^T (SS), O_t += P_t V (A = P from TMEM, B = V from smem)
//   warps 4-7   softmax of tile 0, warps 8-11 softmax of tile 1 (thread == query row == TMEM lane)
// TMEM (512 cols): S0 | S1 | O0 | O1; P_t (bf16, 64 cols) overwrites the head of S_t in place, so P never
// touches shared memory.  The running max is only refreshed when it grew by more than 2^8 (lazy rescale),
// which keeps the O correction off the critical path.
constexpr int FWD_THREADS = 384;
constexpr int FWD_SMEM = 6 * TILE_BYTES + 256 + 1024;

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnFwdParams p,
                int tma_o) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_addr);
  const uint32_t sQ = base, sK = base + 2 * TILE_BYTES, sV = base + 4 * TILE_BYTES;
  const uint32_t bars = base + 6 * TILE_BYTES;
  const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 24, k_empty = bars + 40, v_empty = bars + 56,
                 s_full = bars + 72, p_full = bars + 88, o_done = bars + 104, tmem_slot = bars + 120;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 6 * TILE_BYTES + 120);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, head = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1);
      mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_empty + 8 * s, 1);
      mbar_init(s_full + 8 * s, 1);
      mbar_init(p_full + 8 * s, 128);
      mbar_init(o_done + 8 * s, 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();      // everything above is launch-independent set-up; global inputs may come from the previous kernel
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      load_tile_4d(sQ, &tmQ, q_full, q0, head, b);
      load_tile_4d(sQ + TILE_BYTES, &tmQ, q_full, q0 + 128, head, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(k_empty + 8 * s, ph ^ 1u);
        mbar_expect_tx(k_full + 8 * s, TILE_BYTES);
        load_tile_4d(sK + s * TILE_BYTES, &tmK, k_full + 8 * s, j * 128, head, b);
        mbar_wait(v_empty + 8 * s, ph ^ 1u);
        mbar_expect_tx(v_full + 8 * s, TILE_BYTES);
        load_tile_4d(sV + s * TILE_BYTES, &tmV, v_full + 8 * s, j * 128, head, b);
      }
    }
  } else if (warp == 1) {
    // whole warp runs the (warp-uniform) control flow; one elected lane issues MMAs + commits
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);
    auto issue_s = [&](int t, int s) {   // S_t = Q_t K[s]^T   (called by the elected lane only)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(tmem + t * 128, desc_kmajor(sQ + t * TILE_BYTES, kk), desc_kmajor(sK + s * TILE_BYTES, kk),
                  idesc_qk, kk > 0);
      umma_commit(s_full + 8 * t);
    };
    mbar_wait(q_full, 0);
    mbar_wait(k_full, 0);
    tc_fence_after();
    if (elect_one()) {
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(k_empty);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      mbar_wait(v_full + 8 * s, (j >> 1) & 1);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full + 8 * t, j & 1);
        if (t == 0 && j + 1 < n_kv) mbar_wait(k_full + 8 * (s ^ 1), ((j + 1) >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // O_t += P_t V : A = P (bf16, TMEM, 8 columns per k-step)
            umma_bf16_ts(tmem + 256 + t * 128, tmem + t * 128 + kk * 8, desc_mnmajor(sV + s * TILE_BYTES, kk),
                         idesc_pv, (j > 0 || kk > 0));
          if (t == 1) umma_commit(v_empty + 8 * s);
          if (j + 1 < n_kv) {
            issue_s(t, s ^ 1);
            if (t == 1) umma_commit(k_empty + 8 * (s ^ 1));
          } else {
            umma_commit(o_done + 8 * t);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem + t * 128 + lane_off, tO = tmem + 256 + t * 128 + lane_off;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.Lk - j * 128;             // columns >= valid are padding (last tile only)
      mbar_wait(s_full + 8 * t, j & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      float mxs;
      if (j > 0 && valid >= 128) {
        // Speculative single pass: exponentiate against the running max of the previous tiles while tracking this
        // tile's max.  The lazy-rescale rule (rescale only when the max grows by more than 2^8) is the same as in the
        // two-pass path below, so in the common case — no row of the warp needs a rescale — S is read from TMEM once
        // and the max is off the critical path.  P stays in registers until that is known (it overwrites S).
        uint32_t pk[64];
        const float nm0 = -m_run;
        float2 s2 = make_float2(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        tmem_ld32(tS, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c < 3) tmem_ld32(tS + (c + 1) * 32, (c & 1) ? va : vb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                  make_float2(sl2, sl2), make_float2(nm0, nm0));
            mx0 = fmaxf(mx0, x.x);
            mx1 = fmaxf(mx1, x.y);
            // every other pair on the FMA / ALU pipes (ptx.cuh: ex2_poly2) — the MUFU pipe is the co-bottleneck of this kernel
            const float2 pe = make_float2(ex2(x.x), ex2(x.y));
            s2 = add2(s2, pe);
            pk[c * 16 + (i >> 1)] = pack_bf16x2(pe.x, pe.y);
          }
        }
        const float mxx = fmaxf(mx0, mx1);          // max of (s * scale_log2 - m_run)
        if (!__any_sync(0xffffffffu, mxx > 8.0f)) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pc[i] = pk[c * 16 + i];
            tmem_st16(tS + c * 16, pc);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full + 8 * t);
          l_run += s2.x + s2.y;
          continue;
        }
        mxs = m_run + mxx;                           // some row needs the rescale: redo the tile on the exact path
      } else {
        // TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is reduced
        float mx = -INFINITY;
        tmem_ld32(tS, va);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c < 3) tmem_ld32(tS + (c + 1) * 32, (c & 1) ? va : vb);
          if (valid >= 128) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        mxs = mx * sl2;
      }
      float alpha = 1.0f;
      const bool need = mxs > m_run + 8.0f;         // lazy: p stays <= 2^8 otherwise
      if (need) {
        alpha = ex2(m_run - mxs);                   // 0 on the first tile
        m_run = mxs;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // S_t(j) complete implies P_t V(j-1) complete (in-order tensor pipe): O_t is stable here
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tO + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tO + c * 32, v);
        }
      }
      l_run *= alpha;
      float sum = 0.f;
      const float nm = -m_run;
      tmem_ld32(tS, va);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t (&v)[32] = (c & 1) ? vb : va;
        uint32_t pk[16];
        tmem_ld_wait();
        if (c < 3) tmem_ld32(tS + (c + 1) * 32, (c & 1) ? va : vb);
        if (valid >= 128) {
          float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                  make_float2(sl2, sl2), make_float2(nm, nm));
            const float2 pe = make_float2(ex2(x.x), ex2(x.y));
            s2 = add2(s2, pe);
            pk[i >> 1] = pack_bf16x2(pe.x, pe.y);
          }
          sum += s2.x + s2.y;
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < valid) ? ex2(fmaf(__uint_as_float(v[i]), sl2, nm)) : 0.f;
            const float p1 = (c * 32 + i + 1 < valid) ? ex2(fmaf(__uint_as_float(v[i + 1]), sl2, nm)) : 0.f;
            sum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
        }
        tmem_st16(tS + c * 16, pk);                 // P (bf16 pairs) overwrites consumed S columns
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + 8 * t);
      l_run += sum;
    }
    // epilogue: O / l -> bf16, token-major
    mbar_wait(o_done + 8 * t, 0);
    tc_fence_after();
    const int row = q0 + t * 128 + r;
    const float inv_l = 1.0f / l_run;
    // O leaves through this tile's Q operand tile (every MMA that read it is complete: o_done) in the same two-halves
    // SW128 image and one TMA store per half; per-thread row stores are 32 partial-sector transactions per warp
    // instruction (the whole epilogue of a short cross-attention CTA).  Rows past Lq are clipped by the tensor map.
    uint8_t* otile = gen + t * TILE_BYTES;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + c * 32, v);
      tmem_ld_wait();
      bf16* dst = p.out + ((long long)b * p.Lq + min(row, p.Lq - 1)) * p.ldo + head * HD + c * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
        if (tma_o) st_tile8(otile, r, c * 32 + g * 8, u);
        else if (row < p.Lq) *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
    }
    if (tma_o) {
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (r == 0 && q0 + t * 128 < p.Lq) {
        tma_store_4d(&tmO, sQ + t * TILE_BYTES, 0, q0 + t * 128, head, b);
        tma_store_4d(&tmO, sQ + t * TILE_BYTES + HALF_BYTES, 64, q0 + t * 128, head, b);
        bulk_commit_group();
        bulk_wait_group0();
      }
    }
    if (row < p.Lq && p.lse != nullptr) p.lse[((long long)b * p.nh + head) * p.Lq + row] = m_run + log2f(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ backward
// delta[b, head, row] = sum_d dO * O   (fp32)
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long ldo,
                                     long long lddo, float* __restrict__ delta, int B, int L, int nh) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)B * L * nh;
  if (gw >= total) return;
  const int head = gw % nh;
  const long long tok = gw / nh;           // b*L + l
  const bf16* po = o + tok * ldo + head * HD + lane * 4;
  const bf16* pd = d_o + tok * lddo + head * HD + lane * 4;
  const uint2 a = *reinterpret_cast<const uint2*>(po);
  const uint2 c = *reinterpret_cast<const uint2*>(pd);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), c0 = unpack_bf16x2(c.x), c1 = unpack_bf16x2(c.y);
  float s = a0.x * c0.x + a0.y * c0.y + a1.x * c1.x + a1.y * c1.y;
  s = warp_sum(s);
  if (lane == 0) {
    const int bb = (int)(tok / L), l = (int)(tok % L);
    delta[((long long)bb * nh + head) * L + l] = s;
  }
}

// Backward.  CTA = one 128-row K/V tile of one (b, head), looping over 64-row query sub-tiles:
//   S^T = K Q^T (double-buffered), dP^T = V dO^T (single buffer)   TMEM   M=128 (kv) N=64 (q)  K=128 (d)
//   P^T, dS^T -> bf16 into the retired S^T columns (A operands of dV / dK); dS^T also -> smem (B operand of dQ^T)
//   dV += P^T dO, dK += dS^T Q          (TMEM accumulators)            M=128 (kv) N=128 (d) K=64 (q)
//   dQ^T = K^T dS^T                     (TMEM, own 64 columns)         M=128 (d)  N=64 (q)  K=128 (kv)
//   dQ^T -> fp32 smem tile [q][d] -> cp.reduce.async.bulk.tensor (TMA adds it into the fp32 dq buffer)
// warp 0 TMA producer (K,V once; Q/dO ring of 3) | warp 1 MMA issuer | warps 4-7 compute | warps 8-11 dQ drain.
// Measured (scripts/mma_shapes.cu, scripts/bwd_trace.py): tcgen05.mma issue blocks on a shallow queue, an SS-mode
// M128 N64 K16 MMA costs 53 cycles (6 KiB of operand reads at 128 B/clk) against 37 with A in TMEM, and the loop moves
// ~288 KiB of shared-memory traffic per sub-tile (operands 176, Q/dO fill 32, dS^T 16, dQ staging 2 x 32): the kernel
// runs at ~90% of the shared-memory port, which is what bounds it.
constexpr int BWD_THREADS = 384;
constexpr int QT_BYTES = QSUB * HD * 2;          // 16 KiB: 64 x 128 bf16 (two [64 x 128 B] halves)
constexpr int QT_HALF = QT_BYTES / 2;            // 8 KiB
constexpr int PT_BYTES = 128 * QSUB * 2;         // 16 KiB: [128 kv rows x 128 B]
constexpr int STG_BYTES = QSUB * HD * 4;         // 32 KiB fp32 staging of one dQ sub-tile
constexpr int BWD_OFF_Q = 2 * TILE_BYTES;                    // 3 stages x (Q, dO)
constexpr int BWD_OFF_PT = BWD_OFF_Q + 3 * 2 * QT_BYTES;
constexpr int BWD_OFF_DST = BWD_OFF_PT + PT_BYTES;
constexpr int BWD_OFF_STG = BWD_OFF_DST + PT_BYTES;
constexpr int BWD_OFF_STAT = BWD_OFF_STG + STG_BYTES;       // [2][2][64] floats
constexpr int BWD_OFF_BAR = BWD_OFF_STAT + 1024;
constexpr int BWD_SMEM = BWD_OFF_BAR + 256 + 1024;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmKo,
                const __grid_constant__ CUtensorMap tmVo, const AttnBwdParams p, int out_mode) {
  // out_mode: how dK / dV leave — 0 per-thread stores / red.global; 1 bf16 TMA store (tmKo / tmVo over dk / dv);
  // 2 fp32 TMA reduce-add into dk_acc / dv_acc (tmKo / tmVo); 3 fp32 TMA reduce-add into the compact workspace (tmKo)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_addr);
  const uint32_t sK = base, sV = base + TILE_BYTES, sQ = base + BWD_OFF_Q, sPT = base + BWD_OFF_PT,
                 sDST = base + BWD_OFF_DST, sSTG = base + BWD_OFF_STG;
  uint8_t* gPT = gen + BWD_OFF_PT;
  uint8_t* gDST = gen + BWD_OFF_DST;
  float* gSTG = reinterpret_cast<float*>(gen + BWD_OFF_STG);
  float* s_stat = reinterpret_cast<float*>(gen + BWD_OFF_STAT);   // [buf][lse|delta][64]
  const uint32_t bars = base + BWD_OFF_BAR;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 32, s_full = bars + 56,
                 pds_full = bars + 72, mma_done = bars + 80, dq_drained = bars + 96, tmem_slot = bars + 112,
                 stat_full = bars + 120, dp_full = bars + 136, dp_read = bars + 144, dq_full = bars + 152;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + BWD_OFF_BAR + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid over (local item, split); local item -> (b, head, kv tile) through attn_bwd_decode_item
  const int item_local = blockIdx.x / p.q_splits, split = blockIdx.x % p.q_splits;
  int kv_tile, head, b;
  attn_bwd_decode_item(p, item_local, b, head, kv_tile);
  const int kv0 = kv_tile * 128;
  const int n_q_all = (p.Lq + QSUB - 1) / QSUB;
  const int per = (n_q_all + p.q_splits - 1) / p.q_splits;
  const int qt0 = split * per, qt1 = min(n_q_all, qt0 + per);
  const int n_q = qt1 - qt0;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + 8 * s, 1);
      mbar_init(mma_done + 8 * s, 1);
      mbar_init(dq_drained + 8 * s, 128);
      mbar_init(stat_full + 8 * s, 128);
    }
    mbar_init(pds_full, 128);
    mbar_init(dp_full, 1);
    mbar_init(dp_read, 128);
    mbar_init(dq_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();      // everything above is launch-independent set-up; global inputs may come from the previous kernel
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  // TMEM columns: dV [0,128) | dK [128,256) | S^T buffer 0 / 1 [256,320) / [320,384) (afterwards bf16 P^T in its
  // columns 0..31 and dS^T in 32..63: the A operands of dV / dK) | dP^T [384,448) | dQ^T [448,512)
  const uint32_t tDV = tmem, tDK = tmem + 128, tSTb = tmem + 256, tDPTs = tmem + 384, tDQT = tmem + 448;

  if (n_q > 0) {
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(kv_full, 2 * TILE_BYTES);
        load_tile_4d(sK, &tmK, kv_full, kv0, head, b);
        load_tile_4d(sV, &tmV, kv_full, kv0, head, b);
        for (int i = 0; i < n_q; ++i) {
          const int st = i % 3;
          const uint32_t us = (i / 3) & 1;
          mbar_wait(qdo_empty + 8 * st, us ^ 1u);
          mbar_expect_tx(qdo_full + 8 * st, 2 * QT_BYTES);
          const uint32_t dq = sQ + st * 2 * QT_BYTES, dd = dq + QT_BYTES;
          const int row0 = (qt0 + i) * QSUB;
          tma_load_4d(dq, &tmQ, qdo_full + 8 * st, 0, row0, head, b);
          tma_load_4d(dq + QT_HALF, &tmQ, qdo_full + 8 * st, 64, row0, head, b);
          tma_load_4d(dd, &tmDO, qdo_full + 8 * st, 0, row0, head, b);
          tma_load_4d(dd + QT_HALF, &tmDO, qdo_full + 8 * st, 64, row0, head, b);
        }
      }
    } else if (warp == 1) {
      // The whole warp runs the control flow (warp-uniform => descriptors stay in uniform registers and the
      // UTCHMMA issue needs no per-lane waterfall); one elected lane issues the MMAs and their commits.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);     // S^T, dP^T
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, false, true);   // dV, dK
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, true, true);      // dQ^T
      // tcgen05.mma issue blocks while the (shallow) tensor queue is full, so this warp's program order IS the tensor
      // pipe's schedule.  Per sub-tile i:  S^T(i+1) | dP^T(i+1) | dQ^T(i) | dV(i), dK(i).  The first two only wait for
      // buffers the compute warps released long ago, so the pipe keeps running while sub-tile i is in the softmax math.
      auto issue_s = [&](int k) {
        const int bb = k & 1, st = k % 3;
        mbar_wait(qdo_full + 8 * st, (k / 3) & 1);
        mbar_wait(stat_full + 8 * bb, (k >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          VDS_TRACE(0, k);   // S(k) issue
          const uint32_t q = sQ + st * 2 * QT_BYTES;
          const uint32_t tST = tSTb + bb * 64;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tST, desc_kmajor(sK, kk), umma_smem_desc(q + (kk >> 2) * QT_HALF + (kk & 3) * 32, 16, 1024),
                      idesc_s, kk > 0);
          umma_bf16(tST, desc_k16_noswz(sPT), desc_k16_noswz(sPT + 4096 + (bb * 2 + 0) * 2048), idesc_s, 1);  // - lse
          umma_commit(s_full + 8 * bb);
          VDS_TRACE(7, k);   // S(k) issued
        }
        __syncwarp();
      };
      // dP^T(k) = V dO^T into the single dP^T buffer (free once dp_read says the compute warps hold dP^T(k-1) in registers)
      auto issue_dp = [&](int k) {
        const int bb = k & 1, st = k % 3;
        if (k > 0) mbar_wait(dp_read, (k - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_o = sQ + st * 2 * QT_BYTES + QT_BYTES;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tDPTs, desc_kmajor(sV, kk), umma_smem_desc(d_o + (kk >> 2) * QT_HALF + (kk & 3) * 32, 16, 1024),
                      idesc_s, kk > 0);
          umma_bf16(tDPTs, desc_k16_noswz(sPT), desc_k16_noswz(sPT + 4096 + (bb * 2 + 1) * 2048), idesc_s, 1);  // - delta
          umma_commit(dp_full);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n_q; ++i) {
        const int bb = i & 1, st = i % 3;
        if (i + 1 < n_q) {
          issue_s(i + 1);
          issue_dp(i + 1);
        }
        mbar_wait(pds_full, i & 1);
        if (i > 0) mbar_wait(dq_drained, (i - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          VDS_TRACE(1, i);   // dQ/dV/dK(i) issue
          const uint32_t q = sQ + st * 2 * QT_BYTES, d_o = q + QT_BYTES;
          const uint32_t tPT = tSTb + bb * 64;   // bf16 P^T (columns 0..31) and dS^T (32..63) in the retired S^T columns
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // dQ^T = K^T dS^T : A = K MN-major (M = d), B = dS^T MN-major (N = q)
            umma_bf16(tDQT, desc_mnmajor(sK, kk), umma_smem_desc(sDST + kk * 2048, 16, 1024), idesc_dq, kk > 0);
          umma_commit(dq_full);            // first: the drain warpgroup starts early and the dS^T smem tile is free again
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dV += P^T dO : A = P^T (TMEM, 8 columns per k-step), B = dO MN-major (N = d)
            umma_bf16_ts(tDV, tPT + kk * 8, umma_smem_desc(d_o + kk * 2048, QT_HALF, 1024), idesc_acc,
                         (i > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dK += dS^T Q : A = dS^T (TMEM)
            umma_bf16_ts(tDK, tPT + 32 + kk * 8, umma_smem_desc(q + kk * 2048, QT_HALF, 1024), idesc_acc,
                         (i > 0 || kk > 0));
          umma_commit(qdo_empty + 8 * st);
          umma_commit(mma_done + 8 * bb);
        }
        __syncwarp();
      }
    } else if (warp >= 4 && warp < 8) {
      // ------------------------------------------------------------ compute warpgroup (thread == kv row)
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const int ct = threadIdx.x - 128;    // 0..127
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const bool kv_ok = (kv0 + r) < p.Lk;
      const long long stat_base = ((long long)b * p.nh + head) * p.Lq;
      const float* stat_src = (ct < 64 ? p.lse : p.delta) + stat_base;
      const int sc = ct & 63;
      const float inv_sl2 = 1.0f / p.scale_log2;
      // value this thread contributes to the statistics tile of sub-tile k: -lse/scale_log2 (ct < 64) or -delta
      // The global load is issued an iteration ahead as a bare ld (no dependent instruction until stat_finish at the
      // end of the iteration), so its latency never stalls this warp.
      auto stat_fetch = [&](int k) -> float {
        const int q = min((qt0 + k) * QSUB + sc, p.Lq - 1);
        float raw;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(raw) : "l"(stat_src + q));
        return raw;
      };
      auto stat_finish = [&](int k, float raw) -> float {
        const int q = (qt0 + k) * QSUB + sc;
        if (q >= p.Lq) return ct < 64 ? -INFINITY : 0.f;        // padded query row: exp2(-inf) = 0
        return ct < 64 ? -raw * inv_sl2 : -raw;
      };
      auto stat_value = [&](int k) -> float { return stat_finish(k, stat_fetch(k)); };
      auto stat_write = [&](int k, float v) {
        uint8_t* tile = gPT + 4096 + ((k & 1) * 2 + (ct < 64 ? 0 : 1)) * 2048;
        *reinterpret_cast<uint4*>(tile + k16_off(sc)) = split3_bf16(v);
        *reinterpret_cast<uint4*>(tile + k16_off(sc) + 128) = make_uint4(0u, 0u, 0u, 0u);
      };
      auto stat_store = [&](int k, float v) {
        stat_write(k, v);
        fence_proxy_async_smem();
        mbar_arrive(stat_full + 8 * (k & 1));
      };
      {  // constant A operand of the statistics k-step: ones in columns 0..2 of every kv row
        const float one = 1.0f;
        *reinterpret_cast<uint4*>(gPT + k16_off(ct)) = make_uint4(pack_bf16x2(one, one), pack_bf16x2(one, 0.f), 0u, 0u);
        *reinterpret_cast<uint4*>(gPT + k16_off(ct) + 128) = make_uint4(0u, 0u, 0u, 0u);
      }
      stat_store(0, stat_value(0));
      if (n_q > 1) stat_store(1, stat_value(1));
      const bool kv_full_tile = kv0 + 128 <= p.Lk;
      for (int i = 0; i < n_q; ++i) {
        const int bb = i & 1;
        const float next_raw = (i + 2 < n_q) ? stat_fetch(i + 2) : 0.f;   // prefetched; stored at the end of this iteration
        mbar_wait(s_full + 8 * bb, (i >> 1) & 1);
        tc_fence_after();
        if (ct == 0) VDS_TRACE(2, i);   // compute sees S(i)
        const uint32_t tST = tSTb + bb * 64 + lane_off, tDPT = tDPTs + lane_off;
        // phase 1 (overlaps the dV/dK/dQ MMAs of the previous sub-tile): p = exp2(s'), P^T -> TMEM
        float pf[64];
        {
          uint32_t sv0[32], sv1[32];
          tmem_ld32(tST, sv0);
          tmem_ld32(tST + 32, sv1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float p0 = ex2(__uint_as_float(sv0[e]) * p.scale_log2);
            float p1 = ex2(__uint_as_float(sv1[e]) * p.scale_log2);
            if (!kv_full_tile) {
              p0 = kv_ok ? p0 : 0.f;
              p1 = kv_ok ? p1 : 0.f;
            }
            pf[e] = p0;
            pf[32 + e] = p1;
          }
        }
        {
          uint32_t pp[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) pp[e] = pack_bf16x2(pf[2 * e], pf[2 * e + 1]);
          tmem_st32(tST, pp);
        }
        // phase 2: dS^T = P^T o dP'^T * scale once dP^T(i) has landed
        mbar_wait(dp_full, i & 1);
        tc_fence_after();
        uint32_t dd[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t dv[32];
          tmem_ld32(tDPT + c * 32, dv);
          tmem_ld_wait();
          if (c == 1) {   // dP^T(i) is in registers: the MMA warp may refill the buffer with dP^T(i+1)
            tc_fence_before();
            mbar_arrive(dp_read);
          }
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float d0 = pf[c * 32 + e] * (__uint_as_float(dv[e]) * p.scale);
            const float d1 = pf[c * 32 + e + 1] * (__uint_as_float(dv[e + 1]) * p.scale);
            dd[c * 16 + (e >> 1)] = pack_bf16x2(d0, d1);
          }
        }
        tmem_st32(tST + 32, dd);
        if (ct == 0) VDS_TRACE(3, i);   // math done
        if (i > 0) mbar_wait(dq_full, (i - 1) & 1);   // dS^T smem tile consumed by dQ^T(i-1)
        uint8_t* gDSTb = gDST;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(gDSTb + sw128_offset(r, g)) =
              make_uint4(dd[g * 4], dd[g * 4 + 1], dd[g * 4 + 2], dd[g * 4 + 3]);
        // statistics tile of sub-tile i+2 (buffer bb: S^T(i) and dP^T(i), its readers, are complete) shares the fence
        if (i + 2 < n_q) stat_write(i + 2, stat_finish(i + 2, next_raw));
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pds_full);
        if (i + 2 < n_q) mbar_arrive(stat_full + 8 * bb);
        if (ct == 0) VDS_TRACE(4, i);   // pds arrive   // buffer bb: its MMAs (S^T / dP^T of sub-tile i) are complete
      }
    } else if (warp >= 8) {
      // ------------------------------------------------------------ dQ drain warpgroup (thread == d)
      const int quad = warp & 3;
      const int d = quad * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const bool leader = threadIdx.x == 256;
      for (int i = 0; i < n_q; ++i) {
        const int bb = i & 1;
        mbar_wait(dq_full, i & 1);
        tc_fence_after();
        if (leader) VDS_TRACE(5, i);   // drain sees dQ(i)
        uint32_t v0[32], v1[32];
        tmem_ld32(tDQT + lane_off, v0);
        tmem_ld32(tDQT + lane_off + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_drained);
        if (leader) VDS_TRACE(6, i);   // dq_drained arrive
        if (leader) bulk_wait_group_read0();      // previous reduction has finished reading the staging tile
        named_bar_sync(2, 128);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          gSTG[q * HD + d] = __uint_as_float(v0[q]);
          gSTG[(q + 32) * HD + d] = __uint_as_float(v1[q]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          tma_reduce_add_4d(&tmDQ, sSTG, 0, (qt0 + i) * QSUB, head, b);
          bulk_commit_group();
        }
      }
      if (leader) bulk_wait_group0();
    }
    if (warp >= 4) {
      // dK (compute warpgroup) / dV (drain warpgroup): TMEM lane == kv row
      const int which = warp >= 8 ? 1 : 0;
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      mbar_wait(mma_done + 8 * ((n_q - 1) & 1), ((n_q - 1) >> 1) & 1);
      tc_fence_after();
      const int krow = kv0 + r;
      const uint32_t t = (which == 0 ? tDK : tDV) + lane_off;
      // Every MMA of this CTA is complete (mma_done of the last sub-tile), so the operand tiles are free: with out_mode != 0
      // the tile is staged there (dK over K | V, dV over the Q / dO ring) in TMA box images and leaves by one bulk tensor
      // store / reduce-add per box — row-per-thread global accesses cost 32 line look-ups per warp instruction.
      uint8_t* stage = gen + (which == 0 ? 0 : BWD_OFF_Q);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t + c * 32, v);
        tmem_ld_wait();
        if (out_mode == 1) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            st_tile8(stage, r, c * 32 + g * 8, u);
          }
        } else if (out_mode >= 2) {   // fp32: four [128 rows x 32 floats] SW128 boxes; rows past Lk hold exact zeros
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(stage + c * 16384 + sw128_offset(r, g)) =
                make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                            __uint_as_float(v[g * 4 + 3]));
        } else if (krow < p.Lk) {
          if (p.q_splits == 1) {
            bf16* dst = (which == 0 ? p.dk + ((long long)b * p.Lk + krow) * p.lddk
                                    : p.dv + ((long long)b * p.Lk + krow) * p.lddv) + head * HD + c * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
              u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
              u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
              u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
              *reinterpret_cast<uint4*>(dst + g * 8) = u;
            }
          } else {
            float* dst = p.compact_acc != nullptr
                             ? p.compact_acc + ((long long)item_local * 2 + which) * (128 * HD) + r * HD + c * 32
                             : (which == 0 ? p.dk_acc : p.dv_acc) + ((long long)b * p.Lk + krow) * p.ldkv_acc +
                                   head * HD + c * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4),
                           "f"(__uint_as_float(v[g * 4])), "f"(__uint_as_float(v[g * 4 + 1])),
                           "f"(__uint_as_float(v[g * 4 + 2])), "f"(__uint_as_float(v[g * 4 + 3]))
                           : "memory");
          }
        }
      }
      if (out_mode != 0) {
        fence_proxy_async_smem();
        named_bar_sync(3 + which, 128);
        if (quad == 0 && lane == 0 && kv0 < p.Lk) {
          const uint32_t st = base + (which == 0 ? 0 : BWD_OFF_Q);
          const CUtensorMap* tm = (which == 0 || out_mode == 3) ? &tmKo : &tmVo;
          if (out_mode == 1) {
            tma_store_4d(tm, st, 0, kv0, head, b);
            tma_store_4d(tm, st + HALF_BYTES, 64, kv0, head, b);
          } else if (out_mode == 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_reduce_add_4d(tm, st + c * 16384, c * 32, kv0, head, b);
          } else {
            const int row0 = (item_local * 2 + which) * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_reduce_add_2d(tm, st + c * 16384, c * 32, row0);
          }
          bulk_commit_group();
          bulk_wait_group0();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

long long* g_attn_bwd_trace = nullptr;   // set through vds_debug_attn_bwd_trace (tuning only)
int g_attn_pair_mode = -1;               // vds_debug_attn_pair_mode: -1 = VDS_ATTN_PAIR env (default auto), 0 off, 1 auto, 2 force

int launch_attn_bwd_pairs(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                          int64_t lddo, const AttnBwdParams& p0, int B, int nh, int Lq, int Lk, int pair_base, int n_pairs,
                          int pairs_per_bh, const uint32_t* pieces, int n_pieces, float* compact,
                          cudaStream_t stream);   // attention_bwd2.cu

// Tail plan of the CTA-pair backward.  `P` pairs (fewer than one wave of `C` clusters) of `nq` query sub-tiles each are cut
// into pieces along the query range so that `C` clusters finish together: the P * nq sub-tiles are laid out as one line,
// cut into S equal ranges (cuts within `kMinPiece` of a pair boundary snap to it) and at every pair boundary; a range
// crosses at most one boundary, so a pair ends up in <= 3 pieces.  S is the value whose longest-piece-first schedule on C
// clusters (the order of the returned pieces == the order the hardware hands clusters to free SM pairs) finishes first,
// with `kPieceOverhead` sub-tile units per piece for the prologue, the fp32 red of dK / dV and the exit (measured: 19.4 k
// cycles against 2240 per sub-tile, scripts/bwd2_prof.py).  Piece = pair | first sub-tile << 10 | sub-tile count << 21.
// Returns the number of pieces, 0 if splitting does not pay.
static int attn_bwd_plan_tail(int P, int nq, int C, uint32_t* out, int cap) {
  constexpr int kMinPiece = 8;
  constexpr double kPieceOverhead = 10.0;
  if (P <= 0 || C <= 0 || nq < 2 * kMinPiece || P >= 1024 || nq >= 2048) return 0;
  const long long T = (long long)P * nq;
  struct Piece { int pair, q0, n; };
  std::vector<Piece> best, cur;
  std::vector<long long> cuts;
  std::vector<double> slot;
  double best_cost = (double)((P + C - 1) / C) * (nq + kPieceOverhead);   // unsplit
  const int s_max = (int)std::min<long long>(cap - P, T / kMinPiece);
  for (int S = P; S <= s_max; ++S) {
    cuts.clear();
    for (int pb = 0; pb <= P; ++pb) cuts.push_back((long long)pb * nq);
    for (int k = 1; k < S; ++k) {
      long long c = (2 * k * T + S) / (2LL * S);                   // round(k * T / S)
      const long long pb = ((2 * c + nq) / (2LL * nq)) * nq;         // nearest pair boundary
      if (llabs(c - pb) < kMinPiece) c = pb;
      cuts.push_back(c);
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    if ((int)cuts.size() - 1 > cap) continue;
    cur.clear();
    for (size_t i = 0; i + 1 < cuts.size(); ++i)
      cur.push_back({(int)(cuts[i] / nq), (int)(cuts[i] % nq), (int)(cuts[i + 1] - cuts[i])});
    std::stable_sort(cur.begin(), cur.end(), [](const Piece& a, const Piece& b) { return a.n > b.n; });
    slot.assign(C, 0.0);
    for (const Piece& pc : cur) {   // the next cluster goes to the SM pair that frees first
      auto it = std::min_element(slot.begin(), slot.end());
      *it += pc.n + kPieceOverhead;
    }
    const double cost = *std::max_element(slot.begin(), slot.end());
    if (cost < best_cost - 1e-9) { best_cost = cost; best = cur; }
  }
  for (size_t i = 0; i < best.size(); ++i)
    out[i] = (uint32_t)best[i].pair | ((uint32_t)best[i].q0 << 10) | ((uint32_t)best[i].n << 21);
  return (int)best.size();
}

// tail balancing fix-up: compact fp32 [item][dk|dv][128][128] -> bf16 dk / dv tiles; FIXUP_SPLIT CTAs per item (32 rows
// each).  The workspace is handed back ZEROED (it is zero on entry of every vds_attn_bwd call: allocated zeroed by the
// caller, restored here), which saves a memset node in front of every split launch.
constexpr int FIXUP_SPLIT = 4;
__global__ void __launch_bounds__(256) attn_bwd_tail_fixup_kernel(float* __restrict__ compact, bf16* __restrict__ dk,
                                                                  long long lddk, bf16* __restrict__ dv, long long lddv,
                                                                  const AttnBwdParams p, int Lk) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  int kv_tile, head, b;
  const int item = blockIdx.x / FIXUP_SPLIT, part = blockIdx.x % FIXUP_SPLIT;
  attn_bwd_decode_item(p, item, b, head, kv_tile);
  float* src = compact + (long long)item * 2 * 128 * HD;
  constexpr int ROWS = 128 / FIXUP_SPLIT;
  constexpr int ITERS = 2 * ROWS * (HD / 8) / 256;    // 8-element groups per thread: all loads issued before the first store
  float4 a[ITERS], c[ITERS];
  float4* s4[ITERS];
  bf16* dst[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int idx = it * 256 + threadIdx.x;
    const int which = idx / (ROWS * (HD / 8));
    const int rem = idx % (ROWS * (HD / 8));
    const int r = part * ROWS + rem / (HD / 8), c8 = rem % (HD / 8);
    const int krow = kv_tile * 128 + r;
    s4[it] = nullptr;
    if (krow >= Lk) continue;     // never written by the kernels either: stays zero
    s4[it] = reinterpret_cast<float4*>(src + (which * 128 + r) * HD + c8 * 8);
    dst[it] = (which == 0 ? dk + ((long long)b * Lk + krow) * lddk : dv + ((long long)b * Lk + krow) * lddv) + head * HD + c8 * 8;
    a[it] = __ldcg(s4[it]);       // written by red.global (L2): read there
    c[it] = __ldcg(s4[it] + 1);
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    if (s4[it] == nullptr) continue;
    s4[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
    s4[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 u;
    u.x = pack_bf16x2(a[it].x, a[it].y); u.y = pack_bf16x2(a[it].z, a[it].w);
    u.z = pack_bf16x2(c[it].x, c[it].y); u.w = pack_bf16x2(c[it].z, c[it].w);
    *reinterpret_cast<uint4*>(dst[it]) = u;
  }
}

}  // namespace vds

using namespace vds;

extern "C" {

/* tuning aid: when non-NULL, CTA (0,0,0) of every following attn_bwd launch writes clock64 stamps [iter][8] here */
int vds_debug_attn_bwd_trace(void* buf) {
  vds::g_attn_bwd_trace = (long long*)buf;
  return VDS_OK;
}

/* tests / tuning: selects the backward kernel for self-attention-sized problems. -1: VDS_ATTN_PAIR env (default: auto),
 * 0: 1-CTA kernel only, 1: auto (CTA-pair kernel for whole waves of kv-tile pairs), 2: CTA-pair kernel for every pair */
int vds_debug_attn_pair_mode(int mode) {
  vds::g_attn_pair_mode = mode;
  return VDS_OK;
}

/* q: [B, Lq, ldq] (head h at column h*128), k/v: [B, Lk, ldk|ldv]; out: [B, Lq, ldo]; lse: [B, nh, Lq] fp32 */
int vds_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                 int64_t ldo, float* lse, int B, int nh, int Lq, int Lk, int head_dim, float scale, void* stream) {
  VDS_CHECK_ARG(head_dim == HD, "attn: head_dim=%d unsupported (only 128)", head_dim);
  VDS_CHECK_ARG(B > 0 && nh > 0 && Lq > 0 && Lk > 0, "attn: bad shape");
  VDS_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attn: leading dims must be multiples of 8");
  CUtensorMap tq, tk, tv;
  int r;
  if ((r = make_tmap_tokens(&tq, q, ldq, Lq, nh, B))) return r;
  if ((r = make_tmap_tokens(&tk, k, ldk, Lk, nh, B))) return r;
  if ((r = make_tmap_tokens(&tv, v, ldv, Lk, nh, B))) return r;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    if (e != cudaSuccess) { set_error("attn_fwd: smem attribute: %s", cudaGetErrorString(e)); return VDS_ERR_CUDA; }
    attr = true;
  }
  AttnFwdParams p;
  p.out = (bf16*)out; p.ldo = ldo; p.lse = lse; p.Lq = Lq; p.Lk = Lk; p.nh = nh;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + 255) / 256, nh, B);
  CUtensorMap to = tq;
  int tma_o = 0;
  if (((uintptr_t)out & 15) == 0 && ldo % 8 == 0) {   // else: per-thread row stores
    if ((r = make_tmap_tokens(&to, out, ldo, Lq, nh, B))) return r;
    tma_o = 1;
  }
  launch_k(attn_fwd_kernel, grid, FWD_THREADS, FWD_SMEM, (cudaStream_t)stream, tq, tk, tv, to, p, tma_o);
  VDS_CHECK_LAUNCH("attn_fwd");
  return VDS_OK;
}

/* host-only: the tail plan of the CTA-pair backward for `n_pairs` pairs of `n_qsub` 64-row query sub-tiles on `clusters` SM
 * pairs; writes up to `cap` pieces (pair | first sub-tile << 10 | count << 21, longest first) and returns their number
 * (0: do not split).  Exported for the CPU tests of the schedule. */
int vds_attn_bwd_tail_plan(int n_pairs, int n_qsub, int clusters, uint32_t* pieces, int cap) {
  if (pieces == nullptr || cap <= 0) return 0;
  uint32_t tmp[VDS_BWD2_MAX_PIECES];
  const int n = attn_bwd_plan_tail(n_pairs, n_qsub, clusters, tmp, VDS_BWD2_MAX_PIECES);
  if (n > cap) return 0;
  for (int i = 0; i < n; ++i) pieces[i] = tmp[i];
  return n;
}

int64_t vds_attn_bwd_tail_ws_bytes(int B, int nh, int Lk) {
  (void)B; (void)nh; (void)Lk;
  return (int64_t)(2 * num_sms()) * 2 * 128 * HD * 4;   // remainder items: < SMs tiles of the last wave of pairs + one unpaired tile per (b, head), capped
}

/* delta = rowsum(dO * O) (computed here from o, or passed in precomputed when o == NULL); dq_acc must be zero-initialised fp32 [B, Lq, lddq]; with q_splits > 1 dk/dv are
 * accumulated into zero-initialised fp32 buffers dk_acc/dv_acc [B, Lk, ldkv_acc] instead of dk/dv. */
int vds_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                 int64_t ldo, const void* d_o, int64_t lddo, const float* lse, float* delta, float* dq_acc,
                 int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* dk_acc, float* dv_acc,
                 int64_t ldkv_acc, int q_splits, int B, int nh, int Lq, int Lk, int head_dim, float scale,
                 void* tail_ws, int64_t tail_ws_bytes, void* stream) {
  VDS_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim=%d unsupported (only 128)", head_dim);
  VDS_CHECK_ARG(q_splits >= 1, "attn_bwd: q_splits");
  VDS_CHECK_ARG(q_splits == 1 ? (dk && dv) : (dk_acc && dv_acc), "attn_bwd: missing dk/dv target");
  VDS_CHECK_ARG(lddq % 4 == 0 && ((uintptr_t)dq_acc & 15) == 0, "attn_bwd: dq_acc must be 16-byte aligned, lddq %% 4 == 0");
  CUtensorMap tq, tk, tv, tdo, tdq;
  int r;
  if ((r = make_tmap_tokens(&tq, q, ldq, Lq, nh, B, QSUB))) return r;
  if ((r = make_tmap_tokens(&tk, k, ldk, Lk, nh, B))) return r;
  if ((r = make_tmap_tokens(&tv, v, ldv, Lk, nh, B))) return r;
  if ((r = make_tmap_tokens(&tdo, d_o, lddo, Lq, nh, B, QSUB))) return r;
  if ((r = make_tmap_dq(&tdq, dq_acc, lddq, Lq, nh, B))) return r;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    if (e != cudaSuccess) { set_error("attn_bwd: smem attribute: %s", cudaGetErrorString(e)); return VDS_ERR_CUDA; }
    attr = true;
  }
  if (o != nullptr) {   // o == NULL: delta was produced by the dgrad GEMM's STORE_ROWDOT epilogue
    const long long warps = (long long)B * Lq * nh;
    launch_k(attn_bwd_prep_kernel, (unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream, 
        (const bf16*)o, (const bf16*)d_o, ldo, lddo, delta, B, Lq, nh);
    VDS_CHECK_LAUNCH("attn_bwd_prep");
  }
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc; p.lddq = lddq;
  p.dk = (bf16*)dk; p.lddk = lddk; p.dv = (bf16*)dv; p.lddv = lddv;
  p.dk_acc = dk_acc; p.dv_acc = dv_acc; p.ldkv_acc = ldkv_acc;
  p.Lq = Lq; p.Lk = Lk; p.nh = nh; p.q_splits = q_splits;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.dbg = g_attn_bwd_trace;
  p.compact_acc = nullptr;
  p.rem_pair_base = -1; p.rem_pairs_per_bh = 1; p.rem_pair_tiles = 0;
  const int kv_tiles = (Lk + 127) / 128;
  p.kv_tiles = kv_tiles;
  p.item_base = 0;
  const int total = kv_tiles * nh * B;
  const int sms = num_sms();
  const int n_qsub = (Lq + QSUB - 1) / QSUB;
  cudaStream_t st = (cudaStream_t)stream;

  // `rem` items (local indices 0..rem-1 under the indexing already set in p) on the 1-CTA kernel.  With a workspace and
  // a long query range they are split `s` ways along the query range (fp32 red into a compact workspace + a bf16
  // fix-up), so a partly filled last wave costs ceil(rem*s/SMs)/s waves instead of 1.
  // one launch of the 1-CTA kernel; picks how dK / dV leave (see attn_bwd_kernel: out_mode)
  static const bool tma_env = getenv("VDS_BWD_TMA_OUT") == nullptr || strcmp(getenv("VDS_BWD_TMA_OUT"), "0") != 0;   // tuning switch
  auto launch_1cta = [&](const AttnBwdParams& pk, int items) -> int {
    CUtensorMap tko = tk, tvo = tv;
    int out_mode = 0, rr;
    auto al16 = [](const void* ptr) { return ptr != nullptr && ((uintptr_t)ptr & 15) == 0; };
    if (tma_env && pk.q_splits == 1) {
      if (al16(pk.dk) && al16(pk.dv) && pk.lddk % 8 == 0 && pk.lddv % 8 == 0) {
        if ((rr = make_tmap_tokens(&tko, pk.dk, pk.lddk, Lk, nh, B))) return rr;
        if ((rr = make_tmap_tokens(&tvo, pk.dv, pk.lddv, Lk, nh, B))) return rr;
        out_mode = 1;
      }
    } else if (tma_env && pk.compact_acc != nullptr) {
      if (al16(pk.compact_acc)) {
        uint64_t cdims[2] = {(uint64_t)HD, (uint64_t)items * 2 * 128}, cstr[1] = {(uint64_t)HD * 4};
        uint32_t cbox[2] = {32, 128};
        if ((rr = encode_tmap(&tko, pk.compact_acc, 1, 2, cdims, cstr, cbox, 1))) return rr;
        out_mode = 3;
      }
    } else if (tma_env) {
      if (al16(pk.dk_acc) && al16(pk.dv_acc) && pk.ldkv_acc % 4 == 0) {
        uint64_t dims[4] = {(uint64_t)HD, (uint64_t)Lk, (uint64_t)nh, (uint64_t)B};
        uint64_t strides[3] = {(uint64_t)pk.ldkv_acc * 4, (uint64_t)HD * 4, (uint64_t)Lk * (uint64_t)pk.ldkv_acc * 4};
        uint32_t box[4] = {32, 128, 1, 1};
        if ((rr = encode_tmap(&tko, pk.dk_acc, 1, 4, dims, strides, box, 1))) return rr;
        if ((rr = encode_tmap(&tvo, pk.dv_acc, 1, 4, dims, strides, box, 1))) return rr;
        out_mode = 2;
      }
    }
    launch_k(attn_bwd_kernel, items * pk.q_splits, BWD_THREADS, BWD_SMEM, st, tq, tk, tv, tdo, tdq, tko, tvo, pk, out_mode);
    VDS_CHECK_LAUNCH("attn_bwd");
    return VDS_OK;
  };
  auto launch_items = [&](int rem, bool allow_split) -> int {
    int tail_s = 0;
    if (allow_split && tail_ws != nullptr && rem > 0 && rem % sms != 0 && n_qsub >= 32 &&
        (long long)rem * 2 * 128 * HD * 4 <= tail_ws_bytes) {
      double best = (double)((rem + sms - 1) / sms);
      for (int s = 2; s <= 8; ++s) {
        const double cost = (double)((rem * s + sms - 1) / sms) / s + 0.02 * s;   // + per-split prologue / atomics
        if (cost < best - 0.1) { best = cost; tail_s = s; }
      }
    }
    if (tail_s == 0) {
      if (p.rem_pair_base >= 0) p.dbg = nullptr;
      return launch_1cta(p, rem);
    }
    AttnBwdParams ps = p;
    ps.q_splits = tail_s; ps.compact_acc = (float*)tail_ws;
    if (p.rem_pair_base >= 0) ps.dbg = nullptr;   // the trace buffer belongs to the pair kernel of this call
    { const int rl = launch_1cta(ps, rem); if (rl) return rl; }
    launch_k(attn_bwd_tail_fixup_kernel, rem * FIXUP_SPLIT, 256, 0, st, (float*)tail_ws, (bf16*)dk, lddk, (bf16*)dv, lddv, ps, Lk);
    VDS_CHECK_LAUNCH("attn_bwd_tail_fixup");
    return VDS_OK;
  };

  // CTA-pair kernel (attention_bwd2.cu) for a self-attention-sized problem: 2 adjacent kv tiles of one (b, head) per
  // 2-CTA cluster.  VDS_ATTN_PAIR=0 disables it, VDS_ATTN_PAIR=force uses it regardless of problem size (tests).
  int pair_mode = g_attn_pair_mode;
  if (pair_mode < 0) {
    static int env_mode = -1;
    if (env_mode < 0) {
      const char* e = getenv("VDS_ATTN_PAIR");
      env_mode = (e == nullptr) ? 1 : (strcmp(e, "0") == 0 ? 0 : (strcmp(e, "force") == 0 ? 2 : 1));
    }
    pair_mode = env_mode;
  }
  const int pairs_per_bh = (kv_tiles + 1) / 2;   // the last pair of a (b, head) has a phantom second tile when kv_tiles is odd
  const int total_pairs = pairs_per_bh * nh * B;
  const int clusters = sms / 2;
  bool use_pairs = false;
  if (pair_mode != 0 && q_splits == 1 && dk != nullptr && dv != nullptr && kv_tiles >= 2) {
    // auto: long query ranges only (>= 64 sub-tiles: the pair kernel's longer prologue / epilogue costs more than the
    // ~12 % it gains per sub-tile on short ones: measured at L = 2064) and at least one full wave of pairs
    use_pairs = pair_mode == 2 || (n_qsub >= 64 && total_pairs >= clusters);
  }
  if (use_pairs) {
    // Whole waves of pairs run unsplit (bf16 dK / dV written directly).  The pairs of the last, partly filled wave are cut
    // into pieces along the query range (attn_bwd_plan_tail; fp32 red into the compact workspace + the bf16 fix-up) so that
    // all clusters finish together.
    const int full_p = (total_pairs / clusters) * clusters;
    int rem_p = total_pairs - full_p, main_p = full_p, n_pieces = 0;
    uint32_t pieces[VDS_BWD2_MAX_PIECES];
    if (rem_p > 0) {
      const bool can_split = tail_ws != nullptr && n_qsub >= 16 && (long long)rem_p * 2 * 2 * 128 * HD * 4 <= tail_ws_bytes;
      if (can_split) n_pieces = attn_bwd_plan_tail(rem_p, n_qsub, clusters, pieces, VDS_BWD2_MAX_PIECES);
      static const int uni = getenv("VDS_BWD2_UNIFORM") ? atoi(getenv("VDS_BWD2_UNIFORM")) : 0;   // tuning switch: uniform s-way split
      if (can_split && uni > 1 && rem_p * uni <= VDS_BWD2_MAX_PIECES) {
        const int per = (n_qsub + uni - 1) / uni;
        n_pieces = 0;
        for (int pr = 0; pr < rem_p; ++pr)
          for (int sp = 0; sp < uni; ++sp)
            pieces[n_pieces++] = (uint32_t)pr | ((uint32_t)(sp * per) << 10) | ((uint32_t)std::min(per, n_qsub - sp * per) << 21);
      }
      if (n_pieces == 0) { main_p = total_pairs; rem_p = 0; }   // no split possible / worthwhile: one launch for everything
    }
    static const bool prof_tail = getenv("VDS_B2_PROF_TAIL") != nullptr;   // tuning: the trace buffer goes to the split launch only
    AttnBwdParams pm = p;
    if (prof_tail && rem_p > 0) pm.dbg = nullptr;
    if (main_p > 0 &&
        (r = launch_attn_bwd_pairs(q, ldq, k, ldk, v, ldv, d_o, lddo, pm, B, nh, Lq, Lk, 0, main_p, pairs_per_bh, nullptr, 0, nullptr, st)))
      return r;
    if (rem_p > 0) {
      if ((r = launch_attn_bwd_pairs(q, ldq, k, ldk, v, ldv, d_o, lddo, p, B, nh, Lq, Lk, main_p, rem_p, pairs_per_bh, pieces, n_pieces,
                                     (float*)tail_ws, st)))
        return r;
      AttnBwdParams pf = p;   // fix-up: local item = 2 * (pair - main_p) + cta -> (b, head, kv tile); phantom tiles skip themselves
      pf.rem_pair_base = main_p; pf.rem_pairs_per_bh = pairs_per_bh; pf.rem_pair_tiles = 2 * rem_p;
      launch_k(attn_bwd_tail_fixup_kernel, 2 * rem_p * FIXUP_SPLIT, 256, 0, st, (float*)tail_ws, (bf16*)dk, lddk, (bf16*)dv, lddv, pf, Lk);
      VDS_CHECK_LAUNCH("attn_bwd_tail_fixup");
    }
    return VDS_OK;
  }
  // 1-CTA kernel for everything: full waves in one launch, the remainder of the last wave split along the query range
  const int full = (total / sms) * sms, rem = total - full;
  if (q_splits == 1 && tail_ws != nullptr && full > 0 && rem > 0 && n_qsub >= 32) {
    if ((r = launch_items(full, false))) return r;
    p.item_base = full;
    if ((r = launch_items(rem, true))) return r;
  } else {
    if ((r = launch_items(total, false))) return r;
  }
  return VDS_OK;
}

}  // extern "C"
