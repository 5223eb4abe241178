// 2-CTA tcgen05 GEMM (cta_group::2): a pair of CTAs on one TPC computes a 256 x 256 output tile.
//
// Each CTA keeps its own 128 x 64 A tile and HALF (128 rows) of the 256 x 64 B tile per pipeline stage; the
// leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256), whose tensor cores read both CTAs' shared
// memory, and each CTA's TMEM receives its 128 accumulator rows.  Per CTA and k-block that is 32 KiB of
// operand traffic from L2 for 128 x 256 x 64 MACs — half the B bytes of the 1-CTA kernel, which is what the
// K <= 2048 GEMMs of this model (L2 -> SM bandwidth bound at 128 x 256 tiles) need.
//
//   warp 0  TMA producer (both CTAs; .cta_group::2 loads complete on the LEADER's full barrier)
//   warp 1  MMA issuer (leader only) + TMEM alloc/dealloc (both)
//   warps 2..9 epilogue (both CTAs, own 128 rows; shared code with gemm.cu)
#include <cuda.h>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace vds {

long long* g_gemm2_trace = nullptr;   // tuning aid: CTA 0 accumulates wait / busy cycles per role

constexpr int G2_THREADS = 320;
constexpr int G2_BN = 256;              // tile N of the CTA pair
constexpr int G2_A_BYTES = BM * BK * 2;           // 16 KiB
constexpr int G2_B_BYTES = (G2_BN / 2) * BK * 2;  // 16 KiB: this CTA's half of B
constexpr int G2_STAGE = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_MAX_STAGES = 6;
// Shared memory: [stages x 32 KiB][8 epilogue warps x staging][barriers].  The plain / fp32 epilogues use 6 stages and
// one 4 KiB staging tile per warp; the fused element-wise epilogues (GELU, gate+residual, dGELU) are bound by their
// own latency chain, not by the operand pipeline, and trade one stage for a second staging tile per warp (aux
// operand prefetched by TMA one group ahead / both outputs stored without waiting in between).
template <int EPI>
struct G2Cfg {
  static constexpr bool fused = EPI == VDS_EPI_BIAS_GELU || EPI == VDS_EPI_GATE_RES || EPI == VDS_EPI_DGELU ||
                                EPI == VDS_EPI_STORE_ROWDOT || EPI == VDS_EPI_QKV_ROPE;
  // QKV_ROPE: a third 4 KiB tile per warp receives the cos / sin rows of the warp's 32 tokens by TMA, for one more stage
  static constexpr int stages = EPI == VDS_EPI_QKV_ROPE ? 4 : (fused ? 5 : 6);
  static constexpr int stg_warp = EPI == VDS_EPI_QKV_ROPE ? 12288 : (fused ? 8192 : 4096);
};
constexpr int G2_SMEM = 6 * G2_STAGE + 8 * 4096 + 256 + 1024;   // same total for every variant
// shared::cluster address of the same smem offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}


// Fused element-wise epilogue of one warp (32 accumulator rows x its 128-column half of the tile, as two 64-column
// groups per tile).  Two 4 KiB staging tiles S0 / S1 (1024-byte aligned, XOR swizzle == TMA 128-byte swizzle):
//   BIAS_GELU : act -> S0, Linear output -> S1, both TMA stores issued back to back.
//   DGELU     : aux (pre-activation h) arrives in S1 by a TMA load issued one group ahead; result -> S0 -> TMA store.
//   GATE_RES  : aux (residual x) in S1 as above; x + gate*out -> S0 -> store, then the Linear output through S0 again.
// Every wait on a store's shared-memory read sits behind a full group of math, so it is normally free.
template <int EPI>
__device__ __forceinline__ void fused_epilogue_warp(const GemmDev& p, const CUtensorMap* tmC, const CUtensorMap* tmC2,
                                                    const CUtensorMap* tmAux, uint8_t* S0, uint32_t aux_bar,
                                                    uint32_t tmem_base, uint32_t tfull_bar0, uint32_t leader_tempty0,
                                                    int tile0, int tile_step, int total_tiles, int m_pairs, int n_tiles,
                                                    int crank, int q, int chalf, int lane) {
  constexpr bool kAux = (EPI == VDS_EPI_GATE_RES || EPI == VDS_EPI_DGELU || EPI == VDS_EPI_STORE_ROWDOT);
  uint8_t* S1 = S0 + 4096;
  const uint32_t s0 = smem_u32(S0), s1 = smem_u32(S1);
  auto row_of = [&](int tile) { return ((tile % m_pairs) * 2 + crank) * BM + q * 32; };
  auto col_of = [&](int tile, int gi) { return ((tile / m_pairs) % n_tiles) * G2_BN + (chalf * 2 + gi) * 64; };
  uint32_t aux_phase = 0;
  if (kAux && tile0 < total_tiles && lane == 0) {
    mbar_expect_tx(aux_bar, 4096);
    tma_load_2d(s1, tmAux, aux_bar, col_of(tile0, 0), row_of(tile0));
  }
  const bool has_bias = (EPI != VDS_EPI_DGELU && EPI != VDS_EPI_STORE_ROWDOT) && p.bias != nullptr;
  const bool trace = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
  long long c_tf = 0, c_ld = 0, c_aux = 0, c_math = 0, c_rd = 0, c_sts = 0, c_st = 0;
  int it = 0;
  for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1u;
    long long tq = clock64();
    mbar_wait(tfull_bar0 + 8u * acc, acc_phase);
    c_tf += clock64() - tq;
    tc_fence_after();
    const int row0 = row_of(tile);
    const uint32_t t_base = tmem_base + acc * G2_BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int gi = 0; gi < 2; ++gi) {
      const int col0 = col_of(tile, gi);
      // accumulator columns are read 16 at a time, one load ahead of the math (a full 64-column read up front keeps 64
      // registers live through the whole group and pushes ptxas into a serial, latency-exposed schedule)
      uint32_t ra[16], rb[16];
      tq = clock64();
      tmem_ld16(t_base + (chalf * 2 + gi) * 64, ra);
      c_ld += clock64() - tq;
      tq = clock64();
      uint4 xa[8];      // aux row of this thread
      if constexpr (kAux) {
        mbar_wait(aux_bar, aux_phase);
        aux_phase ^= 1u;
#pragma unroll
        for (int g = 0; g < 8; ++g) xa[g] = *reinterpret_cast<const uint4*>(S1 + stg_off(lane, g));
        fence_proxy_async_smem();   // generic-proxy reads of S1 ordered before the async-proxy (TMA) refill issued below
        __syncwarp();
        // prefetch the aux tile of the next group (this tile's second group or the next tile's first)
        const bool more = gi == 0 || tile + tile_step < total_tiles;
        if (more && lane == 0) {
          const int nt_tile = gi == 0 ? tile : tile + tile_step, ngi = gi == 0 ? 1 : 0;
          mbar_expect_tx(aux_bar, 4096);
          tma_load_2d(s1, tmAux, aux_bar, col_of(nt_tile, ngi), row_of(nt_tile));
        }
      }
      c_aux += clock64() - tq;
      tq = clock64();
      uint4 keep[8];    // second output (bf16 Linear result)
      float dot = 0.f;  // STORE_ROWDOT: this row's partial <C, aux> over the 64 columns of the group
      const int b = (EPI == VDS_EPI_GATE_RES) ? min(row0 + lane, p.M - 1) / p.rows_per_batch : 0;
      uint4 bnext = make_uint4(0u, 0u, 0u, 0u), gnext = make_uint4(0u, 0u, 0u, 0u);
      const bool col_ok = col0 < p.N;        // false: a 64-column group past N (last, narrower tile): math on zeros, stores clipped
      const bool ld_bias = has_bias && col_ok;
      if (ld_bias) bnext = __ldg(reinterpret_cast<const uint4*>(p.bias + col0));
      if constexpr (EPI == VDS_EPI_GATE_RES)
        if (col_ok) gnext = __ldg(reinterpret_cast<const uint4*>(p.gate + (long long)b * p.gate_stride + col0));
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 braw = bnext, graw = gnext;
        if (g < 7) {   // per-column operands one chunk ahead (L1-resident after the first warp touched them)
          if (ld_bias) bnext = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + (g + 1) * 8));
          if constexpr (EPI == VDS_EPI_GATE_RES)
            if (col_ok) gnext = __ldg(reinterpret_cast<const uint4*>(p.gate + (long long)b * p.gate_stride + col0 + (g + 1) * 8));
        }
        uint4 outv;
        float a8[8];
        if ((g & 1) == 0) {
          tmem_ld_wait();
          if (g < 6) {
            tmem_ld16(t_base + (chalf * 2 + gi) * 64 + (g + 2) * 8, (g & 2) ? ra : rb);
          } else if (gi == 1) {   // last read of this accumulator buffer is complete: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = __uint_as_float(((g & 2) ? rb : ra)[(g & 1) * 8 + j]);
        if constexpr (EPI == VDS_EPI_BIAS_GELU || EPI == VDS_EPI_GATE_RES) {
          // unconditional (braw is zero without a bias): a branch here would cut the group into one basic block per
          // chunk and ptxas schedules only inside basic blocks
          float bb[8];
          unpack8(braw, bb);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 t = add2(make_float2(a8[j], a8[j + 1]), make_float2(bb[j], bb[j + 1]));
            a8[j] = t.x; a8[j + 1] = t.y;
          }
        }
        if constexpr (EPI == VDS_EPI_BIAS_GELU) {
          keep[g] = pack8(a8);                                // bf16 Linear output (what the reference's GELU sees)
          unpack8(keep[g], a8);
          gelu_erf8(a8);
          outv = pack8(a8);
        } else if constexpr (EPI == VDS_EPI_GATE_RES) {
          float g8[8], x8[8], o8[8];
          unpack8(graw, g8);
          unpack8(xa[g], x8);
          keep[g] = pack8(a8);                                // Linear output is bf16 in the reference
          unpack8(keep[g], a8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = x8[j] + bf16_round(a8[j] * g8[j]);   // x + (out * gate), each op rounded
          outv = pack8(o8);
        } else if constexpr (EPI == VDS_EPI_STORE_ROWDOT) {
          float x8[8];
          unpack8(xa[g], x8);
          outv = pack8(a8);
          unpack8(outv, a8);                                  // the consumer (attention backward) sees the bf16 values
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = fmaf(a8[j], x8[j], dot);
        } else {   // DGELU
          float h8[8];
          unpack8(xa[g], h8);
          dgelu_mul8(a8, h8);
          outv = pack8(a8);
        }
        // Results leave the registers chunk by chunk (short live ranges keep ptxas from serialising the math pair by
        // pair).  The staging tiles were handed to the TMA store of the previous group one TMEM load and one chunk of
        // math ago; its shared-memory read has normally finished by now.
        if (g == 0) {
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
        }
        *reinterpret_cast<uint4*>(S0 + stg_off(lane, g)) = outv;
        if constexpr (EPI == VDS_EPI_BIAS_GELU) *reinterpret_cast<uint4*>(S1 + stg_off(lane, g)) = keep[g];
      }
      c_math += clock64() - tq;
      tq = clock64();
      fence_proxy_async_smem();
      __syncwarp();
      c_sts += clock64() - tq;
      tq = clock64();
      if constexpr (EPI == VDS_EPI_STORE_ROWDOT) {
        const int row = row0 + lane;
        if (row < p.M && col_ok) {
          const int bb = row / p.rows_per_batch, rr = row % p.rows_per_batch;
          float* dst = reinterpret_cast<float*>(p.C2) + ((long long)bb * (p.N >> 7) + (col0 >> 7)) * p.rows_per_batch + rr;
          atomicAdd(dst, dot);   // two 64-column groups per head: a two-term sum, order-independent
        }
      }
      if (lane == 0) {
        if constexpr (EPI == VDS_EPI_DGELU || EPI == VDS_EPI_STORE_ROWDOT) {
          tma_store_2d(tmC, s0, col0, row0);
        } else {
          tma_store_2d(tmC2, s0, col0, row0);
          if (EPI == VDS_EPI_BIAS_GELU && p.C != nullptr) tma_store_2d(tmC, s1, col0, row0);
        }
        bulk_commit_group();
      }
      if constexpr (EPI == VDS_EPI_GATE_RES) {
        if (p.C != nullptr) {   // S1 is the aux landing tile, so the Linear output reuses S0 after its first store was read
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(S0 + stg_off(lane, g)) = keep[g];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(tmC, s0, col0, row0); bulk_commit_group(); }
        }
      }
      c_st += clock64() - tq;
    }
  }
  if (trace) {
    p.dbg[8] = c_tf; p.dbg[9] = c_ld; p.dbg[10] = c_aux; p.dbg[11] = c_math; p.dbg[12] = c_rd; p.dbg[13] = c_sts; p.dbg[14] = c_st;
  }
  if (lane == 0) bulk_wait_group0();   // all stores complete before the CTA may exit
}

// QKV projection epilogue (VDS_EPI_QKV_ROPE; model.py:124-134): a warp's 128-column half of the 256-wide tile is exactly one
// 128-wide head of q, k or v ("(k h d)" column order), and the thread owns one token row of it.
//   q / k heads: RoPE in fp32 on the bf16-rounded Linear output, half-split over the head (model.py:266-275): columns j and
//                64 + j of the thread's row meet in registers.  The cos / sin values of the warp's 32 token rows arrive by
//                TMA, 16 column pairs at a time, from the packed table of vds_rope_pack ([L + 32 wrap-around rows][4 steps]
//                [16 cos | 16 sin] fp32: one 32-row x 128-byte box per step, also across a sample boundary) into a third
//                staging tile T — per-thread loads of a 512-byte table row cost 32 line look-ups per warp instruction and
//                made the epilogue 2.2 x the mainloop.  x1-half -> S0, x2-half -> S1, one TMA store each.
//   v heads    : v_pre -> C; with the value residual (v0 != NULL) also v = bf16(l*v_pre) + bf16((1-l)*v0) -> C2 [M, N/3].
// Replaces the in-place qkv_post_fwd pass over [B*L, 3h] (one read + one write of the whole qkv buffer per block).
__device__ __forceinline__ void qkv_rope_epilogue_warp(const GemmDev& p, const CUtensorMap* tmC, const CUtensorMap* tmC2,
                                                       const CUtensorMap* tmTab, uint8_t* S0, uint32_t tab_bar,
                                                       uint32_t tmem_base, uint32_t tfull_bar0,
                                                       uint32_t leader_tempty0, int tile0, int tile_step, int total_tiles,
                                                       int m_pairs, int n_tiles, int crank, int q, int chalf, int lane) {
  uint8_t* S1 = S0 + 4096;
  uint8_t* T = S0 + 8192;
  const uint32_t s0 = smem_u32(S0), s1 = smem_u32(S1), st = smem_u32(T);
  const int h = p.N / 3;
  const bool mix = p.v0 != nullptr;
  float lam = 0.f, oml = 0.f;
  if (mix) { lam = __bfloat162float(*p.lambda); oml = bf16_round(1.0f - lam); }
  const bool has_bias = p.bias != nullptr;
  uint32_t tab_phase = 0;
  int it = 0;
  for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1u;
    const int row0 = ((tile % m_pairs) * 2 + crank) * BM + q * 32;
    const int colh = ((tile / m_pairs) % n_tiles) * G2_BN + chalf * 128;   // first column of this warp's head
    const bool col_ok = colh < p.N;          // false: the half past N of the last, narrower tile
    const int which = colh / h;              // 0: q, 1: k, 2: v (warp-uniform)
    const bool rot = col_ok && which < 2;
    const int l0 = min(row0, p.M - 1) % p.rows_per_batch;    // table row of the warp's first token (the table wraps)
    if (rot && lane == 0) {                  // cos / sin of step 0: in flight while the accumulator completes
      mbar_expect_tx(tab_bar, 4096);
      tma_load_3d(st, tmTab, tab_bar, 0, 0, l0);
    }
    mbar_wait(tfull_bar0 + 8u * acc, acc_phase);
    tc_fence_after();
    const uint32_t t_base = tmem_base + acc * G2_BN + (static_cast<uint32_t>(q * 32) << 16) + chalf * 128;
    const int row = min(row0 + lane, p.M - 1);
    auto release_acc = [&]() {               // last TMEM read of this accumulator buffer is complete
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
    };
    auto staging_free = [&]() {              // the TMA stores issued from S0 / S1 have finished reading them
      if (lane == 0) bulk_wait_group_read0();
      __syncwarp();
    };
    if (!col_ok) {
      release_acc();
    } else if (which < 2) {
      // No control flow between a tcgen05.ld and its wait::ld: with the asynchronous destination registers live across the
      // table barrier's spin loop ptxas placed register moves of them ahead of the wait (intermittently stale elements).
      staging_free();
#pragma unroll
      for (int s = 0; s < 4; ++s) {          // 16 column pairs (j, 64 + j) per step
        uint4 b1[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)}, b2[2] = {b1[0], b1[0]};
        if (has_bias) {
          b1[0] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + s * 16));
          b1[1] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + s * 16 + 8));
          b2[0] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + 64 + s * 16));
          b2[1] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + 64 + s * 16 + 8));
        }
        mbar_wait(tab_bar, tab_phase);
        tab_phase ^= 1u;
        uint32_t xa[16], xb[16];
        tmem_ld16(t_base + s * 16, xa);
        tmem_ld16(t_base + 64 + s * 16, xb);
        float c[16], sv[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {     // (overlaps the TMEM read)
          const float4 cc = *reinterpret_cast<const float4*>(T + stg_off(lane, ch));
          const float4 ss = *reinterpret_cast<const float4*>(T + stg_off(lane, 4 + ch));
          c[ch * 4] = cc.x; c[ch * 4 + 1] = cc.y; c[ch * 4 + 2] = cc.z; c[ch * 4 + 3] = cc.w;
          sv[ch * 4] = ss.x; sv[ch * 4 + 1] = ss.y; sv[ch * 4 + 2] = ss.z; sv[ch * 4 + 3] = ss.w;
        }
        tmem_ld_wait();
        // The table reads above (generic proxy) must be ordered before the TMA write of the next step into the same tile
        // (async proxy): without the proxy fence ptxas sank the last two LDS below the TMA issue, and behind a cold bias
        // miss they returned the NEXT step's values (columns 12..15 of the first q tiles, intermittently).
        fence_proxy_async_smem();
        __syncwarp();
        if (s < 3) {
          if (lane == 0) {                   // next step's table rows land while this step is computed
            mbar_expect_tx(tab_bar, 4096);
            tma_load_3d(st, tmTab, tab_bar, 0, s + 1, l0);
          }
        } else {
          release_acc();
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float bb1[8], bb2[8], y1[8], y2[8];
          unpack8(b1[g], bb1);
          unpack8(b2[g], bb2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a1 = bf16_round(__uint_as_float(xa[g * 8 + j]) + bb1[j]);   // what the reference's RoPE sees
            const float a2 = bf16_round(__uint_as_float(xb[g * 8 + j]) + bb2[j]);
            const float cj = c[g * 8 + j], sj = sv[g * 8 + j];
            y1[j] = a1 * cj + a2 * sj;
            y2[j] = a1 * (-sj) + a2 * cj;
          }
          *reinterpret_cast<uint4*>(S0 + stg_off(lane, s * 2 + g)) = pack8(y1);
          *reinterpret_cast<uint4*>(S1 + stg_off(lane, s * 2 + g)) = pack8(y2);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, s0, colh, row0);
        tma_store_2d(tmC, s1, colh + 64, row0);
        bulk_commit_group();
      }
    } else {
      const bf16* v0row = mix ? p.v0 + (long long)row * p.ldv0 + (colh - 2 * h) : nullptr;
#pragma unroll 1
      for (int gi = 0; gi < 2; ++gi) {       // 64-column groups: v_pre -> S0, mixed v -> S1
        staging_free();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint4 b1[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)}, w0[2] = {b1[0], b1[0]};
          if (has_bias) {
            b1[0] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + gi * 64 + s * 16));
            b1[1] = __ldg(reinterpret_cast<const uint4*>(p.bias + colh + gi * 64 + s * 16 + 8));
          }
          if (mix) {
            w0[0] = __ldg(reinterpret_cast<const uint4*>(v0row + gi * 64 + s * 16));
            w0[1] = __ldg(reinterpret_cast<const uint4*>(v0row + gi * 64 + s * 16 + 8));
          }
          uint32_t xa[16];
          tmem_ld16(t_base + gi * 64 + s * 16, xa);
          tmem_ld_wait();
          if (s == 3 && gi == 1) release_acc();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float bb[8], a8[8];
            unpack8(b1[g], bb);
#pragma unroll
            for (int j = 0; j < 8; ++j) a8[j] = __uint_as_float(xa[g * 8 + j]) + bb[j];
            const uint4 vpre = pack8(a8);
            *reinterpret_cast<uint4*>(S0 + stg_off(lane, s * 2 + g)) = vpre;
            if (mix) {
              float v8[8], o8[8];
              unpack8(vpre, a8);
              unpack8(w0[g], v8);
#pragma unroll
              for (int j = 0; j < 8; ++j) o8[j] = bf16_round(lam * a8[j]) + bf16_round(oml * v8[j]);
              *reinterpret_cast<uint4*>(S1 + stg_off(lane, s * 2 + g)) = pack8(o8);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, s0, colh + gi * 64, row0);
          if (mix) tma_store_2d(tmC2, s1, colh - 2 * h + gi * 64, row0);
          bulk_commit_group();
        }
      }
    }
  }
  if (lane == 0) bulk_wait_group0();   // all stores complete before the CTA may exit
}

// (10 warps = 3 on two of the SM sub-partitions: 3 x 32 x regs <= 16K caps the kernel at 168 registers per thread)
template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
             const __grid_constant__ CUtensorMap tmAux, const GemmDev p, long long* dbg, int tma_c, int tma_c2,
             int fast) {
  constexpr int G2_STAGES = G2Cfg<EPI>::stages;
  constexpr int G2_STG = 8 * G2Cfg<EPI>::stg_warp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t bar_base = smem_base + G2_STAGES * G2_STAGE + G2_STG;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G2_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * G2_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * G2_MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * G2_MAX_STAGES + 4);
  auto aux_bar = [&](int w) { return bar_base + 8u * (2 * G2_MAX_STAGES + 5 + w); };   // one per epilogue warp
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + G2_STAGES * G2_STAGE + G2_STG + 8 * (2 * G2_MAX_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar(s), 2);    // one producer arrival per CTA (+ the bytes of both CTAs' loads); leader's is used
      mbar_init(empty_bar(s), 1);   // leader's MMA commit, multicast to both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);   // leader's MMA commit, multicast
      mbar_init(tempty_bar(a), 16); // 8 epilogue warps x 2 CTAs arrive on the LEADER's barrier
    }
    for (int w = 0; w < 8; ++w) mbar_init(aux_bar(w), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  pdl_wait();      // everything above is launch-independent set-up; global inputs may come from the previous kernel
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (p.N + G2_BN - 1) / G2_BN;   // N % 64 == 0; the columns past N of the last tile are TMA zero-fill / clipped
  const int k_iters = (p.K + BK - 1) / BK;
  const int total_tiles = m_pairs * n_tiles * p.splits;
  const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int mt = (tile % m_pairs) * 2 + crank;
        const int rest = tile / m_pairs;
        const int nt = rest % n_tiles;
        const int sp = rest / n_tiles;
        const int kb0 = sp * p.k_per_split;
        const int kb1 = min(k_iters, kb0 + p.k_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t lbar = mapa_rank(full_bar(stage), 0);   // leader's full barrier (shared::cluster address)
          if (leader) mbar_expect_tx(full_bar(stage), 2 * G2_STAGE);
          else mbar_arrive_cluster(lbar);
          const uint32_t a_dst = smem_base + stage * G2_STAGE;
          const uint32_t b_dst = a_dst + G2_A_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(a_dst, &tmA, lbar, kb * BK, mt * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d_2sm(a_dst + i * (BK * 128), &tmA, lbar, mt * BM + i * 64, kb * BK);
          }
          const int n0 = nt * G2_BN + crank * (G2_BN / 2);      // this CTA's half of the B tile
          if constexpr (!B_MN) {
            tma_load_2d_2sm(b_dst, &tmB, lbar, kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < G2_BN / 128; ++i)
              tma_load_2d_2sm(b_dst + i * (BK * 128), &tmB, lbar, n0 + i * 64, kb * BK);
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, G2_BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_tempty = 0, w_full = 0;
      const long long t_begin = clock64();
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int rest = tile / m_pairs;
        const int sp = rest / n_tiles;
        const int kb0 = sp * p.k_per_split;
        const int kb1 = min(k_iters, kb0 + p.k_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        long long tq = clock64();
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        w_tempty += clock64() - tq;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          tq = clock64();
          mbar_wait(full_bar(stage), phase);
          w_full += clock64() - tq;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + stage * G2_STAGE;
            const uint32_t b_addr = a_addr + G2_A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = A_MN ? umma_smem_desc(a_addr + k * 2048, BK * 128, 1024)
                                          : umma_smem_desc(a_addr + k * 32, 16, 1024);
              const uint64_t bdesc = B_MN ? umma_smem_desc(b_addr + k * 2048, BK * 128, 1024)
                                          : umma_smem_desc(b_addr + k * 32, 16, 1024);
              umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(stage), 3);                       // slot free in both CTAs
            if (kb == kb1 - 1) umma_commit_2sm(tfull_bar(acc), 3);      // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (dbg != nullptr && blockIdx.x == 0 && lane == 0) {
        dbg[0] = clock64() - t_begin; dbg[1] = w_tempty; dbg[2] = w_full; dbg[3] = it;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const uint32_t leader_tempty0 = mapa_rank(tempty_bar(0), 0);
    int it = 0;
    long long e_wait = 0, e_busy = 0, e_tmem = 0, e_grp = 0;
    if constexpr (EPI == VDS_EPI_QKV_ROPE) {
      qkv_rope_epilogue_warp(p, &tmC, &tmC2, &tmAux, smem_gen + G2_STAGES * G2_STAGE + (warp - 2) * G2Cfg<EPI>::stg_warp,
                             aux_bar(warp - 2), tmem_base, tfull_bar(0), leader_tempty0, tile0, tile_step, total_tiles,
                             m_pairs, n_tiles, crank, q, chalf, lane);
      it = -1;
    } else if constexpr (G2Cfg<EPI>::fused) {
      if (fast) {
        fused_epilogue_warp<EPI>(p, &tmC, &tmC2, &tmAux, smem_gen + G2_STAGES * G2_STAGE + (warp - 2) * 8192,
                                 aux_bar(warp - 2), tmem_base, tfull_bar(0), leader_tempty0, tile0, tile_step, total_tiles,
                                 m_pairs, n_tiles, crank, q, chalf, lane);
        it = -1;
      }
    }
    for (int tile = tile0; it >= 0 && tile < total_tiles; tile += tile_step, ++it) {
      const int mt = (tile % m_pairs) * 2 + crank;
      const int nt = (tile / m_pairs) % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const long long te0 = clock64();
      mbar_wait(tfull_bar(acc), acc_phase);
      const long long te1 = clock64();
      e_wait += te1 - te0;
      tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const uint32_t t_base = tmem_base + acc * G2_BN + (static_cast<uint32_t>(q * 32) << 16);
      constexpr bool kF32 = (EPI == VDS_EPI_ACCUM_F32 || EPI == VDS_EPI_STORE_F32);
      if constexpr (kF32) {
#pragma unroll 1
        for (int c = chalf * (G2_BN / 64); c < (chalf + 1) * (G2_BN / 64); ++c) {
          uint32_t v[32];
          tmem_ld32(t_base + c * 32, v);
          tmem_ld_wait();
          if (EPI == VDS_EPI_ACCUM_F32 && tma_c)
            accum_f32_tma(&tmC, smem_gen + G2_STAGES * G2_STAGE + (warp - 2) * 4096, mt * BM + q * 32, nt * G2_BN + c * 32, v, lane);
          else
            epilogue_row<EPI>(p, row, nt * G2_BN + c * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
      } else if constexpr (EPI != VDS_EPI_STORE_ROWDOT && EPI != VDS_EPI_QKV_ROPE) {   // (these exist only as fused fast paths)
        uint8_t* stg = smem_gen + G2_STAGES * G2_STAGE + (warp - 2) * 4096;   // 1024-byte aligned (TMA swizzle atom)
        constexpr int GROUPS = G2_BN / 128;
#pragma unroll 1
        for (int gi = 0; gi < GROUPS; ++gi) {
          const int cg = chalf * GROUPS + gi;
          uint32_t r0[32], r1[32];
          const long long tg0 = clock64();
          tmem_ld32(t_base + cg * 64, r0);
          tmem_ld32(t_base + cg * 64 + 32, r1);
          tmem_ld_wait();
          const long long tg1 = clock64();
          if (gi == GROUPS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc);
          }
          epilogue_group64<EPI>(p, stg, mt * BM + q * 32, nt * G2_BN + cg * 64, r0, r1, lane, tma_c ? &tmC : nullptr,
                                tma_c2 ? &tmC2 : nullptr);
          e_tmem += tg1 - tg0;
          e_grp += clock64() - tg1;
        }
      }
      e_busy += clock64() - te1;
    }
    if (lane == 0) bulk_wait_group0();   // all TMA stores of this warp have completed before the CTA may exit
    if (dbg != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0) { dbg[4] = e_wait; dbg[5] = e_busy; dbg[6] = e_tmem; dbg[7] = e_grp; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm2(const vds_gemm_args& a, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!A_MN) { dims[0] = a.K; dims[1] = a.M; box[0] = BK; box[1] = BM; }
    else       { dims[0] = a.M; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a.lda * 2;
    int r = encode_tmap_bf16(&tmA, a.A, 2, dims, strides, box);
    if (r) return r;
    if (!B_MN) { dims[0] = a.K; dims[1] = a.N; box[0] = BK; box[1] = G2_BN / 2; }
    else       { dims[0] = a.N; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a.ldb * 2;
    r = encode_tmap_bf16(&tmB, a.B, 2, dims, strides, box);
    if (r) return r;
  }
  // output tensor maps for the TMA-store epilogue (bf16 outputs without row remap)
  CUtensorMap tmC = tmA, tmC2 = tmA;
  int tma_c = 0, tma_c2 = 0;
  // (measured: helps the plain / GELU epilogues; the aux-reading epilogues are latency-bound elsewhere and get slower)
  if ((EPI == VDS_EPI_STORE || EPI == VDS_EPI_BIAS_GELU) && a.remap_rows == 0) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M}, strides[1];
    uint32_t box[2] = {64, 32};
    if (a.C != nullptr && a.ldc % 8 == 0 && ((uintptr_t)a.C & 15) == 0) {
      strides[0] = (uint64_t)a.ldc * 2;
      int r = encode_tmap_bf16(&tmC, a.C, 2, dims, strides, box);
      if (r) return r;
      tma_c = 1;
    }
    if (a.C2 != nullptr && a.ldc2 % 8 == 0 && ((uintptr_t)a.C2 & 15) == 0) {
      strides[0] = (uint64_t)a.ldc2 * 2;
      int r = encode_tmap_bf16(&tmC2, a.C2, 2, dims, strides, box);
      if (r) return r;
      tma_c2 = 1;
    }
  }
  if (EPI == VDS_EPI_ACCUM_F32) {   // split-K wgrad: fp32 tiles added into C by the TMA unit
    int r = make_tmap_accum_f32(&tmC, a.C, a.ldc, a.M, a.N, &tma_c);
    if (r) return r;
  }
  // fused fast path: every bf16 operand / output of the epilogue reachable by TMA (no row remap, 16-byte aligned)
  CUtensorMap tmAux = tmA;
  int fast = 0;
  if (G2Cfg<EPI>::fused && EPI != VDS_EPI_QKV_ROPE && a.remap_rows == 0) {
    auto ok = [](const void* ptr, long long ld) { return ptr != nullptr && ld % 8 == 0 && ((uintptr_t)ptr & 15) == 0; };
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M}, strides[1];
    uint32_t box[2] = {64, 32};
    auto enc = [&](CUtensorMap* tm, const void* ptr, long long ld) {
      strides[0] = (uint64_t)ld * 2;
      return encode_tmap_bf16(tm, ptr, 2, dims, strides, box);
    };
    const bool need_aux = EPI != VDS_EPI_BIAS_GELU;
    const bool main_c2 = EPI != VDS_EPI_DGELU && EPI != VDS_EPI_STORE_ROWDOT;   // primary output: C2 (act / residual stream), else C
    bool good = main_c2 ? ok(a.C2, a.ldc2) : ok(a.C, a.ldc);
    if (main_c2 && a.C != nullptr) good = good && ok(a.C, a.ldc);
    if (need_aux) good = good && ok(a.aux, a.ldaux);
    if (good) {
      int r = 0;
      if (main_c2) r = enc(&tmC2, a.C2, a.ldc2);
      if (!r && a.C != nullptr) r = enc(&tmC, a.C, a.ldc);
      if (!r && need_aux) r = enc(&tmAux, a.aux, a.ldaux);
      if (r) return r;
      fast = 1;
    }
  }
  if (EPI == VDS_EPI_QKV_ROPE) {
    auto ok = [](const void* ptr, long long ld) { return ptr != nullptr && ld % 8 == 0 && ((uintptr_t)ptr & 15) == 0; };
    const int h = a.N / 3;
    const bool mix = a.v0 != nullptr;
    if (a.N % 384 != 0 || a.rope_tab == nullptr || ((uintptr_t)a.rope_tab & 15) != 0 || a.rows_per_batch <= 0 || a.remap_rows != 0 ||
        !ok(a.C, a.ldc) || (mix && (!ok(a.C2, a.ldc2) || a.lambda_ == nullptr || a.ldv0 % 8 != 0 || ((uintptr_t)a.v0 & 15) != 0))) {
      set_error("gemm2: QKV_ROPE needs N = 3 * (heads * 128), the packed cos / sin table, rows_per_batch, TMA-able C (and C2, v0, lambda with the value residual)");
      return VDS_ERR_UNSUPPORTED;
    }
    {   // packed table [rows_per_batch + 32][4][32] fp32 (vds_rope_pack): box = 32 token rows x one 128-byte step
      uint64_t tdims[3] = {32, 4, (uint64_t)a.rows_per_batch + 32}, tstr[2] = {128, 512};
      uint32_t tbox[3] = {32, 1, 32};
      int r = encode_tmap(&tmAux, a.rope_tab, 1, 3, tdims, tstr, tbox, 1);
      if (r) return r;
    }
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M}, strides[1] = {(uint64_t)a.ldc * 2};
    uint32_t box[2] = {64, 32};
    int r = encode_tmap_bf16(&tmC, a.C, 2, dims, strides, box);
    if (r) return r;
    if (mix) {
      dims[0] = (uint64_t)h; strides[0] = (uint64_t)a.ldc2 * 2;
      if ((r = encode_tmap_bf16(&tmC2, a.C2, 2, dims, strides, box))) return r;
    }
    fast = 1;
  }
  if (EPI == VDS_EPI_STORE_ROWDOT) {
    if (!fast || a.C2 == nullptr || a.rows_per_batch <= 0 || a.N % 128 != 0) {
      set_error("gemm2: STORE_ROWDOT needs TMA-able C / aux, a rowdot buffer (C2), rows_per_batch and N %% 128 == 0");
      return VDS_ERR_UNSUPPORTED;
    }
  }
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  const int k_iters = (a.K + BK - 1) / BK;
  int splits = (EPI == VDS_EPI_ACCUM_F32) ? (a.splits < 1 ? 1 : a.splits) : 1;
  if (splits > k_iters) splits = k_iters;
  p.k_per_split = (k_iters + splits - 1) / splits;
  p.splits = (k_iters + p.k_per_split - 1) / p.k_per_split;
  p.C = a.C; p.ldc = a.ldc; p.C2 = a.C2; p.ldc2 = a.ldc2;
  p.bias = reinterpret_cast<const bf16*>(a.bias);
  p.aux = reinterpret_cast<const bf16*>(a.aux); p.ldaux = a.ldaux;
  p.gate = reinterpret_cast<const bf16*>(a.gate); p.gate_stride = a.gate_stride;
  p.rows_per_batch = a.rows_per_batch > 0 ? a.rows_per_batch : 1;
  p.remap_rows = a.remap_rows; p.remap_stride = a.remap_stride; p.remap_offset = a.remap_offset;
  p.v0 = reinterpret_cast<const bf16*>(a.v0); p.ldv0 = a.ldv0; p.lambda = reinterpret_cast<const bf16*>(a.lambda_);
  p.dbg = g_gemm2_trace;

  auto kern = gemm2_kernel<A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    if (e != cudaSuccess) {
      set_error("gemm2: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return VDS_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long total = (long long)((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + G2_BN - 1) / G2_BN) * p.splits;
  const int max_pairs = num_sms() / 2;
  const int grid = (int)(total < max_pairs ? total : max_pairs) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see common.h: launch_k
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmC2, tmAux, p, g_gemm2_trace, tma_c, tma_c2, fast);
  if (e != cudaSuccess) {
    set_error("gemm2: cluster launch failed: %s", cudaGetErrorString(e));
    return VDS_ERR_CUDA;
  }
  VDS_CHECK_LAUNCH("gemm2");
  return VDS_OK;
}

void gemm2_set_trace(long long* p) { g_gemm2_trace = p; }

// Entry used by gemm.cu's dispatcher: returns VDS_ERR_UNSUPPORTED when the shape / epilogue has no 2-CTA variant.
int gemm2_dispatch(const vds_gemm_args& a, cudaStream_t s) {
  // N needs whole 64-column groups (every epilogue works on those); a last tile narrower than 256 (N = 1152, 3456 of the
  // DiT-XL width) is computed full width on zero-filled B rows and clipped on the way out
  if (a.N % 64 != 0 || a.N < G2_BN) return VDS_ERR_UNSUPPORTED;
  if (!a.a_mn && !a.b_mn) {
    switch (a.epilogue) {
      case VDS_EPI_STORE: return launch_gemm2<false, false, VDS_EPI_STORE>(a, s);
      case VDS_EPI_BIAS_GELU: return launch_gemm2<false, false, VDS_EPI_BIAS_GELU>(a, s);
      case VDS_EPI_GATE_RES: return launch_gemm2<false, false, VDS_EPI_GATE_RES>(a, s);
      case VDS_EPI_QKV_ROPE: return launch_gemm2<false, false, VDS_EPI_QKV_ROPE>(a, s);
      default: return VDS_ERR_UNSUPPORTED;
    }
  }
  if (!a.a_mn && a.b_mn) {
    switch (a.epilogue) {
      case VDS_EPI_STORE: return launch_gemm2<false, true, VDS_EPI_STORE>(a, s);
      case VDS_EPI_DGELU: return launch_gemm2<false, true, VDS_EPI_DGELU>(a, s);
      case VDS_EPI_STORE_ROWDOT: return launch_gemm2<false, true, VDS_EPI_STORE_ROWDOT>(a, s);
      default: return VDS_ERR_UNSUPPORTED;
    }
  }
  if (a.a_mn && a.b_mn && a.epilogue == VDS_EPI_ACCUM_F32) return launch_gemm2<true, true, VDS_EPI_ACCUM_F32>(a, s);
  return VDS_ERR_UNSUPPORTED;
}

}  // namespace vds
