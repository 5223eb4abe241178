// Attention backward on CTA PAIRS (tcgen05.mma.cta_group::2 + distributed shared memory), head_dim = 128
// (autograd of model.py:136; the 1-CTA kernel in attention.cu keeps cross-attention, query-range splits and tails).
//
// Why pairs.  The 1-CTA kernel is bound by its shared-memory port (~288 KiB per 64-row query sub-tile).
// A cluster of two CTAs owns two adjacent 128-row K/V tiles of one (b, head) and walks the query range together:
//   S^T  = K Q^T , dP^T = V dO^T     M = 256 (kv rows of both CTAs) x N = 64 (q) x K = 128 (d)
//   dV  += P^T dO, dK  += dS^T Q     M = 256 x N = 128 (d) x K = 64 (q), A = P^T / dS^T (bf16) from TMEM
//        every CTA feeds its own 128 rows of A and only HALF of each B operand (32 query rows, resp. 64 of the 128
//        d columns), so the Q / dO operand reads and fills per CTA are halved;
//   dQ^T = K^T dS^T                   M = 128 (d: 64 per CTA) x N = 64 (q) x K = 256 (kv rows of BOTH tiles)
//        the contraction runs over the pair's 256 kv rows, so each CTA ends up with a fully pair-summed 64 (d) x 64 (q)
//        block: half the fp32 staging traffic and half the TMA reductions into the fp32 dq buffer.  Its B operand —
//        dS^T of all 256 kv rows for this CTA's 32 query columns — is assembled by the compute warps of both CTAs: every
//        thread (= one kv row) stores one 64-byte half row locally and the other half into the peer's shared memory
//        (st.shared::cluster), then fence.proxy.async + a release.cluster arrive on the leader's barrier.
// Per CTA and sub-tile that is ~224 KiB through the shared-memory port instead of ~288 KiB and ~1560 tensor-pipe cycles
// instead of ~1840 (scripts/mma_shapes2.cu: M256 N64 SS 40-50 cycles for the pair against 53 per CTA; M128 N64 22-28).
// The operand / accumulator layouts of the three cta_group::2 shapes (N-split B halves, the 128-lane x 32-column TMEM
// image of the M128 accumulator, the 64-byte-swizzle MN-major B tile, remote stores feeding the async proxy) were pinned
// on the hardware with scripts/umma_probe.py before this kernel was written.
//
// Roles per CTA (512 threads): warp 0 TMA producer | warp 1 MMA issuer (leader CTA only; relay in the follower) |
// warps 4-11 compute: thread == kv row, warps 4-7 take query columns 0..31 of the sub-tile and warps 8-11 columns 32..63,
// i.e. TWO warps per SM sub-partition — one warp per sub-partition ran the ~640 instructions of a sub-tile at IPC 0.33
// (every FMA / MUFU / conversion latency exposed), two interleave | warps 12-15 dQ drain.  All MMA-completion signals are
// commits multicast to both CTAs' barriers; everything the issuer waits for arrives on the LEADER's barriers (remote
// arrives from the follower).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "attn_common.cuh"
#include "common.h"
#include "ptx.cuh"

namespace vds {

constexpr int B2_THREADS = 512;
constexpr int B2_OFF_K = 0;                               // own K tile, K-major SW128, two 64-wide d halves      32 KiB
constexpr int B2_OFF_V = B2_OFF_K + TILE_BYTES;           // own V tile                                             32 KiB
constexpr int B2_OFF_KT = B2_OFF_V + TILE_BYTES;          // [kv tile 0 | 1] x [128 kv x 64 d of THIS CTA's d half]  32 KiB
constexpr int B2_ROWS_STAGE = 16384;                      // Q rows-half 2 x [32 x 128 B] | dO rows-half
constexpr int B2_OFF_ROWS = B2_OFF_KT + TILE_BYTES;       // 2 stages                                               32 KiB
constexpr int B2_COLS_STAGE = 16384;                      // Q cols-half [64 q x 128 B] | dO cols-half
constexpr int B2_OFF_COLS = B2_OFF_ROWS + 2 * B2_ROWS_STAGE;   // 2 stages                                          32 KiB
constexpr int B2_OFF_ONES = B2_OFF_COLS + 2 * B2_COLS_STAGE;   // ones tile [128 x 16] no swizzle                    4 KiB
constexpr int B2_OFF_STAT = B2_OFF_ONES + 4096;           // [buf 0|1][lse | delta] x [32 q x 16] no swizzle          4 KiB
constexpr int B2_DS_SET = 16384;                          // one set: [kv tile 0 | 1] x [128 kv x 32 q] bf16, 64-byte swizzle
constexpr int B2_OFF_DS = B2_OFF_STAT + 4096;             // 2 sets (sub-tile parity)                                 32 KiB
constexpr int B2_OFF_SEND = B2_OFF_DS + 2 * B2_DS_SET;    // 2 sets x this kv tile's dS^T half for the PEER (DSMEM copy) 16 KiB
constexpr int B2_OFF_STG = B2_OFF_SEND + 2 * 8192;        // fp32 [32 q][64 d] staging of half a dQ block             8 KiB
constexpr int B2_OFF_BAR = B2_OFF_STG + 8192;
constexpr int B2_SMEM = B2_OFF_BAR + 256 + 1024;

__device__ __forceinline__ uint32_t mapa_cta(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
// Arrive on a barrier of (possibly) another CTA of the cluster.  What these arrivals publish lives in TMEM (ordered by
// tcgen05.fence::before_thread_sync) or in this CTA's own shared memory behind fence.proxy.async, so the default
// .release.cta form is used: a .release.cluster arrive costs a cluster-scope memory barrier (measured: it and the generic
// fence.proxy.async after remote stores made the compute warps' hand-off ~3000 cycles per sub-tile).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// shared::cta -> peer shared memory bulk copy (async proxy on both sides); completes `bytes` on the PEER's barrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t peer_dst, uint32_t src, uint32_t bytes, uint32_t peer_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(peer_dst),
               "r"(src), "r"(bytes), "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the barrier at this offset in BOTH CTAs when complete
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// MN-major B tile with 32 elements (64 bytes) per k row, 64-byte swizzle: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(2048 >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// byte offset of 16-byte chunk c (0..3) of row r in a [rows x 64 B] tile with the 64-byte swizzle
__device__ __forceinline__ uint32_t sw64_offset(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// Tuning aid (compile with -DVDS_B2_PROF): cluster 0 accumulates the cycles each role spends in each of its waits and
// writes them at the end: dbg[cta][role 0 issuer | 1 compute warp 4 | 2 drain warp 8][16] (slot 15 = loop total).
#ifdef VDS_B2_PROF
#define B2_PROF_DECL long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64();
#define B2_WAIT(slot, ...) do { const long long t_ = clock64(); __VA_ARGS__; prof_acc[slot] += clock64() - t_; } while (0)
#define B2_PROF_STORE(role)                                                                                     \
  do {                                                                                                          \
    if (p.dbg != nullptr && blockIdx.x < 2 && lane == 0) {                                                      \
      prof_acc[15] = clock64() - prof_t0;                                                                       \
      for (int s_ = 0; s_ < 16; ++s_) p.dbg[((long long)crank * 4 + (role)) * 16 + s_] = prof_acc[s_];          \
    }                                                                                                           \
  } while (0)
#else
#define B2_PROF_DECL
#define B2_WAIT(slot, ...) do { __VA_ARGS__; } while (0)
#define B2_PROF_STORE(role) do { } while (0)
#endif
#define B2_TRACE(slot, it) do { } while (0)
// phase stamps of cluster 0 (PROF build): dbg[(crank * 4 + 3) * 16 + slot] = cycles since kernel entry
#ifdef VDS_B2_PROF
#define B2_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x < 2) p.dbg[((long long)crank * 4 + 3) * 16 + (slot)] = clock64() - t_entry; } while (0)
#else
#define B2_STAMP(slot) do { } while (0)
#endif

struct AttnBwd2Params {
  AttnBwdParams p;
  // pair -> (b*nh + head, kv tiles 2j, 2j+1), pairs_per_bh = ceil(kv_tiles / 2): the second tile of the last pair of a
  // (b, head) may lie entirely past Lk (its loads are TMA zero-fill, its rows masked, nothing of it is written).
  //   n_pieces == 0: cluster c owns pair pair_base + c over the whole query range and writes bf16 dK / dV (TMA store when
  //                  tma_dkdv, else per-thread stores).
  //   n_pieces  > 0: cluster c owns PIECE c of the plan (attn_bwd_plan_tail): pieces[c] = pair_local | first sub-tile << 10 |
  //                  sub-tile count << 21; it adds its dK / dV (fp32 red) into compact[pair_local * 2 + cta][dk | dv][128][128].
  int pair_base, pairs_per_bh, n_pieces, tma_dkdv;
  float* compact;
  uint32_t pieces[VDS_BWD2_MAX_PIECES];
};

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQr,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDOr,
                 const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                 const __grid_constant__ CUtensorMap tmDV, const __grid_constant__ AttnBwd2Params pp) {
  const AttnBwdParams& p = pp.p;
#ifdef VDS_B2_PROF
  const long long t_entry = clock64();
  // per-CTA timeline (every CTA of the launch): dbg[128 + 4 * blockIdx.x + {0: smid, 1: entry, 2: set-up done, 3: exit}] in ns
  auto gtimer = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  if (threadIdx.x == 0 && p.dbg != nullptr) {
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    p.dbg[128 + 4 * (long long)blockIdx.x] = sm;
    p.dbg[128 + 4 * (long long)blockIdx.x + 1] = gtimer();
  }
#endif
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_addr);
  const uint32_t sK = base + B2_OFF_K, sV = base + B2_OFF_V, sKT = base + B2_OFF_KT, sROWS = base + B2_OFF_ROWS,
                 sCOLS = base + B2_OFF_COLS, sONES = base + B2_OFF_ONES, sSTAT = base + B2_OFF_STAT, sDS = base + B2_OFF_DS,
                 sSTG = base + B2_OFF_STG, sSEND = base + B2_OFF_SEND;
  uint8_t* gONES = gen + B2_OFF_ONES;
  uint8_t* gSTAT = gen + B2_OFF_STAT;
  uint8_t* gDS = gen + B2_OFF_DS;
  uint8_t* gSEND = gen + B2_OFF_SEND;
  float* gSTG = reinterpret_cast<float*>(gen + B2_OFF_STG);
  const uint32_t bars = base + B2_OFF_BAR;
  // Barriers the issuer waits on (the LEADER's copies are used; the follower arrives remotely).  Everything one MMA group
  // needs is folded into ONE barrier, because each wait costs the single issuing warp ~80-100 cycles even when it has
  // long completed, and the tensor pipe idles meanwhile (measured: 8 waits = 830 of 2770 cycles per sub-tile):
  //   s_ready[k & 1]    S^T(k), dP^T(k): rows stage landed (2 producer arrivals + bytes) + statistics tiles written (16 warps)
  //   dp_read           dP^T(k+1): the compute warps hold dP^T(k) in registers (16 warps)
  //   dvdk_ready[i & 1] dV(i), dK(i): P^T / dS^T in TMEM (16 warps) + cols stage landed (2 + bytes)
  //   dq_ready[j & 1]   dQ^T(j): peer's dS^T half landed here (arm + bytes), the follower's relay of the same (1), and
  //                     dQ^T(j-1) drained out of TMEM (8 warps; pre-arrived by the issuer for j = 0)
  const uint32_t kv_full = bars, s_ready = bars + 8, dp_read = bars + 24, dvdk_ready = bars + 32, dq_ready = bars + 48;
  // barriers every CTA waits on locally (multicast commits of the leader's MMAs)
  const uint32_t rows_empty = bars + 80, cols_empty = bars + 96, s_full = bars + 112, dp_full = bars + 128,
                 dq_full = bars + 136 /* [2]: by sub-tile parity */, mma_done = bars + 152, tmem_slot = bars + 168;
  // follower only: the leader's dS^T half has landed in my sDS set (relayed to the leader's dq_ready)
  const uint32_t ds_in = bars + 176;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + B2_OFF_BAR + 168);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const bool split_mode = pp.n_pieces > 0;
  const uint32_t piece = split_mode ? pp.pieces[cluster_id] : 0u;
  const int pair_local = split_mode ? (int)(piece & 1023u) : cluster_id;
  const int pid = pp.pair_base + pair_local;
  const int bh = pid / pp.pairs_per_bh, pj = pid % pp.pairs_per_bh;
  const int head = bh % p.nh, b = bh / p.nh;
  const int kv0_pair = pj * 256, kv0 = kv0_pair + (int)crank * 128;
  const int n_q_all = (p.Lq + QSUB - 1) / QSUB;
  const int qt0 = split_mode ? (int)((piece >> 10) & 2047u) : 0;     // first query sub-tile of this cluster
  const int n_q = split_mode ? min((int)(piece >> 21), n_q_all - qt0) : n_q_all;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_ready + 8 * s, 18);      // 2 producers + 8 compute warps x 2 CTAs
      mbar_init(dvdk_ready + 8 * s, 18);
      mbar_init(dq_ready + 8 * s, 10);
      mbar_init(rows_empty + 8 * s, 1);
      mbar_init(cols_empty + 8 * s, 1);
      mbar_init(s_full + 8 * s, 1);
      mbar_init(dq_full + 8 * s, 1);
      mbar_init(ds_in + 8 * s, 1);
    }
    mbar_init(mma_done, 1);
    mbar_init(dp_read, 16);
    mbar_init(dp_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmQr); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDOr); tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // everything above is launch-independent set-up; global inputs may come from the previous kernel
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / multicast commit / 2-SM TMA
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  if (threadIdx.x == 32) B2_STAMP(0);   // set-up done (barriers, TMEM, PDL wait, cluster sync)
#ifdef VDS_B2_PROF
  if (threadIdx.x == 0 && p.dbg != nullptr) p.dbg[128 + 4 * (long long)blockIdx.x + 2] = gtimer();
#endif
  // TMEM columns (same in both CTAs): dV [0,128) | dK [128,256) | S^T buffers [256,320) [320,384) (afterwards, per half of
  // 32 query columns: bf16 P^T in 16 columns, dS^T in the next 16) | dP^T [384,448) | dQ^T [448,480): 128 lanes x 32
  // columns, lane % 64 = d of this CTA's half, lane / 64 = which 32 query columns
  const uint32_t tDV = tmem, tDK = tmem + 128, tSTb = tmem + 256, tDPTs = tmem + 384, tDQT = tmem + 448;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own operand halves)
    if (lane == 0) {
      auto lbar = [&](uint32_t bar) { return mapa_cta(bar, 0); };
      auto arm = [&](uint32_t bar, uint32_t bytes_both) {   // leader: expect both CTAs' bytes; follower: plain arrive
        if (leader) mbar_expect_tx(bar, bytes_both);
        else mbar_arrive_remote(lbar(bar));
      };
      arm(kv_full, 2 * 3 * TILE_BYTES);
      const uint32_t kvb = lbar(kv_full);
      tma_load_4d_2sm(sK, &tmK, kvb, 0, kv0, head, b);
      tma_load_4d_2sm(sK + HALF_BYTES, &tmK, kvb, 64, kv0, head, b);
      tma_load_4d_2sm(sV, &tmV, kvb, 0, kv0, head, b);
      tma_load_4d_2sm(sV + HALF_BYTES, &tmV, kvb, 64, kv0, head, b);
      tma_load_4d_2sm(sKT, &tmK, kvb, (int)crank * 64, kv0_pair, head, b);                 // K of tile 0, my d half
      tma_load_4d_2sm(sKT + HALF_BYTES, &tmK, kvb, (int)crank * 64, kv0_pair + 128, head, b);   // K of tile 1, my d half
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint32_t us = (i >> 1) & 1;
        const int q0 = (qt0 + i) * QSUB;
        {
          mbar_wait(rows_empty + 8 * st, us ^ 1u);
          arm(s_ready + 8 * st, 2 * B2_ROWS_STAGE);
          const uint32_t fb = lbar(s_ready + 8 * st);
          const uint32_t dq = sROWS + st * B2_ROWS_STAGE, dd = dq + 8192;
          const int row0 = q0 + (int)crank * 32;
          tma_load_4d_2sm(dq, &tmQr, fb, 0, row0, head, b);
          tma_load_4d_2sm(dq + 4096, &tmQr, fb, 64, row0, head, b);
          tma_load_4d_2sm(dd, &tmDOr, fb, 0, row0, head, b);
          tma_load_4d_2sm(dd + 4096, &tmDOr, fb, 64, row0, head, b);
        }
        {
          mbar_wait(cols_empty + 8 * st, us ^ 1u);
          arm(dvdk_ready + 8 * st, 2 * B2_COLS_STAGE);
          const uint32_t fb = lbar(dvdk_ready + 8 * st);
          const uint32_t dq = sCOLS + st * B2_COLS_STAGE, dd = dq + 8192;
          tma_load_4d_2sm(dq, &tmQ, fb, (int)crank * 64, q0, head, b);
          tma_load_4d_2sm(dd, &tmDO, fb, (int)crank * 64, q0, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      B2_PROF_DECL
      constexpr uint32_t idesc_s = umma_idesc_bf16(256, 64, false, false);     // S^T, dP^T
      constexpr uint32_t idesc_acc = umma_idesc_bf16(256, 128, false, true);   // dV, dK
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, true, true);      // dQ^T (64 rows of d per CTA)
      auto issue_s = [&](int k) {
        const int bb = k & 1;
        B2_WAIT(0, mbar_wait(s_ready + 8 * bb, (k >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q = sROWS + bb * B2_ROWS_STAGE;
          const uint32_t tST = tSTb + bb * 64;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma2_ss(tST, desc_kmajor(sK, kk), umma_smem_desc(q + (kk >> 2) * 4096 + (kk & 3) * 32, 16, 1024), idesc_s, kk > 0);
          mma2_ss(tST, desc_k16_noswz(sONES), desc_k16_noswz(sSTAT + (bb * 2 + 0) * 1024), idesc_s, 1);   // - lse
          commit2(s_full + 8 * bb);
        }
        __syncwarp();
      };
#ifdef VDS_B2_PROF
      auto issue_s_t = [&](int k) { const long long t_ = clock64(); issue_s(k); prof_acc[8] += clock64() - t_; };
#else
      auto issue_s_t = [&](int k) { issue_s(k); };
#endif
      auto issue_dp = [&](int k) {
        const int bb = k & 1;
        if (k > 0) B2_WAIT(2, mbar_wait(dp_read, (k - 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_o = sROWS + bb * B2_ROWS_STAGE + 8192;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma2_ss(tDPTs, desc_kmajor(sV, kk), umma_smem_desc(d_o + (kk >> 2) * 4096 + (kk & 3) * 32, 16, 1024), idesc_s,
                    kk > 0);
          mma2_ss(tDPTs, desc_k16_noswz(sONES), desc_k16_noswz(sSTAT + (bb * 2 + 1) * 1024), idesc_s, 1);   // - delta
          commit2(dp_full);
          commit2(rows_empty + 8 * bb);   // both users of this rows stage (S^T, dP^T) are complete
        }
        __syncwarp();
      };
#ifdef VDS_B2_PROF
      auto issue_dp_t = [&](int k) { const long long t_ = clock64(); issue_dp(k); prof_acc[9] += clock64() - t_; };
#else
      auto issue_dp_t = [&](int k) { issue_dp(k); };
#endif
      // dV(i) / dK(i): need P^T / dS^T of both CTAs in TMEM and the cols stage (dvdk_ready)
      auto issue_dvdk = [&](int i) {
        const int bb = i & 1;
        B2_WAIT(3, mbar_wait(dvdk_ready + 8 * bb, (i >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q = sCOLS + bb * B2_COLS_STAGE, d_o = q + 8192;
          const uint32_t tPT = tSTb + bb * 64;
#pragma unroll
          // A operands in the retired S^T columns: [P^T q 0..31 | dS^T q 0..31 | P^T q 32..63 | dS^T q 32..63] x 16 columns
          for (int kk = 0; kk < 4; ++kk)   // dV += P^T dO : B = this CTA's 64 d columns of dO, MN-major
            mma2_ts(tDV, tPT + (kk >> 1) * 32 + (kk & 1) * 8, umma_smem_desc(d_o + kk * 2048, 8192, 1024), idesc_acc,
                    (i > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dK += dS^T Q
            mma2_ts(tDK, tPT + (kk >> 1) * 32 + 16 + (kk & 1) * 8, umma_smem_desc(q + kk * 2048, 8192, 1024), idesc_acc,
                    (i > 0 || kk > 0));
          commit2(cols_empty + 8 * bb);
          if (i == n_q - 1) commit2(mma_done);
        }
        __syncwarp();
      };
      // dQ^T(j) over the 256 kv rows of the pair: needs both halves of the dS^T exchange of sub-tile j and the drain of j-1
      auto issue_dq = [&](int j) {
        const int set = j & 1;
        B2_WAIT(5, mbar_wait(dq_ready + 8 * set, (j >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          if (j + 2 < n_q) mbar_expect_tx(dq_ready + 8 * set, 8192);   // arm this set's next exchange (sub-tile j+2)
          const uint32_t ds = sDS + set * B2_DS_SET;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)   // tile kk >> 3, 16 kv rows per k-step
            mma2_ss(tDQT, umma_smem_desc(sKT + (kk >> 3) * HALF_BYTES + (kk & 7) * 2048, 4096, 1024),
                    desc_mn_sw64(ds + (kk >> 3) * 8192 + (kk & 7) * 1024), idesc_dq, kk > 0);
          commit2(dq_full + 8 * set);
        }
        __syncwarp();
      };
      // Program order == tensor-pipe order.  Per sub-tile i:  dP(i+1) | dQ(i-1) | dV(i) dK(i) | S(i+2): dV / dK only need
      // what the compute warps arrive with (TMEM), so they and the next S^T are not held up by the DSMEM exchange of
      // dS^T, whose dQ^T runs one sub-tile late out of double-buffered shared-memory tiles.
      if (elect_one()) {                               // exchanges of sub-tiles 0 and 1
        mbar_expect_tx(dq_ready, 8192);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 8;" ::"r"(dq_ready) : "memory");   // no dQ^T(-1) to drain
        if (n_q > 1) mbar_expect_tx(dq_ready + 8, 8192);
      }
      __syncwarp();
      mbar_wait(kv_full, 0);
      if (lane == 0) B2_STAMP(1);          // K / V tiles landed
      issue_s(0);
      issue_dp(0);
      if (n_q > 1) issue_s(1);
      for (int i = 0; i < n_q; ++i) {
        if (i + 1 < n_q) issue_dp_t(i + 1);   // waits for dp_read(i): the middle of compute(i)
#ifdef VDS_B2_PROF
        { const long long t_ = clock64(); if (i > 0) issue_dq(i - 1); prof_acc[10] += clock64() - t_; }
        { const long long t_ = clock64(); issue_dvdk(i); prof_acc[11] += clock64() - t_; }
#else
        if (i > 0) issue_dq(i - 1);          // its exchange completes around the same time
        issue_dvdk(i);                       // end of compute(i)
#endif
        if (i + 2 < n_q) issue_s_t(i + 2);
      }
      if (lane == 0) B2_STAMP(2);          // last dV / dK issued
      issue_dq(n_q - 1);
      if (lane == 0) B2_STAMP(3);          // last dQ^T issued
      B2_PROF_STORE(0);
    } else {
      // follower: relay "the leader's dS^T half has landed in my sDS slot 0" to the leader's issuer
      const uint32_t l_dq_ready = mapa_cta(dq_ready, 0);
      for (int i = 0; i < n_q; ++i) {
        if (lane == 0) {
          const uint32_t set = i & 1;
          mbar_expect_tx(ds_in + 8 * set, 8192);
          mbar_wait(ds_in + 8 * set, (i >> 1) & 1);
          mbar_arrive_remote(l_dq_ready + 8 * set);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ compute warps (thread == kv row x half the q columns)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;              // kv row of the tile == TMEM lane
    const uint32_t half = (warp - 4) >> 2;       // 0: query columns 0..31 of the sub-tile, 1: columns 32..63
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool kv_ok = (kv0 + r) < p.Lk;
    const bool kv_full_tile = kv0 + 128 <= p.Lk;
    const long long stat_base = ((long long)b * p.nh + head) * p.Lq;
    // statistics tile of this CTA: its 32 query rows of the sub-tile; warp 4 writes -lse/scale, warp 8 -delta
    const bool stat_thread = quad == 0;
    const int which = (int)half, srow = lane;
    const float* stat_src = (which == 0 ? p.lse : p.delta) + stat_base;
    const float inv_sl2 = 1.0f / p.scale_log2;
    const uint32_t l_s_ready = mapa_cta(s_ready, 0), l_dp_read = mapa_cta(dp_read, 0), l_dvdk_ready = mapa_cta(dvdk_ready, 0);
    // My 32 query columns are the B-tile half of CTA `half`: mine -> straight into my sDS; the peer's -> send buffer, then one
    // bulk DSMEM copy per warp.  The bytes I send complete on the peer's barrier: the leader's dq_ready / the follower's ds_in.
    const bool mine = half == crank;
    const uint32_t peer_ds = mapa_cta(sDS, crank ^ 1u), peer_ds_in0 = mapa_cta(leader ? ds_in : dq_ready, crank ^ 1u);
    auto stat_fetch = [&](int k) -> float {
      float raw = 0.f;
      if (stat_thread) {
        const int q = min((qt0 + k) * QSUB + (int)crank * 32 + srow, p.Lq - 1);
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(raw) : "l"(stat_src + q));
      }
      return raw;
    };
    auto stat_finish = [&](int k, float raw) -> float {
      const int q = (qt0 + k) * QSUB + (int)crank * 32 + srow;
      if (q >= p.Lq) return which == 0 ? -INFINITY : 0.f;        // padded query row: exp2(-inf) = 0
      return which == 0 ? -raw * inv_sl2 : -raw;
    };
    auto stat_write = [&](int k, float v) {
      if (stat_thread) {
        uint8_t* tile = gSTAT + ((k & 1) * 2 + which) * 1024;
        *reinterpret_cast<uint4*>(tile + k16_off(srow)) = split3_bf16(v);
        *reinterpret_cast<uint4*>(tile + k16_off(srow) + 128) = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto arrive_leader = [&](uint32_t cluster_bar) {   // one arrival per warp
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(cluster_bar);
    };
    if (half == 0) {  // constant A operand of the statistics k-step: ones in columns 0..2 of every kv row
      const float one = 1.0f;
      *reinterpret_cast<uint4*>(gONES + k16_off(r)) = make_uint4(pack_bf16x2(one, one), pack_bf16x2(one, 0.f), 0u, 0u);
      *reinterpret_cast<uint4*>(gONES + k16_off(r) + 128) = make_uint4(0u, 0u, 0u, 0u);
    }
    B2_PROF_DECL
    for (int k = 0; k < 2 && k < n_q; ++k) {
      stat_write(k, stat_finish(k, stat_fetch(k)));
      fence_proxy_async_smem();
      arrive_leader(l_s_ready + 8 * (k & 1));
    }
    for (int i = 0; i < n_q; ++i) {
      const int bb = i & 1;
      const float next_raw = (i + 2 < n_q) ? stat_fetch(i + 2) : 0.f;
      B2_WAIT(0, mbar_wait(s_full + 8 * bb, (i >> 1) & 1));
      tc_fence_after();
      const uint32_t tST = tSTb + bb * 64 + lane_off, tDPT = tDPTs + lane_off;
      // phase 1 (overlaps the MMAs of the previous sub-tile): p = exp2(s'), P^T -> TMEM
#ifdef VDS_B2_PROF
      const long long tp1 = clock64();
#endif
      float pf[32];
      {
        uint32_t sv[32];
        tmem_ld32(tST + half * 32, sv);
        tmem_ld_wait();
        const float2 sl2 = make_float2(p.scale_log2, p.scale_log2);
        // half of the exponentials on the MUFU pipe (ex2.approx), half as a polynomial on the FMA / ALU pipes: the 8192
        // exp2 per sub-tile are 512 MUFU cycles per SM sub-partition otherwise
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float2 a0 = mul2(make_float2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2);
          const float2 a1 = mul2(make_float2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), sl2);
          const float2 pa1 = ex2_poly2(a1);
          pf[e] = ex2(a0.x); pf[e + 1] = ex2(a0.y); pf[e + 2] = pa1.x; pf[e + 3] = pa1.y;
        }
        if (!kv_full_tile && !kv_ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) pf[e] = 0.f;
        }
      }
      {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(pf[2 * e], pf[2 * e + 1]);
        // bf16 P^T of my 32 query columns (pair (2c, 2c+1) per column) over the FIRST 16 of the 32 S^T columns I have
        // just read; dS^T goes over the other 16.  A warp never writes columns the other half's warp still has to read.
        tmem_st16(tST + half * 32, pk);
      }
      // phase 2: dS^T = P^T o dP'^T once dP^T(i) has landed
#ifdef VDS_B2_PROF
      prof_acc[4] += clock64() - tp1;
#endif
      B2_WAIT(1, mbar_wait(dp_full, i & 1));
      tc_fence_after();
      uint32_t dd[16];
      {
        uint32_t dv[32];
        tmem_ld32(tDPT + half * 32, dv);
        tmem_ld_wait();
        tc_fence_before();      // dP^T(i) is in registers: the issuer may refill the buffer with dP^T(i+1)
        arrive_leader(l_dp_read);
#pragma unroll
        for (int e = 0; e < 32; e += 2) {   // dS^T WITHOUT the softmax scale: the dQ drain and the dK epilogue apply it
          const float2 d = mul2(make_float2(pf[e], pf[e + 1]), make_float2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])));
          dd[e >> 1] = pack_bf16x2(d.x, d.y);
        }
      }
      tmem_st16(tST + half * 32 + 16, dd);
#ifdef VDS_B2_PROF
      const long long tp3 = clock64();
#endif
      const uint32_t set = i & 1;
      if (i >= 2) B2_WAIT(2, mbar_wait(dq_full + 8 * set, ((i - 2) >> 1) & 1));   // dQ^T(i-2) has consumed this set (and its copies)
      {
        // dS^T of my 32 query columns: 64 bytes of row r in the 64-byte-swizzle tile image of kv tile `crank`
        uint8_t* dst = mine ? gDS + set * B2_DS_SET + crank * 8192u : gSEND + set * 8192u;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(dst + sw64_offset(r, c)) = make_uint4(dd[c * 4], dd[c * 4 + 1], dd[c * 4 + 2], dd[c * 4 + 3]);
      }
      // statistics tile of sub-tile i+2 (buffer bb: S^T(i) and dP^T(i), its readers, are complete) shares the fence
      if (i + 2 < n_q) stat_write(i + 2, stat_finish(i + 2, next_raw));
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_remote(l_dvdk_ready + 8 * bb);             // P^T / dS^T in TMEM: dV(i), dK(i) may go
        if (!mine) {                                           // this warp's 32 rows (2 KiB) of the send buffer
          const uint32_t wo = set * 8192u + quad * 2048u;
          dsmem_bulk_copy(peer_ds + set * B2_DS_SET + crank * 8192u + quad * 2048u, sSEND + wo, 2048, peer_ds_in0 + 8 * set);
        }
        if (i + 2 < n_q) mbar_arrive_remote(l_s_ready + 8 * bb);
      }
      __syncwarp();
#ifdef VDS_B2_PROF
      prof_acc[6] += clock64() - tp3;
#endif
    }
    if (warp == 4) B2_PROF_STORE(1);
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ dQ drain warpgroup
    B2_PROF_DECL
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int d_local = (quad & 1) * 32 + lane;   // TMEM lane % 64
    const int qh = quad >> 1;                      // TMEM lane / 64: query columns qh*32 ..
    const bool lead_thread = threadIdx.x == 384;
    const uint32_t l_dq_ready = mapa_cta(dq_ready, 0);
    for (int i = 0; i < n_q; ++i) {
      B2_WAIT(0, mbar_wait(dq_full + 8 * (i & 1), (i >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tDQT + lane_off, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && i + 1 < n_q) mbar_arrive_remote(l_dq_ready + 8 * ((i + 1) & 1));   // dQ^T(i+1) may overwrite the columns
      // the 64 (q) x 64 (d) block leaves in two halves through one 8 KiB staging tile: warps 0,1 hold q 0..31, warps 2,3 q 32..63
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        if (lead_thread) bulk_wait_group_read0();      // previous reduction has finished reading the staging tile
        named_bar_sync(2, 128);
        if (qh == pass) {
#pragma unroll
          for (int c = 0; c < 32; ++c) gSTG[c * 64 + d_local] = __uint_as_float(v[c]) * p.scale;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (lead_thread) {
          tma_reduce_add_4d(&tmDQ, sSTG, (int)crank * 64, (qt0 + i) * QSUB + pass * 32, head, b);
          bulk_commit_group();
        }
      }
    }
    if (lead_thread) { B2_STAMP(6); bulk_wait_group0(); B2_STAMP(7); }   // last reduction issued / complete
    if (warp == 12) B2_PROF_STORE(2);
  }
  if (warp >= 4 && warp < 12) {
    // dK (warps 4-7) / dV (warps 8-11) of this CTA's own kv rows: TMEM lane == kv row
    const int which = warp >= 8 ? 1 : 0;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    if (threadIdx.x == 128) B2_STAMP(8);   // compute warps left the loop
    mbar_wait(mma_done, 0);
    if (threadIdx.x == 128) B2_STAMP(4);   // dV / dK accumulators complete
    tc_fence_after();
    const int krow = kv0 + r;
    const uint32_t t = (which == 0 ? tDK : tDV) + lane_off;
    const float osc = which == 0 ? p.scale : 1.0f;   // dK was accumulated from the unscaled dS^T
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t + c * 32, v);
      tmem_ld_wait();
      if (split_mode && pp.tma_dkdv) {
        // partial sums over this piece's query range, added into the compact fp32 workspace by the TMA unit (fixed up to
        // bf16 later): the 128 x 128 fp32 tile is staged as four [128 rows x 32 floats] SW128 boxes — dK over the K / V operand
        // tiles, dV over the Q / dO stages, all free once mma_done has fired (K^T and the dS^T sets are still read by the last
        // dQ^T).  Rows of kv positions past Lk are exact zeros (P^T and dS^T are masked), so the whole tile may be added.
        uint8_t* box = gen + (which == 0 ? B2_OFF_K : B2_OFF_ROWS) + c * 16384;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(box + sw128_offset(r, g)) =
              make_float4(__uint_as_float(v[g * 4]) * osc, __uint_as_float(v[g * 4 + 1]) * osc,
                          __uint_as_float(v[g * 4 + 2]) * osc, __uint_as_float(v[g * 4 + 3]) * osc);
      } else if (split_mode) {   // (workspace not TMA-able) per-thread fp32 red
        if (krow < p.Lk) {
          float* dst = pp.compact + (((long long)pair_local * 2 + crank) * 2 + which) * (128 * HD) + r * HD + c * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4),
                         "f"(__uint_as_float(v[g * 4]) * osc), "f"(__uint_as_float(v[g * 4 + 1]) * osc),
                         "f"(__uint_as_float(v[g * 4 + 2]) * osc), "f"(__uint_as_float(v[g * 4 + 3]) * osc)
                         : "memory");
        }
      } else if (pp.tma_dkdv) {
        // the tile leaves through the K (dK) / V (dV) operand tile — every MMA that read it is complete (mma_done) — in the
        // same two-halves SW128 image, then one TMA store per half: per-thread row stores are 32 partial sectors per
        // warp instruction (measured: 6400 cycles for the 64 KiB of one CTA)
        uint8_t* tile = gen + (which == 0 ? B2_OFF_K : B2_OFF_V);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * osc, __uint_as_float(v[g * 8 + 1]) * osc);
          u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * osc, __uint_as_float(v[g * 8 + 3]) * osc);
          u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * osc, __uint_as_float(v[g * 8 + 5]) * osc);
          u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * osc, __uint_as_float(v[g * 8 + 7]) * osc);
          st_tile8(tile, r, c * 32 + g * 8, u);
        }
      } else if (krow < p.Lk) {
        bf16* dst = (which == 0 ? p.dk + ((long long)b * p.Lk + krow) * p.lddk
                                : p.dv + ((long long)b * p.Lk + krow) * p.lddv) + head * HD + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * osc, __uint_as_float(v[g * 8 + 1]) * osc);
          u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * osc, __uint_as_float(v[g * 8 + 3]) * osc);
          u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * osc, __uint_as_float(v[g * 8 + 5]) * osc);
          u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * osc, __uint_as_float(v[g * 8 + 7]) * osc);
          *reinterpret_cast<uint4*>(dst + g * 8) = u;
        }
      }
    }
    if (split_mode && pp.tma_dkdv) {
      fence_proxy_async_smem();
      named_bar_sync(3 + which, 128);
      if (quad == 0 && lane == 0 && kv0 < p.Lk) {   // a phantom tile adds nothing
        const uint32_t tile = which == 0 ? sK : sROWS;
        const int row0 = ((pair_local * 2 + (int)crank) * 2 + which) * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_reduce_add_2d(&tmDK, tile + c * 16384, c * 32, row0);
        bulk_commit_group();
        bulk_wait_group0();
      }
    }
    if (!split_mode && pp.tma_dkdv) {
      fence_proxy_async_smem();
      named_bar_sync(3 + which, 128);
      if (quad == 0 && lane == 0 && kv0 < p.Lk) {   // rows past Lk are clipped by the tensor map; a phantom tile stores nothing
        const uint32_t tile = which == 0 ? sK : sV;
        const CUtensorMap* tm = which == 0 ? &tmDK : &tmDV;
        tma_store_4d(tm, tile, 0, kv0, head, b);
        tma_store_4d(tm, tile + HALF_BYTES, 64, kv0, head, b);
        bulk_commit_group();
        bulk_wait_group0();
      }
    }
  }
  if (threadIdx.x == 128) B2_STAMP(5);     // dK epilogue of warp 4 done
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 128) B2_STAMP(9);
  cluster_sync_all();   // nobody leaves while the peer may still store into / arrive on / multicast to this CTA
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  if (threadIdx.x == 32) B2_STAMP(10);
#ifdef VDS_B2_PROF
  if (threadIdx.x == 32 && p.dbg != nullptr) p.dbg[128 + 4 * (long long)blockIdx.x + 3] = gtimer();
#endif
}

// Launches the pair kernel on pairs [pair_base, pair_base + n_pairs) of the (b, head, kv-tile-pair) space: one cluster per
// pair (n_pieces == 0) or one cluster per piece of the tail plan (pieces of those pairs along the query range).
int launch_attn_bwd_pairs(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                          int64_t lddo, const AttnBwdParams& p0, int B, int nh, int Lq, int Lk, int pair_base, int n_pairs,
                          int pairs_per_bh, const uint32_t* pieces, int n_pieces, float* compact, cudaStream_t stream) {
  CUtensorMap tq, tqr, tk, tv, tdo, tdor, tdq, tdk, tdv;
  int r;
  if ((r = make_tmap_tokens(&tq, q, ldq, Lq, nh, B, QSUB))) return r;
  if ((r = make_tmap_tokens(&tqr, q, ldq, Lq, nh, B, 32))) return r;
  if ((r = make_tmap_tokens(&tk, k, ldk, Lk, nh, B))) return r;
  if ((r = make_tmap_tokens(&tv, v, ldv, Lk, nh, B))) return r;
  if ((r = make_tmap_tokens(&tdo, d_o, lddo, Lq, nh, B, QSUB))) return r;
  if ((r = make_tmap_tokens(&tdor, d_o, lddo, Lq, nh, B, 32))) return r;
  if ((r = make_tmap_dq(&tdq, p0.dq_acc, p0.lddq, Lq, nh, B, 64, 32))) return r;
  // bf16 dK / dV leave by TMA store when the buffers qualify (16-byte aligned base and row pitch), else by per-thread stores
  static const bool tma_env = getenv("VDS_BWD2_TMA_OUT") == nullptr || strcmp(getenv("VDS_BWD2_TMA_OUT"), "0") != 0;   // tuning switch
  const bool tma_out = tma_env && n_pieces == 0 && p0.dk != nullptr && p0.dv != nullptr && ((uintptr_t)p0.dk & 15) == 0 &&
                       ((uintptr_t)p0.dv & 15) == 0 && p0.lddk % 8 == 0 && p0.lddv % 8 == 0;
  tdk = tk; tdv = tv;
  bool tma_compact = false;
  if (tma_env && n_pieces > 0 && compact != nullptr && ((uintptr_t)compact & 15) == 0) {
    // compact workspace [n_pairs * 2 tiles][dk | dv][128][128] fp32 as one 2-D matrix of 128 columns: box = 128 rows x 32 floats
    uint64_t cdims[2] = {(uint64_t)HD, (uint64_t)n_pairs * 2 * 2 * 128}, cstr[1] = {(uint64_t)HD * 4};
    uint32_t cbox[2] = {32, 128};
    if ((r = encode_tmap(&tdk, compact, 1, 2, cdims, cstr, cbox, 1))) return r;
    tma_compact = true;
  }
  if (tma_out) {
    if ((r = make_tmap_tokens(&tdk, p0.dk, p0.lddk, Lk, nh, B))) return r;
    if ((r = make_tmap_tokens(&tdv, p0.dv, p0.lddv, Lk, nh, B))) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e != cudaSuccess) { set_error("attn_bwd2: smem attribute: %s", cudaGetErrorString(e)); return VDS_ERR_CUDA; }
    attr = true;
  }
  VDS_CHECK_ARG(n_pieces >= 0 && n_pieces <= VDS_BWD2_MAX_PIECES, "attn_bwd2: %d pieces", n_pieces);
  AttnBwd2Params pp;
  pp.p = p0;
  pp.pair_base = pair_base;
  pp.pairs_per_bh = pairs_per_bh;
  pp.n_pieces = n_pieces;
  pp.tma_dkdv = (tma_out || tma_compact) ? 1 : 0;
  pp.compact = compact;
  for (int i = 0; i < n_pieces; ++i) pp.pieces[i] = pieces[i];
  for (int i = n_pieces; i < VDS_BWD2_MAX_PIECES; ++i) pp.pieces[i] = 0u;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (n_pieces > 0 ? n_pieces : n_pairs));
  cfg.blockDim = dim3(B2_THREADS);
  cfg.dynamicSmemBytes = B2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see common.h: launch_k
  attrs[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, attn_bwd2_kernel, tq, tqr, tk, tv, tdo, tdor, tdq, tdk, tdv, pp);
  if (e != cudaSuccess) {
    set_error("attn_bwd2: cluster launch failed: %s", cudaGetErrorString(e));
    return VDS_ERR_CUDA;
  }
  VDS_CHECK_LAUNCH("attn_bwd2");
  return VDS_OK;
}

}  // namespace vds
