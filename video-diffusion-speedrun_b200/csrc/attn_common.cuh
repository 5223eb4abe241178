// Shared by attention.cu (forward, 1-CTA backward) and attention_bwd2.cu (2-CTA backward): tile geometry, UMMA
// descriptor helpers, the backward parameter block and the tensor maps over token-major [B, L, ld] buffers.
#pragma once
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace vds {

constexpr int QSUB = 64;   // query rows per backward sub-tile
constexpr int HD = 128;
constexpr int TILE_BYTES = 128 * HD * 2;  // 32 KiB: one 128 x 128 bf16 tile = two 64-wide SW128 halves
constexpr int HALF_BYTES = TILE_BYTES / 2;
#define VDS_BWD2_MAX_PIECES 256   // pieces of one tail plan of the CTA-pair backward (kernel-parameter table)

// K-major operand tile (rows x 128 along the contraction): descriptor of 16-wide k-step kk
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) {
  return umma_smem_desc(tile + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
}
// MN-major operand tile (contraction index = smem row): 16 rows per k-step
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kk) {
  return umma_smem_desc(tile + kk * 2048, HALF_BYTES, 1024);
}

__device__ __forceinline__ void load_tile_4d(uint32_t dst, const void* tmap, uint32_t bar, int row0, int head, int b) {
  tma_load_4d(dst, tmap, bar, 0, row0, head, b);
  tma_load_4d(dst + HALF_BYTES, tmap, bar, 64, row0, head, b);
}

// write 8 consecutive bf16 (columns c0..c0+7, c0 % 8 == 0) of row r into a K-major SW128 tile
__device__ __forceinline__ void st_tile8(uint8_t* tile, int r, int c0, uint4 v) {
  *reinterpret_cast<uint4*>(tile + (c0 >> 6) * HALF_BYTES + sw128_offset(r, (c0 & 63) >> 3)) = v;
}


// K-major operand WITHOUT swizzle, 16 columns wide (one k-step): 8-row x 16-byte core matrices, the two 8-column
// halves 128 B apart (LBO), consecutive 8-row groups 256 B apart (SBO).
__device__ __forceinline__ uint64_t desc_k16_noswz(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__device__ __forceinline__ uint32_t k16_off(int row) { return (row >> 3) * 256 + (row & 7) * 16; }
// x = hi + mid + lo with three bf16 terms (fp32-exact to ~2^-24 relative): the per-query statistics ride through
// the tensor core as an extra rank-3 update instead of being re-read from shared memory for every element.
__device__ __forceinline__ uint4 split3_bf16(float x) {
  if (!(fabsf(x) < 3.0e38f)) return make_uint4(pack_bf16x2(x, 0.f), 0u, 0u, 0u);   // +-inf (masked query rows)
  const float hi = bf16_round(x);
  const float mid = bf16_round(x - hi);
  const float lo = bf16_round(x - hi - mid);
  return make_uint4(pack_bf16x2(hi, mid), pack_bf16x2(lo, 0.f), 0u, 0u);
}

struct AttnBwdParams {
  const float* lse; const float* delta;   // [B, nh, Lq]
  float* dq_acc; long long lddq;           // fp32 [B, Lq, lddq] (+=)
  bf16* dk; long long lddk;                // bf16 [B, Lk, lddk], head at head*128 (q_splits == 1)
  bf16* dv; long long lddv;
  float* dk_acc; float* dv_acc; long long ldkv_acc;  // fp32 accumulation targets when q_splits > 1
  float* compact_acc;   // q_splits > 1: [item - item_base][dk|dv][128][128] fp32 (tail balancing), overrides dk_acc/dv_acc
  int item_base, kv_tiles;
  int Lq, Lk, nh, q_splits;
  float scale_log2, scale;
  long long* dbg;   // optional per-iteration clock64 trace of CTA (0,0,0): [iter][8] (debug / tuning only)
  // Remainder mode (rem_pair_base >= 0): this launch covers what the CTA-pair kernel (attention_bwd2.cu) left over —
  // local items [0, rem_pair_tiles) are the two tiles of pairs rem_pair_base, rem_pair_base + 1, ... (pair -> (b*nh + head,
  // kv tiles 2j, 2j+1), rem_pairs_per_bh pairs per (b, head)), the items after that are the unpaired last tile
  // (kv_tiles odd) of (b, head) = 0, 1, ...
  int rem_pair_base, rem_pairs_per_bh, rem_pair_tiles;
};
// local item of a 1-CTA backward launch -> (b, head, kv tile)
__device__ __forceinline__ void attn_bwd_decode_item(const AttnBwdParams& p, int local, int& b, int& head, int& kv_tile) {
  if (p.rem_pair_base >= 0) {
    int bh;
    if (local < p.rem_pair_tiles) {
      const int pid = p.rem_pair_base + (local >> 1);
      bh = pid / p.rem_pairs_per_bh;
      kv_tile = 2 * (pid % p.rem_pairs_per_bh) + (local & 1);
    } else {
      bh = local - p.rem_pair_tiles;
      kv_tile = p.kv_tiles - 1;
    }
    head = bh % p.nh;
    b = bh / p.nh;
  } else {
    const int item = p.item_base + local;
    kv_tile = item % p.kv_tiles;
    head = (item / p.kv_tiles) % p.nh;
    b = item / (p.kv_tiles * p.nh);
  }
}
#define VDS_TRACE(slot, it)                                                                      \
  do {                                                                                           \
    if (p.dbg != nullptr && blockIdx.x == 0 && (it) < 160)                                       \
      p.dbg[(it) * 8 + (slot)] = clock64();                                                      \
  } while (0)


static inline int make_tmap_tokens(CUtensorMap* tm, const void* ptr, long long ld, int L, int nh, int B, int box_rows = 128) {
  uint64_t dims[4] = {(uint64_t)HD, (uint64_t)L, (uint64_t)nh, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)HD * 2, (uint64_t)L * (uint64_t)ld * 2};
  uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return encode_tmap_bf16(tm, ptr, 4, dims, strides, box);
}
// fp32 [B, L, ld] accumulation buffer, box = one [64 rows x box_d floats] sub-tile of one head, no swizzle
static inline int make_tmap_dq(CUtensorMap* tm, const float* ptr, long long ld, int L, int nh, int B, int box_d = 128,
                               int box_rows = QSUB) {
  uint64_t dims[4] = {(uint64_t)HD, (uint64_t)L, (uint64_t)nh, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ld * 4, (uint64_t)HD * 4, (uint64_t)L * (uint64_t)ld * 4};
  uint32_t box[4] = {(uint32_t)box_d, (uint32_t)box_rows, 1, 1};
  return encode_tmap(tm, ptr, 1, 4, dims, strides, box, 0);
}


}  // namespace vds
