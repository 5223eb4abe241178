// Shared by gemm.cu (1-CTA tiles) and gemm2.cu (2-CTA cta_group::2 tiles): device-side argument block and the
// fused epilogues (thread == accumulator row, coalesced through a per-warp swizzled smem transpose).
#pragma once
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "ptx.cuh"

namespace vds {

constexpr int BM = 128;
constexpr int BK = 64;

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
  const bf16* v0; long long ldv0; const bf16* lambda;       // QKV_ROPE value residual (v0 == nullptr: none)
  long long* dbg;   // tuning aid (nullptr in production)
};

// GELU(erf) and its derivative.  Phi(x) = 0.5 (1 + erf(x / sqrt2)) is evaluated as a logistic of an odd polynomial,
//   Phi(x) ~= 1 / (1 + exp(-x p(x^2))),   p of degree 4 in x^2 (coefficients below, fitted minimax on |x| <= 5.6,
//   p > 0 everywhere so the tails saturate to exactly 0 / 1),
// with |gelu err| <= 3.7e-6 and |gelu' err| <= 1.4e-5 in fp32 — 2-3 orders below the bf16 rounding of the outputs
// (fit + fp32 emulation against scipy erf: tests/test_oracle_cpu.py::test_gelu_logistic_fit).  Costs 6 issue slots
// per element on the packed fp32x2 pipe (x^2, 4 FMA, x*q, ex2, 1+e, rcp, x*Phi) against 11 for the
// Abramowitz-Stegun erfc form used before (|err| 1.5e-7) and ~45 for libm erff; the epilogue warps of the K = 512
// MLP GEMMs are issue-bound, so this is what the GEMM's duration follows.  gelu' uses the analytic derivative of the
// same logistic, Phi' = Phi (1 - Phi) (x p(x^2))', so forward and backward stay consistent and need no second ex2.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// q_k = -log2(e) * c_k (exponent of 2^t, t = x q(x^2)) and w_k = (2k + 1) c_k (derivative of x p(x^2))
#define VDS_GELU_Q0 -2.3020565509796143f
#define VDS_GELU_Q1 -0.10519713163375854f
#define VDS_GELU_Q2 0.00034468894591555f
#define VDS_GELU_Q3 9.080363815883175e-05f
#define VDS_GELU_Q4 -3.351275836394052e-06f
#define VDS_GELU_W0 1.5956640243530273f
#define VDS_GELU_W1 0.21875128149986267f
#define VDS_GELU_W2 -0.0011946008307859302f
#define VDS_GELU_W3 -0.00044058202183805406f
#define VDS_GELU_W4 2.0906347344862297e-05f
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// Phi for two elements; s = x^2 (clamped by the caller where the derivative polynomial is also evaluated)
__device__ __forceinline__ float2 phi_logistic2(float2 x, float2 s) {
  float2 q = fma2(splat2(VDS_GELU_Q4), s, splat2(VDS_GELU_Q3));
  q = fma2(q, s, splat2(VDS_GELU_Q2));
  q = fma2(q, s, splat2(VDS_GELU_Q1));
  q = fma2(q, s, splat2(VDS_GELU_Q0));
  const float2 t = mul2(x, q);
  const float2 d = add2(make_float2(ex2(t.x), ex2(t.y)), splat2(1.0f));   // x -> -inf: 2^t = inf, Phi = 0
  return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}
__device__ __forceinline__ float2 gelu_erf2(float2 x) { return mul2(x, phi_logistic2(x, mul2(x, x))); }
__device__ __forceinline__ float2 dgelu_erf2(float2 x) {
  float2 s = mul2(x, x);
  s = make_float2(fminf(s.x, 64.0f), fminf(s.y, 64.0f));   // keeps x * w finite for absurd |x| (Phi (1 - Phi) is 0 there)
  const float2 r = phi_logistic2(x, s);
  float2 w = fma2(splat2(VDS_GELU_W4), s, splat2(VDS_GELU_W3));
  w = fma2(w, s, splat2(VDS_GELU_W2));
  w = fma2(w, s, splat2(VDS_GELU_W1));
  w = fma2(w, s, splat2(VDS_GELU_W0));
  const float2 omr = fma2(splat2(-1.0f), r, splat2(1.0f));
  return fma2(mul2(x, w), mul2(r, omr), r);
}
// Eight elements at a time, written in phases (all polynomials, all ex2, all 1+e, all rcp): the MUFU results are
// consumed ~8 issue slots after they were requested, so one warp hides the MUFU latency by itself (pair-at-a-time
// code stalled twice per pair, which made the epilogue latency- instead of issue-bound).
__device__ __forceinline__ void phi_logistic8(const float2 (&x)[4], const float2 (&s)[4], float2 (&r)[4]) {
  float2 t[4], d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 q = fma2(splat2(VDS_GELU_Q4), s[k], splat2(VDS_GELU_Q3));
    q = fma2(q, s[k], splat2(VDS_GELU_Q2));
    q = fma2(q, s[k], splat2(VDS_GELU_Q1));
    q = fma2(q, s[k], splat2(VDS_GELU_Q0));
    t[k] = mul2(x[k], q);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = make_float2(ex2(t[k].x), ex2(t[k].y));
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = add2(t[k], splat2(1.0f));
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k] = make_float2(rcp_approx(d[k].x), rcp_approx(d[k].y));
}
// v[0..7] <- gelu(v)
__device__ __forceinline__ void gelu_erf8(float (&v)[8]) {
  float2 x[4], s[4], r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { x[k] = make_float2(v[2 * k], v[2 * k + 1]); s[k] = mul2(x[k], x[k]); }
  phi_logistic8(x, s, r);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 g = mul2(x[k], r[k]); v[2 * k] = g.x; v[2 * k + 1] = g.y; }
}
// a[0..7] <- a * gelu'(h)
__device__ __forceinline__ void dgelu_mul8(float (&a)[8], const float (&h)[8]) {
  float2 x[4], s[4], r[4], w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x[k] = make_float2(h[2 * k], h[2 * k + 1]);
    const float2 sq = mul2(x[k], x[k]);
    s[k] = make_float2(fminf(sq.x, 64.0f), fminf(sq.y, 64.0f));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 ww = fma2(splat2(VDS_GELU_W4), s[k], splat2(VDS_GELU_W3));
    ww = fma2(ww, s[k], splat2(VDS_GELU_W2));
    ww = fma2(ww, s[k], splat2(VDS_GELU_W1));
    ww = fma2(ww, s[k], splat2(VDS_GELU_W0));
    w[k] = mul2(x[k], ww);
  }
  phi_logistic8(x, s, r);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 omr = fma2(splat2(-1.0f), r[k], splat2(1.0f));
    const float2 dg = fma2(w[k], mul2(r[k], omr), r[k]);
    const float2 o = mul2(make_float2(a[2 * k], a[2 * k + 1]), dg);
    a[2 * k] = o.x; a[2 * k + 1] = o.y;
  }
}
__device__ __forceinline__ float gelu_erf(float x) { return gelu_erf2(make_float2(x, x)).x; }
__device__ __forceinline__ float dgelu_erf(float x) { return dgelu_erf2(make_float2(x, x)).x; }

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
// tmC / tmC2 != nullptr: the staged 32 x 64 sub-tile leaves through the TMA unit (cp.async.bulk.tensor store, the
// staging buffer's XOR swizzle IS the 128-byte TMA swizzle) instead of 8 LDS + STG per thread: per-thread global
// stores sustain only ~32 B/clk/SM, which made the output write, not the MMAs, the limit of the K = 512 GEMMs.
template <int EPI>
__device__ __forceinline__ void epilogue_group64(const GemmDev& p, uint8_t* stg, int row0, int col0,
                                                 const uint32_t (&r0)[32], const uint32_t (&r1)[32], int lane,
                                                 const void* tmC = nullptr, const void* tmC2 = nullptr) {
  RowMap rm{row0, p.M, p.remap_rows, p.remap_stride, p.remap_offset};
  const int my_row = row0 + lane;
  const bool trace = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
  const long long tq0 = clock64();
  if constexpr (EPI == VDS_EPI_GATE_RES || EPI == VDS_EPI_DGELU) {
    stg_load(stg, p.aux, p.ldaux, rm, col0, p.N, lane);
    __syncwarp();
  }
  uint4 keep[8];  // first output (bf16 Linear result) kept packed while the second goes through the buffer
  const int b = (EPI == VDS_EPI_GATE_RES) ? min(my_row, p.M - 1) / p.rows_per_batch : 0;
  // per-column operands (bias, gate) for all 8 chunks are requested up front: one L2 round trip per group instead
  // of one per chunk (the serialized loads were the longest latency chain of the epilogue)
  uint4 braw[8], graw[8];
  const bool has_bias = (EPI != VDS_EPI_DGELU) && p.bias != nullptr;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int col = col0 + g * 8;
    braw[g] = make_uint4(0u, 0u, 0u, 0u);
    graw[g] = make_uint4(0u, 0u, 0u, 0u);
    if (col < p.N) {
      if (has_bias) braw[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
      if constexpr (EPI == VDS_EPI_GATE_RES)
        graw[g] = __ldg(reinterpret_cast<const uint4*>(p.gate + (long long)b * p.gate_stride + col));
    }
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __uint_as_float(g < 4 ? r0[g * 8 + j] : r1[(g - 4) * 8 + j]);
    if (has_bias) {
      float bb[8];
      unpack8(braw[g], bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += bb[j];
    }
    uint8_t* slot = stg + stg_off(lane, g);
    if constexpr (EPI == VDS_EPI_STORE) {
      *reinterpret_cast<uint4*>(slot) = pack8(acc);
    } else if constexpr (EPI == VDS_EPI_BIAS_GELU) {
      float act[8];
      keep[g] = pack8(acc);                               // bf16 Linear output (what the reference's GELU sees)
      unpack8(keep[g], acc);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float2 a2 = gelu_erf2(make_float2(acc[j], acc[j + 1]));
        act[j] = a2.x;
        act[j + 1] = a2.y;
      }
      *reinterpret_cast<uint4*>(slot) = pack8(act);
    } else if constexpr (EPI == VDS_EPI_GATE_RES) {
      float g8[8], x8[8], o8[8];
      unpack8(graw[g], g8);
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
      for (int j = 0; j < 8; j += 2) {
        const float2 d2 = mul2(make_float2(acc[j], acc[j + 1]), dgelu_erf2(make_float2(h8[j], h8[j + 1])));
        acc[j] = d2.x;
        acc[j + 1] = d2.y;
      }
      *reinterpret_cast<uint4*>(slot) = pack8(acc);
    }
  }
  const long long tq1 = clock64();
  const bool tma_out = (EPI == VDS_EPI_STORE || EPI == VDS_EPI_DGELU) ? tmC != nullptr : tmC2 != nullptr;
  if (tma_out) fence_proxy_async_smem();
  __syncwarp();
  const long long tq2 = clock64();
  if (tma_out) {
    const uint32_t stg_s = smem_u32(stg);
    if constexpr (EPI == VDS_EPI_STORE || EPI == VDS_EPI_DGELU) {
      if (lane == 0) { tma_store_2d(tmC, stg_s, col0, row0); bulk_commit_group(); }
    } else {
      if (lane == 0) { tma_store_2d(tmC2, stg_s, col0, row0); bulk_commit_group(); }
      if (p.C != nullptr) {
        if (lane == 0) bulk_wait_group_read0();     // the TMA unit has read the staging tile
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(stg + stg_off(lane, g)) = keep[g];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d(tmC, stg_s, col0, row0); bulk_commit_group(); }
      }
    }
    if (lane == 0) bulk_wait_group_read0();         // staging reusable by the next group
  } else if constexpr (EPI == VDS_EPI_STORE || EPI == VDS_EPI_DGELU) {
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
  const long long tq3 = clock64();
  __syncwarp();
  if (trace) {
    atomicAdd((unsigned long long*)p.dbg + 8, (unsigned long long)(tq1 - tq0));
    atomicAdd((unsigned long long*)p.dbg + 9, (unsigned long long)(tq2 - tq1));
    atomicAdd((unsigned long long*)p.dbg + 10, (unsigned long long)(tq3 - tq2));
    atomicAdd((unsigned long long*)p.dbg + 11, (unsigned long long)(clock64() - tq3));
  }
}

// fp32 accumulate (VDS_EPI_ACCUM_F32: wgrad split-K) of one warp's 32 rows x 32 columns: transposed through the warp's
// 4 KiB staging tile (thread = row -> 128-byte rows, XOR swizzle == TMA 128-byte swizzle) and added into C by the TMA unit
// (cp.reduce.async.bulk.tensor .add.f32).  Per-thread red.global.add.v4 of row-major data touches 32 lines per warp
// instruction: 64 of them per thread made the epilogue of one 256 x 256 split ~9 us (scripts/wgrad_splits_bench.py).
// Rows / columns past the matrix are clipped by the tensor map.
__device__ __forceinline__ void accum_f32_tma(const CUtensorMap* tmC, uint8_t* stg, int row0, int col0,
                                              const uint32_t (&v)[32], int lane) {
  if (lane == 0) bulk_wait_group_read0();      // the previous reduction has finished reading the tile
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<uint4*>(stg + stg_off(lane, g)) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_reduce_add_2d(tmC, smem_u32(stg), col0, row0);
    bulk_commit_group();
  }
}
// tensor map of an fp32 accumulation target C[M, ldc] for accum_f32_tma (box 32 x 32, 128-byte swizzle); *ok = 0 when C
// does not qualify (alignment): the kernels then fall back to per-thread red.global.add
static inline int make_tmap_accum_f32(CUtensorMap* tm, const void* C, long long ldc, int M, int N, int* ok) {
  *ok = 0;
  static const bool env_on = getenv("VDS_GEMM_TMA_RED") == nullptr || strcmp(getenv("VDS_GEMM_TMA_RED"), "0") != 0;   // tuning switch
  if (!env_on || C == nullptr || ldc % 4 != 0 || ((uintptr_t)C & 15) != 0) return VDS_OK;
  uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, strides[1] = {(uint64_t)ldc * 4};
  uint32_t box[2] = {32, 32};
  int r = encode_tmap(tm, C, 1, 2, dims, strides, box, 1);
  if (r) return r;
  *ok = 1;
  return VDS_OK;
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


}  // namespace vds
