// HBM-bound kernels of the DiT step: patchify / unpatchify index maps, RoPE row gather, timestep
// embedding, SiLU, fused adaLN-modulated RMSNorm (fwd / bwd), gated-residual backward, fused
// RoPE + value-residual post-processing of QKV (fwd / bwd), column sums (bias grads), casts.
// All are coalesced, 16-byte vectorised where the layout allows, and fp32 inside.
#include "common.h"
#include "ptx.cuh"

namespace vds {

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ void ld8(const bf16* p, float (&o)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}
__device__ __forceinline__ void unpack8f(uint4 u, float (&o)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------- patchify
// x [B,C,T,H,W] -> A [B*N, C*pt*p*p]; token l = (h'*W' + w')*T' + t'   (model.py:185, "(h w t)")
// feature k = ((c*pt + dt)*p + dh)*p + dw                              (Conv3d weight, model.py:173)
// Optional fused z_t = x*(1-t) + noise*t in bf16 arithmetic            (train.py:115-116)
__global__ void patchify_kernel(const bf16* __restrict__ x, const bf16* __restrict__ noise,
                                const bf16* __restrict__ tvals, bf16* __restrict__ out, int B, int C, int T,
                                int H, int W, int p, int pt) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long total = (long long)B * C * T * H * W;
  const int Tp = T / pt, Hp = H / p, Wp = W / p;
  const int Kf = C * pt * p * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = r % W; r /= W;
    const int h = r % H; r /= H;
    const int t = r % T; r /= T;
    const int c = r % C; r /= C;
    const int b = (int)r;
    float v = __bfloat162float(x[i]);
    if (noise != nullptr) {
      const float tr = __bfloat162float(tvals[b]);
      const float omt = bf16_round(1.0f - tr);
      v = bf16_round(bf16_round(v * omt) + bf16_round(__bfloat162float(noise[i]) * tr));
    }
    const int tp = t / pt, dt = t % pt, hp = h / p, dh = h % p, wp = w / p, dw = w % p;
    const long long token = ((long long)hp * Wp + wp) * Tp + tp;
    const int k = ((c * pt + dt) * p + dh) * p + dw;
    out[((long long)b * ((long long)Tp * Hp * Wp) + token) * Kf + k] = __float2bfloat16_rn(v);
  }
}

// y [B*N, p*p*pt*C] <-> out [B,C,T,H,W]; feature f = ((p1*p + p2)*pt + p3)*C + c, p1<->H, p2<->W, p3<->T
// (model.py:392-401).  `to_tokens` = backward direction (gather dOut into token rows).
__global__ void unpatchify_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int C, int T, int H,
                                  int W, int p, int pt, int to_tokens) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long total = (long long)B * C * T * H * W;
  const int Tp = T / pt, Hp = H / p, Wp = W / p;
  const int F = C * pt * p * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = r % W; r /= W;
    const int h = r % H; r /= H;
    const int t = r % T; r /= T;
    const int c = r % C; r /= C;
    const int b = (int)r;
    const int tp = t / pt, p3 = t % pt, hp = h / p, p1 = h % p, wp = w / p, p2 = w % p;
    const long long token = ((long long)hp * Wp + wp) * Tp + tp;
    const int f = ((p1 * p + p2) * pt + p3) * C + c;
    const long long j = ((long long)b * ((long long)Tp * Hp * Wp) + token) * F + f;
    if (to_tokens) dst[j] = src[i]; else dst[i] = src[j];
  }
}

// ------------------------------------------------------------------------------------- RoPE rows
// cos/sin [L, D] fp32 gathered from the persistent tables [t_max, h_max, w_max, D] at the random
// start offsets; row l < n_reg is the identity rotation; row n = l - n_reg takes the table entry at
// unravel(n, (T', H', W')) — i.e. the reference's "(t h w)" flattening (model.py:228-261).
template <typename TT>
__global__ void rope_rows_kernel(const TT* __restrict__ tcos, const TT* __restrict__ tsin, float* __restrict__ ocos,
                                 float* __restrict__ osin, int L, int D, int n_reg, int Tp, int Hp, int Wp,
                                 int st, int sh, int sw, int hmax, int wmax, const int* __restrict__ starts_dev) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  if (starts_dev != nullptr) {   // CUDA-graph replays: the (t, h, w) offsets of this step live in device memory
    st = starts_dev[0];
    sh = starts_dev[1];
    sw = starts_dev[2];
  }
  const long long total = (long long)L * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / D), d = (int)(i % D);
    float c = 1.0f, s = 0.0f;
    if (l >= n_reg) {
      int n = l - n_reg;
      const int wi = n % Wp; n /= Wp;
      const int hi = n % Hp; n /= Hp;
      const int ti = n;
      const long long src = ((((long long)(st + ti)) * hmax + (sh + hi)) * wmax + (sw + wi)) * D + d;
      c = (float)tcos[src];
      s = (float)tsin[src];
    }
    ocos[i] = c;
    osin[i] = s;
  }
}

// packed cos / sin table of the fused QKV epilogue (gemm2.cu: qkv_rope_epilogue_warp): [L + 32][4][16 cos | 16 sin]
__global__ void rope_pack_kernel(const float* __restrict__ cos, const float* __restrict__ sin, float* __restrict__ tab, int L) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)(L + 32) * 128;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i >> 7) % L, s = (int)(i >> 5) & 3, j = (int)i & 31;
    tab[i] = j < 16 ? cos[l * 64 + s * 16 + j] : sin[l * 64 + s * 16 + j - 16];
  }
}

// ------------------------------------------------------------------------------------- timestep emb
// out[b, :half] = cos(t*f_i), out[b, half:] = sin(t*f_i), f_i = exp(-ln(max_period)*i/half)  (model.py:12-22)
__global__ void timestep_embedding_kernel(const bf16* __restrict__ t, bf16* __restrict__ out, int B, int dim,
                                          float max_period) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i % half;
  const float f = expf(-logf(max_period) * (float)j / (float)half);
  const float a = __bfloat162float(t[b]) * f;
  out[(long long)b * dim + j] = __float2bfloat16_rn(cosf(a));
  out[(long long)b * dim + half + j] = __float2bfloat16_rn(sinf(a));
}

// ------------------------------------------------------------------------------------- SiLU
__global__ void silu_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  y[i] = __float2bfloat16_rn(v / (1.0f + __expf(-v)));
}
// dx = dy * silu'(x) (+ dx_add)
__global__ void silu_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                long long n) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  const float s = 1.0f / (1.0f + __expf(-v));
  dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * (s * (1.0f + v * (1.0f - s))));
}

// ------------------------------------------------------------------------------------- RMSNorm + modulate
// y = bf16( bf16( bf16(x*rstd [*w]) * bf16(1+scale[b]) ) + shift[b] )     (model.py:34-41, 123)
// One warp per row; the row lives in registers (h <= 256*MAXV).
// NV = ceil(h / 256) 16-byte groups per lane (template), group g = i*32 + lane valid if g < h/8

struct NormArgs {
  const bf16* x; bf16* y; float* rstd;
  const bf16* weight; const bf16* scale; const bf16* shift;
  long long mod_stride;
  int rows_out, h;
  int rows_per_batch_out, in_batch_stride, in_row_offset;
  float eps;
};

template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_mod_fwd_kernel(const NormArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  // Each warp owns RPW consecutive rows: all their 16-byte groups (and the per-sample scale / shift / weight
  // groups) are requested up front, so one DRAM round trip covers RPW rows instead of one.
  constexpr int RPW = (NV <= 3) ? 4 : 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * RPW;
  if (row0 >= a.rows_out) return;
  const int ng = a.h / 8;
  uint4 raw[RPW][NV];
  int bidx[RPW];
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr) {
    const int row = min(row0 + rr, a.rows_out - 1);
    const int b = row / a.rows_per_batch_out, r = row % a.rows_per_batch_out;
    bidx[rr] = b;
    const bf16* xr = a.x + ((long long)b * a.in_batch_stride + a.in_row_offset + r) * a.h;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i * 32 + lane < ng) raw[rr][i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
  }
  // modulation of the first row's sample (rows of one warp straddle samples at most once; handled below), unpacked ONCE per
  // warp: bf16(1 + scale) and shift per column are the same for every row of a sample, and the conversions (XU pipe) were
  // a third of this kernel's issue slots when redone per row
  uint4 opscp[NV], shfp[NV], wfp[NV];   // bf16 pairs
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 + lane < ng) {
      const int c = (i * 32 + lane) * 8;
      if (a.scale != nullptr) {
        float sc[8];
        ld8(a.scale + (long long)bidx[0] * a.mod_stride + c, sc);
        shfp[i] = *reinterpret_cast<const uint4*>(a.shift + (long long)bidx[0] * a.mod_stride + c);
        opscp[i] = make_uint4(pack_bf16x2(1.0f + sc[0], 1.0f + sc[1]), pack_bf16x2(1.0f + sc[2], 1.0f + sc[3]),
                              pack_bf16x2(1.0f + sc[4], 1.0f + sc[5]), pack_bf16x2(1.0f + sc[6], 1.0f + sc[7]));
      }
      if (a.weight != nullptr) wfp[i] = *reinterpret_cast<const uint4*>(a.weight + c);
    }
  }
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr) {
    const int row = row0 + rr;
    if (row >= a.rows_out) break;
    float v[NV][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        unpack8f(raw[rr][i], v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
      }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)a.h + a.eps);
    if (lane == 0 && a.rstd != nullptr) a.rstd[row] = rstd;
    bf16* yr = a.y + (long long)row * a.h;
    const bool same = bidx[rr] == bidx[0];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rstd;
        if (a.weight != nullptr) {
          float w[8];
          unpack8f(wfp[i], w);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= w[j];
        }
        if (a.scale != nullptr) {
          // bf16( bf16( bf16(n) * bf16(1 + scale) ) + shift ): the roundings run on packed pairs (one cvt per two elements)
          uint4 nb;
          nb.x = pack_bf16x2(o[0], o[1]); nb.y = pack_bf16x2(o[2], o[3]);
          nb.z = pack_bf16x2(o[4], o[5]); nb.w = pack_bf16x2(o[6], o[7]);
          float n8[8];
          unpack8f(nb, n8);
          if (same) {
            float os[8];
            unpack8f(opscp[i], os);
#pragma unroll
            for (int j = 0; j < 8; ++j) n8[j] *= os[j];
          } else {
            float sc[8];
            ld8(a.scale + (long long)bidx[rr] * a.mod_stride + c, sc);
#pragma unroll
            for (int j = 0; j < 8; ++j) n8[j] *= bf16_round(1.0f + sc[j]);
          }
          uint4 pb;
          pb.x = pack_bf16x2(n8[0], n8[1]); pb.y = pack_bf16x2(n8[2], n8[3]);
          pb.z = pack_bf16x2(n8[4], n8[5]); pb.w = pack_bf16x2(n8[6], n8[7]);
          unpack8f(pb, o);
          if (same) {
            float sh[8];
            unpack8f(shfp[i], sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += sh[j];
          } else {
            float sh[8];
            ld8(a.shift + (long long)bidx[rr] * a.mod_stride + c, sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += sh[j];
          }
        }
        st8(yr + c, o);
      }
    }
  }
}

// Backward.  Given dy (grad of the modulated output), x, rstd:
//   xn = x*rstd ; xhat = bf16(xn [*w]) ; dxhat = dy*(1+scale)
//   dscale[b] += sum_rows dy*xhat ; dshift[b] += sum_rows dy ; dw += sum_rows dxhat*xn
//   dxn = dxhat [*w] ; dx = rstd*(dxn - xn*mean(dxn*xn)) (+ dx_res)
// Grid (chunks, B): a CTA only touches rows of one sample, so the column sums reduce in registers
// and end in one fp32 atomicAdd per feature per CTA.
struct NormBwdArgs {
  const bf16* dy; const bf16* x; const float* rstd;
  const bf16* weight; const bf16* scale;
  const bf16* dx_res; bf16* dx;
  float* dscale; float* dshift; float* dweight;
  long long mod_stride, dmod_stride;
  int h, rows_per_batch_out, in_batch_stride, in_row_offset, rows_per_cta;
  int dx_full_rows;  // 1: dx indexed like x (in rows); rows outside the normed range untouched
};

template <int NV, bool HAS_W>
__global__ void __launch_bounds__(256, (NV <= 2) ? 2 : 1) rmsnorm_mod_bwd_kernel(const NormBwdArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  constexpr bool PF = false;  // one-row-ahead prefetch costs more in occupancy than it buys (measured)
  extern __shared__ float red[];  // [8 warps][3][h] would be too big: reduce sequentially below
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_cta;
  const int r1 = min(a.rows_per_batch_out, r0 + a.rows_per_cta);
  const int ng = a.h / 8;
  float acc_sc[NV][8], acc_sh[NV][8], acc_w[HAS_W ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_sc[i][j] = 0.f; acc_sh[i][j] = 0.f; if (HAS_W) acc_w[HAS_W ? i : 0][j] = 0.f; }

  // one row per warp per step, the next row's 16-byte groups are already in flight (raw registers)
  uint4 nx[NV], ndy[NV], nrs[NV];
  float nrstd = 0.f;
  auto fetch = [&](int r) {
    const long long orow = (long long)b * a.rows_per_batch_out + r;
    const long long irow = (long long)b * a.in_batch_stride + a.in_row_offset + r;
    const long long drow = a.dx_full_rows ? irow : orow;
    nrstd = a.rstd[orow];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        nx[i] = *reinterpret_cast<const uint4*>(a.x + irow * a.h + c);
        ndy[i] = *reinterpret_cast<const uint4*>(a.dy + orow * a.h + c);
        if (a.dx_res != nullptr) nrs[i] = *reinterpret_cast<const uint4*>(a.dx_res + drow * a.h + c);
      }
    }
  };
  // per-column operands are the same for every row of this CTA (one sample per CTA): unpack / round them once
  // (kept packed as bf16 pairs: 4 registers per group instead of 8, the kernel sits right at the 128-register occupancy step)
  uint4 ops1p[NV], wfp[HAS_W ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 + lane < ng) {
      const int c = (i * 32 + lane) * 8;
      if (a.scale != nullptr) {
        float sc[8];
        ld8(a.scale + (long long)b * a.mod_stride + c, sc);
        ops1p[i] = make_uint4(pack_bf16x2(1.0f + sc[0], 1.0f + sc[1]), pack_bf16x2(1.0f + sc[2], 1.0f + sc[3]),
                              pack_bf16x2(1.0f + sc[4], 1.0f + sc[5]), pack_bf16x2(1.0f + sc[6], 1.0f + sc[7]));
      }
      if (HAS_W) wfp[HAS_W ? i : 0] = *reinterpret_cast<const uint4*>(a.weight + c);
    }
  }
  if (PF && r0 + warp < r1) fetch(r0 + warp);
  for (int r = r0 + warp; r < r1; r += 8) {
    if (!PF) fetch(r);
    const long long orow = (long long)b * a.rows_per_batch_out + r;
    const long long irow = (long long)b * a.in_batch_stride + a.in_row_offset + r;
    const float rstd = nrstd;
    uint4 cx[NV], cdy[NV], crs[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { cx[i] = nx[i]; cdy[i] = ndy[i]; crs[i] = nrs[i]; }
    if (PF && r + 8 < r1) fetch(r + 8);
    float dxn[NV][8], xn[NV][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        float xv[8], dyv[8];
        unpack8f(cx[i], xv);
        unpack8f(cdy[i], dyv);
        float w[8], ops1[8];
        if (HAS_W) unpack8f(wfp[HAS_W ? i : 0], w);
        if (a.scale != nullptr) unpack8f(ops1p[i], ops1);
        (void)c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float n = xv[j] * rstd;
          xn[i][j] = n;
          const float xhat = bf16_round(HAS_W ? n * w[j] : n);
          float dxhat = dyv[j];
          if (a.scale != nullptr) {
            acc_sc[i][j] += dyv[j] * xhat;
            acc_sh[i][j] += dyv[j];
            dxhat = dyv[j] * ops1[j];
          }
          float d = dxhat;
          if (HAS_W) {
            acc_w[HAS_W ? i : 0][j] += dxhat * n;
            d = dxhat * w[j];
          }
          dxn[i][j] = d;
          dot += d * n;
        }
      }
    }
    dot = warp_sum(dot) / (float)a.h;
    const long long drow = a.dx_full_rows ? irow : orow;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dxn[i][j] - xn[i][j] * dot);
        if (a.dx_res != nullptr) {
          float rr[8];
          unpack8f(crs[i], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rr[j];
        }
        st8(a.dx + drow * a.h + c, o);
      }
    }
  }
  // column sums: every warp parks its register partials in shared memory (conflict-free [j][group] order), the
  // CTA adds the 8 warps and issues ONE global atomicAdd per column.
  float* s = red;  // [8 warps][h]
  const int ngr = a.h / 8;
  for (int q = 0; q < 3; ++q) {
    if (q < 2 && a.scale == nullptr) continue;
    if (q == 2 && !HAS_W) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int g = i * 32 + lane;
      if (g < ngr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          s[warp * a.h + j * ngr + g] = q == 0 ? acc_sc[i][j] : (q == 1 ? acc_sh[i][j] : acc_w[HAS_W ? i : 0][j]);
      }
    }
    __syncthreads();
    float* dst = q == 0 ? a.dscale + (long long)b * a.dmod_stride
                        : (q == 1 ? a.dshift + (long long)b * a.dmod_stride : a.dweight);
    for (int col = threadIdx.x; col < a.h; col += blockDim.x) {   // consecutive threads -> consecutive addresses
      const int pos = (col & 7) * ngr + (col >> 3);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s[w * a.h + pos];
      atomicAdd(&dst[col], t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------- gated residual bwd
// x_new = x + o*gate[b]:  do = bf16(dx*gate[b]) ; dgate[b] += sum_rows dx*o      (model.py:139,160,165)
struct GateBwdArgs {
  const bf16* dx; const bf16* o; const bf16* gate; bf16* d_o; float* dgate;
  long long gate_stride, dgate_stride;
  int h, rows_per_batch, rows_per_cta;
};
template <int NV>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const GateBwdArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  extern __shared__ float red[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_cta;
  const int r1 = min(a.rows_per_batch, r0 + a.rows_per_cta);
  const int ng = a.h / 8;
  float acc[NV][8];
  float g[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 + lane < ng) ld8(a.gate + (long long)b * a.gate_stride + (i * 32 + lane) * 8, g[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  uint4 ndx[NV], no[NV];
  auto fetch = [&](int r) {
    const long long row = (long long)b * a.rows_per_batch + r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        ndx[i] = *reinterpret_cast<const uint4*>(a.dx + row * a.h + c);
        no[i] = *reinterpret_cast<const uint4*>(a.o + row * a.h + c);
      }
    }
  };
  if (r0 + warp < r1) fetch(r0 + warp);
  for (int r = r0 + warp; r < r1; r += 8) {
    const long long row = (long long)b * a.rows_per_batch + r;
    uint4 cdx[NV], co[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { cdx[i] = ndx[i]; co[i] = no[i]; }
    if (r + 8 < r1) fetch(r + 8);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < ng) {
        const int c = (i * 32 + lane) * 8;
        float dxv[8], ov[8], out[8];
        unpack8f(cdx[i], dxv);
        unpack8f(co[i], ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] += dxv[j] * ov[j];
          out[j] = dxv[j] * g[i][j];
        }
        st8(a.d_o + row * a.h + c, out);
      }
    }
  }
  const int ngr = a.h / 8;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int g = i * 32 + lane;
    if (g < ngr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp * a.h + j * ngr + g] = acc[i][j];
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < a.h; col += blockDim.x) {
    const int pos = (col & 7) * ngr + (col >> 3);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w * a.h + pos];
    atomicAdd(&a.dgate[(long long)b * a.dgate_stride + col], t);
  }
}

// ------------------------------------------------------------------------------------- QKV post (fwd)
// In place on qkv [B, L, 3h] ("(k h d)" split, model.py:126): RoPE on q and k (fp32 math on the
// bf16-rounded GEMM output, half-split over the whole head, model.py:266-275) and the value residual
// v = l*v + (1-l)*v0 in bf16 arithmetic (model.py:130) written to a separate buffer (v_pre is kept for dl).
struct QkvPostArgs {
  bf16* qkv; const float* cos; const float* sin;
  const bf16* v0; long long v0_ld; bf16* vmix; const bf16* lambda;
  int B, L, h, nh, hd;
};
__global__ void qkv_post_fwd_kernel(const QkvPostArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int half8 = a.hd / 16;                       // 8-element groups in half a head
  const int groups_qk = 2 * a.nh * half8;            // rope groups per token (q and k)
  const int groups_v = (a.v0 != nullptr) ? a.h / 8 : 0;
  const int per_tok = groups_qk + groups_v;
  const long long total = (long long)a.B * a.L * per_tok;
  float lam = 0.f, oml = 0.f;
  if (a.v0 != nullptr) { lam = __bfloat162float(*a.lambda); oml = bf16_round(1.0f - lam); }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gi = (int)(i % per_tok);
    const long long tok = i / per_tok;
    const int l = (int)(tok % a.L);
    bf16* row = a.qkv + tok * 3LL * a.h;
    if (gi < groups_qk) {
      if (a.cos == nullptr) continue;
      const int which = gi / (a.nh * half8);         // 0 = q, 1 = k
      const int rem = gi % (a.nh * half8);
      const int head = rem / half8, g = rem % half8;
      bf16* p1 = row + which * a.h + head * a.hd + g * 8;
      bf16* p2 = p1 + a.hd / 2;
      float x1[8], x2[8], y1[8], y2[8];
      ld8(p1, x1); ld8(p2, x2);
      const float4* c4 = reinterpret_cast<const float4*>(a.cos + (long long)l * (a.hd / 2) + g * 8);
      const float4* s4 = reinterpret_cast<const float4*>(a.sin + (long long)l * (a.hd / 2) + g * 8);
      float c[8], s[8];
      *reinterpret_cast<float4*>(c) = c4[0]; *reinterpret_cast<float4*>(c + 4) = c4[1];
      *reinterpret_cast<float4*>(s) = s4[0]; *reinterpret_cast<float4*>(s + 4) = s4[1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y1[j] = x1[j] * c[j] + x2[j] * s[j];
        y2[j] = x1[j] * (-s[j]) + x2[j] * c[j];
      }
      st8(p1, y1); st8(p2, y2);
    } else {
      const int c0 = (gi - groups_qk) * 8;
      float v[8], v0[8], o[8];
      ld8(row + 2 * a.h + c0, v);
      ld8(a.v0 + tok * a.v0_ld + c0, v0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = bf16_round(lam * v[j]) + bf16_round(oml * v0[j]);
      st8(a.vmix + tok * (long long)a.h + c0, o);
    }
  }
}

// Backward, in place on dqkv [B, L, 3h]:
//   dq (from the fp32 accumulation buffer dq_acc [B,L,h], if given) and dk: inverse rotation;
//   dv_mix -> dl += sum dv_mix*(v_pre - v0); dv0_acc += (1-l)*dv_mix; dv_pre = l*dv_mix  (blocks >= 1)
//   block 0 (v0_acc_in given): dv_pre = dv_mix + dv0_acc
struct QkvPostBwdArgs {
  bf16* dqkv; const float* dq_acc; const float* cos; const float* sin;
  const bf16* qkv_pre;                 // forward qkv buffer (v slot = v_pre)
  const bf16* v0; long long v0_ld; const bf16* lambda;
  float* dlambda; float* dv0_acc; int mode;  // 0: no v handling, 1: blocks>=1 (mix bwd), 2: block 0 (+= dv0_acc)
  int B, L, h, nh, hd;
};
__global__ void qkv_post_bwd_kernel(const QkvPostBwdArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int half8 = a.hd / 16;
  const int groups_qk = 2 * a.nh * half8;
  const int groups_v = (a.mode != 0) ? a.h / 8 : 0;
  const int per_tok = groups_qk + groups_v;
  const long long total = (long long)a.B * a.L * per_tok;
  float lam = 0.f, oml = 0.f;
  if (a.mode == 1) { lam = __bfloat162float(*a.lambda); oml = bf16_round(1.0f - lam); }
  float dl = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gi = (int)(i % per_tok);
    const long long tok = i / per_tok;
    const int l = (int)(tok % a.L);
    bf16* row = a.dqkv + tok * 3LL * a.h;
    if (gi < groups_qk) {
      const int which = gi / (a.nh * half8);
      const int rem = gi % (a.nh * half8);
      const int head = rem / half8, g = rem % half8;
      bf16* p1 = row + which * a.h + head * a.hd + g * 8;
      bf16* p2 = p1 + a.hd / 2;
      float y1[8], y2[8], x1[8], x2[8];
      if (which == 0 && a.dq_acc != nullptr) {
        const float* q1 = a.dq_acc + tok * (long long)a.h + head * a.hd + g * 8;
        const float* q2 = q1 + a.hd / 2;
        *reinterpret_cast<float4*>(y1) = reinterpret_cast<const float4*>(q1)[0];
        *reinterpret_cast<float4*>(y1 + 4) = reinterpret_cast<const float4*>(q1)[1];
        *reinterpret_cast<float4*>(y2) = reinterpret_cast<const float4*>(q2)[0];
        *reinterpret_cast<float4*>(y2 + 4) = reinterpret_cast<const float4*>(q2)[1];
      } else {
        ld8(p1, y1); ld8(p2, y2);
      }
      if (a.cos != nullptr) {
        const float4* c4 = reinterpret_cast<const float4*>(a.cos + (long long)l * (a.hd / 2) + g * 8);
        const float4* s4 = reinterpret_cast<const float4*>(a.sin + (long long)l * (a.hd / 2) + g * 8);
        float c[8], s[8];
        *reinterpret_cast<float4*>(c) = c4[0]; *reinterpret_cast<float4*>(c + 4) = c4[1];
        *reinterpret_cast<float4*>(s) = s4[0]; *reinterpret_cast<float4*>(s + 4) = s4[1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          x1[j] = y1[j] * c[j] - y2[j] * s[j];
          x2[j] = y1[j] * s[j] + y2[j] * c[j];
        }
        st8(p1, x1); st8(p2, x2);
      } else if (which == 0 && a.dq_acc != nullptr) {
        st8(p1, y1); st8(p2, y2);
      }
    } else {
      const int c0 = (gi - groups_qk) * 8;
      float dv[8], o[8];
      ld8(row + 2 * a.h + c0, dv);
      float* acc = a.dv0_acc + tok * (long long)a.h + c0;
      if (a.mode == 1) {
        float vp[8], v0[8];
        ld8(a.qkv_pre + tok * 3LL * a.h + 2 * a.h + c0, vp);
        ld8(a.v0 + tok * a.v0_ld + c0, v0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dl += dv[j] * (vp[j] - v0[j]);
          acc[j] += oml * dv[j];
          o[j] = lam * dv[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = dv[j] + acc[j];
      }
      st8(row + 2 * a.h + c0, o);
    }
  }
  if (a.mode == 1) {
    dl = warp_sum(dl);
    if ((threadIdx.x & 31) == 0 && dl != 0.f) atomicAdd(a.dlambda, dl);
  }
}

// ------------------------------------------------------------------------------------- column sums
// out[n] += sum_rows x[row, n]  (bias gradients), x bf16 [rows, ld], out fp32.
// CTA = 256 columns x `rows_per_cta` rows: a warp reads one 512-byte row segment per step (4 rows in flight),
// the 8 warps meet in shared memory, one atomicAdd per column per CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long rows,
                                                     int n, long long ld, int rows_per_cta) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < n) {
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      float v0[8], v1[8], v2[8], v3[8];
      ld8(x + r * ld + c, v0);
      ld8(x + (r + 8) * ld + c, v1);
      ld8(x + (r + 16) * ld + c, v2);
      ld8(x + (r + 24) * ld + c, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < r1; r += 8) {
      float v[8];
      ld8(x + r * ld + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane + 32 * j] = acc[j];   // conflict-free: [j][lane] order
  __syncthreads();
  const int t = threadIdx.x;          // t = lane' + 32*j'  -> column lane'*8 + j'
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][t];
  const int col = blockIdx.x * 256 + (t & 31) * 8 + (t >> 5);
  if (col < n) atomicAdd(&out[col], s);
}

// out[r, :] (+)= sum_b x[b, r, :]   (register-token gradient: rows 0..15 of every sample)
__global__ void batch_rowsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int B, long long batch_stride,
                                    int rows, int h) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * h) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += __bfloat162float(x[(long long)b * batch_stride + i]);
  out[i] += acc;
}

// ------------------------------------------------------------------------------------- casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n, float scale) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    uint2 u;
    u.x = pack_bf16x2(v.x * scale, v.y * scale);
    u.y = pack_bf16x2(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(y + i) = u;
  } else {
    for (long long j = i; j < n; ++j) y[j] = __float2bfloat16_rn(x[j] * scale);
  }
}
// 2-D variant: x fp32 [rows, cols] (leading dim ldx) -> y bf16 [rows, cols] (leading dim ldy); cols % 4 == 0
__global__ void cast_f32_bf16_2d_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                        long long rows, int cols, float scale) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int c4 = cols / 4;
  const long long total = rows * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int c = (int)(i % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    uint2 u;
    u.x = pack_bf16x2(v.x * scale, v.y * scale);
    u.y = pack_bf16x2(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(y + r * ldy + c) = u;
  }
}
// y(fp32) (+)= float(x(bf16))
__global__ void accum_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n, int accumulate) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  y[i] = accumulate ? y[i] + v : v;
}

}  // namespace vds

using namespace vds;

extern "C" {

int vds_patchify(const void* x, const void* noise, const void* t, void* out, int B, int C, int T, int H, int W,
                 int p, int pt, void* stream) {
  VDS_CHECK_ARG(T % pt == 0 && H % p == 0 && W % p == 0, "patchify: T,H,W must divide by the patch size");
  const long long total = (long long)B * C * T * H * W;
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  launch_k(patchify_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)noise, (const bf16*)t,
                                                          (bf16*)out, B, C, T, H, W, p, pt);
  VDS_CHECK_LAUNCH("patchify");
  return VDS_OK;
}

int vds_unpatchify(const void* src, void* dst, int B, int C, int T, int H, int W, int p, int pt, int to_tokens,
                   void* stream) {
  VDS_CHECK_ARG(T % pt == 0 && H % p == 0 && W % p == 0, "unpatchify: T,H,W must divide by the patch size");
  const long long total = (long long)B * C * T * H * W;
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  launch_k(unpatchify_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)src, (bf16*)dst, B, C, T, H, W, p, pt,
                                                            to_tokens);
  VDS_CHECK_LAUNCH("unpatchify");
  return VDS_OK;
}

int vds_rope_rows(const void* tcos, const void* tsin, int table_is_bf16, float* ocos, float* osin, int L, int D,
                  int n_reg, int Tp, int Hp, int Wp, int st, int sh, int sw, int hmax, int wmax, const int* starts_dev,
                  void* stream) {
  VDS_CHECK_ARG(L == n_reg + Tp * Hp * Wp, "rope_rows: L=%d != %d + %d*%d*%d", L, n_reg, Tp, Hp, Wp);
  const long long total = (long long)L * D;
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  if (table_is_bf16)
    launch_k(rope_rows_kernel<bf16>, grid, 256, 0, (cudaStream_t)stream, (const bf16*)tcos, (const bf16*)tsin, ocos, osin,
                                                                  L, D, n_reg, Tp, Hp, Wp, st, sh, sw, hmax, wmax, starts_dev);
  else
    launch_k(rope_rows_kernel<float>, grid, 256, 0, (cudaStream_t)stream, (const float*)tcos, (const float*)tsin, ocos,
                                                                   osin, L, D, n_reg, Tp, Hp, Wp, st, sh, sw, hmax,
                                                                   wmax, starts_dev);
  VDS_CHECK_LAUNCH("rope_rows");
  return VDS_OK;
}

int vds_rope_pack(const float* cos, const float* sin, float* tab, int L, void* stream) {
  VDS_CHECK_ARG(L > 0 && cos != nullptr && sin != nullptr && tab != nullptr, "rope_pack: bad arguments");
  const long long total = (long long)(L + 32) * 128;
  launch_k(rope_pack_kernel, min(ceil_div(total, 256), num_sms() * 16), 256, 0, (cudaStream_t)stream, cos, sin, tab, L);
  VDS_CHECK_LAUNCH("rope_pack");
  return VDS_OK;
}

int vds_timestep_embedding(const void* t, void* out, int B, int dim, float max_period, void* stream) {
  VDS_CHECK_ARG(dim % 2 == 0, "timestep_embedding: odd dim");
  launch_k(timestep_embedding_kernel, ceil_div((long long)B * dim / 2, 256), 256, 0, (cudaStream_t)stream, 
      (const bf16*)t, (bf16*)out, B, dim, max_period);
  VDS_CHECK_LAUNCH("timestep_embedding");
  return VDS_OK;
}

int vds_silu(const void* x, void* y, int64_t n, void* stream) {
  launch_k(silu_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, n);
  VDS_CHECK_LAUNCH("silu");
  return VDS_OK;
}
int vds_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, void* stream) {
  launch_k(silu_bwd_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)dy, (bf16*)dx, n);
  VDS_CHECK_LAUNCH("silu_bwd");
  return VDS_OK;
}

int vds_rmsnorm_mod_fwd(const void* x, void* y, float* rstd, const void* weight, const void* scale,
                        const void* shift, int64_t mod_stride, int B, int rows_per_batch_out, int in_batch_stride,
                        int in_row_offset, int h, float eps, void* stream) {
  VDS_CHECK_ARG(h % 8 == 0 && h <= 2048, "rmsnorm: h=%d must be a multiple of 8 and <= 2048", h);
  NormArgs a;
  a.x = (const bf16*)x; a.y = (bf16*)y; a.rstd = rstd; a.weight = (const bf16*)weight;
  a.scale = (const bf16*)scale; a.shift = (const bf16*)shift; a.mod_stride = mod_stride;
  a.rows_out = B * rows_per_batch_out; a.h = h; a.rows_per_batch_out = rows_per_batch_out;
  a.in_batch_stride = in_batch_stride; a.in_row_offset = in_row_offset; a.eps = eps;
  const int nvn = (h + 255) / 256;
#define VDS_L(NVV) launch_k(rmsnorm_mod_fwd_kernel<NVV>, ceil_div(a.rows_out, 8 * ((NVV) <= 3 ? 4 : 2)), 256, 0, (cudaStream_t)stream, a)
  if (nvn <= 2) VDS_L(2); else if (nvn <= 3) VDS_L(3); else if (nvn <= 5) VDS_L(5); else VDS_L(8);
#undef VDS_L
  VDS_CHECK_LAUNCH("rmsnorm_mod_fwd");
  return VDS_OK;
}

int vds_rmsnorm_mod_bwd(const void* dy, const void* x, const float* rstd, const void* weight, const void* scale,
                        const void* dx_res, void* dx, float* dscale, float* dshift, float* dweight,
                        int64_t mod_stride, int64_t dmod_stride, int B, int rows_per_batch_out, int in_batch_stride,
                        int in_row_offset, int dx_full_rows, int h, void* stream) {
  VDS_CHECK_ARG(h % 8 == 0 && h <= 2048, "rmsnorm_bwd: h=%d unsupported", h);
  NormBwdArgs a;
  a.dy = (const bf16*)dy; a.x = (const bf16*)x; a.rstd = rstd; a.weight = (const bf16*)weight;
  a.scale = (const bf16*)scale; a.dx_res = (const bf16*)dx_res; a.dx = (bf16*)dx;
  a.dscale = dscale; a.dshift = dshift; a.dweight = dweight; a.mod_stride = mod_stride; a.dmod_stride = dmod_stride;
  a.h = h; a.rows_per_batch_out = rows_per_batch_out; a.in_batch_stride = in_batch_stride;
  a.in_row_offset = in_row_offset; a.dx_full_rows = dx_full_rows;
  // ~2 waves of CTAs over the chip, at least 8 rows (one per warp) each
  int chunks = max(1, (2 * num_sms()) / max(1, B));
  a.rows_per_cta = max(16, ceil_div(rows_per_batch_out, chunks));
  dim3 grid(ceil_div(rows_per_batch_out, a.rows_per_cta), B);
  const int nvn = (h + 255) / 256;
#define VDS_L(NVV, HW)                                                                                      \
  do {                                                                                                     \
    if (8 * h * sizeof(float) > 48 * 1024)                                                                 \
      cudaFuncSetAttribute(rmsnorm_mod_bwd_kernel<NVV, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           (int)(8 * h * sizeof(float)));                                                  \
    launch_k(rmsnorm_mod_bwd_kernel<NVV, HW>, grid, 256, 8 * h * sizeof(float), (cudaStream_t)stream, a);        \
  } while (0)
  if (weight != nullptr) {
    if (nvn <= 2) VDS_L(2, true); else if (nvn <= 3) VDS_L(3, true); else if (nvn <= 5) VDS_L(5, true); else VDS_L(8, true);
  } else {
    if (nvn <= 2) VDS_L(2, false); else if (nvn <= 3) VDS_L(3, false); else if (nvn <= 5) VDS_L(5, false); else VDS_L(8, false);
  }
#undef VDS_L
  VDS_CHECK_LAUNCH("rmsnorm_mod_bwd");
  return VDS_OK;
}

int vds_gate_bwd(const void* dx, const void* o, const void* gate, void* d_o, float* dgate, int64_t gate_stride,
                 int64_t dgate_stride, int B, int rows_per_batch, int h, void* stream) {
  VDS_CHECK_ARG(h % 8 == 0 && h <= 2048, "gate_bwd: h=%d unsupported", h);
  GateBwdArgs a;
  a.dx = (const bf16*)dx; a.o = (const bf16*)o; a.gate = (const bf16*)gate; a.d_o = (bf16*)d_o; a.dgate = dgate;
  a.gate_stride = gate_stride; a.dgate_stride = dgate_stride; a.h = h; a.rows_per_batch = rows_per_batch;
  int chunks = max(1, (2 * num_sms()) / max(1, B));
  a.rows_per_cta = max(16, ceil_div(rows_per_batch, chunks));
  dim3 grid(ceil_div(rows_per_batch, a.rows_per_cta), B);
  const int nvn = (h + 255) / 256;
#define VDS_L(NVV)                                                                                         \
  do {                                                                                                     \
    if (8 * h * sizeof(float) > 48 * 1024)                                                                 \
      cudaFuncSetAttribute(gate_bwd_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                           (int)(8 * h * sizeof(float)));                                                  \
    launch_k(gate_bwd_kernel<NVV>, grid, 256, 8 * h * sizeof(float), (cudaStream_t)stream, a);                   \
  } while (0)
  if (nvn <= 2) VDS_L(2); else if (nvn <= 3) VDS_L(3); else if (nvn <= 5) VDS_L(5); else VDS_L(8);
#undef VDS_L
  VDS_CHECK_LAUNCH("gate_bwd");
  return VDS_OK;
}

int vds_qkv_post_fwd(void* qkv, const float* cos, const float* sin, const void* v0, int64_t v0_ld, void* vmix,
                     const void* lambda, int B, int L, int h, int nh, void* stream) {
  VDS_CHECK_ARG(h % nh == 0 && (h / nh) % 16 == 0, "qkv_post: bad head split h=%d nh=%d", h, nh);
  QkvPostArgs a;
  a.qkv = (bf16*)qkv; a.cos = cos; a.sin = sin; a.v0 = (const bf16*)v0; a.v0_ld = v0_ld; a.vmix = (bf16*)vmix;
  a.lambda = (const bf16*)lambda; a.B = B; a.L = L; a.h = h; a.nh = nh; a.hd = h / nh;
  const long long total = (long long)B * L * (2 * nh * (a.hd / 16) + (v0 ? h / 8 : 0));
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  launch_k(qkv_post_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, a);
  VDS_CHECK_LAUNCH("qkv_post_fwd");
  return VDS_OK;
}

int vds_qkv_post_bwd(void* dqkv, const float* dq_acc, const float* cos, const float* sin, const void* qkv_pre,
                     const void* v0, int64_t v0_ld, const void* lambda, float* dlambda, float* dv0_acc, int mode,
                     int B, int L, int h, int nh, void* stream) {
  VDS_CHECK_ARG(h % nh == 0 && (h / nh) % 16 == 0, "qkv_post_bwd: bad head split h=%d nh=%d", h, nh);
  VDS_CHECK_ARG(mode >= 0 && mode <= 2, "qkv_post_bwd: bad mode %d", mode);
  QkvPostBwdArgs a;
  a.dqkv = (bf16*)dqkv; a.dq_acc = dq_acc; a.cos = cos; a.sin = sin; a.qkv_pre = (const bf16*)qkv_pre;
  a.v0 = (const bf16*)v0; a.v0_ld = v0_ld; a.lambda = (const bf16*)lambda; a.dlambda = dlambda;
  a.dv0_acc = dv0_acc; a.mode = mode; a.B = B; a.L = L; a.h = h; a.nh = nh; a.hd = h / nh;
  const long long total = (long long)B * L * (2 * nh * (a.hd / 16) + (mode ? h / 8 : 0));
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  launch_k(qkv_post_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, a);
  VDS_CHECK_LAUNCH("qkv_post_bwd");
  return VDS_OK;
}

int vds_colsum(const void* x, float* out, int64_t rows, int n, int64_t ld, void* stream) {
  VDS_CHECK_ARG(n % 8 == 0 && ld % 8 == 0, "colsum: n, ld must be multiples of 8");
  const int col_ctas = ceil_div(n, 256);
  int row_chunks = max(1, (2 * num_sms()) / col_ctas);
  int rows_per_cta = max(64, ceil_div(rows, row_chunks));
  dim3 grid(col_ctas, ceil_div(rows, rows_per_cta));
  launch_k(colsum_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, out, rows, n, ld, rows_per_cta);
  VDS_CHECK_LAUNCH("colsum");
  return VDS_OK;
}

int vds_batch_rowsum(const void* x, float* out, int B, int64_t batch_stride, int rows, int h, void* stream) {
  launch_k(batch_rowsum_kernel, ceil_div((long long)rows * h, 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, out, B,
                                                                                           batch_stride, rows, h);
  VDS_CHECK_LAUNCH("batch_rowsum");
  return VDS_OK;
}

int vds_cast_f32_bf16(const float* x, void* y, int64_t n, float scale, void* stream) {
  launch_k(cast_f32_bf16_kernel, ceil_div(ceil_div(n, 4), 256), 256, 0, (cudaStream_t)stream, x, (bf16*)y, n, scale);
  VDS_CHECK_LAUNCH("cast_f32_bf16");
  return VDS_OK;
}
int vds_cast_f32_bf16_2d(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols, float scale,
                         void* stream) {
  VDS_CHECK_ARG(cols % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0,
                "cast_f32_bf16_2d: cols / leading dims must be multiples of 4 and the pointers 16- / 8-byte aligned");
  const long long total = rows * (cols / 4);
  const int grid = min(ceil_div(total, 256), num_sms() * 16);
  launch_k(cast_f32_bf16_2d_kernel, grid, 256, 0, (cudaStream_t)stream, x, (long long)ldx, (bf16*)y, (long long)ldy,
           (long long)rows, cols, scale);
  VDS_CHECK_LAUNCH("cast_f32_bf16_2d");
  return VDS_OK;
}
int vds_accum_bf16_f32(const void* x, float* y, int64_t n, int accumulate, void* stream) {
  launch_k(accum_bf16_f32_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream, (const bf16*)x, y, n, accumulate);
  VDS_CHECK_LAUNCH("accum_bf16_f32");
  return VDS_OK;
}

}  // extern "C"
