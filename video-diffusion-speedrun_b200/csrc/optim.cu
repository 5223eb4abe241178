// Fused denoising-MSE loss (+ its gradient) and fused multi-tensor AdamW — both HBM-bound streams.
#include "common.h"
#include "ptx.cuh"

namespace vds {

// loss = mean_b mean_rest (v - out)^2 with v = bf16(x - noise)            (train.py:117,121-125)
// d_out = 2*(out - v) / (B*per) * gscale  (bf16).  Grid (chunks, B).
__global__ void __launch_bounds__(256) loss_kernel(const bf16* __restrict__ x, const bf16* __restrict__ noise,
                                                   const bf16* __restrict__ out, bf16* __restrict__ d_out,
                                                   float* __restrict__ loss_sum, float* __restrict__ loss_batch,
                                                   long long per, int B, float gscale,
                                                   const float* __restrict__ gscale_dev) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int b = blockIdx.y;
  if (gscale_dev != nullptr) gscale *= *gscale_dev;
  const long long base = (long long)b * per;
  const float k = 2.0f * gscale / ((float)B * (float)per);
  float acc = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8; i < per;
       i += (long long)gridDim.x * blockDim.x * 8) {
    if (i + 8 <= per) {
      const uint4 ux = *reinterpret_cast<const uint4*>(x + base + i);
      const uint4 un = *reinterpret_cast<const uint4*>(noise + base + i);
      const uint4 uo = *reinterpret_cast<const uint4*>(out + base + i);
      const uint32_t* px = &ux.x; const uint32_t* pn = &un.x; const uint32_t* po = &uo.x;
      uint32_t g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fx = unpack_bf16x2(px[j]), fn = unpack_bf16x2(pn[j]), fo = unpack_bf16x2(po[j]);
        const float d0 = fo.x - bf16_round(fx.x - fn.x), d1 = fo.y - bf16_round(fx.y - fn.y);
        acc += d0 * d0 + d1 * d1;
        g[j] = pack_bf16x2(d0 * k, d1 * k);
      }
      if (d_out != nullptr) *reinterpret_cast<uint4*>(d_out + base + i) = make_uint4(g[0], g[1], g[2], g[3]);
    } else {
      for (long long j = i; j < per; ++j) {
        const float d = __bfloat162float(out[base + j]) -
                        bf16_round(__bfloat162float(x[base + j]) - __bfloat162float(noise[base + j]));
        acc += d * d;
        if (d_out != nullptr) d_out[base + j] = __float2bfloat16_rn(d * k);
      }
    }
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    if (loss_sum != nullptr) atomicAdd(loss_sum, s / ((float)B * (float)per));
    if (loss_batch != nullptr) atomicAdd(loss_batch + b, s / (float)per);
  }
}

// torch.optim.AdamW(fused=True) semantics (train.py:340-344), one launch for every parameter:
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// Chunks never straddle tensors; each chunk carries its param-group id (per-group lr / wd by value).
struct AdamArgs {
  float* p; const float* g; float* m; float* v; bf16* p_bf16;
  const long long* chunk_start; const int* chunk_len; const int* chunk_group;
  float lr[16]; float wd[16];
  float beta1, beta2, eps, bc1, bc2_sqrt, gscale;
  int vec_ok;               // all flat buffers 16-byte aligned (bf16 copy 8-byte)
  const float* hyper_dev;   // optional [lr[16] | wd[16] | bc1 | bc2_sqrt] in device memory (CUDA-graph replays)
};
__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a) {
  pdl_wait();      // inputs may come from the previous kernel (common.h: launch_k)
  pdl_trigger();
  const int c = blockIdx.x;
  const long long s = a.chunk_start[c];
  const int n = a.chunk_len[c];
  const int grp = a.chunk_group[c];
  float lr = a.lr[grp], wd = a.wd[grp], bc1 = a.bc1, bc2_sqrt = a.bc2_sqrt;
  if (a.hyper_dev != nullptr) {
    lr = a.hyper_dev[grp]; wd = a.hyper_dev[16 + grp]; bc1 = a.hyper_dev[32]; bc2_sqrt = a.hyper_dev[33];
  }
  const float decay = 1.0f - lr * wd, step = lr / bc1;
  auto update = [&](float g, float& p, float& m, float& v) {
    g *= a.gscale;
    p *= decay;
    m = a.beta1 * m + (1.0f - a.beta1) * g;
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    p -= step * m / (sqrtf(v) / bc2_sqrt + a.eps);
  };
  auto scalar = [&](long long k) {
    float p = a.p[k], m = a.m[k], v = a.v[k];
    update(a.g[k], p, m, v);
    a.p[k] = p; a.m[k] = m; a.v[k] = v;
    if (a.p_bf16 != nullptr) a.p_bf16[k] = __float2bfloat16_rn(p);
  };
  // 16-byte body (the flat buffers are 16-byte aligned; a chunk may start anywhere) with scalar head / tail
  const int lead = a.vec_ok ? min(n, (int)((4 - (s & 3)) & 3)) : n;
  const int nvec = (n - lead) >> 2;
  if ((int)threadIdx.x < lead) scalar(s + threadIdx.x);
  const long long vb = s + lead;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const long long k = vb + 4LL * i;
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(a.g + k));
    float4 p4 = *reinterpret_cast<const float4*>(a.p + k);
    float4 m4 = *reinterpret_cast<const float4*>(a.m + k);
    float4 v4 = *reinterpret_cast<const float4*>(a.v + k);
    update(g4.x, p4.x, m4.x, v4.x);
    update(g4.y, p4.y, m4.y, v4.y);
    update(g4.z, p4.z, m4.z, v4.z);
    update(g4.w, p4.w, m4.w, v4.w);
    *reinterpret_cast<float4*>(a.p + k) = p4;
    *reinterpret_cast<float4*>(a.m + k) = m4;
    *reinterpret_cast<float4*>(a.v + k) = v4;
    if (a.p_bf16 != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(a.p_bf16 + k) = u;
    }
  }
  const int done = lead + 4 * nvec;
  if ((int)threadIdx.x < n - done) scalar(s + done + threadIdx.x);
}

}  // namespace vds

using namespace vds;

extern "C" {

int vds_loss_fwd_bwd(const void* x, const void* noise, const void* out, void* d_out, float* loss_sum,
                     float* loss_batch, int B, int64_t per_sample, float grad_scale, const float* grad_scale_dev,
                     void* stream) {
  VDS_CHECK_ARG(per_sample % 8 == 0, "loss: per-sample element count must be a multiple of 8");
  int chunks = (int)((per_sample / 8 + 255) / 256);
  const int cap = (4 * num_sms() + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, B);
  launch_k(loss_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)noise, (const bf16*)out,
                                                      (bf16*)d_out, loss_sum, loss_batch, per_sample, B, grad_scale,
                                                      grad_scale_dev);
  VDS_CHECK_LAUNCH("loss");
  return VDS_OK;
}

int vds_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const int64_t* chunk_start,
              const int32_t* chunk_len, const int32_t* chunk_group, int n_chunks, const float* lr_host,
              const float* wd_host, int n_groups, float beta1, float beta2, float eps, int step, float grad_scale,
              const float* hyper_dev, void* stream) {
  VDS_CHECK_ARG(n_groups >= 1 && n_groups <= 16, "adamw: 1..16 param groups supported (got %d)", n_groups);
  VDS_CHECK_ARG(step >= 1, "adamw: step must be >= 1");
  if (n_chunks == 0) return VDS_OK;
  AdamArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.p_bf16 = (bf16*)p_bf16;
  a.chunk_start = (const long long*)chunk_start; a.chunk_len = chunk_len; a.chunk_group = chunk_group;
  for (int i = 0; i < 16; ++i) { a.lr[i] = i < n_groups ? lr_host[i] : 0.f; a.wd[i] = i < n_groups ? wd_host[i] : 0.f; }
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.gscale = grad_scale;
  a.hyper_dev = hyper_dev;
  a.vec_ok = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)p_bf16 & 7) == 0;
  launch_k(adamw_kernel, n_chunks, 256, 0, (cudaStream_t)stream, a);
  VDS_CHECK_LAUNCH("adamw");
  return VDS_OK;
}

}  // extern "C"
