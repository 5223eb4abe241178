/* vds_b200 — C ABI of the B200-native (sm_100a) DiT train-step kernels.
 *
 * The reference (fal-ai-community/video-diffusion-speedrun) is 100 % Python on top of PyTorch and
 * has NO native interface of its own.  Each entry point below therefore cites the
 * reference *call site* (file:line in /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*), never allocates,
 *     never synchronises; all buffers (incl. workspaces) are owned by the caller;
 *   - pointers are raw device pointers unless the name says `host`;
 *   - bf16 tensors are `uint16_t`-sized elements (torch.bfloat16), fp32 are `float`;
 *   - return value: 0 = ok, <0 = vds_status error; `vds_last_error()` gives a message;
 *   - no CPU fallback exists: on a box without an sm_100 device every launch returns VDS_ERR_CUDA.
 */
#ifndef VDS_B200_H
#define VDS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum vds_status {
  VDS_OK = 0,
  VDS_ERR_ARG = -1,   /* bad shape / alignment / enum */
  VDS_ERR_CUDA = -2,  /* a CUDA runtime or driver call failed */
  VDS_ERR_UNSUPPORTED = -3
};

const char* vds_last_error(void);
int vds_abi_version(void);
/* number of kernels launched by this library in this process (bench.py's `gpu_launches`). */
int64_t vds_launch_count(void);

/* ------------------------------------------------------------------------------------------ GEMM
 * D[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM), TMA-fed.
 * Replaces every nn.Linear / Conv3d-as-GEMM on the path and their autograd:
 *   fprop  y = x W^T        model.py:62,63,70-74,84-86,90,125,138,147,150,159,165,184,344,390
 *   dgrad  dx = dy W        (a_major K, b_major MN)
 *   wgrad  dW = dy^T x      (a_major MN, b_major MN, fp32 accumulate-into-output)
 * Operand storage: `*_mn == 0`: row-major [rows = M or N][cols = K], leading dim ld* (elements);
 *                  `*_mn == 1`: row-major [rows = K][cols = M or N]            ("MN-major").
 * All leading dims must be multiples of 8 elements, base pointers 16-byte aligned.
 */
enum vds_gemm_epilogue {
  VDS_EPI_STORE = 0,     /* C = bf16(acc + bias?)            ; optional output row remap        */
  VDS_EPI_ACCUM_F32 = 1, /* C(fp32) += acc   (red.global.add; split-K allowed)                   */
  VDS_EPI_BIAS_GELU = 2, /* C = bf16(acc+bias) ; C2 = bf16(gelu_erf(C))        model.py:84-85    */
  VDS_EPI_GATE_RES = 3,  /* C = bf16(acc+bias?) ; C2 = bf16(aux + bf16(C*gate[b]))  model.py:139 */
  VDS_EPI_DGELU = 4,     /* C = bf16(acc * gelu'(aux))                                          */
  VDS_EPI_STORE_F32 = 5, /* C(fp32) = acc + bias?                                               */
  VDS_EPI_QKV_ROPE = 7,  /* QKV projection of a DiT block, N = 3*heads*128 in "(k h d)" column order (model.py:124-134):
                          * C[:, :2N/3] = RoPE(bf16(acc+bias?)) per 128-wide head with the cos / sin of table row (row % rows_per_batch) of rope_tab
                          * (half-split rotation, model.py:266-275); C[:, 2N/3:] = v_pre = bf16(acc+bias?); with v0 != NULL
                          * also C2[M, N/3] = bf16(lambda*v_pre) + bf16((1-lambda)*v0) (model.py:129-130).  2-CTA tile path
                          * only: VDS_ERR_UNSUPPORTED otherwise (callers fall back to VDS_EPI_STORE + vds_qkv_post_fwd). */
  VDS_EPI_STORE_ROWDOT = 6 /* C = bf16(acc) ; rowdot[(b*(N/128) + n/128) * rows_per_batch + r] += sum over the 128-column
                            * head of C*aux (fp32; C2 = float* rowdot, zero-initialised by the caller): the dgrad GEMM
                            * that produces dO also emits delta = rowsum(dO*O) of the attention backward
                            * (model.py:136 backward).  2-CTA tile path only: VDS_ERR_UNSUPPORTED otherwise. */
};

typedef struct vds_gemm_args {
  const void* A;
  const void* B;
  int64_t lda, ldb;
  int32_t M, N, K;
  int32_t a_mn, b_mn;
  int32_t epilogue; /* vds_gemm_epilogue */
  int32_t splits;   /* split-K factor, only with VDS_EPI_ACCUM_F32 (>=1) */
  void* C;
  int64_t ldc;
  void* C2;
  int64_t ldc2;
  const void* bias; /* bf16 [N] or NULL */
  const void* aux;  /* bf16 [M, ldaux] */
  int64_t ldaux;
  const void* gate; /* bf16, gate[b * gate_stride + n], b = row / rows_per_batch */
  int64_t gate_stride;
  int32_t rows_per_batch;
  /* output row remap (VDS_EPI_STORE): out_row = (r / remap_rows) * remap_stride + remap_offset +
   * r % remap_rows; remap_rows == 0 -> identity.  Used to write patch tokens behind the register
   * tokens (model.py:362) without a concat copy. */
  int32_t remap_rows, remap_stride, remap_offset;
  int32_t tile_n; /* 0 = automatic (128 or 256), 128 = force 128-wide tiles (tuning / tests) */
  int32_t cluster; /* 0 = automatic (2-CTA cta_group::2 256x256 tiles when N % 256 == 0 and the grid fills the chip),
                      1 = 1-CTA tiles only, 2 = 1-CTA MMAs in 2-CTA clusters with multicast B (experiment) */
  /* VDS_EPI_QKV_ROPE only (ABI version 2) */
  const float* rope_tab; /* fp32 [rows_per_batch + 32, 4, 32]: vds_rope_pack output */
  const void* v0;        /* bf16 [M, ldv0]: block 0's v (value residual), or NULL */
  int64_t ldv0;
  const void* lambda_;   /* bf16 scalar on the device (lambda_param), with v0 */
} vds_gemm_args;

int vds_gemm(const vds_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------- index / elementwise
 * All replace chains of ATen elementwise / copy kernels at the cited reference lines. */

/* x[B,C,T,H,W] -> A[B*N, C*pt*p*p]; token (h' w' t'), feature ((c*pt+dt)*p+dh)*p+dw.  model.py:173-185.
 * noise/t non-NULL: also forms z_t = x*(1-t[b]) + noise*t[b] in bf16 arithmetic.     train.py:115-116 */
int vds_patchify(const void* x, const void* noise, const void* t, void* out, int B, int C, int T, int H, int W,
                 int p, int pt, void* stream);
/* tokens [B*N, p*p*pt*C] -> [B,C,T,H,W] (to_tokens=0) or the inverse gather (to_tokens=1).  model.py:392-401 */
int vds_unpatchify(const void* src, void* dst, int B, int C, int T, int H, int W, int p, int pt, int to_tokens,
                   void* stream);
/* cos/sin rows [L, D] fp32 from the persistent tables [tmax,hmax,wmax,D] (fp32 or bf16) at the random start
 * offsets; rows < n_reg are the identity rotation; rows flattened "(t h w)".           model.py:219-263 */
int vds_rope_rows(const void* tcos, const void* tsin, int table_is_bf16, float* ocos, float* osin, int L, int D,
                  int n_reg, int Tp, int Hp, int Wp, int st, int sh, int sw, int hmax, int wmax, const int* starts_dev,
                  void* stream);   /* starts_dev != NULL: (t,h,w) offsets read from device memory (graph replays) */
/* cos/sin rows [L, 64] -> the packed table VDS_EPI_QKV_ROPE reads by TMA: tab[l][s][0:16] = cos[l % L][16 s : 16 s + 16],
 * tab[l][s][16:32] = sin[...] for l in [0, L + 32) (32 wrap-around rows: a 32-token box may cross a sample boundary). */
int vds_rope_pack(const float* cos, const float* sin, float* tab, int L, void* stream);
/* [cos(t f_i) | sin(t f_i)], bf16 out.                                                  model.py:12-22 */
int vds_timestep_embedding(const void* t, void* out, int B, int dim, float max_period, void* stream);
/* SiLU and its backward (nn.SiLU at model.py:90,320,340). */
int vds_silu(const void* x, void* y, int64_t n, void* stream);
int vds_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, void* stream);

/* y = bf16(bf16(bf16(x*rstd[*w]) * bf16(1+scale[b])) + shift[b]); optional row map (final norm reads rows
 * in_row_offset.. of each sample and writes compact rows).              model.py:34-41,123,144,164,386-389 */
int vds_rmsnorm_mod_fwd(const void* x, void* y, float* rstd, const void* weight, const void* scale,
                        const void* shift, int64_t mod_stride, int B, int rows_per_batch_out, int in_batch_stride,
                        int in_row_offset, int h, float eps, void* stream);
/* backward of the above: dx (+ dx_res), dscale/dshift [B,h] (fp32, +=), dweight [h] (fp32, +=). */
int vds_rmsnorm_mod_bwd(const void* dy, const void* x, const float* rstd, const void* weight, const void* scale,
                        const void* dx_res, void* dx, float* dscale, float* dshift, float* dweight,
                        int64_t mod_stride, int64_t dmod_stride, int B, int rows_per_batch_out, int in_batch_stride,
                        int in_row_offset, int dx_full_rows, int h, void* stream);
/* backward of x + o*gate[b]: d_o = dx*gate[b]; dgate[b] += sum_rows dx*o.               model.py:139,160,165 */
int vds_gate_bwd(const void* dx, const void* o, const void* gate, void* d_o, float* dgate, int64_t gate_stride,
                 int64_t dgate_stride, int B, int rows_per_batch, int h, void* stream);
/* in place on qkv [B,L,3h]: RoPE(q), RoPE(k) (model.py:132-134,266-275); vmix = l*v + (1-l)*v0 (model.py:130). */
int vds_qkv_post_fwd(void* qkv, const float* cos, const float* sin, const void* v0, int64_t v0_ld, void* vmix,
                     const void* lambda, int B, int L, int h, int nh, void* stream);
/* backward of the above, in place on dqkv (dq optionally read from the fp32 accumulation buffer). */
int vds_qkv_post_bwd(void* dqkv, const float* dq_acc, const float* cos, const float* sin, const void* qkv_pre,
                     const void* v0, int64_t v0_ld, const void* lambda, float* dlambda, float* dv0_acc, int mode,
                     int B, int L, int h, int nh, void* stream);
/* out[n] += sum_rows x[row,n] (bias gradients); out[r,:] += sum_b x[b,r,:] (register-token gradient). */
int vds_colsum(const void* x, float* out, int64_t rows, int n, int64_t ld, void* stream);
int vds_batch_rowsum(const void* x, float* out, int B, int64_t batch_stride, int rows, int h, void* stream);
int vds_cast_f32_bf16(const float* x, void* y, int64_t n, float scale, void* stream);
/* same for a 2-D block with leading dimensions: x fp32 [rows, cols] (ldx) -> y bf16 [rows, cols] (ldy); used to drop a
 * block's dK / dV of the cross-attention (model.py:149-157 backward) into its column slice of the grouped context_kv gradient */
int vds_cast_f32_bf16_2d(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols, float scale,
                         void* stream);
int vds_accum_bf16_f32(const void* x, float* y, int64_t n, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------ attention
 * F.scaled_dot_product_attention (model.py:136 self, model.py:157 cross) and its autograd, head_dim 128,
 * on tcgen05/TMEM with TMA producers.  q/k/v/out are token-major [B, L, ld] buffers with head i at column
 * i*128 (so "(k h d)" / "b h l d -> b l (h d)" rearranges, model.py:126,137, need no copies).
 * lse: [B, nh, Lq] fp32 in the log2 domain.  Backward: dq_acc is a zero-initialised fp32 buffer reduced
 * into with red.global.add; with q_splits > 1 dk/dv are reduced into fp32 dk_acc/dv_acc instead.
 * delta [B, nh, Lq] fp32 = rowsum(dO * O) per head: computed here from o and d_o, or, with o == NULL, taken as
 * given (the dgrad GEMM that produced d_o wrote it through VDS_EPI_STORE_ROWDOT). */
int vds_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                 int64_t ldo, float* lse, int B, int nh, int Lq, int Lk, int head_dim, float scale, void* stream);
int vds_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                 int64_t ldo, const void* d_o, int64_t lddo, const float* lse, float* delta, float* dq_acc,
                 int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* dk_acc, float* dv_acc,
                 int64_t ldkv_acc, int q_splits, int B, int nh, int Lq, int Lk, int head_dim, float scale,
                 void* tail_ws, int64_t tail_ws_bytes, void* stream);
/* tail_ws (optional, q_splits == 1): fp32 workspace of vds_attn_bwd_tail_ws_bytes() bytes that lets the kernel split the
 * items of the partly-filled last wave along the query range (tail balancing).  It must be ZERO-FILLED once after it is
 * allocated; every call hands it back zero-filled (the bf16 fix-up clears what it reads), so it is never memset per call. */
int64_t vds_attn_bwd_tail_ws_bytes(int B, int nh, int Lk);
/* host-only (no GPU work): the tail plan vds_attn_bwd uses for the CTA-pair kernel — `n_pairs` kv-tile pairs (less than one
 * wave) of `n_qsub` 64-row query sub-tiles each, cut along the query range so that `clusters` SM pairs finish together.
 * Writes up to `cap` pieces (pair | first sub-tile << 10 | sub-tile count << 21, longest first = launch order) and returns
 * their number; 0 = the pairs run unsplit. */
int vds_attn_bwd_tail_plan(int n_pairs, int n_qsub, int clusters, uint32_t* pieces, int cap);

/* tuning aid: per-iteration clock64 trace of one CTA of attn_bwd (NULL disables). */
int vds_debug_attn_bwd_trace(void* buf);
/* tests / tuning: which backward kernel serves self-attention-sized problems (q_splits == 1, bf16 dk / dv).
 * -1: VDS_ATTN_PAIR environment variable (default auto); 0: the 1-CTA kernel only; 1: auto — whole waves of kv-tile
 * pairs on the 2-CTA cluster kernel (tcgen05 cta_group::2), the rest on the 1-CTA kernel; 2: every pair on the cluster kernel. */
int vds_debug_attn_pair_mode(int mode);
/* tuning aid: CTA 0 of the 2-CTA GEMM writes {total, wait(tmem empty), wait(smem full), tiles, epi wait, epi busy} cycles. */
int vds_debug_gemm2_trace(void* buf);

/* ------------------------------------------------------------------------------------------ loss / optimizer
 * loss_sum += mean_b mean_rest (bf16(x-noise) - out)^2 ; d_out = 2(out - v)/(B*per) * grad_scale
 * (* grad_scale_dev[0] if non-NULL: the upstream autograd scalar, read on the device).  train.py:117-125 */
int vds_loss_fwd_bwd(const void* x, const void* noise, const void* out, void* d_out, float* loss_sum,
                     float* loss_batch, int B, int64_t per_sample, float grad_scale, const float* grad_scale_dev,
                     void* stream);
/* torch.optim.AdamW(fused=True) math over flat fp32 p/g/m/v with a chunk table (chunks never straddle tensors;
 * chunk_group indexes the host arrays lr/wd), optional bf16 copy of the updated parameters.  train.py:340-344,433 */
int vds_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const int64_t* chunk_start,
              const int32_t* chunk_len, const int32_t* chunk_group, int n_chunks, const float* lr_host,
              const float* wd_host, int n_groups, float beta1, float beta2, float eps, int step, float grad_scale,
              const float* hyper_dev, void* stream); /* hyper_dev != NULL: [lr[16]|wd[16]|bc1|sqrt(bc2)] on the device */

#ifdef __cplusplus
}
#endif
#endif /* VDS_B200_H */
