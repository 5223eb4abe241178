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
  VDS_EPI_STORE_F32 = 5  /* C(fp32) = acc + bias?                                               */
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
} vds_gemm_args;

int vds_gemm(const vds_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDS_B200_H */
