// Host-side helpers shared by all translation units of libvds_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vds_b200.h"

namespace vds {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link against libcuda).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1; bf16 elements, 128-B swizzle.
int encode_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

// general form: fp32 or bf16 elements, 128-B swizzle or none
int encode_tmap(CUtensorMap* tm, const void* base, int is_f32, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

// Every kernel of the library is launched with programmatic dependent launch allowed (env VDS_PDL=0 turns it off):
// the next kernel's CTAs may be scheduled, run their prologue (barrier init, TMEM alloc, tensor-map prefetch) and
// then block in griddepcontrol.wait (ptx.cuh: pdl_wait) until the previous kernel has completed and flushed, so the
// launch + prologue latency between the ~1200 dependent kernels of a step overlaps the previous kernel's tail.
// Contract for kernels: no global-memory access that depends on earlier kernels before pdl_wait().
int pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define VDS_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      vds::set_error(__VA_ARGS__);      \
      return VDS_ERR_ARG;               \
    }                                   \
  } while (0)

#define VDS_CHECK_LAUNCH(name)                                                         \
  do {                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) {                                                           \
      vds::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));           \
      return VDS_ERR_CUDA;                                                             \
    }                                                                                  \
    vds::count_launch();                                                               \
  } while (0)

}  // namespace vds
