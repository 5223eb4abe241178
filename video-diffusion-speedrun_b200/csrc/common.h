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
