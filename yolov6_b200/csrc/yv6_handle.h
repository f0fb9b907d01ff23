// yv6_handle.h -- per-device context behind the opaque `yv6_handle` of include/yv6.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>

typedef CUresult (*yv6_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

struct yv6_handle {
  int device;
  int num_sms;
  int max_smem_optin;              // bytes of dynamic shared memory a CTA may opt in to
  yv6_encode_tiled_fn encode_tiled;  // resolved through cudaGetDriverEntryPoint (no -lcuda)
  void* scratch;                   // device scratch for kernels that need counters / partials
  size_t scratch_bytes;
};
