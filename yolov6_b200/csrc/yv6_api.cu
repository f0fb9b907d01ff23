// yv6_api.cu -- handle lifetime and error plumbing of the C ABI (include/yv6.h).
#include <cstdarg>
#include <cstdlib>

#include "yv6_common.cuh"
#include "yv6_handle.h"

static thread_local char g_err[1024] = "";

void yv6_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* yv6_last_error(void) { return g_err; }
extern "C" int yv6_abi_version(void) { return YV6_ABI_VERSION; }

extern "C" int yv6_create(int device, yv6_handle** out) {
  YV6_REQUIRE(out != nullptr, "yv6_create: null out pointer");
  *out = nullptr;
  int count = 0;
  YV6_CHECK_CUDA(cudaGetDeviceCount(&count));
  YV6_REQUIRE(device >= 0 && device < count, "yv6_create: device %d out of range (%d visible)", device, count);
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  struct Restore {           // the caller's current device is left as it was (torch keeps its own notion of it)
    int d;
    ~Restore() { if (d >= 0) cudaSetDevice(d); }
  } restore{prev_device};
  YV6_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  YV6_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    yv6_set_error("yv6_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
    return YV6_ERR_STATE;
  }
  yv6_handle* h = static_cast<yv6_handle*>(calloc(1, sizeof(yv6_handle)));
  YV6_REQUIRE(h != nullptr, "yv6_create: out of host memory");
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    free(h);
    yv6_set_error("yv6_create: cannot resolve cuTensorMapEncodeTiled (%s)", cudaGetErrorString(e));
    return YV6_ERR_CUDA;
  }
  h->encode_tiled = reinterpret_cast<yv6_encode_tiled_fn>(fn);
  h->scratch_bytes = 1 << 20;
  e = cudaMalloc(&h->scratch, h->scratch_bytes);
  if (e != cudaSuccess) {
    free(h);
    yv6_set_error("yv6_create: cudaMalloc(scratch) failed: %s", cudaGetErrorString(e));
    return YV6_ERR_CUDA;
  }
  cudaMemset(h->scratch, 0, h->scratch_bytes);
  *out = h;
  return YV6_OK;
}

extern "C" int yv6_destroy(yv6_handle* h) {
  if (h == nullptr) return YV6_OK;
  if (h->scratch) cudaFree(h->scratch);
  free(h);
  return YV6_OK;
}
