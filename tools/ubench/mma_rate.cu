// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, bf16 operands from shared memory,
// M = 128) as a function of N, of the accumulator pattern and of the descriptor pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I. -o /tmp/mma_rate tools/ubench/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolov6_b200/csrc/yv6_common.cuh"
void yv6_set_error(const char*, ...) {}
using namespace yv6;

struct Cfg { int N, accs, stages, same_desc, sbo, L, kstep_units; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)c.N);
    const uint64_t dconst = umma_smem_desc(0, (uint32_t)c.sbo, 2u);
    const uint32_t a_base = smem_u32(smem) >> 4;
    const uint32_t a_stage = (24 * 1024) >> 4, b_stage = (uint32_t)(c.N * 128) >> 4;
    const uint32_t b_base = a_base + 3 * a_stage;     // A: 3 x 24 KB, B: stages x N*128 B after it
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {              // rep 0 warms up
      __syncwarp();
      if (lane == 0) t0 = clock64();
      uint32_t st = 0, acc = 0;
      const uint64_t bconst = umma_smem_desc(0, 1024u, 2u);
      for (int i = 0; i < c.L; i += 4) {             // one "k-block": four K=16 steps inside a 128-byte swizzle row
        const uint64_t ad = dconst | (uint64_t)(a_base + st * a_stage);
        const uint64_t bd = bconst | (uint64_t)(b_base + st * b_stage);
        const uint32_t d = tmem + acc * (uint32_t)c.N;
        if (elect_one()) {
          umma_bf16(d, ad, bd, idesc, 1u);
          umma_bf16(d, ad + c.kstep_units, bd + 2, idesc, 1u);
          umma_bf16(d, ad + 2 * c.kstep_units, bd + 4, idesc, 1u);
          umma_bf16(d, ad + 3 * c.kstep_units, bd + 6, idesc, 1u);
        }
        __syncwarp();
        if (!c.same_desc && ++st == 3u) st = 0;
        if (++acc == (uint32_t)c.accs) acc = 0;
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, (uint32_t)rep & 1u);
      tc_fence_after();
      if (lane == 0) t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("%-6s %-5s %-7s %-9s %-6s %-6s | cycles/MMA (floor N/2)\n", "N", "accs", "stages", "same_desc", "sbo", "grid");
  const int Ns[] = {64, 128, 256};
  for (int grid : {1, 148})
    for (int N : Ns)
      for (int accs : {1, 2})
        for (int same : {1, 0})
          for (int sbo : {1024, 1280}) {
            if (accs * N > 512) continue;
            Cfg c{N, accs, 3, same, sbo, 2048, 2};
            mma_rate_kernel<<<grid, 128, 216 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long cyc;
            cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
            printf("%-6d %-5d %-7d %-9d %-6d %-6d | %.1f  (%d)\n", N, accs, c.stages, same, sbo, grid, (double)cyc / c.L, N / 2);
          }
  return 0;
}
