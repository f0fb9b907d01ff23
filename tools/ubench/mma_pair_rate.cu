// Microbenchmark (NOT YET RUN ON HARDWARE -- written at the end of round 1 for the first GPU call of round 2):
// issue rate of tcgen05.mma.cta_group::2 (CTA pair, M = 256 over two SMs, bf16 operands from shared memory; each CTA holds its
// 128 rows of A and N/2 columns of B) as a function of N.  Expected floor per the B300 notes: max(M_atom,128)*N/(256*2) clocks
// = N/2 for a pair, i.e. twice the per-SM work of cta_group::1 in the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mma_pair_rate tools/ubench/mma_pair_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolov6_b200/csrc/yv6_common.cuh"
void yv6_set_error(const char*, ...) {}
using namespace yv6;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of both CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

struct Cfg { int N, L; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_pair_kernel(Cfg c, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  cluster_sync_all();
  if (warp == 0) tmem_alloc2(&tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (rank == 0) {                                   // only the leader CTA issues; operands are read from both CTAs' smem
      const uint32_t idesc = umma_idesc_bf16(256, (uint32_t)c.N);
      const uint64_t dconst = umma_smem_desc(0, 1024u, 2u);
      const uint32_t a_base = smem_u32(smem) >> 4;                // A: 3 stages x 16 KB (this CTA's 128 rows x K = 64)
      const uint32_t b_base = a_base + ((3 * 16 * 1024) >> 4);    // B: 3 stages x (N/2 rows x 128 B)
      const uint32_t a_stage = (16 * 1024) >> 4, b_stage = (uint32_t)(c.N / 2 * 128) >> 4;
      for (int rep = 0; rep < 2; ++rep) {
        __syncwarp();
        if (lane == 0) t0 = clock64();
        uint32_t st = 0;
        for (int i = 0; i < c.L; i += 4) {
          const uint64_t ad = dconst | (uint64_t)(a_base + st * a_stage);
          const uint64_t bd = dconst | (uint64_t)(b_base + st * b_stage);
          if (elect_one()) {
            umma2_bf16(tmem, ad, bd, idesc, 1u);
            umma2_bf16(tmem, ad + 2, bd + 2, idesc, 1u);
            umma2_bf16(tmem, ad + 4, bd + 4, idesc, 1u);
            umma2_bf16(tmem, ad + 6, bd + 6, idesc, 1u);
          }
          __syncwarp();
          if (++st == 3u) st = 0;
        }
        if (elect_one()) umma2_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, (uint32_t)rep & 1u);
        tc_fence_after();
        if (lane == 0) t1 = clock64();
      }
      if (lane == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    } else {
      for (int rep = 0; rep < 2; ++rep) mbar_wait(&bar, (uint32_t)rep & 1u);   // the peer just waits for the pair's commits
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%-6s %-6s | cycles per tcgen05.mma.cta_group::2 (M=256, K=16); floor N/2\n", "N", "grid");
  for (int grid : {2, 148})
    for (int N : {64, 128, 256}) {
      Cfg c{N, 2048};
      mma_pair_kernel<<<grid, 128, 196 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long cyc;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("%-6d %-6d | %.1f  (%d)\n", N, grid, (double)cyc / c.L, N / 2);
    }
  return 0;
}
