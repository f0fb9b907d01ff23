// yv6_conv_wgrad.cu -- weight gradient of a 1x1 / 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[co][r][s][ci] = sum over output pixels (n,ho,wo) of dY[n,ho,wo,co] * X[n, ho*st + r - pad, wo*st + s - pad, ci]
//
// i.e. what autograd computes for `F.conv2d` in the reference's training step (Trainer.train_in_steps,
// yolov6/core/engine.py:142-176 -> cuDNN wgrad).  GEMM view per filter tap: M = Cout, N = Cin, K = pixels.
// Both operands are read straight from the NHWC activations by TMA as [128 pixels][64 channels] boxes;
// that shared-memory image is the canonical *MN-major* 128-byte-swizzled UMMA layout (channels contiguous,
// 8 pixel rows per swizzle atom), so no transpose is ever materialised:
//   A = dY box  (K = pixels x M = 64-wide Cout blocks),
//   B = X box shifted by the tap (zero fill = padding, elementStrides = conv stride).
// One CTA owns a (Cout tile, Cin tile, tap, pixel range) unit: it accumulates 128 x N fp32 in TMEM over its
// pixel tiles (8 tcgen05.mma per 128-pixel tile) and adds the result to dW with fp32 reductions
// (split-K over the pixel ranges).  warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#include <algorithm>
#include <cstdlib>

#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;
constexpr int kWgMaxTaps = 3;          // filter taps accumulated by one CTA (they share the dY tile)

struct WgParams {
  int32_t BW, BH, PT;                 // pixel box, BW*BH == PT (128 or 64): the K extent of one pipeline stage
  int32_t tiles_w, tiles_h, tiles_i, ptiles;
  int32_t co_tiles, ci_tiles, taps, kw, ksplit;
  int32_t T, tap_groups;              // taps per unit (1, or 3 = one filter row) and taps / T
  int32_t Cout, Cin, stride, pad;
  int32_t a_blocks, a_loaded;         // 64-wide dY blocks in the operand layout (2) and the ones TMA really fetches (1 if Cout <= 64)
  int32_t b_blocks, b_blk_elems, b_blk_bytes;  // X blocks of 64/32/16 channels
  int32_t NT;                         // UMMA N = b_blocks * b_blk_elems (<= 256)
  int32_t b_layout, b_sbo;            // swizzle mode / 8-row group stride of the X blocks
  int32_t stages, a_stage_bytes, b_tap_bytes, b_stage_bytes;
  int32_t tmem_cols;
  float* dw;                          // fp32 [Cout][kh*kw][Cin]
};

__device__ __forceinline__ void wg_tma_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// MN-major operand descriptor: LBO = bytes between channel blocks, SBO = bytes between 8-pixel-row groups
__device__ __forceinline__ uint64_t wg_desc(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * p.a_stage_bytes;
  float* sT = reinterpret_cast<float*>(sB + (size_t)p.stages * p.b_stage_bytes);   // epilogue transpose: 4 warps x [32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + 4 * 32 * 33);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* done = bars + 2 * kWgMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // unit decode: tap group fastest, then channel tiles, then the pixel range -- CTAs that run at the same time read the SAME
  // pixels of dY / X for different taps / channel tiles, so each operand byte is fetched from DRAM once and then served by L2
  int u = blockIdx.x;
  const int tg = u % p.tap_groups; u /= p.tap_groups;
  const int ci_t = u % p.ci_tiles; u /= p.ci_tiles;
  const int co_t = u % p.co_tiles; u /= p.co_tiles;
  const int ks = u;
  const int tap0 = tg * p.T;
  const int per = (p.ptiles + p.ksplit - 1) / p.ksplit;
  const int pt0 = ks * per, pt1 = min(p.ptiles, pt0 + per);
  const int npt = max(0, pt1 - pt0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t a_blk_bytes = (uint32_t)p.PT * 128u;                  // one 64-channel dY block: PT pixel rows of 128 bytes

  if (warp == 0) {
    const uint32_t tx = (uint32_t)p.a_loaded * a_blk_bytes + (uint32_t)(p.T * p.b_blocks * p.PT * p.b_blk_bytes);
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      int m = pt;
      const int tw = m % p.tiles_w; m /= p.tiles_w;
      const int th = m % p.tiles_h;
      const int ti = m / p.tiles_h;
      const int w0 = tw * p.BW, h0 = th * p.BH;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[stage], tx);
        for (int j = 0; j < p.a_loaded; ++j)
          wg_tma_4d(sA + (size_t)stage * p.a_stage_bytes + (size_t)j * a_blk_bytes, &tmDY, &full[stage], co_t * 128 + j * 64, w0, h0, ti);
        for (int t = 0; t < p.T; ++t) {
          const int tap = tap0 + t;
          const int r = tap / p.kw, sx = tap - r * p.kw;
          uint8_t* dst = sB + (size_t)stage * p.b_stage_bytes + (size_t)t * p.b_tap_bytes;
          for (int j = 0; j < p.b_blocks; ++j)
            wg_tma_4d(dst + (size_t)j * p.PT * p.b_blk_bytes, &tmX, &full[stage], ci_t * p.NT + j * p.b_blk_elems,
                      w0 * p.stride + sx - p.pad, h0 * p.stride + r - p.pad, ti);
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // instruction descriptor: bf16 x bf16 -> fp32, both operands MN-major (bits 15 / 16), M = 128, N = NT
    // When a tap's X tile is a single channel block (Cin tile <= 64), the T taps of the stage are just consecutive blocks of
    // one wider MN-major operand (block stride = LBO = one tap tile): ONE instruction of N = T * NT per K step replaces T
    // instructions of N = NT, so the dY operand is read once instead of T times (64-channel layers are bound by exactly that
    // shared-memory traffic).  The accumulator columns are the same (tap t at t * NT), so the epilogue does not change.
    const bool merged = (p.b_blocks == 1 && p.T > 1 && p.T * p.NT <= 256 && p.b_tap_bytes == p.PT * p.b_blk_bytes);
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)(merged ? p.T * p.NT : p.NT)) | (1u << 15) | (1u << 16);
    const uint64_t da = wg_desc(a_blk_bytes, 1024u, 2u);
    const uint64_t db = wg_desc((uint32_t)(p.PT * p.b_blk_bytes), (uint32_t)p.b_sbo, (uint32_t)p.b_layout);
    const uint32_t a_base = smem_u32(sA) >> 4, b_base = smem_u32(sB) >> 4;
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4, b_step = (uint32_t)p.b_stage_bytes >> 4, b_tap = (uint32_t)p.b_tap_bytes >> 4;
    const uint32_t a_k = 2048u >> 4, b_k = (uint32_t)(2 * p.b_sbo) >> 4;   // 16 pixels = two 8-row groups
    const int ksteps = p.PT / 16;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < npt; ++it) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint64_t ad = da | (uint64_t)(a_base + (uint32_t)stage * a_step);
      const uint64_t bd = db | (uint64_t)(b_base + (uint32_t)stage * b_step);
      if (elect_one()) {
        for (int t = 0; t < (merged ? 1 : p.T); ++t) {
          const uint64_t bt = bd + (uint64_t)((uint32_t)t * b_tap);
          const uint32_t acc = tmem_base + (uint32_t)(t * p.NT);
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(acc, ad + (uint64_t)(k * a_k), bt + (uint64_t)(k * b_k), idesc, (uint32_t)((it | k) != 0));
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // Epilogue: warp q owns TMEM lanes (= output channels) 32q..32q+31.  32x32 blocks go through a per-warp shared-memory
    // transpose so that every reduction instruction adds 32 CONSECUTIVE floats of one dW row (one 128-byte line) instead of
    // one float in each of 32 rows.
    const int q = warp & 3;
    float* tr = sT + q * 32 * 33;
    mbar_wait(done, 0);
    tc_fence_after();
    if (npt > 0) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int co_base = co_t * 128 + q * 32;
      for (int t = 0; t < p.T; ++t) {
        const int tap = tap0 + t;
        for (int c0 = 0; c0 < p.NT; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)(t * p.NT + c0), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          const int ci = ci_t * p.NT + c0 + lane;
          if (ci < p.Cin && c0 + lane < p.NT) {
            float* out = p.dw + ((int64_t)co_base * p.taps + tap) * p.Cin + ci;
            const int rows = min(32, p.Cout - co_base);
            for (int row = 0; row < rows; ++row) atomicAdd(out + (int64_t)row * p.taps * p.Cin, tr[row * 33 + lane]);
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace yv6

using namespace yv6;

extern "C" int yv6_conv_wgrad(yv6_handle* h, const yv6_wgrad_desc* d, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && d && d->x && d->dy && d->dw, "wgrad: null argument");
  YV6_REQUIRE(h->encode_tiled != nullptr, "wgrad: cuTensorMapEncodeTiled unavailable");
  YV6_REQUIRE(d->Cin > 0 && d->Cin % 16 == 0 && d->x_c_total % 8 == 0 && d->dy_c_total % 8 == 0, "wgrad: channel counts");
  YV6_REQUIRE(d->kh == d->kw && (d->kh == 1 || d->kh == 3), "wgrad: kernel %dx%d unsupported", d->kh, d->kw);
  YV6_REQUIRE(d->stride == 1 || d->stride == 2, "wgrad: stride %d", d->stride);
  const int Ho = (d->H + 2 * d->pad - d->kh) / d->stride + 1, Wo = (d->W + 2 * d->pad - d->kw) / d->stride + 1;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.Cout = d->Cout; p.Cin = d->Cin; p.stride = d->stride; p.pad = d->pad;
  p.taps = d->kh * d->kw; p.kw = d->kw;
  p.dw = d->dw;
  p.b_blk_elems = (d->Cin % 64 == 0) ? 64 : (d->Cin % 32 == 0) ? 32 : 16;
  p.b_blk_bytes = p.b_blk_elems * 2;
  p.b_layout = (p.b_blk_bytes == 128) ? 2 : (p.b_blk_bytes == 64) ? 4 : 6;
  p.b_sbo = 8 * p.b_blk_bytes;
  // 3x3: one CTA accumulates the three taps of a filter row -- they share the dY tile, which cuts the L2 -> SM operand
  // stream per MMA by 1.5x (128 channels) to 2.3x (64 channels); the Cin tile is then at most 128 wide (3 x 128 TMEM columns).
  p.T = (p.taps == 9 && d->force_taps != 1) ? 3 : 1;
  const int nt_max = (p.T == 3) ? 128 : 256;
  p.ci_tiles = (d->Cin + nt_max - 1) / nt_max;
  const int per_tile = (d->Cin + p.ci_tiles - 1) / p.ci_tiles;
  p.b_blocks = (per_tile + p.b_blk_elems - 1) / p.b_blk_elems;
  p.NT = p.b_blocks * p.b_blk_elems;
  YV6_REQUIRE(p.NT % 16 == 0 && p.NT <= nt_max, "wgrad: N tile %d", p.NT);
  p.tap_groups = p.taps / p.T;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.T * p.NT) p.tmem_cols <<= 1;
  p.PT = (p.T == 3) ? 64 : 128;
  // pixel box: exactly PT rows (partial boxes are zero filled by TMA, so the K sum stays exact)
  long best = -1;
  for (int bw = 1; bw <= p.PT; bw <<= 1) {
    const int bh = p.PT / bw;
    if (bw * d->stride > 256 || bh * d->stride > 256) continue;
    const long t = (long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * d->N;
    if (best < 0 || t < best || (t == best && bw > p.BW)) { best = t; p.BW = bw; p.BH = bh; }
  }
  p.tiles_w = (Wo + p.BW - 1) / p.BW; p.tiles_h = (Ho + p.BH - 1) / p.BH; p.tiles_i = d->N;
  p.ptiles = p.tiles_w * p.tiles_h * p.tiles_i;
  p.a_blocks = 2;
  p.a_loaded = (d->Cout <= 64) ? 1 : 2;     // rows 64..127 of the accumulator are never stored for Cout <= 64
  p.co_tiles = (d->Cout + 127) / 128;
  // Cout <= 64: only the first 64-channel block exists; the MMA still reads M = 128 rows, the upper 64 from whatever follows in
  // shared memory (the next stage / the X tiles) -- they land in accumulator rows that are never stored
  p.a_stage_bytes = p.a_loaded * p.PT * 128;
  p.b_tap_bytes = ((p.b_blocks * p.PT * p.b_blk_bytes + 1023) / 1024) * 1024;
  p.b_stage_bytes = p.T * p.b_tap_bytes;
  const int fixed = 1024 /* alignment */ + 4 * 32 * 33 * 4 /* epilogue transpose */ + 512 /* barriers */;
  const int budget = h->max_smem_optin - fixed;
  p.stages = std::min(kWgMaxStages, budget / (p.a_stage_bytes + p.b_stage_bytes));
  YV6_REQUIRE(p.stages >= 2, "wgrad: not enough shared memory");
  const int base_units = p.co_tiles * p.ci_tiles * p.tap_groups;
  // split-K over pixel ranges: the kernel is bound by the L2 -> shared-memory operand stream of each CTA (measured: the time
  // per 64-pixel stage does not depend on how many CTAs run), so it wants exactly two full waves of CTAs -- one unit more than
  // a multiple of the SM count costs a whole extra wave -- with at least ~8 stages of work each (every extra split adds a
  // pass of fp32 reductions over the weight tensor)
  const int min_tiles = 8 * 128 / p.PT;
  p.ksplit = std::max(1, std::min((p.ptiles + min_tiles - 1) / min_tiles, std::max(1, 2 * h->num_sms / base_units)));
  if (d->force_ksplit > 0) p.ksplit = std::min(p.ptiles, d->force_ksplit);
  const int units = base_units * p.ksplit;

  CUtensorMap tmDY, tmX;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d->N};
    const uint64_t pix = (uint64_t)d->dy_c_total * 2;
    cuuint64_t strides[3] = {pix, pix * Wo, pix * Wo * Ho};
    cuuint32_t box[4] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = h->encode_tiled(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->dy), dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { yv6_set_error("wgrad: cuTensorMapEncodeTiled(dY) failed with %d", (int)cr); return YV6_ERR_CUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    const uint64_t pix = (uint64_t)d->x_c_total * 2;
    cuuint64_t strides[3] = {pix, pix * d->W, pix * d->W * d->H};
    cuuint32_t box[4] = {(cuuint32_t)p.b_blk_elems, (cuuint32_t)(p.BW * d->stride), (cuuint32_t)(p.BH * d->stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    const CUtensorMapSwizzle sw = (p.b_blk_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (p.b_blk_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult cr = h->encode_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { yv6_set_error("wgrad: cuTensorMapEncodeTiled(X) failed with %d", (int)cr); return YV6_ERR_CUDA; }
  }
  if (!(h->configured & YV6_CFG_WGRAD)) {
    YV6_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem_optin));
    h->configured |= YV6_CFG_WGRAD;
  }
  const size_t smem = (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + fixed;
  conv_wgrad_kernel<<<units, kWgThreads, smem, (cudaStream_t)stream>>>(tmDY, tmX, p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
