// yv6_conv_igemm.cu -- fused implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Computes, for NHWC bf16 activations and KRSC bf16 weights,
//     y = act(conv(x, w) + bias) [+ alpha * residual]
// which in deploy form is every 3x3 / 1x1 conv of the YOLOv6 backbones, necks and head:
// ConvModule.forward_fuse (reference yolov6/layers/common.py:50-54), RepVGGBlock with
// rbr_reparam (common.py:247-248), BottleRep's shortcut (common.py:605-608), the 1x1s of
// BiFusion/BepC3/CSPSPPF, the head stems and preds (effidehead.py:72-118) and, through the
// output strides, torch.cat (reppan.py:228,232) and ConvTranspose2d k2s2 (common.py:181-194).
//
// Design (B200-first, im2col-free):
//   * GEMM view: M = output pixels, N = Cout, K = taps x Cin.  One M tile is a BW x BH x BI box of
//     output pixels (<= 128 rows).  For filter tap (r,s) the A tile is that same box of the INPUT
//     shifted by (r-pad, s-pad): one TMA box load from a 5-D tensor map (C, W, H, N, plane) with
//     zero fill out of bounds (= conv padding) and elementStrides = conv stride.  The box lands in
//     shared memory as `rows` consecutive K-major rows in the 128/64/32-byte swizzle that
//     tcgen05.mma consumes directly -- no im2col buffer, no index math on the SMs.
//   * B tile = [BN x KB] slice of the weights, one TMA load from a 3-D map (K, Cout, plane).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (single thread, tcgen05.mma cta_group::1,
//     M=128, N=BN<=256, K=16), warps 2..5 = epilogue (tcgen05.ld -> bias/act/residual -> global).
//   * persistent CTAs (grid = #SMs) over a static tile schedule; smem ring of `stages` K-blocks
//     (producer runs ahead across tiles), double-buffered TMEM accumulators so the epilogue of
//     tile i overlaps the mainloop of tile i+1.
//   * bf16x3 mode (nsplit=3): same kernel, K loop additionally runs over six (plane_a, plane_b)
//     pairs, giving fp32-equivalent products with fp32 accumulation.
//   * halo variants (MODE 1 / 2: 3x3 stride 1; MODE 3 / 4: 3x3 stride 2 on the column-pair view of the input): ONE input box
//     per channel block feeds every tap through shifted UMMA descriptors; weights streamed through their own ring or resident.
//   * CTA-pair variants (CP = true): clusters of two CTAs, tcgen05 cta_group::2, M = 256 per instruction, half a weight tile
//     staged per CTA.  DESIGN.md section 4 has the measurements behind each of these.
#include <algorithm>
#include <cstdarg>

#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

// warp 0 producer, warp 1 MMA, then G epilogue groups of four warps, one per TMEM accumulator:
// G = 2 (320 threads) in general, G = 4 (576 threads) when 4 accumulators fit (BN <= 128): small-K tiles are
// bound by the instruction latency of the epilogue, which more resident warps hide.
constexpr int kMaxGroups = 4;
constexpr int kMaxStages = 12;
constexpr int kTileRows = 128;

struct ConvKParams {
  int32_t BW, BH, BI, rows;
  int32_t tiles_w, tiles_h, tiles_i, tiles_n, num_tiles;
  int32_t N, Ho, Wo, Cout, BN;
  int32_t taps, kw, stride, stride_w, pad, pad_w, Cin;   // stride / pad = rows (h), stride_w / pad_w = columns
  int32_t cin_blocks, kb_elems, kb_bytes, ksteps;
  int32_t sbo_bytes, layout_type;
  int32_t npairs;
  int32_t stages, a_stage_bytes, b_stage_bytes;
  int32_t tmem_cols, groups;
  int32_t act, y_dtype, out_planes, res_planes;
  void* y;
  int64_t y_img_stride, y_h_stride, y_w_stride, y_plane_stride;
  const __nv_bfloat16* res;
  float alpha;
  int64_t res_img_stride, res_h_stride, res_w_stride, res_plane_stride;
  const float* bias;
  int32_t tma_store;      // 1: stage the tile in smem and TMA-store it, 0: direct global stores
  int32_t c_chunk;        // output columns per staged chunk: 64 (bf16) or 32 (fp32) -> 128-byte rows
  // halo mode (3x3 stride-1, Cin % 64 == 0): one (BH+2)x(BW+2) input box per channel block feeds all
  // nine taps through UMMA descriptors offset into it; B tiles ride their own ring (or stay resident).
  int32_t halo, a_stages, b_stages, b_resident;   // halo: 1 = 3x3 stride 1, 2 = column-pair view of a 3x3 stride-2 conv
  int32_t skip_cb;                                // halo 2: the first skip_cb channel blocks of the s = 0 taps are all-zero weights (skipped)
  int32_t a_region_bytes, b_region_bytes;  // smem carve: [A ring][B ring][2 C buffers][barriers]
  // CTA-pair mode (cluster of two CTAs, tcgen05 cta_group::2): one schedule unit = two consecutive M tiles (one per CTA)
  // x one N tile; every CTA stages its own input tile and HALF of the weight tile (b_rows = BN / 2 rows), one
  // M256 x BN x K16 instruction issued by the even CTA feeds both accumulators.  Halves the shared-memory traffic of
  // the weight operand per FLOP, which is what bounds the single-CTA kernel (A + B operand reads plus the TMA fills
  // exceed 128 B/clk per SM at every BN).
  int32_t cpair, m_tiles, b_rows;
  int32_t bias_smem;                       // bias[0 .. tiles_n*BN) is staged in shared memory by the epilogue warps
  int32_t res_aligned;                     // residual rows are 16-byte aligned (channel offset / pitches % 8 == 0)
  int32_t fast_act;                        // bf16 outputs: SiLU through tanh.approx (rel. error 2^-11 < bf16 ulp)
  unsigned long long* trace;               // debug: clock64 stamps of CTA 0 (16 slots) or null
};
#define YV6_TRACE(slot) do { if (p.trace != nullptr && blockIdx.x == 0) p.trace[slot] = (unsigned long long)clock64(); } while (0)

constexpr int kHaloW = 10, kHaloH = 18;                   // BW = 8, BH = 16 output tile + 1-pixel border
constexpr int kHaloBytes = kHaloW * kHaloH * 128;         // 23040
constexpr int kHaloStageBytes = 24 * 1024;                // rounded up to the 1024-byte swizzle atom
// Halo geometry of a kernel variant.  S2 = false: 3x3 stride 1 (above).  S2 = true: 3x2 kernel with stride (2, 1) and padding
// (1, 1) -- a 3x3 stride-2 conv on the column-pair view [N, H, W/2, 2C] of its input: the same 8 x 16 output tile reads a
// 9 x 33 box; tap (r, s) starts (9 r + s) pixels into it and consecutive output rows are TWO box rows apart (the UMMA
// descriptor's stride between 8-row groups), so the stride costs nothing at issue time.
template <bool S2>
struct HaloGeom {
  static constexpr int W = S2 ? 9 : kHaloW, H = S2 ? 33 : kHaloH, KW = S2 ? 2 : 3, TAPS = S2 ? 6 : 9, SH = S2 ? 2 : 1;
  static constexpr int kBytes = W * H * 128;
  static constexpr int kStageBytes = ((kBytes + 1023) / 1024) * 1024;     // 24 KB / 38 KB
};
constexpr int kMaxAStages = 6, kMaxBStages = 40;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int group) {  // the 128 threads of one epilogue group
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

struct TileCoord {
  int w0, h0, i0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  int nt = tile % p.tiles_n;
  int m = tile / p.tiles_n;
  int tw = m % p.tiles_w;
  m /= p.tiles_w;
  int th = m % p.tiles_h;
  int ti = m / p.tiles_h;
  t.w0 = tw * p.BW;
  t.h0 = th * p.BH;
  t.i0 = ti * p.BI;
  t.n0 = nt * p.BN;
  return t;
}

// CTA-pair mode: unit -> (M tile 2u + rank, N tile).  A missing second tile (odd tile count) is placed on image N: its TMA
// loads are zero filled and its TMA stores dropped, so it needs no special case anywhere.
__device__ __forceinline__ TileCoord decode_unit(const ConvKParams& p, int unit, int rank) {
  TileCoord t;
  const int nt = unit % p.tiles_n;
  int m = (unit / p.tiles_n) * 2 + rank;
  t.n0 = nt * p.BN;
  if (m >= p.m_tiles) {
    t.w0 = 0;
    t.h0 = 0;
    t.i0 = p.N;
    return t;
  }
  const int tw = m % p.tiles_w;
  m /= p.tiles_w;
  const int th = m % p.tiles_h;
  const int ti = m / p.tiles_h;
  t.w0 = tw * p.BW;
  t.h0 = th * p.BH;
  t.i0 = ti * p.BI;
  return t;
}

// bf16x3 plane pairs, smallest products first; the plain bf16 mode uses the last entry only.
__constant__ int kPairA[6] = {0, 1, 2, 0, 1, 0};
__constant__ int kPairB[6] = {2, 1, 0, 1, 0, 0};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Store 16 consecutive output channels of one pixel.
__device__ __forceinline__ void store_chunk(const ConvKParams& p, int64_t off, int n, int ncol,
                                            const float (&v)[16]) {
  if (p.y_dtype == YV6_DT_F32) {
    float* y = reinterpret_cast<float*>(p.y) + off + n;
    if (ncol == 16 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
      float4* y4 = reinterpret_cast<float4*>(y);
#pragma unroll
      for (int j = 0; j < 4; ++j) y4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      for (int j = 0; j < ncol; ++j) y[j] = v[j];
    }
    return;
  }
  float rem[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rem[j] = v[j];
  for (int pl = 0; pl < p.out_planes; ++pl) {
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + pl * p.y_plane_stride + off + n;
    if (ncol == 16 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
      uint4 q0, q1;
      q0.x = pack_bf16x2(rem[0], rem[1]);
      q0.y = pack_bf16x2(rem[2], rem[3]);
      q0.z = pack_bf16x2(rem[4], rem[5]);
      q0.w = pack_bf16x2(rem[6], rem[7]);
      q1.x = pack_bf16x2(rem[8], rem[9]);
      q1.y = pack_bf16x2(rem[10], rem[11]);
      q1.z = pack_bf16x2(rem[12], rem[13]);
      q1.w = pack_bf16x2(rem[14], rem[15]);
      reinterpret_cast<uint4*>(y)[0] = q0;
      reinterpret_cast<uint4*>(y)[1] = q1;
    } else {
      for (int j = 0; j < ncol; ++j) y[j] = __float2bfloat16_rn(rem[j]);
    }
    if (pl + 1 < p.out_planes) {
#pragma unroll
      for (int j = 0; j < 16; ++j) rem[j] -= __bfloat162float(__float2bfloat16_rn(rem[j]));
    }
  }
}

constexpr int kCBufBytes = kTileRows * 128;  // one staged output chunk: 128 rows x 128 B
constexpr int kCBufCount = 4;               // two per epilogue group

constexpr int kBiasSmemFloats = 1024;        // bias staged in shared memory when the layer's channels fit

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// bias + activation (+ residual) for 16 consecutive output channels of one pixel
__device__ __forceinline__ void epilogue_math(const ConvKParams& p, const float* sbias, const uint32_t (&r)[16], int n, int ncol,
                                              bool valid, int64_t roff, float (&v)[16]) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // shared-memory copy (broadcast LDS) when staged; the global path costs an exposed L2 round trip per chunk
      const float4 b = sbias ? *reinterpret_cast<const float4*>(sbias + n + 4 * j)
                             : __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
      v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  }
  if (p.act == YV6_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (p.act == YV6_ACT_SILU) {
    if (p.fast_act) {                       // x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU per element
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = 0.5f * v[j];
        v[j] = fmaf(h, tanh_fast(h), h);
      }
    } else {                                // ex2 / rcp approximations are ~1e-7 relative
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] * rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * v[j]));
    }
  } else if (p.act == YV6_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * v[j]));
  }
  if (p.res != nullptr && valid && ncol > 0) {
    for (int pl = 0; pl < p.res_planes; ++pl) {
      const __nv_bfloat16* rp = p.res + pl * p.res_plane_stride + roff + n;
      if (ncol == 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(rp));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
          v[2 * j] += p.alpha * __low2float(b2);
          v[2 * j + 1] += p.alpha * __high2float(b2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncol) v[j] += p.alpha * __bfloat162float(rp[j]);
      }
    }
  }
}

// Fast epilogue step: 32 accumulator columns of this thread's row -> bias, activation, optional residual ->
// one staged (swizzled) row segment.  Straight-line code: a single TMEM load, bias from shared memory,
// the activation chosen once per call, 128-bit shared stores.  `base` = shared address of the row,
// `rxor` = row & 7 (128-byte swizzle), `unit0` = first 16-byte unit of the segment within the row.
template <bool F32>
__device__ __forceinline__ void epi_cols32(const ConvKParams& p, const float* sbias, uint32_t taddr, int n,
                                           const __nv_bfloat16* res_row, uint32_t base, uint32_t rxor, int unit0) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  uint4 rq[4];
  if (res_row != nullptr) {               // issued before the TMEM wait: the L2 round trip overlaps it
#pragma unroll
    for (int j = 0; j < 4; ++j) rq[j] = __ldg(reinterpret_cast<const uint4*>(res_row + n) + j);
  }
  tmem_ld_wait();
  float v[32];
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = sbias ? *reinterpret_cast<const float4*>(sbias + n + 4 * j)
                             : __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
      v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  }
  if (p.act == YV6_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (p.act == YV6_ACT_SILU) {
    if (p.fast_act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float h = 0.5f * v[j];
        v[j] = fmaf(h, tanh_fast(h), h);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] * rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * v[j]));
    }
  } else if (p.act == YV6_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * v[j]));
  }
  if (res_row != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {rq[j].x, rq[j].y, rq[j].z, rq[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[8 * j + 2 * k] = fmaf(p.alpha, __low2float(b2), v[8 * j + 2 * k]);
        v[8 * j + 2 * k + 1] = fmaf(p.alpha, __high2float(b2), v[8 * j + 2 * k + 1]);
      }
    }
  }
  if (F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = base + ((((uint32_t)(unit0 + j)) ^ rxor) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]),
                   "f"(v[4 * j + 3])
                   : "memory");
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t a = base + ((((uint32_t)(unit0 + j)) ^ rxor) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])),
                   "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                   "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                   : "memory");
    }
  }
}

// MODE: 0 = one TMA box per (tap, channel block); 1 = halo input tiles, weights streamed through a ring;
//       2 = halo input tiles, the layer's weights resident in shared memory.  Compile-time so that each
//       variant carries only its own producer / issue loops (instruction-cache footprint, issue-slot count).
//   CP: CTA-pair variant (launched as clusters of two CTAs): see ConvKParams::cpair.
template <int G, int MODE, bool CP>
__global__ void __launch_bounds__(64 + 128 * G, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.a_region_bytes;
  uint8_t* sC = sB + (size_t)p.b_region_bytes;
  float* sBiasBuf = reinterpret_cast<float*>(sC + kCBufCount * kCBufBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + kCBufCount * kCBufBytes + kBiasSmemFloats * sizeof(float));
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + kMaxGroups;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + kMaxGroups);
  uint64_t* a_full = tempty + kMaxGroups + 2;            // halo mode rings
  uint64_t* a_empty = a_full + kMaxAStages;
  uint64_t* b_full = a_empty + kMaxAStages;
  uint64_t* b_empty = b_full + kMaxBStages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool HALO = (MODE != 0);
  constexpr bool BRES = (MODE == 2 || MODE == 4);
  using HG = HaloGeom<(MODE >= 3)>;
  // schedule: CTA (or CTA pair) takes units unit0, unit0 + ustep, ...; in pair mode this CTA computes M tile `rank` of a unit
  const int rank = CP ? (int)cluster_ctarank() : 0;
  const int unit0 = CP ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = CP ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool leader = (rank == 0);
  if (threadIdx.x == 0) YV6_TRACE(0);

  // Prologue, kept short because ~50 of YOLOv6-S's 73 conv launches run a single wave: the tensor-map fetches start first
  // (their latency overlaps everything below), the 32 lanes of warp 0 initialise the barrier array side by side (the halo
  // variants have up to 2 x 40 weight-stage barriers), warp 1 allocates TMEM meanwhile.
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (p.tma_store) tma_prefetch_desc(&tmC);
    }
    constexpr int kBarSlots = 2 * kMaxStages + 2 * kMaxGroups + 2 + 2 * kMaxAStages + 2 * kMaxBStages;
    constexpr int kTemptyFirst = 2 * kMaxStages + kMaxGroups, kTmemSlot = 2 * kMaxStages + 2 * kMaxGroups;
    for (int i = lane; i < kBarSlots; i += 32) {
      if (i == kTmemSlot || i == kTmemSlot + 1) continue;     // the TMEM base address lives here (written by warp 1's alloc)
      // tempty: every epilogue thread arrives -- pair mode: one arrival per epilogue warp of either CTA, on the leader's barrier
      const bool is_tempty = (i >= kTemptyFirst && i < kTemptyFirst + kMaxGroups);
      mbar_init(&bars[i], is_tempty ? (CP ? 8u : 128u) : 1u);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CP) tmem_alloc_pair(tmem_ptr, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  if (CP) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) YV6_TRACE(1);

  const int kblocks = p.npairs * p.taps * p.cin_blocks;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the schedule (keeps control flow convergent so addresses / coordinates
    // live in uniform registers); one elected lane arms the barrier and issues the two TMA loads.
    if constexpr (HALO) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      bool first = true;
      // pair mode: both CTAs fill their own rings; all completion bytes land on the LEADER's full barriers, which the
      // leader's producer arms with the byte count of both CTAs (a peer's bytes may arrive before the arming: the
      // transaction count is signed, the phase cannot complete before the leader's own arrival)
      const uint32_t a_tx = (uint32_t)HG::kBytes * (CP ? 2u : 1u), b_tx = (uint32_t)(p.b_rows * 128) * (CP ? 2u : 1u);
      const int nrow0 = rank * p.b_rows;                      // this CTA's half of the weight tile's rows
      for (int tile = unit0; tile < p.num_tiles; tile += ustep) {
        const TileCoord t = CP ? decode_unit(p, tile, rank) : decode_tile(p, tile);
        int slot = 0;
        for (int pi = 0; pi < p.npairs; ++pi) {
          const int pa = kPairA[6 - p.npairs + pi], pb = kPairB[6 - p.npairs + pi];
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&a_empty[sa], pha ^ 1);
            if (pi == 0 && cb == 0 && lane == 0) {
              if (tile == unit0 + 4 * ustep) YV6_TRACE(12);
              if (tile == unit0 + 8 * ustep) YV6_TRACE(13);
            }
            if (elect_one()) {
              if (leader) mbar_expect_tx(&a_full[sa], a_tx);
              if (CP) tma_load_5d_pair(sA + (size_t)sa * HG::kStageBytes, &tmA, &a_full[sa], cb * 64, t.w0 - 1, t.h0 * HG::SH - 1, t.i0, pa);
              else tma_load_5d(sA + (size_t)sa * HG::kStageBytes, &tmA, &a_full[sa], cb * 64, t.w0 - 1, t.h0 * HG::SH - 1, t.i0, pa);
            }
            __syncwarp();
            if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
            for (int tap = 0; tap < HG::TAPS; ++tap) {
              if (HG::SH == 2 && (tap % HG::KW) == 0 && cb < p.skip_cb) continue;     // all-zero weight block
              if constexpr (BRES) {
                if (first && elect_one()) {
                  if (leader) mbar_expect_tx(&b_full[slot], b_tx);
                  if (CP) tma_load_3d_pair(sB + (size_t)slot * p.b_stage_bytes, &tmB, &b_full[slot], tap * p.Cin + cb * 64, t.n0 + nrow0, pb);
                  else tma_load_3d(sB + (size_t)slot * p.b_stage_bytes, &tmB, &b_full[slot], tap * p.Cin + cb * 64, t.n0, pb);
                }
                __syncwarp();
              } else {
                mbar_wait(&b_empty[sb], phb ^ 1);
                if (elect_one()) {
                  if (leader) mbar_expect_tx(&b_full[sb], b_tx);
                  if (CP) tma_load_3d_pair(sB + (size_t)sb * p.b_stage_bytes, &tmB, &b_full[sb], tap * p.Cin + cb * 64, t.n0 + nrow0, pb);
                  else tma_load_3d(sB + (size_t)sb * p.b_stage_bytes, &tmB, &b_full[sb], tap * p.Cin + cb * 64, t.n0, pb);
                }
                __syncwarp();
                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
              }
              ++slot;
            }
          }
        }
        first = false;
      }
    } else {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = (uint32_t)(p.rows * p.kb_bytes + p.b_rows * p.kb_bytes) * (CP ? 2u : 1u);
    const int nrow0 = rank * p.b_rows;
    for (int tile = unit0; tile < p.num_tiles; tile += ustep) {
      const TileCoord t = CP ? decode_unit(p, tile, rank) : decode_tile(p, tile);
      for (int pi = 0; pi < p.npairs; ++pi) {
        const int pa = kPairA[6 - p.npairs + pi], pb = kPairB[6 - p.npairs + pi];
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.kw, s = tap - r * p.kw;
          const int cx = t.w0 * p.stride_w + s - p.pad_w;
          const int cy = t.h0 * p.stride + r - p.pad;
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&full[stage], tx);
              if (CP) {
                tma_load_5d_pair(sA + (size_t)stage * p.a_stage_bytes, &tmA, &full[stage], cb * p.kb_elems, cx, cy, t.i0, pa);
                tma_load_3d_pair(sB + (size_t)stage * p.b_stage_bytes, &tmB, &full[stage], tap * p.Cin + cb * p.kb_elems,
                                 t.n0 + nrow0, pb);
              } else {
                tma_load_5d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &full[stage], cb * p.kb_elems, cx, cy,
                            t.i0, pa);
                tma_load_3d(sB + (size_t)stage * p.b_stage_bytes, &tmB, &full[stage],
                            tap * p.Cin + cb * p.kb_elems, t.n0, pb);
              }
              if (tile == unit0 && pi == 0 && tap == 0 && cb == 0) YV6_TRACE(2);
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer ================================
    // Warp-convergent loop; the elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors are
    // a constant high word plus (smem address >> 4), advanced by 2 (= 32 bytes) per K=16 step.
    // Pair mode: the leader CTA issues M256 instructions over both CTAs' operands (same shared-memory offsets in
    // both) and its commits arrive on the barriers of both CTAs.
    const uint32_t idesc = umma_idesc_bf16(CP ? 256u : 128u, (uint32_t)p.BN);
    const uint64_t desc_const = umma_smem_desc(0, (uint32_t)p.sbo_bytes, (uint32_t)p.layout_type);
    const uint32_t a_base = (smem_u32(sA) & 0x3ffffu) >> 4, b_base = (smem_u32(sB) & 0x3ffffu) >> 4;
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t accumulate) {
      if (CP) umma_bf16_pair(d, ad, bd, idesc, accumulate);
      else umma_bf16(d, ad, bd, idesc, accumulate);
    };
    auto commit = [&](uint64_t* bar) {
      if (CP) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4, b_step = (uint32_t)p.b_stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // Loop-invariant parameters live in registers: the asm statements below clobber "memory", which would
    // otherwise force a constant-bank reload (and its latency) in front of every tcgen05.mma group.
    const int num_tiles = p.num_tiles, BN = p.BN, ksteps = p.ksteps, nstages = p.stages;
    const int a_stages = p.a_stages, b_stages = p.b_stages;
    constexpr bool b_resident = BRES;
    const int pcs = p.npairs * p.cin_blocks;
    if constexpr (HALO) {
      // A descriptors walk the halo tile: 8-row groups are the 8 pixels of one output row, one halo row
      // (10 pixels = 1280 bytes) apart; tap (r,s) just shifts the start address by (10 r + s) pixels.
      const uint64_t desc_a = umma_smem_desc(0, (uint32_t)(HG::W * 128 * HG::SH), 2u);
      const uint64_t desc_b = umma_smem_desc(0, 1024u, 2u);
      const uint32_t halo_step = (uint32_t)HG::kStageBytes >> 4;
      const int skip_cb = p.skip_cb, cin_blocks = p.cin_blocks;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      bool first = true;
      for (int tile = unit0; tile < num_tiles; tile += ustep) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        int slot = 0;
        uint32_t started = 0;
        for (int pc = 0; pc < pcs; ++pc) {
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a0 = a_base + (uint32_t)sa * halo_step;
          const bool skip_s0 = (HG::SH == 2) && (pc % cin_blocks) < skip_cb;      // this channel block of the s = 0 taps is all zero
#pragma unroll
          for (int tap = 0; tap < HG::TAPS; ++tap) {
            if (HG::SH == 2 && (tap % HG::KW) == 0 && skip_s0) continue;
            const int bs = b_resident ? slot : sb;
            // resident weights were loaded (and waited for) during this CTA's first tile
            if (!b_resident || first) {
              mbar_wait(&b_full[bs], b_resident ? 0u : phb);
              tc_fence_after();
            }
            const uint64_t ad = desc_a | (uint64_t)(a0 + (uint32_t)((tap / HG::KW) * HG::W + (tap % HG::KW)) * 8u);
            const uint64_t bd = desc_b | (uint64_t)(b_base + (uint32_t)bs * b_step);
            if (elect_one()) {
              mma(d_tmem, ad, bd, started);
              mma(d_tmem, ad + 2, bd + 2, 1u);
              mma(d_tmem, ad + 4, bd + 4, 1u);
              mma(d_tmem, ad + 6, bd + 6, 1u);
              if (!b_resident) commit(&b_empty[sb]);
            }
            __syncwarp();
            started = 1u;
            if (!b_resident && ++sb == b_stages) { sb = 0; phb ^= 1; }
            ++slot;
          }
          if (elect_one()) commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == a_stages) { sa = 0; pha ^= 1; }
        }
        if (elect_one()) commit(&tfull[acc]);
        __syncwarp();
        if (lane == 0) {
          if (tile == unit0) YV6_TRACE(4);
          if (tile == unit0 + 4 * ustep) YV6_TRACE(14);
          if (tile == unit0 + 8 * ustep) YV6_TRACE(15);
        }
        first = false;
        if (++acc == G) { acc = 0; acc_phase ^= 1; }
      }
    } else
    for (int tile = unit0; tile < num_tiles; tile += ustep) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (kb == 0 && tile == unit0 && lane == 0) YV6_TRACE(3);
        const uint64_t ad = desc_const | (uint64_t)(a_base + (uint32_t)stage * a_step);
        const uint64_t bd = desc_const | (uint64_t)(b_base + (uint32_t)stage * b_step);
        if (elect_one()) {
          mma(d_tmem, ad, bd, (uint32_t)(kb != 0));
          if (ksteps > 1) mma(d_tmem, ad + 2, bd + 2, 1u);
          if (ksteps > 2) {
            mma(d_tmem, ad + 4, bd + 4, 1u);
            mma(d_tmem, ad + 6, bd + 6, 1u);
          }
          commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) commit(&tfull[acc]);
      __syncwarp();
      if (tile == unit0 && lane == 0) YV6_TRACE(4);
      if (++acc == G) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ================================ epilogue ================================
    // Two groups of four warps; group g owns TMEM accumulator g, i.e. every second tile of this CTA, so
    // an epilogue may take up to two mainloop times before it stalls the MMA warp.
    const int group = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int bw = row % p.BW;
    const int tq = row / p.BW;
    const int bh = tq % p.BH;
    const int bi = tq / p.BH;
    constexpr int NBUF = kCBufCount / G;          // staging buffers per group (2 when G = 2, 1 when G = 4)
    uint8_t* gC = sC + group * NBUF * kCBufBytes;
    const bool f32 = (p.y_dtype == YV6_DT_F32);
    const float* sbias = nullptr;
    if (p.bias_smem) {   // the epilogue warps are idle until the first accumulator is ready: stage the bias now
      for (int i = (int)threadIdx.x - 64; i < p.tiles_n * p.BN; i += 128 * G) sBiasBuf[i] = __ldg(p.bias + i);
      asm volatile("bar.sync 7, %0;" ::"r"(128 * G) : "memory");
      sbias = sBiasBuf;
    }
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int tile = unit0 + group * ustep; tile < p.num_tiles; tile += G * ustep) {
      const TileCoord t = CP ? decode_unit(p, tile, rank) : decode_tile(p, tile);
      const int img = t.i0 + bi, ho = t.h0 + bh, wo = t.w0 + bw;
      const bool valid = (row < p.rows) && (img < p.N) && (ho < p.Ho) && (wo < p.Wo);
      const int64_t off = (int64_t)img * p.y_img_stride + (int64_t)ho * p.y_h_stride +
                          (int64_t)wo * p.y_w_stride;
      const int64_t roff = (int64_t)img * p.res_img_stride + (int64_t)ho * p.res_h_stride +
                           (int64_t)wo * p.res_w_stride;
      // straight-line path: single output plane, no residual or a 16-byte aligned bf16 one
      const __nv_bfloat16* res_row = (p.res != nullptr && valid) ? p.res + roff : nullptr;
      const bool fast_tile = (p.out_planes == 1) && (p.res == nullptr || (p.res_planes == 1 && p.res_aligned));
      mbar_wait(&tfull[group], acc_phase);
      tc_fence_after();
      if (tile == unit0 && q == 0 && lane == 0) YV6_TRACE(5);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * p.BN);
      if (p.tma_store == 2) {
        // ---- per-warp stores: the 32 rows of this warp form a box of the output, so each warp stages
        //      its rows (4 KB, swizzled) and issues its own TMA store -- no block-level barrier at all.
        const int wrow0 = q * 32;
        const int sw0 = t.w0 + wrow0 % p.BW, sh0 = t.h0 + wrow0 / p.BW;
        const uint32_t lrow = (uint32_t)lane * 128u, lxor = (uint32_t)(lane & 7);
        for (int c0 = 0; c0 < p.BN; c0 += p.c_chunk) {
          const int nsub = min(p.c_chunk, p.BN - c0) >> 4;
          for (int pl = 0; pl < p.out_planes; ++pl) {
            uint8_t* buf = gC + cbuf * kCBufBytes + q * 4096;
            if (lane == 0) tma_store_wait_read<NBUF - 1>();  // the store that last read this buffer is done
            __syncwarp();
            const uint32_t base = smem_u32(buf) + lrow;
            if (fast_tile && nsub * 16 == p.c_chunk) {
              if (f32) {
                epi_cols32<true>(p, sbias, taddr + (uint32_t)c0, t.n0 + c0, res_row, base, lxor, 0);
              } else {
                epi_cols32<false>(p, sbias, taddr + (uint32_t)c0, t.n0 + c0, res_row, base, lxor, 0);
                epi_cols32<false>(p, sbias, taddr + (uint32_t)(c0 + 32), t.n0 + c0 + 32, res_row, base, lxor, 4);
              }
            } else {
            uint32_t r[2][16];
            tmem_ld16(taddr + (uint32_t)c0, r[0]);
#pragma unroll 1
            for (int sb = 0; sb < 4; ++sb) {
              if (sb < nsub) {
                tmem_ld_wait();
                if (sb + 1 < nsub) tmem_ld16(taddr + (uint32_t)(c0 + 16 * (sb + 1)), r[(sb + 1) & 1]);
                const int n = t.n0 + c0 + 16 * sb;
                float v[16];
                epilogue_math(p, sbias, r[sb & 1], n, min(16, p.Cout - n), valid, roff, v);
                for (int k = 0; k < pl; ++k) {  // bf16x3: peel the planes already written
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
                }
                if (f32) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t a = base + ((((uint32_t)(4 * sb + j)) ^ lxor) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                 : "memory");
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const uint32_t a = base + ((((uint32_t)(2 * sb + j)) ^ lxor) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                                 "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                                 "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                                 : "memory");
                  }
                }
              }
            }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_5d(&tmC, buf, t.n0 + c0, sw0, sh0, t.i0, pl);
              tma_store_commit();
            }
            if (++cbuf == NBUF) cbuf = 0;
          }
        }
      } else if (p.tma_store == 1) {
        // ---- block-level store: tiles whose warps do not map to boxes (e.g. 20x5): the group stages the
        //      whole 128-row chunk, synchronises on its named barrier, one thread issues the TMA store.
        const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);  // first thread of the group
        const uint32_t row_smem = (uint32_t)row * 128u, row_xor = (uint32_t)(row & 7);
        for (int c0 = 0; c0 < p.BN; c0 += p.c_chunk) {
          const int nsub = min(p.c_chunk, p.BN - c0) >> 4;
          for (int pl = 0; pl < p.out_planes; ++pl) {
            uint8_t* buf = gC + cbuf * kCBufBytes;
            if (issuer) tma_store_wait_read<NBUF - 1>();
            epi_bar_sync(group);
            const uint32_t base = smem_u32(buf) + row_smem;
            if (fast_tile && nsub * 16 == p.c_chunk) {
              if (f32) {
                epi_cols32<true>(p, sbias, taddr + (uint32_t)c0, t.n0 + c0, res_row, base, row_xor, 0);
              } else {
                epi_cols32<false>(p, sbias, taddr + (uint32_t)c0, t.n0 + c0, res_row, base, row_xor, 0);
                epi_cols32<false>(p, sbias, taddr + (uint32_t)(c0 + 32), t.n0 + c0 + 32, res_row, base, row_xor, 4);
              }
            } else {
            uint32_t r[2][16];
            tmem_ld16(taddr + (uint32_t)c0, r[0]);
#pragma unroll 1
            for (int sb = 0; sb < 4; ++sb) {
              if (sb < nsub) {
                tmem_ld_wait();
                if (sb + 1 < nsub) tmem_ld16(taddr + (uint32_t)(c0 + 16 * (sb + 1)), r[(sb + 1) & 1]);
                const int n = t.n0 + c0 + 16 * sb;
                float v[16];
                epilogue_math(p, sbias, r[sb & 1], n, min(16, p.Cout - n), valid, roff, v);
                for (int k = 0; k < pl; ++k) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
                }
                if (row < p.rows) {
                  if (f32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const uint32_t a = base + ((((uint32_t)(4 * sb + j)) ^ row_xor) << 4);
                      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                   "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                   : "memory");
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      const uint32_t a = base + ((((uint32_t)(2 * sb + j)) ^ row_xor) << 4);
                      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                                   "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                                   "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                                   : "memory");
                    }
                  }
                }
              }
            }
            }
            fence_proxy_async_smem();
            epi_bar_sync(group);
            if (issuer) {
              tma_store_5d(&tmC, buf, t.n0 + c0, t.w0, t.h0, t.i0, pl);
              tma_store_commit();
            }
            if (++cbuf == NBUF) cbuf = 0;
          }
        }
      } else {
        // ---- fallback: direct global stores (outputs whose strides TMA cannot express) ----
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + (uint32_t)c0, r);
          tmem_ld_wait();
          const int n = t.n0 + c0;
          const int ncol = min(16, p.Cout - n);
          float v[16];
          epilogue_math(p, sbias, r, n, ncol, valid, roff, v);
          if (valid && ncol > 0) store_chunk(p, off, n, ncol, v);
        }
      }
      tc_fence_before();
      if (CP) {     // the accumulator of BOTH CTAs is free again once every epilogue warp of the pair has read its lanes
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[group]);
      } else {
        mbar_arrive(&tempty[group]);
      }
      if (q == 0 && lane == 0) {
        if (tile == unit0) YV6_TRACE(6);
        if (tile == unit0 + 4 * ustep) YV6_TRACE(9);
        if (tile == unit0 + 8 * ustep) YV6_TRACE(10);
      }
      acc_phase ^= 1;
    }
    // the staging buffers must outlive the TMA engine's READS only; the global writes complete with the grid
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (group == 0 && q == 0 && lane == 0) YV6_TRACE(7);
  }

  tc_fence_before();
  if (CP) cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer can still signal it or issue into it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CP) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (lane == 0) YV6_TRACE(8);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tile planning, tensor-map encoding, launch
// ------------------------------------------------------------------------------------------------
struct ConvPlan {
  ConvKParams k;
  int grid;
  size_t smem_bytes;
  CUtensorMapSwizzle swz;
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static int plan_conv(const yv6_handle* h, const yv6_conv_desc* d, ConvPlan* plan) {
  YV6_REQUIRE(d != nullptr, "conv: null descriptor");
  YV6_REQUIRE(d->x && d->w && d->y, "conv: null tensor pointer");
  YV6_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "conv: bad input shape %dx%dx%d", d->N, d->H, d->W);
  YV6_REQUIRE(d->Cin > 0 && d->Cin % 16 == 0, "conv: Cin=%d must be a positive multiple of 16", d->Cin);
  YV6_REQUIRE(d->x_c_total >= d->Cin && d->x_c_total % 8 == 0, "conv: bad x_c_total=%d", d->x_c_total);
  YV6_REQUIRE(d->Cout > 0, "conv: Cout=%d", d->Cout);
  YV6_REQUIRE(d->kh >= 1 && d->kh <= 3 && d->kw >= 1 && d->kw <= 3, "conv: kernel %dx%d unsupported", d->kh, d->kw);
  YV6_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride %d unsupported", d->stride);
  YV6_REQUIRE(d->stride_w >= 0 && d->stride_w <= 2, "conv: stride_w %d unsupported", d->stride_w);
  const int stride_w = d->stride_w > 0 ? d->stride_w : d->stride;
  YV6_REQUIRE(d->nsplit == 1 || d->nsplit == 3, "conv: nsplit must be 1 or 3");
  YV6_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
              "conv: x / w must be 16-byte aligned");
  YV6_REQUIRE(d->y_dtype == YV6_DT_BF16 || d->y_dtype == YV6_DT_F32, "conv: bad y_dtype");

  ConvKParams& k = plan->k;
  memset(&k, 0, sizeof(k));
  k.N = d->N;
  // pad_w < 0 means "same as pad" (square padding); out_h/out_w > 0 override the conv arithmetic (used by
  // the parity sub-problems of a stride-2 dgrad, whose far-side reads rely on TMA zero fill)
  const int pad_w = (d->pad_w == YV6_PAD_SAME) ? d->pad : d->pad_w;
  k.Ho = d->out_h > 0 ? d->out_h : (d->H + 2 * d->pad - d->kh) / d->stride + 1;
  k.Wo = d->out_w > 0 ? d->out_w : (d->W + 2 * pad_w - d->kw) / stride_w + 1;
  YV6_REQUIRE(k.Ho > 0 && k.Wo > 0, "conv: empty output");
  k.Cout = d->Cout;
  k.Cin = d->Cin;
  k.taps = d->kh * d->kw;
  k.kw = d->kw;
  k.stride = d->stride;
  k.stride_w = stride_w;
  k.pad = d->pad;
  k.pad_w = pad_w;

  // K block = largest of 64/32/16 channels dividing Cin -> 128/64/32-byte swizzle
  k.kb_elems = (d->Cin % 64 == 0) ? 64 : (d->Cin % 32 == 0) ? 32 : 16;
  k.kb_bytes = k.kb_elems * 2;
  k.ksteps = k.kb_elems / 16;
  k.cin_blocks = d->Cin / k.kb_elems;
  k.sbo_bytes = 8 * k.kb_bytes;
  k.layout_type = (k.kb_bytes == 128) ? 2 : (k.kb_bytes == 64) ? 4 : 6;
  plan->swz = (k.kb_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
              : (k.kb_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                   : CU_TENSOR_MAP_SWIZZLE_32B;
  k.npairs = (d->nsplit == 3) ? 6 : 1;

  // output staging: 128-byte rows -> 64 bf16 or 32 fp32 columns per TMA-stored chunk
  const int esz = (d->y_dtype == YV6_DT_F32) ? 4 : 2;
  k.c_chunk = 128 / esz;
  k.tma_store = (d->force_direct != 1) && ((d->y_w_stride * esz) % 16 == 0) && ((d->y_h_stride * esz) % 16 == 0) &&
                ((d->y_img_stride * esz) % 16 == 0) && ((d->y_plane_stride * esz) % 16 == 0) &&
                ((reinterpret_cast<uintptr_t>(d->y) & 15) == 0);
  // N tiling is chosen together with the M tiling below
  const int bn_align = k.c_chunk;  // with several N tiles, a tile must end on a staged-chunk boundary
  // M tiling: BW x BH x BI box of output pixels, <= 128 rows, fewest tiles wins
  if (d->force_bw > 0) {
    k.BW = d->force_bw;
    k.BH = std::max(1, d->force_bh);
    k.BI = std::max(1, d->force_bi);
    YV6_REQUIRE(k.BW * k.BH * k.BI <= kTileRows && k.BW * stride_w <= 256 && k.BH * d->stride <= 256,
                "conv: forced tile %dx%dx%d invalid", k.BW, k.BH, k.BI);
  } else {
    long best_tiles = -1, best_box_tiles = -1;
    int bestw = 1, besth = 1, besti = 1, boxw = 0, boxh = 0;
    const int maxbw = std::min(std::min(k.Wo, kTileRows), 256 / stride_w);
    for (int bw = 1; bw <= maxbw; ++bw) {
      const int maxbh = std::min(std::min(k.Ho, kTileRows / bw), 256 / d->stride);
      for (int bh = 1; bh <= maxbh; ++bh) {
        int bi = 1;
        if (bw >= k.Wo && bh >= k.Ho) bi = std::max(1, std::min(d->N, kTileRows / (bw * bh)));
        long tiles = (long)ceil_div(k.Wo, bw) * ceil_div(k.Ho, bh) * ceil_div(d->N, bi);
        if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && bw > bestw)) {
          best_tiles = tiles;
          bestw = bw;
          besth = bh;
          besti = bi;
        }
      }
    }
    // "warp-box" tiles: 128 rows whose 32-row quarters are boxes of the output (BW | 32 or 32 | BW), which
    // lets every epilogue warp TMA-store its own rows without a block barrier; preferred within 8 % of the best
    for (int bw = 1; bw <= std::min(kTileRows, 256 / stride_w); bw <<= 1) {
      const int bh = kTileRows / bw;
      if (bh * d->stride > 256) continue;
      long tiles = (long)ceil_div(k.Wo, bw) * ceil_div(k.Ho, bh) * d->N;
      if (best_box_tiles < 0 || tiles < best_box_tiles || (tiles == best_box_tiles && bw > boxw)) {
        best_box_tiles = tiles;
        boxw = bw;
        boxh = bh;
      }
    }
    if (best_box_tiles > 0 && best_box_tiles * 100 <= best_tiles * 108 && d->force_bi == 0) {
      bestw = boxw;
      besth = boxh;
      besti = 1;
    }
    k.BW = bestw;
    k.BH = besth;
    k.BI = besti;
  }
  // halo mode: 3x3 stride-1 with 64-channel K blocks and an 8x16 output tile; taken when its fixed tile
  // shape costs at most 25% more tiles than the best free-form box (it moves ~6x fewer A bytes)
  k.halo = 0;
  if (d->kh == 3 && d->kw == 3 && d->stride == 1 && stride_w == 1 && d->pad == 1 && pad_w == 1 && d->out_h == 0 && d->out_w == 0 &&
      d->Cin % 64 == 0 && d->force_bw == 0 && d->force_halo >= 0) {
    const long generic = (long)ceil_div(k.Wo, k.BW) * ceil_div(k.Ho, k.BH) * ceil_div(d->N, k.BI);
    const long halo_tiles = (long)ceil_div(k.Wo, 8) * ceil_div(k.Ho, 16) * d->N;
    if (d->force_halo > 0 || halo_tiles * 4 <= generic * 5) {
      k.halo = 1;
      k.BW = 8;
      k.BH = 16;
      k.BI = 1;
    }
  }
  // halo 2: the column-pair view of a 3x3 stride-2 conv (3x2 kernel, stride (2, 1), pad (1, 1), out_w = W), same 8 x 16 tile
  // over a 9 x 33 box: ~2.6x fewer A bytes from L2 than one box per tap.  pair_view additionally promises that the weights of
  // the left tap are zero over the even pixel's channels [0, Cin/2): whole 64-channel blocks of zeros are skipped.
  k.skip_cb = 0;
  if (d->kh == 3 && d->kw == 2 && d->stride == 2 && stride_w == 1 && d->pad == 1 && pad_w == 1 && d->out_w == d->W && d->out_h == 0 &&
      d->Cin % 64 == 0 && d->force_bw == 0 && d->force_halo >= 0 && d->H % 2 == 0) {
    const long generic = (long)ceil_div(k.Wo, k.BW) * ceil_div(k.Ho, k.BH) * ceil_div(d->N, k.BI);
    const long halo_tiles = (long)ceil_div(k.Wo, 8) * ceil_div(k.Ho, 16) * d->N;
    // auto (profiles/r02_s2_halo_sweep.md): up to 128 real input channels and where the fixed tile wastes < 25 %; beyond that the
    // weight operand dominates the L2 -> shared-memory traffic and the plain stride-2 mainloop (fewer, wider K blocks) wins
    if (d->force_halo > 0 || (halo_tiles * 4 <= generic * 5 && d->Cin <= 256)) {
      k.halo = 2;
      k.BW = 8;
      k.BH = 16;
      k.BI = 1;
      if (d->pair_view && (d->Cin / 2) % 64 == 0) k.skip_cb = d->Cin / 128;
    }
  }
  const long m_tiles = (long)ceil_div(k.Wo, k.BW) * ceil_div(k.Ho, k.BH) * ceil_div(d->N, k.BI);
  // N tiling: BN <= 256, multiple of 16; pick the split whose wave count x tile cost is smallest
  // (a 448-tile layer on 148 SMs runs 4 waves at BN=256 but 7 half-cost waves at BN=128).
  if (d->force_bn > 0) {
    YV6_REQUIRE(d->force_bn % 16 == 0 && d->force_bn <= 256, "conv: force_bn must be a multiple of 16 <= 256");
    k.BN = d->force_bn;
  } else {
    double best_cost = -1;
    int best_bn = 0;
    const int nt_min = ceil_div(d->Cout, 256);
    for (int nt = nt_min; nt <= nt_min * 4; ++nt) {
      int bn = ceil_div(ceil_div(d->Cout, nt), 16) * 16;
      if (nt > 1) bn = ceil_div(bn, bn_align) * bn_align;
      if (bn > 256 || (nt > 1 && bn < 64)) continue;
      const long tiles = m_tiles * ceil_div(d->Cout, bn);
      const double waves = (double)ceil_div((int)std::min<long>(tiles, 1l << 30), h->num_sms);
      const double cost = waves * (bn + 48.0);  // +48: per-tile fixed cost and A re-reads favour wide tiles
      if (best_cost < 0 || cost < best_cost - 1e-9) {
        best_cost = cost;
        best_bn = bn;
      }
    }
    k.BN = best_bn;
  }
  YV6_REQUIRE(k.BN >= 16, "conv: could not choose BN for Cout=%d", d->Cout);
  k.tiles_n = ceil_div(d->Cout, k.BN);
  if (k.tiles_n > 1 && k.tma_store && (k.BN % bn_align) != 0) k.tma_store = 0;
  k.rows = k.BW * k.BH * k.BI;
  if (k.tma_store && k.rows == kTileRows && k.BI == 1 && (32 % k.BW == 0 || k.BW % 32 == 0) && d->force_direct != 2)
    k.tma_store = 2;  // per-warp stores
  k.tiles_w = ceil_div(k.Wo, k.BW);
  k.tiles_h = ceil_div(k.Ho, k.BH);
  k.tiles_i = ceil_div(d->N, k.BI);
  YV6_REQUIRE(m_tiles * k.tiles_n < (1l << 30), "conv: too many tiles");
  k.m_tiles = (int)m_tiles;
  // CTA pairs (force_pair: 1 = on whenever there are two M tiles, -1 = off, 0 = auto).  Measured on B200, bs32
  // (profiles/r02_pair_sweep.md): pairs gain where the weight operand dominates the shared-memory traffic -- 3x3 stride-1
  // layers over >= 128 input channels (128->128 @80x80: 57.5 -> 52.0 us) -- are neutral on the stride-2 layers and LOSE on
  // the HBM-bound 1x1 layers (two SMs in lockstep hide less latency; 64->64 @160x160: 50 -> 67 us) and on the
  // resident-weight 64-channel 3x3 layers (73 -> 93 us), so auto mode takes them only for the first kind.
  // ... and for the stride-2 halo mainloop from 128 output channels on: half a weight tile per CTA keeps the 64->128 layer's
  // weights resident and halves the streamed weight bytes of the wider ones (64->128 @160: 46.4 -> 43.7 us, 128->256 @80: 42.8 -> 36.7).
  const bool pair_auto = (d->kh == 3 && d->kw == 3 && d->stride == 1 && stride_w == 1 && d->Cin >= 128 && d->Cout <= d->Cin) ||
                         (k.halo == 2 && d->Cout >= 128);
  k.cpair = (m_tiles >= 2 && h->max_clusters > 0 && (d->force_pair > 0 || (d->force_pair == 0 && pair_auto))) ? 1 : 0;
  k.b_rows = k.cpair ? k.BN / 2 : k.BN;
  // schedule units: tiles, or (two consecutive M tiles) x N tile for CTA pairs
  k.num_tiles = k.cpair ? (int)(((m_tiles + 1) / 2) * k.tiles_n) : (int)(m_tiles * k.tiles_n);

  // smem ring(s)
  k.a_stage_bytes = kTileRows * k.kb_bytes;
  k.b_stage_bytes = ((k.b_rows * k.kb_bytes + 1023) / 1024) * 1024;
  const int budget = h->max_smem_optin - 1024 - 1024 - kCBufCount * kCBufBytes - kBiasSmemFloats * (int)sizeof(float);
  if (k.halo) {
    const int halo_stage = (k.halo == 2) ? HaloGeom<true>::kStageBytes : HaloGeom<false>::kStageBytes;
    const int min_a = (k.halo == 2) ? 2 : 3;          // A stages that must fit next to resident weights
    // B tiles one output tile consumes (halo 2: six taps, minus the all-zero blocks of the three left taps)
    const int b_tiles = (k.halo == 2) ? k.npairs * (k.cin_blocks * 6 - 3 * k.skip_cb) : k.npairs * k.cin_blocks * 9;
    k.b_resident = (k.tiles_n == 1 && b_tiles <= kMaxBStages && d->force_stages == 0 &&
                    (long)b_tiles * k.b_stage_bytes + min_a * halo_stage <= budget) ? 1 : 0;
    if (k.b_resident) {
      k.b_stages = b_tiles;
      k.a_stages = std::min(kMaxAStages, (budget - b_tiles * k.b_stage_bytes) / halo_stage);
    } else {
      k.a_stages = (k.halo == 1 && budget - 3 * halo_stage >= 3 * k.b_stage_bytes) ? 3 : 2;
      k.b_stages = std::min(kMaxBStages, (budget - k.a_stages * halo_stage) / k.b_stage_bytes);
      if (d->force_stages > 0) k.b_stages = std::min(k.b_stages, std::max(2, d->force_stages));
    }
    YV6_REQUIRE(k.a_stages >= 2 && k.b_stages >= 2, "conv(halo): not enough shared memory");
    k.stages = k.b_stages;
    k.a_region_bytes = k.a_stages * halo_stage;
    k.b_region_bytes = k.b_stages * k.b_stage_bytes;
  } else {
    const int stage_bytes = k.a_stage_bytes + k.b_stage_bytes;
    int stages = std::min(kMaxStages, budget / stage_bytes);
    if (d->force_stages > 0) stages = std::min(stages, d->force_stages);
    YV6_REQUIRE(stages >= 2, "conv: not enough shared memory for a 2-stage pipeline");
    k.stages = stages;
    k.a_region_bytes = stages * k.a_stage_bytes;
    k.b_region_bytes = stages * k.b_stage_bytes;
  }
  k.trace = reinterpret_cast<unsigned long long*>(d->trace);
  plan->smem_bytes = (size_t)k.a_region_bytes + k.b_region_bytes + kCBufCount * kCBufBytes + kBiasSmemFloats * sizeof(float) + 1024 + 1024;

  // four groups only pay off when the tile's mainloop is shorter than its epilogue (1x1 / small-K layers)
  const int kblocks_per_tile = k.npairs * k.taps * k.cin_blocks;
  k.groups = (4 * k.BN <= 512 && kblocks_per_tile <= 8 && d->force_groups != 2) ? 4 : 2;
  if (d->force_groups == 4 && 4 * k.BN <= 512) k.groups = 4;
  int cols = 32;
  while (cols < k.groups * k.BN) cols *= 2;
  k.tmem_cols = cols;

  k.act = d->act;
  k.y_dtype = d->y_dtype;
  k.bias_smem = (d->bias != nullptr && k.tiles_n * k.BN <= kBiasSmemFloats) ? 1 : 0;
  k.fast_act = (d->y_dtype == YV6_DT_BF16 && d->nsplit != 3) ? 1 : 0;
  k.res_aligned = (d->res != nullptr && (reinterpret_cast<uintptr_t>(d->res) & 15) == 0 && d->res_img_stride % 8 == 0 &&
                   d->res_h_stride % 8 == 0 && d->res_w_stride % 8 == 0) ? 1 : 0;
  k.out_planes = (d->nsplit == 3 && d->y_dtype == YV6_DT_BF16) ? 3 : 1;
  k.res_planes = (d->nsplit == 3) ? 3 : 1;
  k.y = d->y;
  k.y_img_stride = d->y_img_stride;
  k.y_h_stride = d->y_h_stride;
  k.y_w_stride = d->y_w_stride;
  k.y_plane_stride = d->y_plane_stride;
  k.res = reinterpret_cast<const __nv_bfloat16*>(d->res);
  k.alpha = d->alpha;
  k.res_img_stride = d->res_img_stride;
  k.res_h_stride = d->res_h_stride;
  k.res_w_stride = d->res_w_stride;
  k.res_plane_stride = d->res_plane_stride;
  k.bias = d->bias;

  if (k.cpair) {   // grid counts CTAs: two per unit
    int clusters = std::min(k.num_tiles, h->max_clusters);
    if (d->force_grid > 0) clusters = std::min(k.num_tiles, std::max(1, d->force_grid / 2));
    plan->grid = 2 * clusters;
  } else {
    plan->grid = std::min(k.num_tiles, h->num_sms);
    if (d->force_grid > 0) plan->grid = std::min(k.num_tiles, d->force_grid);
  }
  return YV6_OK;
}

}  // namespace yv6

using namespace yv6;

using ConvKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvKParams);
static const ConvKernelFn kConvKernels[2][2][5] = {
    {{conv_igemm_kernel<2, 0, false>, conv_igemm_kernel<2, 1, false>, conv_igemm_kernel<2, 2, false>, conv_igemm_kernel<2, 3, false>,
      conv_igemm_kernel<2, 4, false>},
     {conv_igemm_kernel<4, 0, false>, conv_igemm_kernel<4, 1, false>, conv_igemm_kernel<4, 2, false>, conv_igemm_kernel<4, 3, false>,
      conv_igemm_kernel<4, 4, false>}},
    {{conv_igemm_kernel<2, 0, true>, conv_igemm_kernel<2, 1, true>, conv_igemm_kernel<2, 2, true>, conv_igemm_kernel<2, 3, true>,
      conv_igemm_kernel<2, 4, true>},
     {conv_igemm_kernel<4, 0, true>, conv_igemm_kernel<4, 1, true>, conv_igemm_kernel<4, 2, true>, conv_igemm_kernel<4, 3, true>,
      conv_igemm_kernel<4, 4, true>}}};

// Once per device: opt the kernels in to the large dynamic shared memory and ask how many 2-CTA clusters of the conv
// kernel (one CTA per SM) can be co-resident -- the grid of the pair variants.
static int conv_configure(yv6_handle* h) {
  if (h->configured & YV6_CFG_CONV) return YV6_OK;
  for (int c = 0; c < 2; ++c)
    for (int g = 0; g < 2; ++g)
      for (int m = 0; m < 5; ++m)
        YV6_CHECK_CUDA(cudaFuncSetAttribute(kConvKernels[c][g][m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem_optin));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * h->num_sms));
  cfg.blockDim = dim3(64 + 128 * 2);
  cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kConvKernels[1][0][1], &cfg);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    n = 0;                                   // pairs unavailable: the single-CTA variants are used
  }
  h->max_clusters = std::min(n, h->num_sms / 2);
  h->configured |= YV6_CFG_CONV;
  return YV6_OK;
}

extern "C" int yv6_conv_plan(yv6_handle* h, const yv6_conv_desc* d, int32_t* out8) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h != nullptr && out8 != nullptr, "conv_plan: null argument");
  { const int rc0 = conv_configure(h); if (rc0 != YV6_OK) return rc0; }
  ConvPlan plan;
  int rc = plan_conv(h, d, &plan);
  if (rc != YV6_OK) return rc;
  out8[0] = plan.k.BW;
  out8[1] = plan.k.BH;
  out8[2] = plan.k.BI;
  out8[3] = plan.k.BN;
  out8[4] = plan.k.kb_elems;
  out8[5] = plan.k.stages;
  out8[6] = plan.grid;
  out8[7] = plan.k.num_tiles;
  out8[8] = plan.k.halo;
  out8[9] = (plan.k.halo ? plan.k.a_stages * 100 + plan.k.b_resident : 0) + 10 * plan.k.cpair;
  return YV6_OK;
}

// Host-only twin of yv6_conv_plan: plans against stated device properties instead of a handle, makes no CUDA call.
extern "C" int yv6_conv_plan_host(int num_sms, int max_smem_optin, int max_clusters, const yv6_conv_desc* d, int32_t* out12) {
  YV6_REQUIRE(out12 != nullptr, "conv_plan_host: null argument");
  YV6_REQUIRE(num_sms > 0 && max_smem_optin > 0 && max_clusters >= 0, "conv_plan_host: bad device properties");
  yv6_handle fake;
  memset(&fake, 0, sizeof(fake));
  fake.device = -1;
  fake.num_sms = num_sms;
  fake.max_smem_optin = max_smem_optin;
  fake.max_clusters = max_clusters;
  ConvPlan plan;
  int rc = plan_conv(&fake, d, &plan);
  if (rc != YV6_OK) return rc;
  YV6_REQUIRE(plan.smem_bytes <= (size_t)max_smem_optin, "conv_plan_host: plan needs %zu bytes of shared memory, device offers %d",
              plan.smem_bytes, max_smem_optin);
  YV6_REQUIRE(plan.k.tmem_cols <= 512, "conv_plan_host: plan needs %d TMEM columns", plan.k.tmem_cols);
  out12[0] = plan.k.BW;
  out12[1] = plan.k.BH;
  out12[2] = plan.k.BI;
  out12[3] = plan.k.BN;
  out12[4] = plan.k.kb_elems;
  out12[5] = plan.k.stages;
  out12[6] = plan.grid;
  out12[7] = plan.k.num_tiles;
  out12[8] = plan.k.halo;
  out12[9] = (plan.k.halo ? plan.k.a_stages * 100 + plan.k.b_resident : 0) + 10 * plan.k.cpair;
  out12[10] = (int32_t)plan.smem_bytes;
  out12[11] = plan.k.tmem_cols;
  return YV6_OK;
}

extern "C" int yv6_conv_fwd(yv6_handle* h, const yv6_conv_desc* d, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h != nullptr, "conv_fwd: null handle");
  { const int rc0 = conv_configure(h); if (rc0 != YV6_OK) return rc0; }
  ConvPlan plan;
  int rc = plan_conv(h, d, &plan);
  if (rc != YV6_OK) return rc;
  const ConvKParams& k = plan.k;
  YV6_REQUIRE(h->encode_tiled != nullptr, "conv_fwd: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");

  // A: (C, W, H, N, plane) over the input (channel slice of a wider NHWC buffer allowed)
  CUtensorMap tmA, tmB, tmC;
  {
    const uint64_t planes = (d->nsplit == 3) ? 3 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N, planes};
    const uint64_t pix = (uint64_t)d->x_c_total * 2;
    uint64_t plane_stride = (d->nsplit == 3) ? (uint64_t)d->x_plane_stride * 2
                                             : (uint64_t)d->N * d->H * d->W * pix;
    YV6_REQUIRE(plane_stride % 16 == 0, "conv: x_plane_stride must be a multiple of 8 elements");
    cuuint64_t strides[4] = {pix, pix * d->W, pix * d->W * d->H, plane_stride};
    cuuint32_t box[5] = {(cuuint32_t)k.kb_elems, (cuuint32_t)(k.BW * k.stride_w), (cuuint32_t)(k.BH * d->stride),
                         (cuuint32_t)k.BI, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)k.stride_w, (cuuint32_t)d->stride, 1, 1};
    if (k.halo) {       // one dense box per channel block
      box[1] = (k.halo == 2) ? HaloGeom<true>::W : HaloGeom<false>::W;
      box[2] = (k.halo == 2) ? HaloGeom<true>::H : HaloGeom<false>::H;
      estr[1] = estr[2] = 1;
    }
    CUresult cr = h->encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->x), dims,
                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, plan.swz,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      yv6_set_error("conv: cuTensorMapEncodeTiled(A) failed with %d (C=%d W=%d H=%d N=%d box=%u,%u,%u,%u)", (int)cr,
                    d->Cin, d->W, d->H, d->N, box[0], box[1], box[2], box[3]);
      return YV6_ERR_CUDA;
    }
  }
  {
    const uint64_t planes = (d->nsplit == 3) ? 3 : 1;
    const uint64_t ktot = (uint64_t)k.taps * d->Cin;
    cuuint64_t dims[3] = {ktot, (cuuint64_t)d->Cout, planes};
    uint64_t plane_stride = (d->nsplit == 3) ? (uint64_t)d->w_plane_stride * 2 : ktot * 2 * d->Cout;
    YV6_REQUIRE(plane_stride % 16 == 0, "conv: w_plane_stride must be a multiple of 8 elements");
    cuuint64_t strides[2] = {ktot * 2, plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)k.kb_elems, (cuuint32_t)k.b_rows, 1};   // pair mode: each CTA loads half of the N tile
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = h->encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), dims,
                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, plan.swz,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      yv6_set_error("conv: cuTensorMapEncodeTiled(B) failed with %d", (int)cr);
      return YV6_ERR_CUDA;
    }
  }

  // C: (Cout, Wo, Ho, N, plane) over the output view; one box = one staged chunk of one tile
  if (k.tma_store) {
    const uint64_t esz = (d->y_dtype == YV6_DT_F32) ? 4 : 2;
    const uint64_t planes = (uint64_t)k.out_planes;
    cuuint64_t dims[5] = {(cuuint64_t)d->Cout, (cuuint64_t)k.Wo, (cuuint64_t)k.Ho, (cuuint64_t)d->N, planes};
    uint64_t plane_stride = (planes > 1) ? (uint64_t)d->y_plane_stride * esz : (uint64_t)d->y_img_stride * esz * d->N;
    if (plane_stride == 0) plane_stride = 16;
    cuuint64_t strides[4] = {(uint64_t)d->y_w_stride * esz, (uint64_t)d->y_h_stride * esz,
                             (uint64_t)d->y_img_stride * esz, plane_stride};
    cuuint32_t box[5] = {(cuuint32_t)k.c_chunk, (cuuint32_t)k.BW, (cuuint32_t)k.BH, (cuuint32_t)k.BI, 1};
    if (k.tma_store == 2) {  // one warp = 32 consecutive rows of the tile
      box[1] = (cuuint32_t)std::min(k.BW, 32);
      box[2] = (cuuint32_t)std::max(1, 32 / k.BW);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = h->encode_tiled(&tmC, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                                  5, d->y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      yv6_set_error("conv: cuTensorMapEncodeTiled(C) failed with %d (Cout=%d Wo=%d Ho=%d N=%d strides=%llu,%llu,%llu)",
                    (int)cr, d->Cout, k.Wo, k.Ho, d->N, (unsigned long long)strides[0],
                    (unsigned long long)strides[1], (unsigned long long)strides[2]);
      return YV6_ERR_CUDA;
    }
  } else {
    tmC = tmA;  // unused by the kernel
  }

  const int mode = k.halo ? (k.halo == 2 ? 3 : 1) + (k.b_resident ? 1 : 0) : 0;
  ConvKernelFn fn = kConvKernels[k.cpair][k.groups == 4 ? 1 : 0][mode];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3((unsigned)(64 + 128 * k.groups));
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  if (k.cpair) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  YV6_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmC, k));
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
