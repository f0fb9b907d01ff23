// yv6_train.cu -- the HBM-bound training kernels around the conv / wgrad GEMMs (sm_100a, CUDA cores).
//
// The reference trains in *train form*: every ConvModule is conv -> BatchNorm(batch stats) -> act
// (yolov6/layers/common.py:26-49) and every RepVGGBlock is relu(BN(conv3x3) + BN(conv1x1) + BN(x))
// with three independent BatchNorms (common.py:245-255); autograd + cuDNN then run the backward.
// Here each block is: raw conv(s) on tensor cores (yv6_conv_fwd with no bias/act) and
//   yv6_bn_stats      per-channel sum / sum-of-squares of an NHWC bf16 tensor            (fwd)
//   yv6_bn_finalize   mean / invstd / running-stat update (momentum .03, eps 1e-3)        (fwd)
//   yv6_bn_apply_fwd  y = act(sum_b (x_b * scale_b + shift_b)) over up to 3 branches      (fwd)
//   yv6_bn_bwd_reduce per branch: sum(dz), sum(dz * xhat_b) with dz = dY * act'(.)         (bwd)
//   yv6_bn_bwd_apply  per branch: dx_b = scale_b * (dz - S1/M - xhat_b * S2_b/M)           (bwd)
// plus yv6_head_grad_prep (sigmoid backward + repack of the head gradients), yv6_maxpool5_bwd (SPPF)
// and yv6_stem_wgrad (the 3-channel first conv).  All activations are NHWC bf16 with a channel pitch
// (slices of concat buffers); statistics and parameter gradients are float64 / float32.
#include <algorithm>

#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kTrThreads = 256;

struct View {          // NHWC bf16 tensor slice
  const __nv_bfloat16* p;
  int64_t pitch;       // elements between consecutive pixels
};

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    v[2 * j] = __low2float(b2);
    v[2 * j + 1] = __high2float(b2);
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&b2);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------- bn_stats
// grid: (channel groups of 8 per block.x dimension folded into threads) -- thread = (pixel lane, channel group)
__global__ void __launch_bounds__(kTrThreads) bn_stats_kernel(View x, int64_t pixels, int C, double* sum, double* sumsq) {
  const int cgs = C / 8;
  const int cg = threadIdx.x % cgs;                      // requires cgs | blockDim (host picks block = cgs * rows)
  const int prow = threadIdx.x / cgs, prows = blockDim.x / cgs;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (int64_t px = (int64_t)blockIdx.x * prows + prow; px < pixels; px += (int64_t)gridDim.x * prows) {
    float v[8];
    ld8(x.p + px * x.pitch + cg * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] += v[j] * v[j]; }
  }
  // block-level reduction in shared memory first: one global atomic per channel per block instead of one per
  // thread (the per-thread version serialised on C addresses and cost 56 % of a training step)
  extern __shared__ double sred[];                      // [2][C]
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sred[c] = 0.0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sred[cg * 8 + j], (double)s[j]);
    atomicAdd(&sred[C + cg * 8 + j], (double)q[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&sum[c], sred[c]);
    atomicAdd(&sumsq[c], sred[C + c]);
  }
}

// mean, biased var -> invstd, scale/shift; running stats with unbiased var (torch BatchNorm2d semantics)
__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                                   float eps, float momentum, float* running_mean, float* running_var, float* mean_out,
                                   float* invstd_out, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sum[c] / count;
  double var = sumsq[c] / count - m * m;
  if (var < 0) var = 0;
  const double inv = 1.0 / sqrt(var + (double)eps);
  mean_out[c] = (float)m;
  invstd_out[c] = (float)inv;
  const double g = gamma[c];
  scale[c] = (float)(g * inv);
  shift[c] = (float)((double)beta[c] - m * g * inv);
  if (running_mean != nullptr) {
    const double unb = (count > 1) ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  }
}

// ---------------------------------------------------------------------------------- bn_apply_fwd
struct ApplyParams {
  View x[3];
  const float* scale[3];
  const float* shift[3];
  int nb, act, C;
  int64_t pixels;
  __nv_bfloat16* y;
  int64_t y_pitch;
  View res;                  // optional shortcut: y = act(z) + alpha * res (BottleRep, common.py:600-617)
  float alpha;
};
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == YV6_ACT_RELU) return fmaxf(z, 0.f);
  if (act == YV6_ACT_SILU) return z / (1.f + __expf(-z));
  return z;
}
__global__ void __launch_bounds__(kTrThreads) bn_apply_fwd_kernel(const ApplyParams p) {
  const int cgs = p.C / 8;
  const int64_t total = p.pixels * cgs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t px = i / cgs;
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
    for (int b = 0; b < p.nb; ++b) {
      float v[8];
      ld8(p.x[b].p + px * p.x[b].pitch + cg * 8, v);
      const float4 s0 = *reinterpret_cast<const float4*>(p.scale[b] + cg * 8), s1 = *reinterpret_cast<const float4*>(p.scale[b] + cg * 8 + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(p.shift[b] + cg * 8), h1 = *reinterpret_cast<const float4*>(p.shift[b] + cg * 8 + 4);
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += v[j] * sc[j] + sh[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = act_fwd(z[j], p.act);
    if (p.res.p != nullptr) {
      float r[8];
      ld8(p.res.p + px * p.res.pitch + cg * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += p.alpha * r[j];
    }
    st8(p.y + px * p.y_pitch + cg * 8, z);
  }
}

// ---------------------------------------------------------------------------------- bn backward
struct BwdParams {
  View x[3];                 // branch inputs to their BN (raw conv outputs / block input for the identity branch)
  const float* mean[3];
  const float* invstd[3];
  const float* scale[3];     // gamma * invstd
  const float* shift[3];
  View dy;                   // gradient w.r.t. the block output
  View res;                  // optional shortcut input (y = act(z) + alpha * res)
  float alpha;
  __nv_bfloat16* dres;       // g(res) += alpha * dy
  int64_t dres_pitch;
  double* dalpha;            // += sum dy * res
  int nb, act, C;
  int64_t pixels;
  double* s1;                // [C]      sum dz            (shared by the branches)
  double* s2[3];             // [C] each sum dz * xhat_b
  // apply
  __nv_bfloat16* dx[3];
  int64_t dx_pitch[3];
  int accumulate[3];         // dx_b += ... instead of =
  double inv_count;
};

__device__ __forceinline__ void bwd_dz(const BwdParams& p, int64_t px, int cg, float (&dz)[8]) {
  ld8(p.dy.p + px * p.dy.pitch + cg * 8, dz);
  if (p.act != YV6_ACT_NONE) {   // the pre-activation is recomputed from the branch inputs (same fp32 arithmetic as the forward)
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
    for (int b = 0; b < p.nb; ++b) {
      float v[8];
      ld8(p.x[b].p + px * p.x[b].pitch + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += v[j] * p.scale[b][cg * 8 + j] + p.shift[b][cg * 8 + j];
    }
    if (p.act == YV6_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = (z[j] > 0.f) ? dz[j] : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sg = 1.f / (1.f + __expf(-z[j]));
        dz[j] *= sg * (1.f + z[j] * (1.f - sg));
      }
    }
  }
}

__global__ void __launch_bounds__(kTrThreads) bn_bwd_reduce_kernel(const BwdParams p) {
  const int cgs = p.C / 8;
  const int cg = threadIdx.x % cgs;
  const int prow = threadIdx.x / cgs, prows = blockDim.x / cgs;
  float a1[8], a2[3][8];
  float da = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[0][j] = a2[1][j] = a2[2][j] = 0.f; }
  for (int64_t px = (int64_t)blockIdx.x * prows + prow; px < p.pixels; px += (int64_t)gridDim.x * prows) {
    float dz[8];
    if (p.dalpha != nullptr) {
      float g[8], r[8];
      ld8(p.dy.p + px * p.dy.pitch + cg * 8, g);
      ld8(p.res.p + px * p.res.pitch + cg * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) da += g[j] * r[j];
    }
    bwd_dz(p, px, cg, dz);
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] += dz[j];
    for (int b = 0; b < p.nb; ++b) {
      float v[8];
      ld8(p.x[b].p + px * p.x[b].pitch + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[b][j] += dz[j] * (v[j] - p.mean[b][cg * 8 + j]) * p.invstd[b][cg * 8 + j];
    }
  }
  extern __shared__ double sred[];                      // [1 + nb][C]: block-level sums before the global atomics
  for (int c = threadIdx.x; c < (1 + p.nb) * p.C; c += blockDim.x) sred[c] = 0.0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sred[cg * 8 + j], (double)a1[j]);
    for (int b = 0; b < p.nb; ++b) atomicAdd(&sred[(1 + b) * p.C + cg * 8 + j], (double)a2[b][j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    atomicAdd(&p.s1[c], sred[c]);
    for (int b = 0; b < p.nb; ++b) atomicAdd(&p.s2[b][c], sred[(1 + b) * p.C + c]);
  }
  if (p.dalpha != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(p.dalpha, (double)da);
  }
}

__global__ void __launch_bounds__(kTrThreads) bn_bwd_apply_kernel(const BwdParams p) {
  const int cgs = p.C / 8;
  const int64_t total = p.pixels * cgs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const int64_t px = i / cgs;
    float dz[8];
    if (p.dres != nullptr) {
      float g[8], old[8];
      ld8(p.dy.p + px * p.dy.pitch + cg * 8, g);
      __nv_bfloat16* dst = p.dres + px * p.dres_pitch + cg * 8;
      ld8(dst, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) old[j] += p.alpha * g[j];
      st8(dst, old);
    }
    bwd_dz(p, px, cg, dz);
    for (int b = 0; b < p.nb; ++b) {
      float v[8], o[8];
      ld8(p.x[b].p + px * p.x[b].pitch + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        const float xh = (v[j] - p.mean[b][c]) * p.invstd[b][c];
        o[j] = p.scale[b][c] * (dz[j] - (float)(p.s1[c] * p.inv_count) - xh * (float)(p.s2[b][c] * p.inv_count));
      }
      __nv_bfloat16* dst = p.dx[b] + px * p.dx_pitch[b] + cg * 8;
      if (p.accumulate[b]) {
        float old[8];
        ld8(dst, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      st8(dst, o);
    }
  }
}

// ---------------------------------------------------------------------------------- head grad prep
// dlogit = dscore * s * (1 - s) (sigmoid backward, effidehead.py:85) or dreg; repacks level `l` of the
// [B, A, ch] fp32 head tensors into a dense NHWC bf16 tensor [B, H_l, W_l, ch_pad] (zero padded channels).
__global__ void __launch_bounds__(kTrThreads) head_grad_prep_kernel(const float* grad, const float* scores, int B, int A, int ch,
                                                                   int off, int hw, int ch_pad, __nv_bfloat16* out) {
  const int64_t total = (int64_t)B * hw * ch_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ch_pad);
    const int64_t px = i / ch_pad;
    const int b = (int)(px / hw), a = off + (int)(px % hw);
    float v = 0.f;
    if (c < ch) {
      const int64_t src = ((int64_t)b * A + a) * ch + c;
      v = grad[src];
      if (scores != nullptr) { const float s = scores[src]; v *= s * (1.f - s); }
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------- maxpool5 backward
// y = MaxPool2d(5, 1, 2)(x): scatter dy to the arg-max of each window (first maximum in row-major window
// order).  dx is fp32 [N,H,W,C] scratch (zeroed by the caller); x / dy are NHWC bf16 slices.
__global__ void __launch_bounds__(kTrThreads) maxpool5_bwd_kernel(View x, View dy, int N, int H, int W, int C, float* dx) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int w = (int)((i / C) % W), h = (int)((i / ((int64_t)C * W)) % H), n = (int)(i / ((int64_t)C * W * H));
    float best = -INFINITY;
    int bh = h, bw = w;
    for (int dyy = -2; dyy <= 2; ++dyy) {
      const int hh = h + dyy;
      if (hh < 0 || hh >= H) continue;
      for (int dxx = -2; dxx <= 2; ++dxx) {
        const int ww = w + dxx;
        if (ww < 0 || ww >= W) continue;
        const float v = __bfloat162float(x.p[(((int64_t)n * H + hh) * W + ww) * x.pitch + c]);
        if (v > best) { best = v; bh = hh; bw = ww; }
      }
    }
    const float g = __bfloat162float(dy.p[(((int64_t)n * H + h) * W + w) * dy.pitch + c]);
    atomicAdd(&dx[(((int64_t)n * H + bh) * W + bw) * C + c], g);
  }
}

// dst (bf16 slice) (+)= src (fp32 dense [pixels, C])
__global__ void __launch_bounds__(kTrThreads) add_f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, int64_t dst_pitch, int64_t pixels,
                                                                    int C, int accumulate) {
  const int64_t total = pixels * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t px = i / C;
    __nv_bfloat16* d = dst + px * dst_pitch + c;
    float v = src[i];
    if (accumulate) v += __bfloat162float(*d);
    *d = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------- stem wgrad
// dW[co][r][s][c] = sum_pixels dY[p][co] * x[n, c, 2ho + r - 1, 2wo + s - 1]; 27 * Cout outputs.
__global__ void __launch_bounds__(kTrThreads) stem_wgrad_kernel(const void* x, int x_u8, float in_scale, const __nv_bfloat16* dy,
                                                               int64_t dy_pitch, int N, int H, int W, int Cout, float* dw) {
  extern __shared__ float acc[];  // [27 * Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t pixels = (int64_t)N * Ho * Wo;
  // thread = (pixel lane, tap k): each thread owns one of the 27 input taps for a strided set of pixels
  const int k = threadIdx.x % 27, lane = threadIdx.x / 27, lanes = blockDim.x / 27;
  if (lane < lanes) {
    const int r = k / 9, s = (k / 3) % 3, c = k % 3;
    float part[64];
    for (int co = 0; co < Cout; ++co) part[co] = 0.f;
    for (int64_t px = (int64_t)blockIdx.x * lanes + lane; px < pixels; px += (int64_t)gridDim.x * lanes) {
      const int wo = (int)(px % Wo), ho = (int)((px / Wo) % Ho), n = (int)(px / ((int64_t)Wo * Ho));
      const int hi = 2 * ho - 1 + r, wi = 2 * wo - 1 + s;
      if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
      const int64_t idx = (((int64_t)n * 3 + c) * H + hi) * W + wi;
      const float xv = x_u8 ? (float)reinterpret_cast<const uint8_t*>(x)[idx] * in_scale : reinterpret_cast<const float*>(x)[idx];
      const __nv_bfloat16* g = dy + px * dy_pitch;
      for (int co = 0; co < Cout; ++co) part[co] += xv * __bfloat162float(g[co]);
    }
    for (int co = 0; co < Cout; ++co) atomicAdd(&acc[co * 27 + (r * 3 + s) * 3 + c], part[co]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) atomicAdd(&dw[i], acc[i]);
}

}  // namespace yv6

using namespace yv6;

static inline unsigned grid_for(int64_t total, int threads, int num_sms) {
  const int64_t want = (total + threads - 1) / threads;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * 16));
}

extern "C" int yv6_bn_stats(yv6_handle* h, const void* x, int64_t pixels, int32_t C, int64_t pitch, double* sum, double* sumsq,
                            void* stream) {
  YV6_REQUIRE(h && x && sum && sumsq, "bn_stats: null argument");
  YV6_REQUIRE(C % 8 == 0 && C <= 2048 && pitch % 8 == 0, "bn_stats: C=%d pitch=%lld", C, (long long)pitch);
  const int cgs = C / 8;
  int rows = std::max(1, kTrThreads / cgs);
  const int threads = cgs * rows;
  YV6_REQUIRE(threads <= 1024, "bn_stats: too many channels");
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * C, s));
  YV6_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * C, s));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((pixels + rows - 1) / rows, (int64_t)h->num_sms * 8));
  bn_stats_kernel<<<grid, threads, sizeof(double) * 2 * C, s>>>(View{reinterpret_cast<const __nv_bfloat16*>(x), pitch}, pixels, C, sum, sumsq);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_bn_finalize(yv6_handle* h, const double* sum, const double* sumsq, double count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* mean_out, float* invstd_out, float* scale, float* shift, int32_t C, void* stream) {
  YV6_REQUIRE(h && sum && sumsq && gamma && beta && mean_out && invstd_out && scale && shift, "bn_finalize: null argument");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, count, gamma, beta, eps, momentum, running_mean,
                                                                      running_var, mean_out, invstd_out, scale, shift, C);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_bn_apply_fwd(yv6_handle* h, const yv6_bn_desc* d, void* stream) {
  YV6_REQUIRE(h && d && d->nb >= 1 && d->nb <= 3 && d->y, "bn_apply_fwd: bad descriptor");
  YV6_REQUIRE(d->C % 8 == 0, "bn_apply_fwd: C must be a multiple of 8");
  ApplyParams p;
  for (int b = 0; b < d->nb; ++b) {
    p.x[b] = View{reinterpret_cast<const __nv_bfloat16*>(d->x[b]), d->x_pitch[b]};
    p.scale[b] = d->scale[b];
    p.shift[b] = d->shift[b];
  }
  p.nb = d->nb; p.act = d->act; p.C = d->C; p.pixels = d->pixels;
  p.y = reinterpret_cast<__nv_bfloat16*>(d->y); p.y_pitch = d->y_pitch;
  p.res = View{reinterpret_cast<const __nv_bfloat16*>(d->res), d->res_pitch};
  p.alpha = d->res_alpha;
  bn_apply_fwd_kernel<<<grid_for(d->pixels * (d->C / 8), kTrThreads, h->num_sms), kTrThreads, 0, (cudaStream_t)stream>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_bn_bwd(yv6_handle* h, const yv6_bn_desc* d, void* stream) {
  YV6_REQUIRE(h && d && d->nb >= 1 && d->nb <= 3 && d->dy && d->s1, "bn_bwd: bad descriptor");
  YV6_REQUIRE(d->C % 8 == 0 && d->C <= 2048, "bn_bwd: C");
  BwdParams p;
  for (int b = 0; b < d->nb; ++b) {
    p.x[b] = View{reinterpret_cast<const __nv_bfloat16*>(d->x[b]), d->x_pitch[b]};
    p.mean[b] = d->mean[b]; p.invstd[b] = d->invstd[b]; p.scale[b] = d->scale[b]; p.shift[b] = d->shift[b];
    p.s2[b] = d->s2[b];
    p.dx[b] = reinterpret_cast<__nv_bfloat16*>(d->dx[b]); p.dx_pitch[b] = d->dx_pitch[b]; p.accumulate[b] = d->accumulate[b];
  }
  p.dy = View{reinterpret_cast<const __nv_bfloat16*>(d->dy), d->dy_pitch};
  p.res = View{reinterpret_cast<const __nv_bfloat16*>(d->res), d->res_pitch};
  p.alpha = d->res_alpha;
  p.dres = d->res ? reinterpret_cast<__nv_bfloat16*>(d->dres) : nullptr;
  p.dres_pitch = d->dres_pitch;
  p.dalpha = d->res ? d->dalpha : nullptr;
  YV6_REQUIRE(!d->res || (d->dres && d->dalpha), "bn_bwd: shortcut without dres / dalpha");
  p.nb = d->nb; p.act = d->act; p.C = d->C; p.pixels = d->pixels;
  p.s1 = d->s1;
  p.inv_count = 1.0 / (double)d->pixels;
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(d->s1, 0, sizeof(double) * d->C, s));
  for (int b = 0; b < d->nb; ++b) YV6_CHECK_CUDA(cudaMemsetAsync(d->s2[b], 0, sizeof(double) * d->C, s));
  if (p.dalpha) YV6_CHECK_CUDA(cudaMemsetAsync(p.dalpha, 0, sizeof(double), s));
  const int cgs = d->C / 8;
  const int rows = std::max(1, kTrThreads / cgs);
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((d->pixels + rows - 1) / rows, (int64_t)h->num_sms * 8));
  const size_t red_smem = sizeof(double) * (1 + d->nb) * d->C;
  static bool configured = false;
  if (!configured) {
    YV6_CHECK_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * (int)sizeof(double)));
    configured = true;
  }
  bn_bwd_reduce_kernel<<<grid, cgs * rows, red_smem, s>>>(p);
  bn_bwd_apply_kernel<<<grid_for(d->pixels * cgs, kTrThreads, h->num_sms), kTrThreads, 0, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_head_grad_prep(yv6_handle* h, const float* grad, const float* scores_or_null, int32_t B, int32_t A, int32_t ch,
                                  int32_t level_off, int32_t level_hw, int32_t ch_pad, void* out_bf16, void* stream) {
  YV6_REQUIRE(h && grad && out_bf16 && ch_pad >= ch, "head_grad_prep: bad argument");
  const int64_t total = (int64_t)B * level_hw * ch_pad;
  head_grad_prep_kernel<<<grid_for(total, kTrThreads, h->num_sms), kTrThreads, 0, (cudaStream_t)stream>>>(
      grad, scores_or_null, B, A, ch, level_off, level_hw, ch_pad, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_maxpool5_bwd(yv6_handle* h, const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch, int32_t N,
                                int32_t H, int32_t W, int32_t C, float* dx_scratch, void* dx, int64_t dx_pitch, int32_t accumulate,
                                void* stream) {
  YV6_REQUIRE(h && x && dy && dx_scratch && dx, "maxpool5_bwd: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t total = (int64_t)N * H * W * C;
  YV6_CHECK_CUDA(cudaMemsetAsync(dx_scratch, 0, sizeof(float) * total, s));
  maxpool5_bwd_kernel<<<grid_for(total, kTrThreads, h->num_sms), kTrThreads, 0, s>>>(
      View{reinterpret_cast<const __nv_bfloat16*>(x), x_pitch}, View{reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch}, N, H, W, C,
      dx_scratch);
  add_f32_to_bf16_kernel<<<grid_for(total, kTrThreads, h->num_sms), kTrThreads, 0, s>>>(
      dx_scratch, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch, (int64_t)N * H * W, C, accumulate);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_stem_wgrad(yv6_handle* h, const void* x, int32_t x_dtype, float in_scale, const void* dy, int64_t dy_pitch,
                              int32_t N, int32_t H, int32_t W, int32_t Cout, float* dw, void* stream) {
  YV6_REQUIRE(h && x && dy && dw && Cout <= 64, "stem_wgrad: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * Cout, s));
  const int threads = 27 * 9;  // 9 pixel lanes x 27 taps
  stem_wgrad_kernel<<<h->num_sms * 4, threads, sizeof(float) * 27 * Cout, s>>>(x, x_dtype == YV6_DT_U8, in_scale,
                                                                               reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, N, H, W,
                                                                               Cout, dw);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
