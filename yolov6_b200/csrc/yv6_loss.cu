// yv6_loss.cu -- fused detection loss, forward AND backward in one pass.
//
// Reference: ComputeLoss.__call__ (yolov6/models/losses/loss.py:157-182), VarifocalLoss (:201-211),
// BboxLoss (:214-278), IOUloss (yolov6/utils/figure_iou.py:23-100), dist2bbox / bbox2dist
// (yolov6/utils/general.py:32-52).  The reference materialises one-hot labels [B,A,nc+1], runs
// masked_select with dynamic shapes and three host syncs, then autograd walks it all again.  Here:
//   * yv6_box_decode  : pred_distri -> pred_bboxes (xyxy in stride units) for the assigner
//                       (loss.py:194-198, DFL softmax . proj when use_dfl).
//   * yv6_det_loss    : consumes the compact assignment (gt_idx, fg, norm) and writes the three loss
//                       sums plus dL/dpred_scores [B,A,nc] and dL/dpred_distri [B,A,R]:
//       - class term: one warp per anchor row streams the nc scores once (VFL value + gradient,
//         including the gradient through the focal weight);
//       - box term: one thread per positive anchor evaluates the IoU loss (giou/siou/ciou/diou) in
//         float64 forward-mode dual numbers (exact autograd gradients w.r.t. the 4 box coordinates),
//         chains through ltrb decode / DFL softmax, and adds the DFL cross-entropy terms;
//       - normalisation by target_scores_sum (if > 1) and the loss weights are applied on the device
//         (no host sync); partial sums are reduced in a fixed order (deterministic).
#include <algorithm>

#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kLossThreads = 256;
constexpr int kMaxBins = 32;

// ---------------------------------------------------------------------------------- dual numbers
struct D4 {
  double v;
  double d[4];
};
__device__ __forceinline__ D4 dconst(double c) { return D4{c, {0, 0, 0, 0}}; }
__device__ __forceinline__ D4 dvar(double c, int i) {
  D4 r = dconst(c);
  r.d[i] = 1.0;
  return r;
}
__device__ __forceinline__ D4 operator+(const D4& a, const D4& b) {
  return D4{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2], a.d[3] + b.d[3]}};
}
__device__ __forceinline__ D4 operator-(const D4& a, const D4& b) {
  return D4{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2], a.d[3] - b.d[3]}};
}
__device__ __forceinline__ D4 operator*(const D4& a, const D4& b) {
  return D4{a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2],
                        a.d[3] * b.v + a.v * b.d[3]}};
}
__device__ __forceinline__ D4 operator/(const D4& a, const D4& b) {
  const double q = a.v / b.v, ib = 1.0 / b.v;
  return D4{q, {(a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib,
                (a.d[3] - q * b.d[3]) * ib}};
}
__device__ __forceinline__ D4 operator+(const D4& a, double c) { D4 r = a; r.v += c; return r; }
__device__ __forceinline__ D4 operator-(const D4& a, double c) { D4 r = a; r.v -= c; return r; }
__device__ __forceinline__ D4 operator*(const D4& a, double c) {
  return D4{a.v * c, {a.d[0] * c, a.d[1] * c, a.d[2] * c, a.d[3] * c}};
}
__device__ __forceinline__ D4 dchain(const D4& a, double fv, double fp) {  // f(a) with f'(a.v) = fp
  return D4{fv, {a.d[0] * fp, a.d[1] * fp, a.d[2] * fp, a.d[3] * fp}};
}
// torch.min / torch.max of two tensors: the gradient goes to the selected operand, ties split evenly
__device__ __forceinline__ D4 dmin(const D4& a, const D4& b) {
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  return (a + b) * 0.5;
}
__device__ __forceinline__ D4 dmax(const D4& a, const D4& b) {
  if (a.v > b.v) return a;
  if (b.v > a.v) return b;
  return (a + b) * 0.5;
}
__device__ __forceinline__ D4 dclamp0(const D4& a) { return (a.v >= 0.0) ? a : dconst(0.0); }  // clamp(min=0)
__device__ __forceinline__ D4 dabs(const D4& a) { return dchain(a, fabs(a.v), (a.v > 0) - (a.v < 0)); }
__device__ __forceinline__ D4 dsqrt(const D4& a) { const double s = sqrt(a.v); return dchain(a, s, 0.5 / s); }
__device__ __forceinline__ D4 dexp(const D4& a) { const double e = exp(a.v); return dchain(a, e, e); }
__device__ __forceinline__ D4 dcos(const D4& a) { return dchain(a, cos(a.v), -sin(a.v)); }
__device__ __forceinline__ D4 dasin(const D4& a) { return dchain(a, asin(a.v), 1.0 / sqrt(1.0 - a.v * a.v)); }
__device__ __forceinline__ D4 datan(const D4& a) { return dchain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
__device__ __forceinline__ D4 dpowi(const D4& a, int n) {
  const double pn1 = pow(a.v, (double)(n - 1));
  return dchain(a, pn1 * a.v, n * pn1);
}

enum { IOU_GIOU = 0, IOU_SIOU = 1, IOU_CIOU = 2, IOU_DIOU = 3 };

// IOUloss(box_format='xyxy', eps=1e-10) on one pair; b1 = prediction (dual), b2 = target (constants)
__device__ D4 iou_loss_dual(const D4 (&b1)[4], const double (&t)[4], int type, double eps) {
  const D4 b2x1 = dconst(t[0]), b2y1 = dconst(t[1]), b2x2 = dconst(t[2]), b2y2 = dconst(t[3]);
  const D4 inter = dclamp0(dmin(b1[2], b2x2) - dmax(b1[0], b2x1)) * dclamp0(dmin(b1[3], b2y2) - dmax(b1[1], b2y1));
  const D4 w1 = b1[2] - b1[0], h1 = b1[3] - b1[1] + eps;
  const D4 w2 = b2x2 - b2x1, h2 = b2y2 - b2y1 + eps;
  const D4 uni = w1 * h1 + w2 * h2 - inter + eps;
  D4 iou = inter / uni;
  const D4 cw = dmax(b1[2], b2x2) - dmin(b1[0], b2x1);
  const D4 ch = dmax(b1[3], b2y2) - dmin(b1[1], b2y1);
  if (type == IOU_GIOU) {
    const D4 c_area = cw * ch + eps;
    iou = iou - (c_area - uni) / c_area;
  } else if (type == IOU_DIOU || type == IOU_CIOU) {
    const D4 c2 = cw * cw + ch * ch + eps;
    const D4 dx = b2x1 + b2x2 - b1[0] - b1[2], dy = b2y1 + b2y2 - b1[1] - b1[3];
    const D4 rho2 = (dx * dx + dy * dy) * 0.25;
    if (type == IOU_DIOU) {
      iou = iou - rho2 / c2;
    } else {
      const D4 da = datan(w2 / h2) - datan(w1 / h1);
      const D4 v = (da * da) * (4.0 / (M_PI * M_PI));
      const double alpha = v.v / (v.v - iou.v + (1.0 + eps));  // computed under no_grad in the reference
      iou = iou - (rho2 / c2 + v * alpha);
    }
  } else {  // SIoU, figure_iou.py:75-92
    const D4 s_cw = (b2x1 + b2x2 - b1[0] - b1[2]) * 0.5 + eps;
    const D4 s_ch = (b2y1 + b2y2 - b1[1] - b1[3]) * 0.5 + eps;
    const D4 sigma = dsqrt(s_cw * s_cw + s_ch * s_ch);
    const D4 sin1 = dabs(s_cw) / sigma, sin2 = dabs(s_ch) / sigma;
    const D4 sin_a = (sin1.v > sqrt(2.0) / 2.0) ? sin2 : sin1;
    const D4 angle_cost = dcos(dasin(sin_a) * 2.0 - M_PI / 2.0);
    const D4 rx = s_cw / cw, ry = s_ch / ch;
    const D4 gamma = angle_cost - 2.0;
    const D4 dist_cost = dconst(2.0) - dexp(gamma * (rx * rx)) - dexp(gamma * (ry * ry));
    const D4 ow = dabs(w1 - w2) / dmax(w1, w2), oh = dabs(h1 - h2) / dmax(h1, h2);
    const D4 shape_cost = dpowi(dconst(1.0) - dexp(ow * -1.0), 4) + dpowi(dconst(1.0) - dexp(oh * -1.0), 4);
    iou = iou - (dist_cost + shape_cost) * 0.5;
  }
  return dconst(1.0) - iou;
}

// ---------------------------------------------------------------------------------- decode
struct DecodeBoxParams {
  const float* distri;   // [B,A,R]
  const float* points;   // [A,2] anchor centres in pixels
  const float* strides;  // [A]
  float* boxes;          // [B,A,4] xyxy in stride units (scale_out = 0) or pixels (scale_out = 1)
  int32_t B, A, R, reg_max, scale_out;
};

// ltrb distances of one anchor: plain (R == 4) or DFL expectation (softmax . [0..reg_max]) in f32 like torch
__device__ __forceinline__ void ltrb_f32(const float* reg, int R, int reg_max, float (&d)[4]) {
  if (R == 4) {
#pragma unroll
    for (int s = 0; s < 4; ++s) d[s] = reg[s];
    return;
  }
  const int nb = reg_max + 1;
  for (int s = 0; s < 4; ++s) {
    const float* l = reg + s * nb;
    float mx = -INFINITY;
    for (int i = 0; i < nb; ++i) mx = fmaxf(mx, l[i]);
    float den = 0.f;
    for (int i = 0; i < nb; ++i) den += expf(l[i] - mx);
    float e = 0.f;
    for (int i = 0; i < nb; ++i) e += (expf(l[i] - mx) / den) * (float)i;
    d[s] = e;
  }
}

__global__ void __launch_bounds__(kLossThreads) box_decode_kernel(const DecodeBoxParams p) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)p.B * p.A) return;
  const int a = (int)(o % p.A);
  float d[4];
  ltrb_f32(p.distri + o * p.R, p.R, p.reg_max, d);
  const float s = p.strides[a];
  const float cx = __fdiv_rn(p.points[2 * a], s), cy = __fdiv_rn(p.points[2 * a + 1], s);  // loss.py:82
  float x1 = __fsub_rn(cx, d[0]), y1 = __fsub_rn(cy, d[1]), x2 = __fadd_rn(cx, d[2]), y2 = __fadd_rn(cy, d[3]);
  if (p.scale_out) { x1 = __fmul_rn(x1, s); y1 = __fmul_rn(y1, s); x2 = __fmul_rn(x2, s); y2 = __fmul_rn(y2, s); }
  reinterpret_cast<float4*>(p.boxes)[o] = make_float4(x1, y1, x2, y2);
}

// ---------------------------------------------------------------------------------- loss
struct LossParams {
  const float* scores;    // [B,A,nc] post-sigmoid
  const float* distri;    // [B,A,R]
  const float* points;    // [A,2] pixels
  const float* strides;   // [A]
  const double* gt;       // [B,G,5]
  const int32_t* gt_idx;  // [B,A]
  const uint8_t* fg;      // [B,A]
  const double* norm;     // [B,A]
  int32_t B, A, G, nc, R, reg_max, use_dfl, iou_type;
  double w_cls, w_iou, w_dfl, grad_scale;
  float* grad_scores;     // [B,A,nc]
  float* grad_distri;     // [B,A,R]
  double* partial;        // [4][nblk_max] : tss, cls, iou, dfl partial sums
  double norm_thr;        // divide by target_scores_sum when it exceeds this: 1 (loss.py:168-169) or 0 (loss_fuseab.py:139, 203-206)
  double* out;            // [8]: loss, w_iou*iou, w_dfl*dfl, w_cls*cls, tss, num_pos, raw sums...
  int32_t nblk_rows, nblk_cls, nblk_box;
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
  __syncthreads();
  return s;  // valid on thread 0
}

// pass 0: target_scores_sum and number of positives (loss.py:165, 225)
__global__ void __launch_bounds__(kLossThreads) loss_tss_kernel(const LossParams p) {
  __shared__ double sh[kLossThreads / 32];
  const int64_t n = (int64_t)p.B * p.A;
  double s = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (p.fg[i]) { s += p.norm[i]; c += 1.0; }
  const double bs = block_sum(s, sh);
  const double bc = block_sum(c, sh);
  if (threadIdx.x == 0) {
    p.partial[blockIdx.x] = bs;
    p.partial[p.nblk_rows + blockIdx.x] = bc;
  }
}
__global__ void loss_tss_final_kernel(const LossParams p) {
  if (threadIdx.x == 0) {
    double s = 0.0, c = 0.0;
    for (int i = 0; i < p.nblk_rows; ++i) { s += p.partial[i]; c += p.partial[p.nblk_rows + i]; }
    p.out[4] = s;
    p.out[5] = c;
  }
}

// pass 1: varifocal loss value + gradient, one warp per anchor row (loss.py:205-211)
__global__ void __launch_bounds__(kLossThreads) loss_cls_kernel(const LossParams p) {
  __shared__ double sh[kLossThreads / 32];
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int64_t rows = (int64_t)p.B * p.A;
  const double tss = p.out[4];
  const double denom = (tss > p.norm_thr) ? tss : 1.0;                       // loss.py:168-169
  const double gscale = p.w_cls * p.grad_scale / denom;
  double acc = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * warps) {
    const bool fg = p.fg[row] != 0;
    int label = -1;
    double q64 = 0.0;
    if (fg) {
      const int b = (int)(row / p.A);
      label = (int)p.gt[((int64_t)b * p.G + p.gt_idx[row]) * 5];
      q64 = p.norm[row];
    }
    const float* s = p.scores + row * p.nc;
    float* g = p.grad_scores + row * p.nc;
    for (int c = lane; c < p.nc; c += 32) {
      const float pr = __ldg(s + c);
      const bool pos = (c == label);
      const float q = pos ? (float)q64 : 0.f;
      // F.binary_cross_entropy in f32 with log clamped at -100
      const float lp = fmaxf(logf(pr), -100.f), l1p = fmaxf(log1pf(-pr), -100.f);
      const float bce = -(q * lp + (1.f - q) * l1p);
      const float wneg = 0.75f * (pr * pr);                           // alpha * p^gamma * (1 - label), f32
      const double w = pos ? q64 : (double)wneg;
      acc += (double)bce * w;
      const float dbce = (pr - q) / fmaxf((1.f - pr) * pr, 1e-12f);   // torch's BCE backward
      const double dw = pos ? 0.0 : 1.5 * (double)pr;
      g[c] = (float)(((double)dbce * w + (double)bce * dw) * gscale);
    }
  }
  const double bs = block_sum(acc, sh);
  if (threadIdx.x == 0) p.partial[2 * p.nblk_rows + blockIdx.x] = bs;
}

// pass 2: IoU + DFL terms and the gradient w.r.t. pred_distri, one thread per anchor (loss.py:222-278)
__global__ void __launch_bounds__(kLossThreads) loss_box_kernel(const LossParams p) {
  __shared__ double sh[kLossThreads / 32];
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t rows = (int64_t)p.B * p.A;
  double l_iou = 0.0, l_dfl = 0.0;
  if (o < rows) {
    float* gd = p.grad_distri + o * p.R;
    if (!p.fg[o]) {
      for (int j = 0; j < p.R; ++j) gd[j] = 0.f;
    } else {
      const int a = (int)(o % p.A), b = (int)(o / p.A);
      const double tss = p.out[4];
      const double denom = (tss > p.norm_thr) ? tss : 1.0;
      const double bw = p.norm[o];                                    // bbox_weight = sum_c target_scores
      const float* reg = p.distri + o * p.R;
      const float s = p.strides[a];
      const float cx = __fdiv_rn(p.points[2 * a], s), cy = __fdiv_rn(p.points[2 * a + 1], s);
      float d[4];
      ltrb_f32(reg, p.R, p.reg_max, d);
      const float pbx[4] = {__fsub_rn(cx, d[0]), __fsub_rn(cy, d[1]), __fadd_rn(cx, d[2]), __fadd_rn(cy, d[3])};
      const double* gr = p.gt + ((int64_t)b * p.G + p.gt_idx[o]) * 5;
      const double t[4] = {gr[1] / (double)s, gr[2] / (double)s, gr[3] / (double)s, gr[4] / (double)s};  // loss.py:158
      D4 b1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b1[j] = dvar((double)pbx[j], j);
      const D4 li = iou_loss_dual(b1, t, p.iou_type, 1e-10);
      l_iou = li.v * bw;
      const double ci = p.w_iou * p.grad_scale * bw / denom;
      // d loss / d(l, t, r, b): x1 = cx - l, y1 = cy - t, x2 = cx + r, y2 = cy + b
      double gl[4] = {-li.d[0] * ci, -li.d[1] * ci, li.d[2] * ci, li.d[3] * ci};
      if (p.R == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gd[j] = (float)gl[j];
      } else {
        const int nb = p.reg_max + 1;
        const double cd = p.w_dfl * p.grad_scale * bw / denom * 0.25;  // mean over the 4 sides
        const double ctr[2] = {(double)cx, (double)cy};
        for (int side = 0; side < 4; ++side) {
          const float* l = reg + side * nb;
          float mx = -INFINITY;
          for (int i = 0; i < nb; ++i) mx = fmaxf(mx, l[i]);
          float den = 0.f;
          for (int i = 0; i < nb; ++i) den += expf(l[i] - mx);
          const float lse = mx + logf(den);
          // bbox2dist (general.py:46-52): target ltrb clipped to [0, reg_max - 0.01]
          double tv = (side < 2) ? (ctr[side] - t[side]) : (t[side] - ctr[side - 2]);
          tv = fmin(fmax(tv, 0.0), (double)p.reg_max - 0.01);
          const int tl = (int)tv, tr = tl + 1;
          const double wl = (double)(float)tr - tv, wr = 1.0 - wl;     // loss.py:270-271
          const float ce_l = lse - l[tl], ce_r = lse - l[tr];           // F.cross_entropy in f32
          l_dfl += ((double)ce_l * wl + (double)ce_r * wr) * 0.25 * bw;
          for (int i = 0; i < nb; ++i) {
            const double sm = (double)(expf(l[i] - mx) / den);
            double gi = gl[side] * sm * ((double)i - (double)d[side]);  // through the DFL expectation
            gi += cd * (sm * (wl + wr) - (i == tl ? wl : 0.0) - (i == tr ? wr : 0.0));
            gd[side * nb + i] = (float)gi;
          }
        }
      }
    }
  }
  const double bi = block_sum(l_iou, sh);
  const double bd = block_sum(l_dfl, sh);
  if (threadIdx.x == 0) {
    p.partial[3 * p.nblk_rows + blockIdx.x] = bi;
    p.partial[3 * p.nblk_rows + p.nblk_box + blockIdx.x] = bd;
  }
}

__global__ void loss_final_kernel(const LossParams p) {
  if (threadIdx.x != 0) return;
  double cls = 0.0, iou = 0.0, dfl = 0.0;
  for (int i = 0; i < p.nblk_cls; ++i) cls += p.partial[2 * p.nblk_rows + i];
  for (int i = 0; i < p.nblk_box; ++i) {
    iou += p.partial[3 * p.nblk_rows + i];
    dfl += p.partial[3 * p.nblk_rows + p.nblk_box + i];
  }
  const double tss = p.out[4];
  const double denom = (tss > p.norm_thr) ? tss : 1.0;
  cls /= denom;
  iou /= denom;
  dfl = p.use_dfl ? dfl / denom : 0.0;
  p.out[0] = p.w_cls * cls + p.w_iou * iou + p.w_dfl * dfl;  // loss.py:175-177
  p.out[1] = p.w_iou * iou;                                  // loss_items order iou, dfl, cls (loss.py:179-182)
  p.out[2] = p.w_dfl * dfl;
  p.out[3] = p.w_cls * cls;
}

}  // namespace yv6

using namespace yv6;

extern "C" int yv6_box_decode(yv6_handle* h, const float* pred_distri, const float* anc_points, const float* strides,
                              int32_t B, int32_t A, int32_t reg_ch, int32_t scale_to_pixels, float* boxes, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && pred_distri && anc_points && strides && boxes, "box_decode: null argument");
  YV6_REQUIRE(reg_ch == 4 || (reg_ch % 4 == 0 && reg_ch / 4 <= kMaxBins), "box_decode: reg_ch=%d", reg_ch);
  DecodeBoxParams p{pred_distri, anc_points, strides, boxes, B, A, reg_ch, reg_ch / 4 - 1, scale_to_pixels};
  const int64_t n = (int64_t)B * A;
  box_decode_kernel<<<(unsigned)((n + kLossThreads - 1) / kLossThreads), kLossThreads, 0, (cudaStream_t)stream>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int64_t yv6_det_loss_workspace_bytes(int32_t B, int32_t A) {
  const int64_t nblk = ((int64_t)B * A + kLossThreads - 1) / kLossThreads;
  return (5 * nblk + 16) * 8;
}

extern "C" int yv6_det_loss(yv6_handle* h, const yv6_loss_desc* d, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && d, "det_loss: null argument");
  YV6_REQUIRE(d->pred_scores && d->pred_distri && d->anc_points && d->strides && d->gt_idx && d->fg && d->norm &&
                  d->grad_scores && d->grad_distri && d->out && d->workspace,
              "det_loss: null tensor");
  YV6_REQUIRE(d->G == 0 || d->gt, "det_loss: null gt");
  YV6_REQUIRE(d->reg_ch == 4 || (d->reg_ch % 4 == 0 && d->reg_ch / 4 <= kMaxBins), "det_loss: reg_ch=%d", d->reg_ch);
  YV6_REQUIRE(d->iou_type >= 0 && d->iou_type <= 3, "det_loss: iou_type");
  YV6_REQUIRE(d->workspace_bytes >= yv6_det_loss_workspace_bytes(d->B, d->A), "det_loss: workspace too small");
  LossParams p;
  p.scores = d->pred_scores; p.distri = d->pred_distri; p.points = d->anc_points; p.strides = d->strides;
  p.gt = d->gt; p.gt_idx = d->gt_idx; p.fg = d->fg; p.norm = d->norm;
  p.B = d->B; p.A = d->A; p.G = d->G; p.nc = d->nc; p.R = d->reg_ch; p.reg_max = d->reg_ch / 4 - 1;
  p.use_dfl = (d->reg_ch > 4); p.iou_type = d->iou_type;
  p.w_cls = d->w_cls; p.w_iou = d->w_iou; p.w_dfl = d->w_dfl; p.grad_scale = d->grad_scale;
  p.norm_thr = d->norm_gt_zero ? 0.0 : 1.0;
  p.grad_scores = d->grad_scores; p.grad_distri = d->grad_distri;
  p.partial = reinterpret_cast<double*>(d->workspace);
  p.out = d->out;
  const int64_t rows = (int64_t)d->B * d->A;
  p.nblk_box = (int)((rows + kLossThreads - 1) / kLossThreads);
  p.nblk_rows = p.nblk_box;
  p.nblk_cls = std::min(p.nblk_rows, h->num_sms * 8);
  const int nblk_tss = std::min(p.nblk_rows, h->num_sms * 4);
  cudaStream_t s = (cudaStream_t)stream;
  LossParams pt = p;
  pt.nblk_rows = nblk_tss;  // tss partials live in the first 2*nblk_tss slots
  loss_tss_kernel<<<nblk_tss, kLossThreads, 0, s>>>(pt);
  loss_tss_final_kernel<<<1, 32, 0, s>>>(pt);
  loss_cls_kernel<<<p.nblk_cls, kLossThreads, 0, s>>>(p);
  loss_box_kernel<<<p.nblk_box, kLossThreads, 0, s>>>(p);
  loss_final_kernel<<<1, 32, 0, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
