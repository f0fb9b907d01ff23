// yv6_assign.cu -- label assignment on the GPU without any [B,G,A] tensor in HBM.
//
//   * yv6_targets_pad : ComputeLoss.preprocess (reference yolov6/models/losses/loss.py:184-192):
//                       ragged [n,6] targets -> padded [B,G,5] float64 (cls, xyxy pixels), on device.
//   * yv6_tal_assign  : TaskAlignedAssigner.forward (yolov6/assigners/tal_assigner.py:22-173 with
//                       assigner_utils.py:25-89): per (image, gt) block streams the anchors once,
//                       keeps the top-k alignment metrics in registers, second pass resolves anchors
//                       claimed by several gts, third pass normalises.  Output is compact
//                       (gt_idx, fg, norm) -- `yv6_assign_expand` materialises the reference's dense
//                       (labels i64, bboxes f64, scores f64, fg) tensors for the drop-in API.
//   * yv6_atss_assign : ATSSAssigner.forward (yolov6/assigners/atss_assigner.py:18-161).
//
// Everything derived from the ground truth is float64 exactly as in the reference (its targets are
// numpy float64), and mixed f32/f64 operations promote the way torch does, so integer
// outputs (gt_idx, fg, labels) match away from exact metric ties.  Ties in top-k go to the lower
// anchor index (torch.topk leaves them unspecified).
#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kTopkMax = 32;   // 13 in ComputeLoss (loss.py:41), 26 in the fuse_ab loss (loss_fuseab.py:40)
constexpr int kAssignThreads = 256;

// ------------------------------------------------------------------------------------------------
// targets -> padded gt
// ------------------------------------------------------------------------------------------------
__global__ void targets_pad_kernel(const float* __restrict__ t, int n, int B, int G, float sw, float sh,
                                   double* __restrict__ gt, int32_t* __restrict__ gt_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int img = (int)t[i * 6];
  if (img < 0 || img >= B) return;
  int rank = 0;  // position among the rows of the same image, in input order (loss.py:186-187)
  for (int j = 0; j < i; ++j) rank += ((int)t[j * 6] == img);
  atomicAdd(&gt_count[img], 1);
  if (rank >= G) return;
  // loss.py:190-191: xywh (f64) * scale (f32 tensor, promoted), then general.py:55-61 in place
  const double cx = (double)t[i * 6 + 2] * (double)sw, cy = (double)t[i * 6 + 3] * (double)sh;
  const double w = (double)t[i * 6 + 4] * (double)sw, h = (double)t[i * 6 + 5] * (double)sh;
  const double x1 = cx - w * 0.5, y1 = cy - h * 0.5;
  double* o = gt + ((int64_t)img * G + rank) * 5;
  o[0] = (double)t[i * 6 + 1];
  o[1] = x1;
  o[2] = y1;
  o[3] = x1 + w;
  o[4] = y1 + h;
}

__global__ void targets_fill_kernel(double* gt, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double* o = gt + i * 5;
  o[0] = -1.0;  // loss.py:189 pad rows [-1, 0, 0, 0, 0]
  o[1] = o[2] = o[3] = o[4] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// shared geometry (float64 with torch's promotion rules)
// ------------------------------------------------------------------------------------------------
struct GtBox {
  double x1, y1, x2, y2;
};

// iou_calculator (assigner_utils.py:69-89): box1 = gt (f64), box2 = pred (f32; its area is an f32 product)
__device__ __forceinline__ double pair_iou(const GtBox& g, const float4& p, double eps) {
  const double ix1 = fmax(g.x1, (double)p.x), iy1 = fmax(g.y1, (double)p.y);
  const double ix2 = fmin(g.x2, (double)p.z), iy2 = fmin(g.y2, (double)p.w);
  const double overlap = fmax(ix2 - ix1, 0.0) * fmax(iy2 - iy1, 0.0);
  const double area1 = fmax(g.x2 - g.x1, 0.0) * fmax(g.y2 - g.y1, 0.0);
  const float area2 = __fmul_rn(fmaxf(__fsub_rn(p.z, p.x), 0.f), fmaxf(__fsub_rn(p.w, p.y), 0.f));
  const double uni = area1 + (double)area2 - overlap + eps;
  return overlap / uni;
}

// select_candidates_in_gts (assigner_utils.py:25-44): min(ltrb deltas) > eps, centre f32 vs gt f64
__device__ __forceinline__ bool centre_in_gt(const GtBox& g, float cx, float cy, double eps = 1e-9) {
  const double d = fmin(fmin((double)cx - g.x1, (double)cy - g.y1), fmin(g.x2 - (double)cx, g.y2 - (double)cy));
  return d > eps;
}

struct TalParams {
  const float* scores;   // [B,A,nc]
  const float* boxes;    // [B,A,4] pixels
  const float* points;   // [A,2] pixels
  const double* gt;      // [B,G,5]
  const uint8_t* mask;   // [B,G]
  int32_t B, A, G, nc, topk;
  double alpha, beta, eps;
  int32_t* cnt;          // [B,A] scratch (zeroed)
  int32_t* sel;          // [B,A] scratch
  double* al_a;          // [B,A] scratch
  double* ov_a;          // [B,A] scratch
  unsigned long long* pos_al;  // [B,G] scratch (zeroed), bit patterns of non-negative doubles
  unsigned long long* pos_ov;  // [B,G]
  int32_t* gt_idx;       // [B,A] out
  uint8_t* fg;           // [B,A] out
  double* norm;          // [B,A] out: target score at the assigned class (0 for background)
};

__device__ __forceinline__ double align_metric(double sc, double ov, double alpha, double beta) {
  return pow(sc, alpha) * pow(ov, beta);  // tal_assigner.py:131
}

__device__ __forceinline__ bool better(double v, int i, double bv, int bi) { return v > bv || (v == bv && i < bi); }

// one block per (gt, image): top-k of (metric * in_gt) over the anchors
__global__ void __launch_bounds__(kAssignThreads) tal_topk_kernel(const TalParams p) {
  const int g = blockIdx.x, b = blockIdx.y;
  if (!p.mask[b * p.G + g]) return;  // padded gt: its top-k is zeroed by the duplicate-index rule (tal_assigner.py:146-149)
  const double* gr = p.gt + ((int64_t)b * p.G + g) * 5;
  int label = (int)gr[0];
  if (label < 0) label += p.nc;  // torch negative index (tal_assigner.py:123-128)
  label = min(max(label, 0), p.nc - 1);
  const GtBox gb{gr[1], gr[2], gr[3], gr[4]};
  const int K = p.topk;
  double lv[kTopkMax];
  int li[kTopkMax];
#pragma unroll
  for (int k = 0; k < kTopkMax; ++k) { lv[k] = -1.0; li[k] = 0x7fffffff; }
  for (int a = threadIdx.x; a < p.A; a += kAssignThreads) {
    const float2 pt = reinterpret_cast<const float2*>(p.points)[a];
    double v = 0.0;
    if (centre_in_gt(gb, pt.x, pt.y)) {
      const float4 pb = reinterpret_cast<const float4*>(p.boxes)[(int64_t)b * p.A + a];
      const double ov = pair_iou(gb, pb, p.eps);
      const double sc = (double)p.scores[((int64_t)b * p.A + a) * p.nc + label];
      v = align_metric(sc, ov, p.alpha, p.beta);
    }
    if (better(v, a, lv[K - 1], li[K - 1])) {  // insertion into the sorted local list
      int k = K - 1;
      while (k > 0 && better(v, a, lv[k - 1], li[k - 1])) { lv[k] = lv[k - 1]; li[k] = li[k - 1]; --k; }
      lv[k] = v;
      li[k] = a;
    }
  }
  // K rounds of block-wide arg-best over the heads of the local lists
  __shared__ double s_v[kAssignThreads / 32];
  __shared__ int s_i[kAssignThreads / 32];
  __shared__ int s_t[kAssignThreads / 32];
  __shared__ int winner_t, winner_i;
  __shared__ int chosen[kTopkMax];
  int head = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < K; ++r) {
    double v = (head < K) ? lv[head] : -2.0;
    int i = (head < K) ? li[head] : 0x7fffffff;
    int t = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      const int ot = __shfl_xor_sync(0xffffffffu, t, o);
      if (better(ov, oi, v, i)) { v = ov; i = oi; t = ot; }
    }
    if (lane == 0) { s_v[warp] = v; s_i[warp] = i; s_t[warp] = t; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double bv = s_v[0];
      int bi = s_i[0], bt = s_t[0];
      for (int w = 1; w < kAssignThreads / 32; ++w)
        if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; bt = s_t[w]; }
      winner_t = bt;
      winner_i = bi;
      chosen[r] = (bv >= 0.0) ? bi : -1;
    }
    __syncthreads();
    if (threadIdx.x == winner_t && winner_i != 0x7fffffff) ++head;
    __syncthreads();
  }
  if (threadIdx.x < K) {
    const int a = chosen[threadIdx.x];
    if (a >= 0 && a < p.A) {
      const float2 pt = reinterpret_cast<const float2*>(p.points)[a];
      if (centre_in_gt(gb, pt.x, pt.y)) {  // mask_pos = in_topk * in_gts * mask_gt (tal_assigner.py:113)
        atomicAdd(&p.cnt[(int64_t)b * p.A + a], 1);
        atomicExch(&p.sel[(int64_t)b * p.A + a], g);
      }
    }
  }
}

// per anchor: resolve multi-gt claims by highest IoU over ALL G rows (assigner_utils.py:46-67)
__global__ void __launch_bounds__(kAssignThreads) tal_resolve_kernel(const TalParams p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const int64_t o = (int64_t)b * p.A + a;
  const int c = p.cnt[o];
  if (c == 0) {
    p.gt_idx[o] = 0;
    p.fg[o] = 0;
    p.norm[o] = 0.0;
    return;
  }
  const float4 pb = reinterpret_cast<const float4*>(p.boxes)[o];
  int gi = p.sel[o];
  if (c > 1) {
    double best = -1.0;
    gi = 0;
    for (int g = 0; g < p.G; ++g) {
      const double* gr = p.gt + ((int64_t)b * p.G + g) * 5;
      const double ov = pair_iou(GtBox{gr[1], gr[2], gr[3], gr[4]}, pb, p.eps);
      if (ov > best) { best = ov; gi = g; }  // first max
    }
  }
  const double* gr = p.gt + ((int64_t)b * p.G + gi) * 5;
  int label = (int)gr[0];
  if (label < 0) label += p.nc;
  label = min(max(label, 0), p.nc - 1);
  const double ov = pair_iou(GtBox{gr[1], gr[2], gr[3], gr[4]}, pb, p.eps);
  const double al = align_metric((double)p.scores[o * p.nc + label], ov, p.alpha, p.beta);
  p.gt_idx[o] = gi;
  p.fg[o] = 1;
  p.al_a[o] = al;
  p.ov_a[o] = ov;
  atomicMax(&p.pos_al[b * p.G + gi], (unsigned long long)__double_as_longlong(al));   // >= 0: bit order = value order
  atomicMax(&p.pos_ov[b * p.G + gi], (unsigned long long)__double_as_longlong(ov));
}

// per anchor: target score = metric * max-IoU-of-its-gt / (max-metric-of-its-gt + eps) (tal_assigner.py:77-81)
__global__ void __launch_bounds__(kAssignThreads) tal_finalize_kernel(const TalParams p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const int64_t o = (int64_t)b * p.A + a;
  if (!p.fg[o]) return;
  const int gi = p.gt_idx[o];
  const double pa = __longlong_as_double((long long)p.pos_al[b * p.G + gi]);
  const double po = __longlong_as_double((long long)p.pos_ov[b * p.G + gi]);
  p.norm[o] = p.al_a[o] * po / (pa + p.eps);
}

// dense expansion for the drop-in assigner API (tal_assigner.py:151-173 / atss_assigner.py:136-161)
struct ExpandParams {
  const double* gt;
  const int32_t* gt_idx;
  const uint8_t* fg;
  const double* norm;
  int32_t B, A, G, nc, bg_label;  // bg_label < 0: TAL convention (label of gt_idx, clamped); else ATSS (bg = nc)
  int64_t* labels;
  double* bboxes;
  double* scores;  // [B,A,nc] zero-filled by the caller
  uint8_t* fg_out;
};
__global__ void __launch_bounds__(kAssignThreads) assign_expand_kernel(const ExpandParams p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const int64_t o = (int64_t)b * p.A + a;
  const int gi = p.gt_idx[o];
  const double* gr = p.gt + ((int64_t)b * p.G + gi) * 5;
  int label = (int)gr[0];
  const bool fg = p.fg[o] != 0;
  if (p.bg_label >= 0) label = fg ? label : p.bg_label;
  else if (label < 0) label = 0;  // tal_assigner.py:165
  p.labels[o] = label;
  p.bboxes[o * 4 + 0] = gr[1];
  p.bboxes[o * 4 + 1] = gr[2];
  p.bboxes[o * 4 + 2] = gr[3];
  p.bboxes[o * 4 + 3] = gr[4];
  p.fg_out[o] = fg;
  if (fg && label >= 0 && label < p.nc) p.scores[o * p.nc + label] = p.norm[o];
}

// ------------------------------------------------------------------------------------------------
// ATSS
// ------------------------------------------------------------------------------------------------
constexpr int kAtssMaxLevels = 6;
constexpr int kAtssTopkMax = 9;
struct AtssParams {
  const float* anchors;   // [A,4] anchor boxes (f32)
  const float* pd_boxes;  // [B,A,4] pixels (may be null -> no soft label)
  const double* gt;
  const uint8_t* mask;
  int32_t B, A, G, nc, topk, nl;
  int32_t lvl_off[kAtssMaxLevels + 1];
  int32_t* cnt;
  int32_t* sel;
  int32_t* gt_idx;
  uint8_t* fg;
  double* norm;           // soft label = IoU(gt, pred) of the assigned gt
};

// bbox_overlaps(mode='iou') (iou2d_calculator.py:201-243): gt f64 vs anchor f32 promoted, union = max(.., 1e-6)
__device__ __forceinline__ double anchor_iou(const GtBox& g, const float4& an) {
  const double area1 = (g.x2 - g.x1) * (g.y2 - g.y1);
  const double ax1 = an.x, ay1 = an.y, ax2 = an.z, ay2 = an.w;
  const double area2 = (ax2 - ax1) * (ay2 - ay1);
  const double w = fmax(fmin(g.x2, ax2) - fmax(g.x1, ax1), 0.0), h = fmax(fmin(g.y2, ay2) - fmax(g.y1, ay1), 0.0);
  const double inter = w * h;
  return inter / fmax(area1 + area2 - inter, 1e-6);
}

// one block per (gt, image): per level the `topk` nearest anchor centres, then mean+std IoU threshold
__global__ void __launch_bounds__(kAssignThreads) atss_candidates_kernel(const AtssParams p) {
  const int g = blockIdx.x, b = blockIdx.y;
  if (!p.mask[b * p.G + g]) return;
  const double* gr = p.gt + ((int64_t)b * p.G + g) * 5;
  const GtBox gb{gr[1], gr[2], gr[3], gr[4]};
  const double gcx = (gb.x1 + gb.x2) / 2.0, gcy = (gb.y1 + gb.y2) / 2.0;
  __shared__ double s_v[kAssignThreads / 32];
  __shared__ int s_i[kAssignThreads / 32];
  __shared__ int s_t[kAssignThreads / 32];
  __shared__ int winner_t, winner_i;
  __shared__ int cand[kAtssMaxLevels * kAtssTopkMax];
  __shared__ double cand_ov[kAtssMaxLevels * kAtssTopkMax];
  __shared__ int n_cand;
  if (threadIdx.x == 0) n_cand = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int l = 0; l < p.nl; ++l) {
    const int a0 = p.lvl_off[l], a1 = p.lvl_off[l + 1];
    const int K = min(p.topk, a1 - a0);
    double lv[kAtssTopkMax];
    int li[kAtssTopkMax];
#pragma unroll
    for (int k = 0; k < kAtssTopkMax; ++k) { lv[k] = -1e300; li[k] = 0x7fffffff; }
    for (int a = a0 + threadIdx.x; a < a1; a += kAssignThreads) {
      const float4 an = reinterpret_cast<const float4*>(p.anchors)[a];
      // dist_calculator (assigner_utils.py:4-23): centres of f32 anchors are f32, distance promotes to f64
      const float acx = __fdiv_rn(__fadd_rn(an.x, an.z), 2.f), acy = __fdiv_rn(__fadd_rn(an.y, an.w), 2.f);
      const double dx = gcx - (double)acx, dy = gcy - (double)acy;
      const double v = -sqrt(dx * dx + dy * dy);  // negate: "smallest distance" = best
      if (better(v, a, lv[K - 1], li[K - 1])) {
        int k = K - 1;
        while (k > 0 && better(v, a, lv[k - 1], li[k - 1])) { lv[k] = lv[k - 1]; li[k] = li[k - 1]; --k; }
        lv[k] = v;
        li[k] = a;
      }
    }
    int head = 0;
    for (int r = 0; r < K; ++r) {
      double v = (head < K) ? lv[head] : -1e301;
      int i = (head < K) ? li[head] : 0x7fffffff;
      int t = threadIdx.x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        const int ot = __shfl_xor_sync(0xffffffffu, t, o);
        if (better(ov, oi, v, i)) { v = ov; i = oi; t = ot; }
      }
      if (lane == 0) { s_v[warp] = v; s_i[warp] = i; s_t[warp] = t; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double bv = s_v[0];
        int bi = s_i[0], bt = s_t[0];
        for (int w = 1; w < kAssignThreads / 32; ++w)
          if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; bt = s_t[w]; }
        winner_t = bt;
        winner_i = bi;
        if (bi != 0x7fffffff) {
          cand[n_cand] = bi;
          cand_ov[n_cand] = anchor_iou(gb, reinterpret_cast<const float4*>(p.anchors)[bi]);
          ++n_cand;
        }
      }
      __syncthreads();
      if (threadIdx.x == winner_t && winner_i != 0x7fffffff) ++head;
      __syncthreads();
    }
  }
  // threshold = mean + unbiased std of the candidate IoUs (atss_assigner.py:132-134)
  __shared__ double thr;
  if (threadIdx.x == 0) {
    const int n = n_cand;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += cand_ov[i];
    const double mean = s / n;
    double ss = 0.0;
    for (int i = 0; i < n; ++i) ss += (cand_ov[i] - mean) * (cand_ov[i] - mean);
    thr = mean + sqrt(ss / (n - 1));
  }
  __syncthreads();
  if (threadIdx.x < n_cand) {
    const int a = cand[threadIdx.x];
    if (cand_ov[threadIdx.x] > thr) {
      const float4 an = reinterpret_cast<const float4*>(p.anchors)[a];
      const float acx = __fdiv_rn(__fadd_rn(an.x, an.z), 2.f), acy = __fdiv_rn(__fadd_rn(an.y, an.w), 2.f);
      if (centre_in_gt(gb, acx, acy)) {
        atomicAdd(&p.cnt[(int64_t)b * p.A + a], 1);
        atomicExch(&p.sel[(int64_t)b * p.A + a], g);
      }
    }
  }
}

__global__ void __launch_bounds__(kAssignThreads) atss_resolve_kernel(const AtssParams p) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const int64_t o = (int64_t)b * p.A + a;
  const int c = p.cnt[o];
  if (c == 0) {
    p.gt_idx[o] = 0;
    p.fg[o] = 0;
    p.norm[o] = 0.0;
    return;
  }
  int gi = p.sel[o];
  if (c > 1) {  // highest anchor-box IoU over all G rows
    const float4 an = reinterpret_cast<const float4*>(p.anchors)[a];
    double best = -1.0;
    gi = 0;
    for (int g = 0; g < p.G; ++g) {
      const double* gr = p.gt + ((int64_t)b * p.G + g) * 5;
      const double ov = anchor_iou(GtBox{gr[1], gr[2], gr[3], gr[4]}, an);
      if (ov > best) { best = ov; gi = g; }
    }
  }
  p.gt_idx[o] = gi;
  p.fg[o] = 1;
  double soft = 1.0;
  if (p.pd_boxes != nullptr) {  // atss_assigner.py:81-84
    const double* gr = p.gt + ((int64_t)b * p.G + gi) * 5;
    soft = pair_iou(GtBox{gr[1], gr[2], gr[3], gr[4]}, reinterpret_cast<const float4*>(p.pd_boxes)[o], 1e-9);
  }
  p.norm[o] = soft;
}

}  // namespace yv6

using namespace yv6;

extern "C" int yv6_targets_pad(yv6_handle* h, const float* targets, int32_t n, int32_t B, int32_t G, float scale_w,
                               float scale_h, double* gt, int32_t* gt_count, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && gt && gt_count && (targets || n == 0), "targets_pad: null argument");
  YV6_REQUIRE(B > 0 && G >= 0 && n >= 0, "targets_pad: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(gt_count, 0, sizeof(int32_t) * B, s));
  const int64_t rows = (int64_t)B * G;
  if (rows > 0) targets_fill_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(gt, rows);
  if (n > 0) targets_pad_kernel<<<(n + 127) / 128, 128, 0, s>>>(targets, n, B, G, scale_w, scale_h, gt, gt_count);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int64_t yv6_assign_workspace_bytes(int32_t B, int32_t A, int32_t G) {
  return (int64_t)B * A * (4 + 4 + 8 + 8) + (int64_t)B * (G > 0 ? G : 1) * 16 + 1024;
}

extern "C" int yv6_tal_assign(yv6_handle* h, const float* pd_scores, const float* pd_bboxes, const float* anc_points,
                              const double* gt, const uint8_t* mask_gt, int32_t B, int32_t A, int32_t G, int32_t nc,
                              int32_t topk, double alpha, double beta, double eps, int32_t* gt_idx, uint8_t* fg,
                              double* norm, void* workspace, int64_t workspace_bytes, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && pd_scores && pd_bboxes && anc_points && gt_idx && fg && norm && workspace, "tal: null argument");
  YV6_REQUIRE(B > 0 && A > 0 && nc > 0 && G >= 0, "tal: bad sizes");
  YV6_REQUIRE(topk >= 1 && topk <= kTopkMax, "tal: topk=%d out of range (1..%d)", topk, kTopkMax);
  YV6_REQUIRE(workspace_bytes >= yv6_assign_workspace_bytes(B, A, G), "tal: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (G == 0) {  // tal_assigner.py:48-53: everything background
    YV6_CHECK_CUDA(cudaMemsetAsync(gt_idx, 0, sizeof(int32_t) * B * A, s));
    YV6_CHECK_CUDA(cudaMemsetAsync(fg, 0, (size_t)B * A, s));
    YV6_CHECK_CUDA(cudaMemsetAsync(norm, 0, sizeof(double) * B * A, s));
    return YV6_OK;
  }
  YV6_REQUIRE(gt && mask_gt, "tal: null gt");
  TalParams p;
  p.scores = pd_scores; p.boxes = pd_bboxes; p.points = anc_points; p.gt = gt; p.mask = mask_gt;
  p.B = B; p.A = A; p.G = G; p.nc = nc; p.topk = topk; p.alpha = alpha; p.beta = beta; p.eps = eps;
  char* w = reinterpret_cast<char*>(workspace);
  p.cnt = reinterpret_cast<int32_t*>(w); w += (int64_t)B * A * 4;
  p.sel = reinterpret_cast<int32_t*>(w); w += (int64_t)B * A * 4;
  p.al_a = reinterpret_cast<double*>(w); w += (int64_t)B * A * 8;
  p.ov_a = reinterpret_cast<double*>(w); w += (int64_t)B * A * 8;
  p.pos_al = reinterpret_cast<unsigned long long*>(w); w += (int64_t)B * G * 8;
  p.pos_ov = reinterpret_cast<unsigned long long*>(w);
  p.gt_idx = gt_idx; p.fg = fg; p.norm = norm;
  YV6_CHECK_CUDA(cudaMemsetAsync(p.cnt, 0, (size_t)B * A * 4, s));
  YV6_CHECK_CUDA(cudaMemsetAsync(p.pos_al, 0, (size_t)B * G * 16, s));
  tal_topk_kernel<<<dim3(G, B), kAssignThreads, 0, s>>>(p);
  const dim3 ga((A + kAssignThreads - 1) / kAssignThreads, B);
  tal_resolve_kernel<<<ga, kAssignThreads, 0, s>>>(p);
  tal_finalize_kernel<<<ga, kAssignThreads, 0, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_atss_assign(yv6_handle* h, const float* anc_bboxes, const int32_t* n_level_bboxes, int32_t nl,
                               const double* gt, const uint8_t* mask_gt, const float* pd_bboxes, int32_t B, int32_t A,
                               int32_t G, int32_t nc, int32_t topk, int32_t* gt_idx, uint8_t* fg, double* norm,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && anc_bboxes && n_level_bboxes && gt_idx && fg && norm && workspace, "atss: null argument");
  YV6_REQUIRE(nl >= 1 && nl <= kAtssMaxLevels, "atss: nl=%d out of range", nl);
  YV6_REQUIRE(topk >= 1 && topk <= kAtssTopkMax, "atss: topk=%d out of range (1..%d)", topk, kAtssTopkMax);
  YV6_REQUIRE(workspace_bytes >= yv6_assign_workspace_bytes(B, A, G), "atss: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (G == 0) {
    YV6_CHECK_CUDA(cudaMemsetAsync(gt_idx, 0, sizeof(int32_t) * B * A, s));
    YV6_CHECK_CUDA(cudaMemsetAsync(fg, 0, (size_t)B * A, s));
    YV6_CHECK_CUDA(cudaMemsetAsync(norm, 0, sizeof(double) * B * A, s));
    return YV6_OK;
  }
  YV6_REQUIRE(gt && mask_gt, "atss: null gt");
  AtssParams p;
  p.anchors = anc_bboxes; p.pd_boxes = pd_bboxes; p.gt = gt; p.mask = mask_gt;
  p.B = B; p.A = A; p.G = G; p.nc = nc; p.topk = topk; p.nl = nl;
  int off = 0;
  for (int l = 0; l < nl; ++l) { p.lvl_off[l] = off; off += n_level_bboxes[l]; }
  p.lvl_off[nl] = off;
  YV6_REQUIRE(off == A, "atss: level sizes sum to %d, expected A=%d", off, A);
  char* w = reinterpret_cast<char*>(workspace);
  p.cnt = reinterpret_cast<int32_t*>(w); w += (int64_t)B * A * 4;
  p.sel = reinterpret_cast<int32_t*>(w);
  p.gt_idx = gt_idx; p.fg = fg; p.norm = norm;
  YV6_CHECK_CUDA(cudaMemsetAsync(p.cnt, 0, (size_t)B * A * 4, s));
  atss_candidates_kernel<<<dim3(G, B), kAssignThreads, 0, s>>>(p);
  atss_resolve_kernel<<<dim3((A + kAssignThreads - 1) / kAssignThreads, B), kAssignThreads, 0, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_assign_expand(yv6_handle* h, const double* gt, const int32_t* gt_idx, const uint8_t* fg,
                                 const double* norm, int32_t B, int32_t A, int32_t G, int32_t nc, int32_t bg_label,
                                 int64_t* labels, double* bboxes, double* scores, uint8_t* fg_out, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && gt && gt_idx && fg && norm && labels && bboxes && scores && fg_out, "assign_expand: null argument");
  YV6_REQUIRE(G > 0, "assign_expand: needs at least one gt row");
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(scores, 0, sizeof(double) * (size_t)B * A * nc, s));
  ExpandParams p{gt, gt_idx, fg, norm, B, A, G, nc, bg_label, labels, bboxes, scores, fg_out};
  assign_expand_kernel<<<dim3((A + kAssignThreads - 1) / kAssignThreads, B), kAssignThreads, 0, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
