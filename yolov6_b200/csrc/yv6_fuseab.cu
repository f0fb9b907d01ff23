// yv6_fuseab.cu -- the anchor-aided ("fuse_ab") training branch of the decoupled head.
//
// Reference: yolov6/models/heads/effidehead_fuseab.py:94-140 (Detect.forward, training branch) and
// yolov6/models/losses/loss_fuseab.py:58-76 (box construction in ComputeLoss.__call__).  Per level l the two extra
// 1x1 pred convs emit, per pixel, na = 3 anchors x nc class logits and na x 4 box values; the reference reshapes them to
// (b, na, h, w, .) and concatenates the levels, i.e. row (level l, anchor a, pixel p) of the [B, na*A, .] tensors sits at
// na * off_l + a * hw_l + p.  The pred convs here write the natural NHWC order [B, hw, na * ch] (fp32; sigmoid fused for
// the class branch); the kernels below move between the two orders and apply the box transform
//     wh = (2 * sigmoid(r_wh))^2 * anchors_init[l][a] / stride_l          (effidehead_fuseab.py:117-119)
// and, for the loss, turn (x_off, y_off, w, h) around the cell centre into the equivalent (l, t, r, b) distances so that
// the anchor-free loss kernel (yv6_det_loss) can be reused unchanged:
//     x1 = (ax + xo) - w/2, x2 = x1 + w  (loss_fuseab.py:73-76, xywh2xyxy general.py:54-61)  ->  l = ax - x1, r = x2 - ax.
#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kAbThreads = 256;

__device__ __forceinline__ float sigmoidf_rn(float x) { return 1.f / (1.f + expf(-x)); }

// one thread per (b, pixel, a, j): j in [0, nc) copies a class score, j in [nc, nc + 4) a box value
__global__ void __launch_bounds__(kAbThreads) ab_pack_kernel(const float* __restrict__ raw_cls, const float* __restrict__ raw_reg, int B, int hw,
                                                             int na, int nc, float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
                                                             int off3, int A3, float* __restrict__ cls_ab, float* __restrict__ reg_ab) {
  const int per = nc + 4;
  const int64_t total = (int64_t)B * hw * na * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % per);
    int64_t r = i / per;
    const int p = (int)(r % hw); r /= hw;          // destination-major order: consecutive threads write consecutive floats
    const int a = (int)(r % na);
    const int b = (int)(r / na);
    const int64_t row = (int64_t)b * A3 + off3 + (int64_t)a * hw + p;
    if (j < nc) {
      cls_ab[row * nc + j] = raw_cls[((int64_t)b * hw + p) * (na * nc) + a * nc + j];
    } else {
      const int k = j - nc;
      const float v = raw_reg[((int64_t)b * hw + p) * (na * 4) + a * 4 + k];
      float o = v;
      if (k >= 2) {
        const float anc = (a == 0) ? (k == 2 ? aw0 : ah0) : (a == 1) ? (k == 2 ? aw1 : ah1) : (k == 2 ? aw2 : ah2);
        const float t = __fmul_rn(sigmoidf_rn(v), 2.f);
        o = __fmul_rn(__fmul_rn(t, t), anc);
      }
      reg_ab[row * 4 + k] = o;
    }
  }
}

// backward of the above + sigmoid backward of the class branch; outputs are the dense NHWC bf16 gradients w.r.t. the raw
// conv outputs ([B, hw, ch_pad], zero padded channels) that the pred convs' wgrad / dgrad consume
__global__ void __launch_bounds__(kAbThreads) ab_grad_kernel(const float* __restrict__ g_cls, const float* __restrict__ cls_ab,
                                                             const float* __restrict__ g_reg, const float* __restrict__ raw_reg, int B, int hw,
                                                             int na, int nc, float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
                                                             int off3, int A3, int cpad, int rpad, __nv_bfloat16* __restrict__ dl_cls,
                                                             __nv_bfloat16* __restrict__ dl_reg) {
  const int per = cpad + rpad;
  const int64_t total = (int64_t)B * hw * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % per);
    const int64_t px = i / per;
    const int b = (int)(px / hw), p = (int)(px % hw);
    if (j < cpad) {
      float v = 0.f;
      if (j < na * nc) {
        const int a = j / nc, c = j - a * nc;
        const int64_t src = ((int64_t)b * A3 + off3 + (int64_t)a * hw + p) * nc + c;
        const float s = cls_ab[src];
        v = g_cls[src] * s * (1.f - s);
      }
      dl_cls[px * cpad + j] = __float2bfloat16_rn(v);
    } else {
      const int q = j - cpad;
      float v = 0.f;
      if (q < na * 4) {
        const int a = q >> 2, k = q & 3;
        const int64_t src = ((int64_t)b * A3 + off3 + (int64_t)a * hw + p) * 4 + k;
        v = g_reg[src];
        if (k >= 2) {
          const float anc = (a == 0) ? (k == 2 ? aw0 : ah0) : (a == 1) ? (k == 2 ? aw1 : ah1) : (k == 2 ? aw2 : ah2);
          const float sg = sigmoidf_rn(raw_reg[px * (na * 4) + q]);
          v *= anc * 8.f * sg * sg * (1.f - sg);           // d/dr [(2 s)^2] = 8 s^2 (1 - s)
        }
      }
      dl_reg[px * rpad + q] = __float2bfloat16_rn(v);
    }
  }
}

// (x_off, y_off, w, h) around the cell centre -> boxes in pixels for the assigner and equivalent ltrb distances for the loss
__global__ void __launch_bounds__(kAbThreads) ab_boxes_kernel(const float* __restrict__ reg_ab, const float* __restrict__ pts,
                                                              const float* __restrict__ strides, int B, int A3, float* __restrict__ ltrb,
                                                              float* __restrict__ boxes_px) {
  const int64_t total = (int64_t)B * A3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i % A3);
    const float s = strides[a];
    const float ax = __fdiv_rn(pts[2 * a], s), ay = __fdiv_rn(pts[2 * a + 1], s);        // anchor_points / stride_tensor, loss_fuseab.py:71
    const float4 r = reinterpret_cast<const float4*>(reg_ab)[i];
    const float cx = __fadd_rn(r.x, ax), cy = __fadd_rn(r.y, ay);                        // pred_distri[..., :2] += anchor_points_s
    const float x1 = __fsub_rn(cx, __fmul_rn(r.z, 0.5f)), y1 = __fsub_rn(cy, __fmul_rn(r.w, 0.5f));   // xywh2xyxy, general.py:54-61
    const float x2 = __fadd_rn(x1, r.z), y2 = __fadd_rn(y1, r.w);
    reinterpret_cast<float4*>(boxes_px)[i] = make_float4(__fmul_rn(x1, s), __fmul_rn(y1, s), __fmul_rn(x2, s), __fmul_rn(y2, s));
    reinterpret_cast<float4*>(ltrb)[i] = make_float4(__fsub_rn(ax, x1), __fsub_rn(ay, y1), __fsub_rn(x2, ax), __fsub_rn(y2, ay));
  }
}

// l = ax - x1 = -xo + w/2 ... : d/d(xo, yo, w, h) of a loss given its gradient w.r.t. (l, t, r, b)
__global__ void __launch_bounds__(kAbThreads) ab_boxes_bwd_kernel(const float* __restrict__ g_ltrb, int64_t rows, float* __restrict__ g_reg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(g_ltrb)[i];
    // x1 = cx - w/2, x2 = x1 + w = cx + w/2  ->  l = ax - cx + w/2, r = cx - ax + w/2
    reinterpret_cast<float4*>(g_reg)[i] = make_float4(g.z - g.x, g.w - g.y, 0.5f * (g.x + g.z), 0.5f * (g.y + g.w));
  }
}

static inline unsigned ab_grid(int64_t total, int sms) {
  const int64_t b = (total + kAbThreads - 1) / kAbThreads;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sms * 16));
}

}  // namespace yv6

using namespace yv6;

extern "C" int yv6_head_ab_pack(yv6_handle* h, const float* raw_cls, const float* raw_reg, int32_t B, int32_t hw, int32_t na, int32_t nc,
                                const float* anchors_wh, int32_t row_off, int32_t rows_total, float* cls_ab, float* reg_ab, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && raw_cls && raw_reg && anchors_wh && cls_ab && reg_ab, "head_ab_pack: null argument");
  YV6_REQUIRE(na == 3 && nc >= 1 && B >= 1 && hw >= 1, "head_ab_pack: na=%d (the reference head has 3 anchors per level), nc=%d", na, nc);
  const int64_t total = (int64_t)B * hw * na * (nc + 4);
  ab_pack_kernel<<<ab_grid(total, h->num_sms), kAbThreads, 0, (cudaStream_t)stream>>>(raw_cls, raw_reg, B, hw, na, nc, anchors_wh[0], anchors_wh[1],
                                                                                  anchors_wh[2], anchors_wh[3], anchors_wh[4], anchors_wh[5],
                                                                                  row_off, rows_total, cls_ab, reg_ab);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_head_ab_grad(yv6_handle* h, const float* grad_cls_ab, const float* cls_ab, const float* grad_reg_ab, const float* raw_reg,
                                int32_t B, int32_t hw, int32_t na, int32_t nc, const float* anchors_wh, int32_t row_off, int32_t rows_total,
                                int32_t cls_pad, int32_t reg_pad, void* dl_cls, void* dl_reg, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && grad_cls_ab && cls_ab && grad_reg_ab && raw_reg && anchors_wh && dl_cls && dl_reg, "head_ab_grad: null argument");
  YV6_REQUIRE(na == 3 && cls_pad >= na * nc && reg_pad >= na * 4, "head_ab_grad: bad padding (%d, %d)", cls_pad, reg_pad);
  const int64_t total = (int64_t)B * hw * (cls_pad + reg_pad);
  ab_grad_kernel<<<ab_grid(total, h->num_sms), kAbThreads, 0, (cudaStream_t)stream>>>(
      grad_cls_ab, cls_ab, grad_reg_ab, raw_reg, B, hw, na, nc, anchors_wh[0], anchors_wh[1], anchors_wh[2], anchors_wh[3], anchors_wh[4],
      anchors_wh[5], row_off, rows_total, cls_pad, reg_pad, reinterpret_cast<__nv_bfloat16*>(dl_cls), reinterpret_cast<__nv_bfloat16*>(dl_reg));
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_ab_boxes(yv6_handle* h, const float* reg_ab, const float* anc_points, const float* strides, int32_t B, int32_t A,
                            float* ltrb, float* boxes_px, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && reg_ab && anc_points && strides && ltrb && boxes_px, "ab_boxes: null argument");
  ab_boxes_kernel<<<ab_grid((int64_t)B * A, h->num_sms), kAbThreads, 0, (cudaStream_t)stream>>>(reg_ab, anc_points, strides, B, A, ltrb, boxes_px);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

extern "C" int yv6_ab_boxes_bwd(yv6_handle* h, const float* grad_ltrb, int64_t rows, float* grad_reg_ab, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && grad_ltrb && grad_reg_ab, "ab_boxes_bwd: null argument");
  ab_boxes_bwd_kernel<<<ab_grid(rows, h->num_sms), kAbThreads, 0, (cudaStream_t)stream>>>(grad_ltrb, rows, grad_reg_ab);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

// ------------------------------------------------------------------------------------------------
// Self-distillation terms of yolov6/models/losses/loss_distill.py (second half): both
// `distill_loss_cls` (:213-222, rows = all B*A anchors, C = classes, "logits" = the post-sigmoid scores) and
// `distill_loss_dfl` (:351-361, rows = the 4 sides of every positive anchor, C = reg_max + 1 = 17) are
//     T^2 * KL( softmax(teacher / T) || softmax(student / T) )   summed over rows (the DFL one averaged over its rows).
// One warp per row; value and gradient d/ds_j = (p_j - q_j) / T in the same pass.  The factor in front -- loss weights, the
// cosine decay, and for the DFL term 1 / (4 * num_pos) with num_pos on the device -- is `scale` x optional device scalars.
// ------------------------------------------------------------------------------------------------
namespace yv6 {
__global__ void __launch_bounds__(256) kl_rows_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t rows, int C, float inv_T,
                                                      const uint8_t* __restrict__ mask, int rows_per_mask, double scale,
                                                      const double* __restrict__ count_ptr, const double* __restrict__ gate_ptr,
                                                      double* __restrict__ acc, float* __restrict__ grad) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  double eff = scale;
  if (count_ptr != nullptr) {        // mean over the active rows: rows_per_mask * (number of active mask entries)
    const double n = *count_ptr * (double)rows_per_mask;
    eff = (n > 0.0) ? scale / n : 0.0;
  }
  if (gate_ptr != nullptr && !(*gate_ptr > 0.0)) eff = 0.0;    // loss_distill.py:318-323: the term is multiplied by weights that sum to 0
  double local = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * warps + warp; r < rows; r += (int64_t)gridDim.x * warps) {
    if (mask != nullptr && !mask[r / rows_per_mask]) continue;
    const float* sr = s + r * C;
    const float* tr = t + r * C;
    // The row's KL (~1e-4 at T = 20 over sigmoid scores) is a difference of two log-sum-exps of ~4.4: in fp32 their rounding
    // (~3e-7 each) is a 0.4 % error per row, so the log-sum-exps and log-probabilities are formed in double.
    const double iT = (double)inv_T;
    double ms = -INFINITY, mt = -INFINITY;
    for (int j = lane; j < C; j += 32) { ms = fmax(ms, (double)sr[j] * iT); mt = fmax(mt, (double)tr[j] * iT); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ms = fmax(ms, __shfl_xor_sync(0xffffffffu, ms, o)); mt = fmax(mt, __shfl_xor_sync(0xffffffffu, mt, o)); }
    double zs = 0.0, zt = 0.0;
    for (int j = lane; j < C; j += 32) { zs += exp((double)sr[j] * iT - ms); zt += exp((double)tr[j] * iT - mt); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { zs += __shfl_xor_sync(0xffffffffu, zs, o); zt += __shfl_xor_sync(0xffffffffu, zt, o); }
    const double ls = ms + log(zs), lt = mt + log(zt);           // log-sum-exp of student / teacher
    double kl = 0.0;
    for (int j = lane; j < C; j += 32) {
      const double a = (double)sr[j] * iT - ls, b = (double)tr[j] * iT - lt;  // log p_j, log q_j
      const double q = exp(b), pj = exp(a);
      kl += q * (b - a);
      if (grad != nullptr) grad[r * C + j] += (float)(eff * (pj - q) * iT);
    }
    local += kl;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane == 0) sh[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < warps; ++w) tot += sh[w];
    if (tot != 0.0) atomicAdd(acc, eff * tot);
  }
}
}  // namespace yv6

extern "C" int yv6_kl_rows(yv6_handle* h, const float* student, const float* teacher, int64_t rows, int32_t C, float temperature,
                           const uint8_t* row_mask, int32_t rows_per_mask, double scale, const double* count_ptr, const double* gate_ptr,
                           double* acc, float* grad, void* stream) {
  yv6_device_guard _dev(h);
  YV6_REQUIRE(h && student && teacher && acc, "kl_rows: null argument");
  YV6_REQUIRE(C >= 1 && temperature > 0.f && rows_per_mask >= 1, "kl_rows: C=%d T=%f", C, (double)temperature);
  if (rows <= 0) return YV6_OK;
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, (int64_t)h->num_sms * 8);
  kl_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(student, teacher, rows, C, 1.f / temperature, row_mask, rows_per_mask, scale,
                                                                    count_ptr, gate_ptr, acc, grad);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}
