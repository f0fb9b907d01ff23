// yv6_nms.cu -- batched non-maximum suppression for all images of a batch in four launches.
//
// Reference: `non_max_suppression` (yolov6/utils/nms.py:31-105) which loops over images in Python,
// compacts with boolean masks and calls torchvision.ops.nms per image (nms.py:96).  Here:
//   1. nms_scan   : warp-per-anchor-row pass over pred[B,A,5+nc] (coalesced class reads): candidate
//                   rule of nms.py:48, obj*cls (nms.py:69), best class (nms.py:79-80) or per-class
//                   count (nms.py:75-77); per-anchor entry counts + per-tile totals.
//   2. nms_emit   : order-preserving compaction (anchor-major, class-minor = `nonzero` order): boxes
//                   xywh->xyxy (nms.py:21-28), 64-bit sort keys.
//   3. nms_sort   : one block per image, bitonic sort of (descending score, ascending position) keys
//                   -- a total order equal to torchvision's stable descending sort.
//   4. nms_greedy : one block per image, greedy suppression in chunks of 64 sorted candidates against
//                   the kept set; class offset boxes + cls*4096 (nms.py:94-95); float IoU promoted to
//                   double against the double threshold exactly like torchvision's CPU kernel; stops
//                   at max_det (nms.py:97-98); at most max_nms = 30000 candidates enter (nms.py:90-91).
// All fp32 arithmetic that feeds a comparison uses explicit round-to-nearest intrinsics (no FMA
// contraction) in the reference's operation order, so kept indices and class ids are bit-exact.
#include <algorithm>

#include "yv6_common.cuh"
#include "yv6_handle.h"

namespace yv6 {

constexpr int kNmsTile = 256;       // anchors per block in scan/emit
constexpr int kMaxNms = 30000;      // nms.py:55
constexpr int kMultiCap = 65536;    // candidate capacity per image in multi-label mode
constexpr int kSortSmemMax = 16384; // keys sorted in shared memory up to this many

struct NmsWs {
  int32_t* anchor_cnt;   // [B][A]
  float* anchor_score;   // [B][A] best score (single-label)
  int32_t* anchor_cls;   // [B][A] best class (single-label)
  int32_t* tile_cnt;     // [B][T]
  int32_t* cand_count;   // [B]
  int32_t* overflow;     // [1]
  uint64_t* keys;        // [B][cap2]
  float4* boxes;         // [B][cap]
  float* scores;         // [B][cap]
  int32_t* cls;          // [B][cap]
  int32_t* anchors;      // [B][cap]
  int32_t cap, cap2, T;
};

struct NmsParams {
  const float* pred;
  int32_t B, A, nc, no;
  float conf;
  double iou;
  int32_t agnostic, multi_label, max_det;
  const uint8_t* class_mask;
  float* out;
  int32_t* out_count;
  int32_t* out_src;
  NmsWs ws;
};

__global__ void __launch_bounds__(kNmsTile) nms_scan_kernel(const NmsParams p) {
  const int b = blockIdx.y, tile = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int warp_tot[kNmsTile / 32];
  int my_cnt = 0;
  float my_score = 0.f;
  int my_cls = 0;
  for (int i = 0; i < 32; ++i) {
    const int a = tile * kNmsTile + warp * 32 + i;
    if (a >= p.A) break;  // warp-uniform
    const float* row = p.pred + ((int64_t)b * p.A + a) * p.no;
    const float obj = __ldg(row + 4);
    float raw_max = -INFINITY, best = -INFINITY;
    int best_c = 0x7fffffff, cnt = 0;
    for (int c = lane; c < p.nc; c += 32) {
      const float v = __ldg(row + 5 + c);
      raw_max = fmaxf(raw_max, v);
      const float s = __fmul_rn(v, obj);                                   // nms.py:69
      const bool cls_ok = (p.class_mask == nullptr) || (p.class_mask[c] != 0);
      if (s > best) { best = s; best_c = c; }                              // first max within the lane
      if (s > p.conf && cls_ok) ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      raw_max = fmaxf(raw_max, __shfl_xor_sync(0xffffffffu, raw_max, o));
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }  // first max overall
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const bool cand = (obj > p.conf) && (raw_max > p.conf);                // nms.py:48
    int entries;
    if (p.multi_label) {
      entries = cand ? cnt : 0;                                            // nms.py:75-77
    } else {
      const bool cls_ok = (p.class_mask == nullptr) || (best_c < p.nc && p.class_mask[best_c] != 0);
      entries = (cand && best > p.conf && cls_ok) ? 1 : 0;                 // nms.py:79-84
    }
    if (lane == i) { my_cnt = entries; my_score = best; my_cls = best_c; }
  }
  const int a = tile * kNmsTile + threadIdx.x;
  if (a < p.A) {
    const int64_t o = (int64_t)b * p.A + a;
    p.ws.anchor_cnt[o] = my_cnt;
    p.ws.anchor_score[o] = my_score;
    p.ws.anchor_cls[o] = my_cls;
  }
  int tot = my_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) warp_tot[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kNmsTile / 32; ++w) s += warp_tot[w];
    p.ws.tile_cnt[b * p.ws.T + tile] = s;
  }
}

__device__ __forceinline__ float4 xywh2xyxy_rn(const float* r) {
  const float x = __ldg(r), y = __ldg(r + 1), w = __ldg(r + 2), h = __ldg(r + 3);
  const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);               // nms.py:24-27
  return make_float4(__fsub_rn(x, hw), __fsub_rn(y, hh), __fadd_rn(x, hw), __fadd_rn(y, hh));
}

__device__ __forceinline__ uint64_t make_key(float score, int pos) {
  return ((uint64_t)(0xffffffffu - __float_as_uint(score)) << 32) | (uint32_t)pos;  // score > 0
}

__global__ void __launch_bounds__(kNmsTile) nms_emit_kernel(const NmsParams p) {
  const int b = blockIdx.y, tile = blockIdx.x;
  __shared__ int red[kNmsTile];
  // offset of this tile = sum of the preceding tiles' counts
  int part = 0;
  for (int t = threadIdx.x; t < tile; t += kNmsTile) part += p.ws.tile_cnt[b * p.ws.T + t];
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = kNmsTile / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const int tile_off = red[0];
  __syncthreads();
  const int a = tile * kNmsTile + threadIdx.x;
  const int cnt = (a < p.A) ? p.ws.anchor_cnt[(int64_t)b * p.A + a] : 0;
  // exclusive scan of cnt over the block
  red[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 1; o < kNmsTile; o <<= 1) {
    const int v = (threadIdx.x >= o) ? red[threadIdx.x - o] : 0;
    __syncthreads();
    red[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = tile_off + red[threadIdx.x] - cnt;
  if (tile == gridDim.x - 1 && threadIdx.x == kNmsTile - 1) {
    const int total = tile_off + red[kNmsTile - 1];
    p.ws.cand_count[b] = min(total, p.ws.cap);
    if (total > p.ws.cap) atomicExch(p.ws.overflow, 1);
  }
  if (cnt == 0) return;
  const float* row = p.pred + ((int64_t)b * p.A + a) * p.no;
  const float4 box = xywh2xyxy_rn(row);
  const int64_t base = (int64_t)b * p.ws.cap;
  if (!p.multi_label) {
    if (pos < p.ws.cap) {
      const float s = p.ws.anchor_score[(int64_t)b * p.A + a];
      p.ws.boxes[base + pos] = box;
      p.ws.scores[base + pos] = s;
      p.ws.cls[base + pos] = p.ws.anchor_cls[(int64_t)b * p.A + a];
      p.ws.anchors[base + pos] = a;
      p.ws.keys[(int64_t)b * p.ws.cap2 + pos] = make_key(s, pos);
    }
  } else {
    const float obj = __ldg(row + 4);
    for (int c = 0; c < p.nc; ++c) {
      const float s = __fmul_rn(__ldg(row + 5 + c), obj);
      const bool cls_ok = (p.class_mask == nullptr) || (p.class_mask[c] != 0);
      if (s > p.conf && cls_ok) {
        if (pos < p.ws.cap) {
          p.ws.boxes[base + pos] = box;
          p.ws.scores[base + pos] = s;
          p.ws.cls[base + pos] = c;
          p.ws.anchors[base + pos] = a;
          p.ws.keys[(int64_t)b * p.ws.cap2 + pos] = make_key(s, pos);
        }
        ++pos;
      }
    }
  }
}

__global__ void __launch_bounds__(1024) nms_sort_kernel(const NmsParams p) {
  extern __shared__ uint64_t skeys[];
  const int b = blockIdx.x;
  const int n = p.ws.cand_count[b];
  if (n <= 1) return;
  int P = 2;
  while (P < n) P <<= 1;
  uint64_t* g = p.ws.keys + (int64_t)b * p.ws.cap2;
  const bool in_smem = (P <= kSortSmemMax);
  uint64_t* k = in_smem ? skeys : g;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const uint64_t v = (i < n) ? g[i] : ~0ull;  // pad sorts last
    k[i] = v;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int j = size >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const uint64_t x = k[i], y = k[l];
        const bool up = ((i & size) == 0);
        if ((x > y) == up) { k[i] = y; k[l] = x; }
      }
      __syncthreads();
    }
  }
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = k[i];
}

constexpr int kGreedyThreads = 512;

__global__ void __launch_bounds__(kGreedyThreads) nms_greedy_kernel(const NmsParams p) {
  extern __shared__ float4 gsm[];
  float4* kept_box = gsm;                                              // [max_det] offset boxes
  float* kept_area = reinterpret_cast<float*>(kept_box + p.max_det);   // [max_det]
  __shared__ float4 ch_box[64];
  __shared__ float ch_area[64];
  __shared__ int ch_idx[64];
  __shared__ int ch_alive[64];
  __shared__ unsigned int ch_mask[64][2];
  __shared__ int s_kept;
  const int b = blockIdx.x;
  const int n = min(p.ws.cand_count[b], kMaxNms);
  const uint64_t* keys = p.ws.keys + (int64_t)b * p.ws.cap2;
  const int64_t base = (int64_t)b * p.ws.cap;
  const double thr = p.iou;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 64) {
    const int m = min(64, n - c0);
    const int kept = s_kept;
    if (kept >= p.max_det) break;
    if (threadIdx.x < 64) {
      const int c = threadIdx.x;
      ch_alive[c] = (c < m);
      ch_mask[c][0] = ch_mask[c][1] = 0u;
      if (c < m) {
        const int idx = (int)(keys[c0 + c] & 0xffffffffu);
        float4 bx = p.ws.boxes[base + idx];
        if (!p.agnostic) {
          const float off = __fmul_rn((float)p.ws.cls[base + idx], 4096.f);  // nms.py:94
          bx = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
        }
        ch_box[c] = bx;
        ch_area[c] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        ch_idx[c] = idx;
      }
    }
    __syncthreads();
    // (a) candidates of this chunk vs. boxes kept from earlier chunks
    {
      const int c = threadIdx.x & 63;
      if (c < m) {
        const float4 cb = ch_box[c];
        const float ca = ch_area[c];
        for (int k = threadIdx.x >> 6; k < kept; k += kGreedyThreads / 64) {
          const float4 kb = kept_box[k];
          const float w = fmaxf(0.f, __fsub_rn(fminf(kb.z, cb.z), fmaxf(kb.x, cb.x)));
          const float h = fmaxf(0.f, __fsub_rn(fminf(kb.w, cb.w), fmaxf(kb.y, cb.y)));
          const float inter = __fmul_rn(w, h);
          const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(kept_area[k], ca), inter));
          if ((double)ovr > thr) { ch_alive[c] = 0; break; }
        }
      }
    }
    // (b) pairwise suppression inside the chunk: bit j of mask[i] = (i < j and IoU(i,j) > thr)
    for (int pr = threadIdx.x; pr < 64 * 64; pr += kGreedyThreads) {
      const int i = pr >> 6, j = pr & 63;
      if (i < j && j < m) {
        const float4 a = ch_box[i], c = ch_box[j];
        const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
        const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
        const float inter = __fmul_rn(w, h);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ch_area[i], ch_area[j]), inter));
        if ((double)ovr > thr) atomicOr(&ch_mask[i][j >> 5], 1u << (j & 31));
      }
    }
    __syncthreads();
    // (c) sequential resolve of the chunk, append survivors
    if (threadIdx.x == 0) {
      unsigned int rem0 = 0u, rem1 = 0u;
      int k = kept;
      for (int c = 0; c < m && k < p.max_det; ++c) {
        const bool removed = (c < 32) ? ((rem0 >> c) & 1u) : ((rem1 >> (c - 32)) & 1u);
        if (ch_alive[c] && !removed) {
          rem0 |= ch_mask[c][0];
          rem1 |= ch_mask[c][1];
          kept_box[k] = ch_box[c];
          kept_area[k] = ch_area[c];
          const int idx = ch_idx[c];
          const float4 bx = p.ws.boxes[base + idx];
          float* o = p.out + ((int64_t)b * p.max_det + k) * 6;
          o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
          o[4] = p.ws.scores[base + idx];
          o[5] = (float)p.ws.cls[base + idx];
          p.out_src[((int64_t)b * p.max_det + k) * 2] = p.ws.anchors[base + idx];
          p.out_src[((int64_t)b * p.max_det + k) * 2 + 1] = p.ws.cls[base + idx];
          ++k;
        }
      }
      s_kept = k;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) p.out_count[b] = s_kept;
}

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

static void nms_layout(int32_t B, int32_t A, int32_t nc, int32_t multi_label, NmsWs* ws, int64_t* total, char* base) {
  int64_t cap = multi_label ? (int64_t)A * nc : A;
  if (cap > kMultiCap && multi_label) cap = kMultiCap;
  int64_t cap2 = 2;
  while (cap2 < cap) cap2 <<= 1;
  const int T = (A + kNmsTile - 1) / kNmsTile;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += align256(bytes); return base ? base + o : (char*)nullptr; };
  ws->anchor_cnt = (int32_t*)take((int64_t)B * A * 4);
  ws->anchor_score = (float*)take((int64_t)B * A * 4);
  ws->anchor_cls = (int32_t*)take((int64_t)B * A * 4);
  ws->tile_cnt = (int32_t*)take((int64_t)B * T * 4);
  ws->cand_count = (int32_t*)take((int64_t)B * 4);
  ws->overflow = (int32_t*)take(4);
  ws->keys = (uint64_t*)take((int64_t)B * cap2 * 8);
  ws->boxes = (float4*)take((int64_t)B * cap * 16);
  ws->scores = (float*)take((int64_t)B * cap * 4);
  ws->cls = (int32_t*)take((int64_t)B * cap * 4);
  ws->anchors = (int32_t*)take((int64_t)B * cap * 4);
  ws->cap = (int32_t)cap;
  ws->cap2 = (int32_t)cap2;
  ws->T = T;
  *total = off;
}

}  // namespace yv6

using namespace yv6;

extern "C" int64_t yv6_nms_workspace_bytes(int32_t B, int32_t A, int32_t nc, int32_t multi_label) {
  NmsWs ws;
  int64_t total = 0;
  nms_layout(B, A, nc, multi_label, &ws, &total, nullptr);
  return total;
}

extern "C" int yv6_nms_batched(yv6_handle* h, const float* pred, int32_t B, int32_t A, int32_t nc, float conf_thres,
                               double iou_thres, int32_t agnostic, int32_t multi_label, const uint8_t* class_mask,
                               int32_t max_det, float* out, int32_t* out_count, int32_t* out_src, int32_t* overflow,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  YV6_REQUIRE(h && pred && out && out_count && out_src && workspace, "nms: null argument");
  YV6_REQUIRE(B > 0 && A > 0 && nc > 0, "nms: bad shape B=%d A=%d nc=%d", B, A, nc);
  YV6_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, "nms: conf_thres must be in [0,1]");       // nms.py:50
  YV6_REQUIRE(iou_thres >= 0.0 && iou_thres <= 1.0, "nms: iou_thres must be in [0,1]");          // nms.py:51
  YV6_REQUIRE(max_det > 0 && max_det <= 4096, "nms: max_det=%d out of range (1..4096)", max_det);
  NmsParams p;
  int64_t need = 0;
  const int ml = (multi_label && nc > 1) ? 1 : 0;                                                // nms.py:57
  nms_layout(B, A, nc, ml, &p.ws, &need, reinterpret_cast<char*>(workspace));
  YV6_REQUIRE(workspace_bytes >= need, "nms: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  p.pred = pred;
  p.B = B;
  p.A = A;
  p.nc = nc;
  p.no = nc + 5;
  p.conf = conf_thres;
  p.iou = iou_thres;
  p.agnostic = agnostic;
  p.multi_label = ml;
  p.max_det = max_det;
  p.class_mask = class_mask;
  p.out = out;
  p.out_count = out_count;
  p.out_src = out_src;
  if (overflow != nullptr) p.ws.overflow = overflow;
  cudaStream_t s = (cudaStream_t)stream;
  YV6_CHECK_CUDA(cudaMemsetAsync(p.ws.overflow, 0, 4, s));
  dim3 grid(p.ws.T, B);
  nms_scan_kernel<<<grid, kNmsTile, 0, s>>>(p);
  nms_emit_kernel<<<grid, kNmsTile, 0, s>>>(p);
  static bool configured = false;
  if (!configured) {
    YV6_CHECK_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemMax * 8));
    YV6_CHECK_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 20));
    configured = true;
  }
  const size_t sort_smem = (size_t)std::min<int64_t>(p.ws.cap2, kSortSmemMax) * 8;
  nms_sort_kernel<<<B, 1024, sort_smem, s>>>(p);
  nms_greedy_kernel<<<B, kGreedyThreads, (size_t)max_det * 20, s>>>(p);
  YV6_CHECK_CUDA(cudaGetLastError());
  return YV6_OK;
}

