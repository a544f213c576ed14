// Batched RPN proposal selection around the NMS: the two data-dependent stages of
// detectron2/modeling/proposal_generator/proposal_utils.py:22-135 as fixed-capacity kernels, one CTA per image.
//
//   d2b_rpn_prepare   gathers the per-level top-k candidates, clips them to the image (Boxes.clip, :112), marks non-finite
//                     (:104-110) and too-small (:115-119) boxes as IGNORED (category -1) instead of removing them, and
//                     applies torchvision's batched-NMS coordinate offsets per image -- level * (max coordinate of that
//                     image's surviving boxes + 1), fp32 -- so that every IoU rounds like the reference's;
//   d2b_rpn_select    walks the score-ordered keep list of ONE NMS over all images and hands every image its first
//                     post_nms_topk survivors (:129) in a fixed [N, post_nms_topk] layout + a count.
//
// Together with d2b_nms (category = image * L + level, per-category bound = pre_nms_topk) the whole selection is a
// sync-free launch sequence with static shapes: capturable in a CUDA graph; the reference loops over images in Python
// with boolean indexing and one `.item()` per image.  Compiled with -fmad=false like nms.cu (bit-exact clip / offsets).
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;

struct RpnLevels {
  int L;
  const float* proposals[D2B_MAX_LEVELS];    // [N, A_l, 4]
  const int64_t* topk_idx[D2B_MAX_LEVELS];   // [N, k_l]
  const float* topk_scores[D2B_MAX_LEVELS];  // [N, k_l]
  int A[D2B_MAX_LEVELS], k[D2B_MAX_LEVELS], t0[D2B_MAX_LEVELS + 1];  // t0: prefix of k
};

__device__ __forceinline__ bool finitef(float v) { return fabsf(v) <= 3.402823466e38f; }  // false for inf and NaN

__global__ void __launch_bounds__(kThreads) rpn_prepare_kernel(const RpnLevels P, int T, const float* __restrict__ image_hw,
                                                               float min_box_size, int use_offsets,
                                                               float* __restrict__ flat_boxes, float* __restrict__ nms_boxes,
                                                               float* __restrict__ nms_scores, float* __restrict__ raw_scores,
                                                               long long* __restrict__ cat_ids, int* __restrict__ nonfinite) {
  __shared__ float s_red[32];
  __shared__ float s_max;
  const int n = blockIdx.x, tid = threadIdx.x;
  const float ih = image_hw[2 * n], iw = image_hw[2 * n + 1];
  float mx = -INFINITY;
  int bad = 0;
  for (int t = tid; t < T; t += kThreads) {
    int l = 0;
    while (l + 1 < P.L && t >= P.t0[l + 1]) ++l;
    const int j = t - P.t0[l];
    const long long a = P.topk_idx[l][(size_t)n * P.k[l] + j];
    const float s = P.topk_scores[l][(size_t)n * P.k[l] + j];
    const float4 b = *reinterpret_cast<const float4*>(P.proposals[l] + ((size_t)n * P.A[l] + a) * 4);
    const bool fin = finitef(b.x) && finitef(b.y) && finitef(b.z) && finitef(b.w) && finitef(s);
    // Boxes.clip: x to [0, w], y to [0, h]   (torch.clamp(min=0) then minimum with the size, like the host restatement)
    const float x1 = fminf(fmaxf(b.x, 0.f), iw), y1 = fminf(fmaxf(b.y, 0.f), ih);
    const float x2 = fminf(fmaxf(b.z, 0.f), iw), y2 = fminf(fmaxf(b.w, 0.f), ih);
    const bool valid = fin && (x2 - x1) > min_box_size && (y2 - y1) > min_box_size;
    const size_t o = (size_t)n * T + t;
    *reinterpret_cast<float4*>(flat_boxes + o * 4) = valid ? make_float4(x1, y1, x2, y2) : make_float4(0.f, 0.f, 0.f, 0.f);
    raw_scores[o] = s;
    nms_scores[o] = valid ? s : -INFINITY;
    cat_ids[o] = valid ? (long long)n * P.L + l : -1LL;
    if (valid) mx = fmaxf(mx, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
    bad |= fin ? 0 : 1;
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  if (bad) atomicOr(nonfinite, 1);
  __syncthreads();
  if (tid < 32) {
    mx = s_red[tid];
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (tid == 0) s_max = mx;
  }
  __syncthreads();
  const float scale = s_max + 1.0f;  // torchvision _batched_nms_coordinate_trick: idxs * (boxes.max() + 1)
  for (int t = tid; t < T; t += kThreads) {
    int l = 0;
    while (l + 1 < P.L && t >= P.t0[l + 1]) ++l;
    const size_t o = (size_t)n * T + t;
    float4 b = *reinterpret_cast<const float4*>(flat_boxes + o * 4);
    if (use_offsets && cat_ids[o] >= 0) {
      const float off = (float)l * scale;
      b.x += off;
      b.y += off;
      b.z += off;
      b.w += off;
    }
    *reinterpret_cast<float4*>(nms_boxes + o * 4) = b;
  }
}

// Exclusive prefix sum of one int per thread over a 1024-thread CTA; `total` = sum.
__device__ __forceinline__ int block_scan(int v, int* __restrict__ warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // warp_tot reuse between calls
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  const int wt = warp_tot[lane];
  int winc = wt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  total = __shfl_sync(0xffffffffu, winc, 31);
  const int wbase = __shfl_sync(0xffffffffu, winc, warp) - __shfl_sync(0xffffffffu, wt, warp);
  return wbase + inc - v;
}

__global__ void __launch_bounds__(kThreads) rpn_select_kernel(const long long* __restrict__ keep,
                                                              const long long* __restrict__ num_keep, int T, int post_topk,
                                                              const float* __restrict__ flat_boxes,
                                                              const float* __restrict__ raw_scores,
                                                              const long long* __restrict__ cat_ids,
                                                              float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                              long long* __restrict__ out_index, long long* __restrict__ counts) {
  __shared__ int warp_tot[32];
  const int n = blockIdx.x, tid = threadIdx.x;
  const long long nk = max(0LL, *num_keep);
  int have = 0;
  for (long long j0 = 0; j0 < nk && have < post_topk; j0 += kThreads) {
    const long long j = j0 + tid;
    long long kidx = -1;
    int mine = 0;
    if (j < nk) {
      kidx = keep[j];
      mine = (kidx / T == n && cat_ids[kidx] >= 0) ? 1 : 0;
    }
    int total;
    const int rank = have + block_scan(mine, warp_tot, total);
    if (mine && rank < post_topk) {
      const size_t o = (size_t)n * post_topk + rank;
      *reinterpret_cast<float4*>(out_boxes + o * 4) = *reinterpret_cast<const float4*>(flat_boxes + (size_t)kidx * 4);
      out_scores[o] = raw_scores[kidx];
      out_index[o] = kidx;
    }
    have += total;
  }
  const int cnt = min(have, post_topk);
  for (int r = cnt + tid; r < post_topk; r += kThreads) {  // deterministic padding
    const size_t o = (size_t)n * post_topk + r;
    *reinterpret_cast<float4*>(out_boxes + o * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    out_scores[o] = 0.f;
    out_index[o] = 0;
  }
  if (tid == 0) counts[n] = cnt;
}

}  // namespace

D2B_API int d2b_rpn_prepare(const d2b_rpn_levels* lv, int N, const float* image_hw, float min_box_size, int use_offsets,
                            float* flat_boxes, float* nms_boxes, float* nms_scores, float* raw_scores, int64_t* cat_ids,
                            int* nonfinite, void* stream) {
  if (!lv || lv->num_levels < 1 || lv->num_levels > D2B_MAX_LEVELS || N < 0) return D2B_EINVAL;
  if (!nonfinite) return D2B_EINVAL;
  D2B_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int), (cudaStream_t)stream));
  if (N == 0) return D2B_OK;
  if (!image_hw || !flat_boxes || !nms_boxes || !nms_scores || !raw_scores || !cat_ids) return D2B_EINVAL;
  RpnLevels P = {};
  P.L = lv->num_levels;
  int T = 0;
  for (int l = 0; l < P.L; ++l) {
    if (!lv->proposals[l] || !lv->topk_idx[l] || !lv->topk_scores[l] || lv->A[l] <= 0 || lv->k[l] < 0 || lv->k[l] > lv->A[l])
      return D2B_EINVAL;
    if ((reinterpret_cast<uintptr_t>(lv->proposals[l]) & 15) != 0) return D2B_EINVAL;
    P.proposals[l] = lv->proposals[l];
    P.topk_idx[l] = lv->topk_idx[l];
    P.topk_scores[l] = lv->topk_scores[l];
    P.A[l] = lv->A[l];
    P.k[l] = lv->k[l];
    P.t0[l] = T;
    T += lv->k[l];
  }
  P.t0[P.L] = T;
  if (T == 0) return D2B_OK;
  rpn_prepare_kernel<<<N, kThreads, 0, (cudaStream_t)stream>>>(P, T, image_hw, min_box_size, use_offsets, flat_boxes, nms_boxes,
                                                               nms_scores, raw_scores, (long long*)cat_ids, nonfinite);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

D2B_API int d2b_rpn_select(const int64_t* keep, const int64_t* num_keep, int N, int T, int post_nms_topk,
                           const float* flat_boxes, const float* raw_scores, const int64_t* cat_ids, float* out_boxes,
                           float* out_scores, int64_t* out_index, int64_t* counts, void* stream) {
  if (N < 0 || T < 0 || post_nms_topk < 0) return D2B_EINVAL;
  if (N == 0) return D2B_OK;
  if (!counts) return D2B_EINVAL;
  if (T == 0 || post_nms_topk == 0) {
    D2B_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * N, (cudaStream_t)stream));
    return D2B_OK;
  }
  if (!keep || !num_keep || !flat_boxes || !raw_scores || !cat_ids || !out_boxes || !out_scores || !out_index) return D2B_EINVAL;
  rpn_select_kernel<<<N, kThreads, 0, (cudaStream_t)stream>>>((const long long*)keep, (const long long*)num_keep, T, post_nms_topk,
                                                              flat_boxes, raw_scores, (const long long*)cat_ids, out_boxes,
                                                              out_scores, (long long*)out_index, (long long*)counts);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

// ================================================================================================ Fast R-CNN / dense-head candidates
// The candidate stages of the two inference post-processors that sit on the same NMS, as fixed-capacity
// kernels with one CTA per image; d2b_rpn_select above then hands every image its first `topk` survivors.
//
//   d2b_frcnn_prepare   fast_rcnn_inference_single_image (detectron2/modeling/roi_heads/fast_rcnn.py:117-173) up to the NMS:
//                       rows with a non-finite box or score are dropped (:137-140), the (row, class) pairs with
//                       score > score_thresh (:150-154) are written IN ROW-MAJOR ORDER (the order `nonzero()` gives the
//                       reference; it decides ties inside NMS) by an ordered block compaction -- no nonzero(), no sync --
//                       with their box clipped to the image (:146-147) and torchvision's batched-NMS coordinate offsets.
//   d2b_dense_prepare   DenseDetector._decode_per_level_predictions (meta_arch/dense_detector.py:186-235) after the per-level
//                       top-k: Box2BoxTransform.apply_deltas (box_regression.py:78-116, same fp32 expression order) on the
//                       selected (anchor, class) pairs only, class ids, coordinate offsets.
namespace {

struct FrcnnImages {
  int N;
  int row_start[D2B_MAX_IMAGES + 1];
};

__device__ __forceinline__ float block_max(float v, float* s_red, float* s_out) {
  const int tid = threadIdx.x;
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // s_red reuse
  if ((tid & 31) == 0) s_red[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    v = s_red[tid];
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0) *s_out = v;
  }
  __syncthreads();
  return *s_out;
}

__global__ void __launch_bounds__(kThreads) frcnn_prepare_kernel(const FrcnnImages I, const float* __restrict__ boxes,
                                                                 const float* __restrict__ scores, int K, int kreg,
                                                                 const float* __restrict__ image_hw, float score_thresh, int cap,
                                                                 float* __restrict__ cand_boxes, float* __restrict__ nms_boxes,
                                                                 float* __restrict__ nms_scores, float* __restrict__ raw_scores,
                                                                 long long* __restrict__ cand_flat, long long* __restrict__ cat_ids,
                                                                 long long* __restrict__ n_cand, long long* __restrict__ row_map) {
  __shared__ int warp_tot[32];
  __shared__ float s_red[32];
  __shared__ float s_max;
  const int n = blockIdx.x, tid = threadIdx.x;
  const int rs = I.row_start[n], R = I.row_start[n + 1] - rs;
  const float ih = image_hw[2 * n], iw = image_hw[2 * n + 1];
  const size_t obase = (size_t)n * cap;
  int have = 0, have_rows = 0;
  float mx = -INFINITY;
  for (int r0 = 0; r0 < R; r0 += kThreads) {
    const int r = r0 + tid;
    int cnt = 0, valid = 0;
    const float* __restrict__ srow = scores + (size_t)(rs + (r < R ? r : 0)) * (K + 1);
    const float* __restrict__ brow = boxes + (size_t)(rs + (r < R ? r : 0)) * kreg * 4;
    if (r < R) {
      valid = 1;
      for (int c = 0; c <= K; ++c) valid &= finitef(srow[c]) ? 1 : 0;
      for (int c = 0; c < kreg * 4; ++c) valid &= finitef(brow[c]) ? 1 : 0;
      if (valid)
        for (int c = 0; c < K; ++c) cnt += srow[c] > score_thresh ? 1 : 0;
    }
    int total, total_rows;
    const int off = have + block_scan(cnt, warp_tot, total);
    const int vrank = have_rows + block_scan(valid, warp_tot, total_rows);
    if (r < R) row_map[rs + r] = valid ? vrank : -1;  // index of the row among the valid rows (:138-140)
    if (cnt) {
      int pos = off;
      for (int c = 0; c < K && pos < cap; ++c) {
        const float sc = srow[c];
        if (!(sc > score_thresh)) continue;
        const float* __restrict__ b = brow + (kreg == 1 ? 0 : c * 4);
        // Boxes.clip: clamp(min=0, max=w / h)
        const float x1 = fminf(fmaxf(b[0], 0.f), iw), y1 = fminf(fmaxf(b[1], 0.f), ih);
        const float x2 = fminf(fmaxf(b[2], 0.f), iw), y2 = fminf(fmaxf(b[3], 0.f), ih);
        *reinterpret_cast<float4*>(cand_boxes + (obase + pos) * 4) = make_float4(x1, y1, x2, y2);
        raw_scores[obase + pos] = sc;
        nms_scores[obase + pos] = sc;
        cand_flat[obase + pos] = (long long)r * K + c;
        cat_ids[obase + pos] = (long long)n * (K + 1) + c;
        mx = fmaxf(mx, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
        ++pos;
      }
    }
    have += total;
    have_rows += total_rows;
  }
  if (tid == 0) n_cand[n] = have;  // > cap: the list was truncated, the caller redoes the image
  const int live = min(have, cap);
  mx = block_max(mx, s_red, &s_max);
  const float scale = (live > 0 ? mx : 0.f) + 1.0f;  // torchvision _batched_nms_coordinate_trick: idxs * (boxes.max() + 1)
  __threadfence_block();
  for (int t = tid; t < cap; t += kThreads) {
    const size_t o = obase + t;
    if (t < live) {
      float4 b = *reinterpret_cast<const float4*>(cand_boxes + o * 4);
      const float offv = (float)(cat_ids[o] - (long long)n * (K + 1)) * scale;
      b.x += offv;
      b.y += offv;
      b.z += offv;
      b.w += offv;
      *reinterpret_cast<float4*>(nms_boxes + o * 4) = b;
    } else {  // dead slot: ignored by the NMS kernels
      *reinterpret_cast<float4*>(cand_boxes + o * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(nms_boxes + o * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      raw_scores[o] = 0.f;
      nms_scores[o] = -INFINITY;
      cand_flat[o] = 0;
      cat_ids[o] = -1;
    }
  }
}

struct DenseLevels {
  int L;
  const float* anchors[D2B_MAX_LEVELS];      // [R_l, 4]
  const float* deltas[D2B_MAX_LEVELS];       // [N, R_l, 4]
  const int64_t* topk_idx[D2B_MAX_LEVELS];   // [N, k_l]  flat (anchor * K + class)
  const float* topk_scores[D2B_MAX_LEVELS];  // [N, k_l]  -inf = dead slot
  int R[D2B_MAX_LEVELS], k[D2B_MAX_LEVELS], t0[D2B_MAX_LEVELS + 1];
};

__global__ void __launch_bounds__(kThreads) dense_prepare_kernel(const DenseLevels P, int T, int K, float wx, float wy, float ww,
                                                                 float wh, float scale_clamp, float* __restrict__ flat_boxes,
                                                                 float* __restrict__ nms_boxes, float* __restrict__ nms_scores,
                                                                 float* __restrict__ raw_scores, long long* __restrict__ classes,
                                                                 long long* __restrict__ cat_ids) {
  __shared__ float s_red[32];
  __shared__ int s_redi[32];
  __shared__ float s_max;
  __shared__ int s_live;
  const int n = blockIdx.x, tid = threadIdx.x;
  float mx = -INFINITY;
  int nlive = 0;
  for (int t = tid; t < T; t += kThreads) {
    int l = 0;
    while (l + 1 < P.L && t >= P.t0[l + 1]) ++l;
    const int j = t - P.t0[l];
    const long long f = P.topk_idx[l][(size_t)n * P.k[l] + j];
    const float s = P.topk_scores[l][(size_t)n * P.k[l] + j];
    const bool live = s > -INFINITY;
    const long long a = f / K;
    const long long cls = f - a * K;
    const float4 an = *reinterpret_cast<const float4*>(P.anchors[l] + (size_t)a * 4);
    const float4 d = *reinterpret_cast<const float4*>(P.deltas[l] + ((size_t)n * P.R[l] + a) * 4);
    // Box2BoxTransform.apply_deltas, op for op (this file is compiled with -fmad=false)
    const float widths = an.z - an.x, heights = an.w - an.y;
    const float ctr_x = an.x + 0.5f * widths, ctr_y = an.y + 0.5f * heights;
    const float dx = d.x / wx, dy = d.y / wy;
    float dw = d.z / ww, dh = d.w / wh;
    dw = dw > scale_clamp ? scale_clamp : dw;  // torch.clamp(max=): NaN stays NaN
    dh = dh > scale_clamp ? scale_clamp : dh;
    const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
    const float pw = expf(dw) * widths, ph = expf(dh) * heights;
    const float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph, x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
    const size_t o = (size_t)n * T + t;
    *reinterpret_cast<float4*>(flat_boxes + o * 4) = make_float4(x1, y1, x2, y2);
    raw_scores[o] = s;
    nms_scores[o] = live ? s : -INFINITY;
    classes[o] = cls;
    cat_ids[o] = live ? (long long)n * (K + 1) + cls : -1LL;
    if (live) {
      mx = fmaxf(mx, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
      ++nlive;
    }
  }
  for (int o = 16; o; o >>= 1) nlive += __shfl_xor_sync(0xffffffffu, nlive, o);
  if ((tid & 31) == 0) s_redi[tid >> 5] = nlive;
  mx = block_max(mx, s_red, &s_max);
  if (tid < 32) {
    int v = s_redi[tid];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tid == 0) s_live = v;
  }
  __syncthreads();
  // torchvision/ops/boxes.py batched_nms: coordinate trick while the image has at most 100 000 box elements on CUDA
  const bool trick = (long long)s_live * 4 <= 100000;
  const float scale = (s_live > 0 ? mx : 0.f) + 1.0f;
  for (int t = tid; t < T; t += kThreads) {
    const size_t o = (size_t)n * T + t;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cat_ids[o] >= 0) {
      b = *reinterpret_cast<const float4*>(flat_boxes + o * 4);
      if (trick) {
        const float offv = (float)classes[o] * scale;
        b.x += offv;
        b.y += offv;
        b.z += offv;
        b.w += offv;
      }
    }
    *reinterpret_cast<float4*>(nms_boxes + o * 4) = b;
  }
}

}  // namespace

D2B_API int d2b_frcnn_prepare(const float* boxes, const float* scores, const int* row_start, int N, int num_classes, int kreg,
                              const float* image_hw, float score_thresh, int cap, float* cand_boxes, float* nms_boxes,
                              float* nms_scores, float* raw_scores, int64_t* cand_flat, int64_t* cat_ids, int64_t* n_cand,
                              int64_t* row_map, void* stream) {
  if (N < 0 || N > D2B_MAX_IMAGES || num_classes <= 0 || (kreg != 1 && kreg != num_classes) || cap < 0 || !row_start)
    return D2B_EINVAL;
  if (N == 0) return D2B_OK;
  FrcnnImages I = {};
  I.N = N;
  for (int i = 0; i <= N; ++i) {
    I.row_start[i] = row_start[i];
    if (i && row_start[i] < row_start[i - 1]) return D2B_EINVAL;
  }
  if (!image_hw || !n_cand) return D2B_EINVAL;
  if (row_start[N] > row_start[0] && (!boxes || !scores || !row_map)) return D2B_EINVAL;
  if (cap > 0 && (!cand_boxes || !nms_boxes || !nms_scores || !raw_scores || !cand_flat || !cat_ids)) return D2B_EINVAL;
  if ((reinterpret_cast<uintptr_t>(cand_boxes) & 15) != 0 || (reinterpret_cast<uintptr_t>(nms_boxes) & 15) != 0) return D2B_EINVAL;
  frcnn_prepare_kernel<<<N, kThreads, 0, (cudaStream_t)stream>>>(I, boxes, scores, num_classes, kreg, image_hw, score_thresh, cap,
                                                                 cand_boxes, nms_boxes, nms_scores, raw_scores,
                                                                 (long long*)cand_flat, (long long*)cat_ids, (long long*)n_cand,
                                                                 (long long*)row_map);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

D2B_API int d2b_dense_prepare(const d2b_dense_levels* lv, int N, int num_classes, const float* weights, float scale_clamp,
                              float* flat_boxes, float* nms_boxes, float* nms_scores, float* raw_scores, int64_t* classes,
                              int64_t* cat_ids, void* stream) {
  if (!lv || lv->num_levels < 1 || lv->num_levels > D2B_MAX_LEVELS || N < 0 || num_classes <= 0 || !weights) return D2B_EINVAL;
  if (N == 0) return D2B_OK;
  DenseLevels P = {};
  P.L = lv->num_levels;
  int T = 0;
  for (int l = 0; l < P.L; ++l) {
    if (lv->R[l] < 0 || lv->k[l] < 0) return D2B_EINVAL;
    if (lv->k[l] > 0 && (!lv->anchors[l] || !lv->deltas[l] || !lv->topk_idx[l] || !lv->topk_scores[l])) return D2B_EINVAL;
    if ((reinterpret_cast<uintptr_t>(lv->anchors[l]) & 15) != 0 || (reinterpret_cast<uintptr_t>(lv->deltas[l]) & 15) != 0)
      return D2B_EINVAL;
    P.anchors[l] = lv->anchors[l];
    P.deltas[l] = lv->deltas[l];
    P.topk_idx[l] = lv->topk_idx[l];
    P.topk_scores[l] = lv->topk_scores[l];
    P.R[l] = lv->R[l];
    P.k[l] = lv->k[l];
    P.t0[l] = T;
    T += lv->k[l];
  }
  P.t0[P.L] = T;
  if (T == 0) return D2B_OK;
  if (!flat_boxes || !nms_boxes || !nms_scores || !raw_scores || !classes || !cat_ids) return D2B_EINVAL;
  dense_prepare_kernel<<<N, kThreads, 0, (cudaStream_t)stream>>>(P, T, num_classes, weights[0], weights[1], weights[2], weights[3],
                                                                 scale_clamp, flat_boxes, nms_boxes, nms_scores, raw_scores,
                                                                 (long long*)classes, (long long*)cat_ids);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

// ================================================================================================ mask targets + loss
// Mask-head training target and loss in one pass:
//   BitMasks.crop_and_resize (detectron2/structures/masks.py:193-224): RoIAlign(S x S, scale 1, sampling_ratio 0, aligned)
//   of the proposal's ground-truth bitmask, thresholded at 0.5
//   + the per-class gather and binary_cross_entropy_with_logits of mask_rcnn_loss (modeling/roi_heads/mask_head.py:60-112).
// The reference first materialises one H x W byte mask PER PROPOSAL (BitMasks indexing, K x H x W bytes), converts it to
// fp32, pools it, thresholds, gathers the class channel of the logits and reduces.  Here a CTA per proposal samples the
// ground-truth mask it is matched to (mask_index) straight from the [G,H,W] byte tensor, one thread per output bin with the
// reference's accumulation order (torchvision roi_align: sum over iy, ix of w1 v1 + w2 v2 + w3 v3 + w4 v4, then / count),
// writes the 0/1 target (kept for the backward and the accuracy statistics) and the proposal's loss sum.
namespace {

struct Tap1 {
  int lo, hi;
  float wl, wh;
};

__device__ __forceinline__ Tap1 make_tap1(float v, int size) {  // torchvision roi_align bilinear_interpolate
  Tap1 t;
  if (v < -1.0f || v > (float)size) {
    t.lo = t.hi = 0;
    t.wl = t.wh = 0.f;
    return t;
  }
  v = fmaxf(v, 0.f);
  int lo = (int)v, hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = v - (float)lo;
  t.lo = lo;
  t.hi = hi;
  t.wh = l;
  t.wl = 1.f - l;
  return t;
}

__global__ void __launch_bounds__(256) mask_loss_fwd_kernel(const float* __restrict__ logits, int C, int S,
                                                            const unsigned char* __restrict__ gt, int G, int H, int W,
                                                            const float* __restrict__ boxes,
                                                            const long long* __restrict__ mask_index,
                                                            const long long* __restrict__ classes,
                                                            float* __restrict__ loss_per_roi,
                                                            unsigned char* __restrict__ targets) {
  __shared__ float s_red[8];
  const int k = blockIdx.x, tid = threadIdx.x;
  const float* b = boxes + (size_t)k * 4;
  // aligned = True, spatial_scale = 1: box - 0.5, no minimum size
  const float sw = b[0] * 1.0f - 0.5f, sh = b[1] * 1.0f - 0.5f, ew = b[2] * 1.0f - 0.5f, eh = b[3] * 1.0f - 0.5f;
  const float rw = ew - sw, rh = eh - sh;
  const float bin_h = rh / (float)S, bin_w = rw / (float)S;
  int gh = (int)ceilf(rh / (float)S), gw = (int)ceilf(rw / (float)S);
  gh = max(gh, 0);
  gw = max(gw, 0);
  const float count = (float)max(gh * gw, 1);
  long long mi = mask_index ? mask_index[k] : k;
  const bool have_mask = mi >= 0 && mi < G;
  const unsigned char* __restrict__ m = gt + (size_t)(have_mask ? mi : 0) * H * W;
  const long long cls = classes ? classes[k] : 0;
  const bool cls_ok = cls >= 0 && cls < C;
  const float* __restrict__ lg = logits + ((size_t)k * C + (cls_ok ? cls : 0)) * S * S;
  float acc_loss = 0.f;
  for (int bin = tid; bin < S * S; bin += 256) {
    const int ph = bin / S, pw = bin - ph * S;
    float v = 0.f;
    if (have_mask) {
      for (int iy = 0; iy < gh; ++iy) {
        const Tap1 ty = make_tap1(sh + (float)ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh, H);
        for (int ix = 0; ix < gw; ++ix) {
          const Tap1 tx = make_tap1(sw + (float)pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw, W);
          const float v1 = m[(size_t)ty.lo * W + tx.lo] ? 1.f : 0.f, v2 = m[(size_t)ty.lo * W + tx.hi] ? 1.f : 0.f;
          const float v3 = m[(size_t)ty.hi * W + tx.lo] ? 1.f : 0.f, v4 = m[(size_t)ty.hi * W + tx.hi] ? 1.f : 0.f;
          v += (ty.wl * tx.wl) * v1 + (ty.wl * tx.wh) * v2 + (ty.wh * tx.wl) * v3 + (ty.wh * tx.wh) * v4;
        }
      }
      v /= count;
    }
    const float t = v >= 0.5f ? 1.f : 0.f;
    targets[(size_t)k * S * S + bin] = (unsigned char)(v >= 0.5f);
    if (cls_ok) {
      const float x = lg[bin];
      // binary_cross_entropy_with_logits: (1 - t) x + max(-x, 0) + log(exp(-max) + exp(-x - max)),  max = max(-x, 0)
      const float mxv = fmaxf(-x, 0.f);
      acc_loss += (1.f - t) * x + mxv + logf(expf(-mxv) + expf(-x - mxv));
    }
  }
  for (int o = 16; o; o >>= 1) acc_loss += __shfl_xor_sync(0xffffffffu, acc_loss, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = acc_loss;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    loss_per_roi[k] = s;
  }
}

// grad_logits[k, c, :] = (c == class_k) ? (sigmoid(x) - t) * grad_scale[k] : 0      grid (K, C)
__global__ void __launch_bounds__(256) mask_loss_bwd_kernel(const float* __restrict__ logits, int C, int S,
                                                            const unsigned char* __restrict__ targets,
                                                            const long long* __restrict__ classes,
                                                            const float* __restrict__ grad_scale,
                                                            float* __restrict__ grad_logits) {
  const int k = blockIdx.x, c = blockIdx.y;
  const long long cls = classes ? classes[k] : 0;
  const size_t base = ((size_t)k * C + c) * S * S;
  const float scale = grad_scale[k];  // d loss / d loss_per_roi[k]
  for (int bin = threadIdx.x; bin < S * S; bin += 256) {
    float gval = 0.f;
    if (c == cls) {
      const float x = logits[base + bin];
      const float t = targets[(size_t)k * S * S + bin] ? 1.f : 0.f;
      gval = (1.f / (1.f + expf(-x)) - t) * scale;
    }
    grad_logits[base + bin] = gval;
  }
}

}  // namespace

D2B_API int d2b_mask_loss_forward(const float* logits, int K, int C, int S, const uint8_t* gt_masks, int G, int H, int W,
                                  const float* boxes, const int64_t* mask_index, const int64_t* classes,
                                  float* loss_per_roi, uint8_t* targets, void* stream) {
  if (K < 0 || C <= 0 || S <= 0 || G < 0 || H <= 0 || W <= 0) return D2B_EINVAL;
  if (K == 0) return D2B_OK;
  if (!logits || !gt_masks || !boxes || !loss_per_roi || !targets) return D2B_EINVAL;
  mask_loss_fwd_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(logits, C, S, gt_masks, G, H, W, boxes, (const long long*)mask_index,
                                                            (const long long*)classes, loss_per_roi, targets);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

D2B_API int d2b_mask_loss_backward(const float* logits, int K, int C, int S, const uint8_t* targets, const int64_t* classes,
                                   const float* grad_scale, float* grad_logits, void* stream) {
  if (K < 0 || C <= 0 || S <= 0 || C > 65535) return D2B_EINVAL;
  if (K == 0) return D2B_OK;
  if (!logits || !targets || !grad_scale || !grad_logits) return D2B_EINVAL;
  dim3 grid(K, C);
  mask_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, C, S, targets, (const long long*)classes, grad_scale,
                                                               grad_logits);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}
