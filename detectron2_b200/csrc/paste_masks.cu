// paste_masks_in_image for sm_100a -- one fused kernel instead of the reference's
// meshgrid + grid_sample + compare + copy chain (detectron2/layers/mask_ops.py:17-69,74-147).
//
// HBM-bound byte kernel: the output (N*H*W bytes) dominates; each thread produces 16 consecutive output
// bytes and stores them with one 128-bit st.global.  The 28x28 soft mask of the current instance is staged in
// shared memory once per CTA.  Pixels whose sample point falls outside the mask support are written as 0
// without touching the mask (most of the image).
//
// Arithmetic mirrors the reference expression order (no FMA contraction: this file is compiled with -fmad=false):
//   g  = ((p + 0.5 - b0) / (b1 - b0)) * 2 - 1          (mask_ops.py:53-54)
//   i  = ((g + 1) * M - 1) / 2                           (grid_sample, align_corners=False)
//   v  = nw*w_nw + ne*w_ne + sw*w_sw + se*w_se           (zeros padding)
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPix = 16;       // output bytes per thread
constexpr int kMaxM = 64;      // largest mask side staged in smem (28 in every shipped config)

__device__ __forceinline__ float sample_coord(float p, float b0, float b1, float M) {
  float g = (p + 0.5f - b0) / (b1 - b0) * 2.f - 1.f;
  return ((g + 1.f) * M - 1.f) / 2.f;
}

__global__ void __launch_bounds__(kThreads) paste_masks_kernel(const float* __restrict__ masks,
                                                               const float* __restrict__ boxes, int M, int H, int W,
                                                               float threshold, uint8_t* __restrict__ out,
                                                               int chunks_per_mask) {
  __shared__ float smask[kMaxM * kMaxM];
  const int n = blockIdx.y;
  const float* __restrict__ mk = masks + (size_t)n * M * M;
  for (int i = threadIdx.x; i < M * M; i += kThreads) smask[i] = mk[i];
  const float x0 = boxes[4 * n], y0 = boxes[4 * n + 1], x1 = boxes[4 * n + 2], y1 = boxes[4 * n + 3];
  __syncthreads();
  const float fM = (float)M;
  const long long plane = (long long)H * W;
  uint8_t* __restrict__ obase = out + (size_t)n * plane;
  // obase may be misaligned w.r.t. 16 B when H*W is not a multiple of 16: chunk 0 starts at the first aligned byte,
  // the (<16 byte) head is handled by the last chunk id.
  const int head = (int)((16 - ((uintptr_t)obase & 15)) & 15);
  for (long long chunk = (long long)blockIdx.x * kThreads + threadIdx.x; chunk <= chunks_per_mask;
       chunk += (long long)gridDim.x * kThreads) {
    long long start, end;
    if (chunk == chunks_per_mask) {  // head
      start = 0;
      end = head < plane ? head : plane;
    } else {
      start = head + chunk * kPix;
      end = start + kPix;
      if (end > plane) end = plane;
    }
    if (start >= end) continue;
    int py = (int)(start / W);
    int px = (int)(start - (long long)py * W);
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    float iy = sample_coord((float)py, y0, y1, fM);
    const int cnt = (int)(end - start);
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      if (j >= cnt) break;
      float ix = sample_coord((float)px, x0, x1, fM);
      float v = 0.f;
      // in-range test written so that NaN / inf (degenerate boxes) fall through to 0
      if (ix > -1.f && ix < fM && iy > -1.f && iy < fM) {
        float fx = floorf(ix), fy = floorf(iy);
        int xw = (int)fx, yn = (int)fy, xe = xw + 1, ys = yn + 1;
        float wx1 = ix - fx, wx0 = (float)xe - ix, wy1 = iy - fy, wy0 = (float)ys - iy;
        bool okw = xw >= 0, oke = xe < M, okn = yn >= 0, oks = ys < M;
        if (okn && okw) v += smask[yn * M + xw] * (wx0 * wy0);
        if (okn && oke) v += smask[yn * M + xe] * (wx1 * wy0);
        if (oks && okw) v += smask[ys * M + xw] * (wx0 * wy1);
        if (oks && oke) v += smask[ys * M + xe] * (wx1 * wy1);
      }
      uint32_t byte = threshold >= 0.f ? (v >= threshold ? 1u : 0u) : (uint32_t)(uint8_t)(v * 255.f);
      pk[j >> 2] |= byte << ((j & 3) * 8);
      if (++px == W) {
        px = 0;
        ++py;
        iy = sample_coord((float)py, y0, y1, fM);
      }
    }
    uint8_t* dst = obase + start;
    if (cnt == kPix) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kPix; ++j)
        if (j < cnt) dst[j] = (uint8_t)(pk[j >> 2] >> ((j & 3) * 8));
    }
  }
}

}  // namespace

D2B_API int d2b_paste_masks(const float* masks, const float* boxes, int N, int M, int H, int W, float threshold,
                            uint8_t* out, void* stream) {
  if (N == 0 || H == 0 || W == 0) return D2B_OK;
  if (!masks || !boxes || !out || N < 0 || M <= 0 || H < 0 || W < 0) return D2B_EINVAL;
  if (M > kMaxM) return D2B_EUNSUPPORTED;
  long long plane = (long long)H * W;
  int chunks = (int)((plane + kPix - 1) / kPix);
  int gx = d2b_cdiv(chunks + 1, kThreads);
  // enough CTAs per mask to fill the machine even for a single mask, capped to keep the smem mask staging amortised
  int want = d2b_cdiv(8LL * kNumSMs, N);
  if (gx > want) gx = want < 1 ? 1 : want;
  dim3 grid(gx, N);
  paste_masks_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(masks, boxes, M, H, W, threshold, out, chunks);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}
