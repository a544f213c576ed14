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

// exact value of one output pixel (reference expression order, see file header)
__device__ __forceinline__ uint32_t paste_pixel(const float* __restrict__ smask, int M, float fM, int px, int py,
                                                float x0, float y0, float x1, float y1, float threshold) {
  const float ix = sample_coord((float)px, x0, x1, fM);
  const float iy = sample_coord((float)py, y0, y1, fM);
  float v = 0.f;
  // in-range test written so that NaN / inf (degenerate boxes) fall through to 0
  if (ix > -1.f && ix < fM && iy > -1.f && iy < fM) {
    const float fx = floorf(ix), fy = floorf(iy);
    const int xw = (int)fx, yn = (int)fy, xe = xw + 1, ys = yn + 1;
    const float wx1 = ix - fx, wx0 = (float)xe - ix, wy1 = iy - fy, wy0 = (float)ys - iy;
    const bool okw = xw >= 0, oke = xe < M, okn = yn >= 0, oks = ys < M;
    if (okn && okw) v += smask[yn * M + xw] * (wx0 * wy0);
    if (okn && oke) v += smask[yn * M + xe] * (wx1 * wy0);
    if (oks && okw) v += smask[ys * M + xw] * (wx0 * wy1);
    if (oks && oke) v += smask[ys * M + xe] * (wx1 * wy1);
  }
  return threshold >= 0.f ? (v >= threshold ? 1u : 0u) : (uint32_t)(uint8_t)(v * 255.f);
}

// grid (gx, N).  Two phases per mask:
//   1. every 16-byte chunk of the output plane that cannot see the mask (conservative rectangle test) is written as
//      one 128-bit store of the "outside" value -- this is ~90% of the bytes and runs at store bandwidth;
//   2. the rows/columns of the conservative rectangle, widened to whole 16-byte chunks, are evaluated exactly with one
//      pixel per lane, so a warp works on 32 neighbouring pixels (no divergence between inside / outside lanes).
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
  // Conservative support of the pasted mask: outside [cx0,cx1] x [ry0,ry1] every sample point lies >= 1.5 px beyond the
  // mask's (-1, M) support, far more than fp32 rounding can move it, so the value there is exactly that of v = 0.
  // Degenerate / non-finite boxes disable the shortcut (everything is evaluated exactly).
  int cx0 = 0, cx1 = W - 1, ry0 = 0, ry1 = H - 1;
  {
    const float bw = x1 - x0, bh = y1 - y0;
    if (W >= 2 * kPix && bw > 0.f && bh > 0.f && bw < 1e8f && bh < 1e8f && fabsf(x0) < 1e8f && fabsf(y0) < 1e8f) {
      const float fx0 = floorf(x0 - bw / fM) - 2.f, fx1 = ceilf(x1 + bw / fM) + 2.f;
      const float fy0 = floorf(y0 - bh / fM) - 2.f, fy1 = ceilf(y1 + bh / fM) + 2.f;
      cx0 = (int)fmaxf(fx0, 0.f);
      cx1 = (int)fminf(fx1, (float)(W - 1));
      ry0 = (int)fmaxf(fy0, 0.f);
      ry1 = (int)fminf(fy1, (float)(H - 1));
    }
  }
  const bool empty = cx1 < cx0 || ry1 < ry0;  // rectangle entirely off the image
  const uint32_t zbyte = threshold >= 0.f ? ((0.f >= threshold) ? 1u : 0u) : 0u;
  const uint32_t zword = zbyte * 0x01010101u;
  const long long plane = (long long)H * W;
  uint8_t* __restrict__ obase = out + (size_t)n * plane;
  // obase may be misaligned w.r.t. 16 B when H*W is not a multiple of 16: chunk c covers [head + 16c, head + 16c + 16)
  const int head = (int)((16 - ((uintptr_t)obase & 15)) & 15);
  const long long stride = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;

  // ---- phase 1: chunks that cannot see the mask
  for (long long chunk = gtid; chunk < chunks_per_mask; chunk += stride) {
    const long long start = head + chunk * kPix;
    if (start + kPix > plane) continue;  // ragged tail: phase 2b
    const int py = (int)(start / W);
    const int px = (int)(start - (long long)py * W);
    bool active = false;
    if (!empty) {
      const int pxe = px + kPix - 1;
      if (pxe < W) {
        active = (py >= ry0 && py <= ry1 && pxe >= cx0 && px <= cx1);
      } else {  // chunk wraps into the next row
        active = (py >= ry0 && py <= ry1 && px <= cx1) || (py + 1 >= ry0 && py + 1 <= ry1 && pxe - W >= cx0);
      }
    }
    if (!active) *reinterpret_cast<uint4*>(obase + start) = make_uint4(zword, zword, zword, zword);
  }
  // ---- phase 2a: the rectangle, row by row, widened to chunk boundaries (one pixel per lane)
  if (!empty) {
    const int RL = (cx1 - cx0 + 1) + 2 * (kPix - 1) + 1;
    const long long items = (long long)(ry1 - ry0 + 1) * RL;
    for (long long it = gtid; it < items; it += stride) {
      const int r = ry0 + (int)(it / RL);
      const int t = (int)(it - (long long)(r - ry0) * RL);
      const long long lo = (long long)r * W + cx0, hi = (long long)r * W + cx1;  // inclusive flat range of this row
      long long A = lo - head;
      A = (A >= 0 ? (A / kPix) * kPix : 0) + head;
      if (lo < head) A = 0;
      long long B = ((hi - head) / kPix + 1) * kPix + head;
      if (hi < head) B = head;
      if (B > plane) B = plane;
      const long long byte = A + t;
      if (byte >= B) continue;
      const int py = (int)(byte / W), px = (int)(byte - (long long)py * W);
      obase[byte] = (uint8_t)paste_pixel(smask, M, fM, px, py, x0, y0, x1, y1, threshold);
    }
  }
  // ---- phase 2b: unaligned head and ragged tail bytes (< 32 bytes per mask)
  if (blockIdx.x == 0) {
    const long long tail0 = head + (long long)((plane - head) / kPix) * kPix;
    for (long long byte = threadIdx.x; byte < head && byte < plane; byte += kThreads) {
      const int py = (int)(byte / W), px = (int)(byte - (long long)py * W);
      obase[byte] = (uint8_t)paste_pixel(smask, M, fM, px, py, x0, y0, x1, y1, threshold);
    }
    for (long long byte = tail0 + threadIdx.x; byte < plane; byte += kThreads) {
      if (byte < 0) continue;
      const int py = (int)(byte / W), px = (int)(byte - (long long)py * W);
      obase[byte] = (uint8_t)paste_pixel(smask, M, fM, px, py, x0, y0, x1, y1, threshold);
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
  int chunks = (int)((plane + kPix - 1) / kPix);  // upper bound; chunks past the plane are skipped in-kernel
  int gx = d2b_cdiv(chunks + 1, kThreads);
  // enough CTAs per mask to fill the machine even for a single mask, capped to keep the smem mask staging amortised
  int want = d2b_cdiv(8LL * kNumSMs, N);
  if (gx > want) gx = want < 1 ? 1 : want;
  dim3 grid(gx, N);
  paste_masks_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(masks, boxes, M, H, W, threshold, out, chunks);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}
