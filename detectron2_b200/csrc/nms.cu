// NMS (axis-aligned + rotated) and pairwise rotated-box IoU for sm_100a.
//
// Replaces torchvision::nms as reached from detectron2/layers/nms.py:5-22, and
// detectron2/layers/csrc/{nms_rotated/nms_rotated_cuda.cu, box_iou_rotated/box_iou_rotated_cuda.cu}.
// This file is compiled with -fmad=false: every float expression rounds like the reference's CPU build
// (x86-64, no FMA), which is the bit-exact parity target (see DESIGN.md "bit-exactness").
//
// Pipeline of d2b_nms (all on the caller's stream, no host round trip -- the reference copies the N x N/64 bitmask
// to the host and scans it there, nms_rotated_cuda.cu:114-137):
//   1. stable descending radix sort of the scores (CUB)                              -> order[rank]
//   2. batched NMS: stable radix sort of the ranks by category (CUB) -> class-major order with descending scores
//      inside each class; a single-CTA kernel derives the segment table
//   3. gather boxes in that order, applying the batched-NMS coordinate offsets in fp32 on the fly
//   4. IoU bitmask, 64x64 tiles, upper triangle and same-class tiles only, stored COLUMN-WORD-MAJOR maskT[w][i] so that
//      both the tile writes and the scan's reads are coalesced
//   5. greedy scan, one CTA per class segment in parallel: per 64-box block one thread resolves the intra-block chain
//      (branch-free) from the diagonal word, then 16 warps OR the kept rows into the `removed` words (rows prefetched
//      one block ahead in ping-pong registers); kept boxes are flagged at their global score rank
//   6. single-CTA compaction of the flags in score order -> kept original indices + device-side count.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// rotated IoU, after box_iou_rotated_utils.h (CPU branch).  float / double promotions follow the reference's
// C++ expression types exactly; see oracle/d2_oracle.c for the line-by-line citations.
// ------------------------------------------------------------------------------------------------
struct P2 {
  float x, y;
};
__device__ __forceinline__ float crs(P2 a, P2 b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ float dt(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ P2 sub(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }

__device__ __forceinline__ void rot_vertices(float xc, float yc, float w, float h, float a, P2* p) {
  double theta = (double)a * 0.01745329251;
  float c2 = (float)cos(theta) * 0.5f, s2 = (float)sin(theta) * 0.5f;
  p[0].x = xc + s2 * h + c2 * w;
  p[0].y = yc + c2 * h - s2 * w;
  p[1].x = xc - s2 * h + c2 * w;
  p[1].y = yc - c2 * h - s2 * w;
  p[2].x = 2 * xc - p[0].x;
  p[2].y = 2 * yc - p[0].y;
  p[3].x = 2 * xc - p[1].x;
  p[3].y = 2 * yc - p[1].y;
}

__device__ float rotated_iou(const float* __restrict__ b1, const float* __restrict__ b2) {
  const double sx = (double)(b1[0] + b2[0]) / 2.0, sy = (double)(b1[1] + b2[1]) / 2.0;
  const float x1 = (float)((double)b1[0] - sx), y1 = (float)((double)b1[1] - sy);
  const float x2 = (float)((double)b2[0] - sx), y2 = (float)((double)b2[1] - sy);
  const float area1 = b1[2] * b1[3], area2 = b2[2] * b2[3];
  if ((double)area1 < 1e-14 || (double)area2 < 1e-14) return 0.f;

  P2 p1[4], p2[4], v1[4], v2[4];
  rot_vertices(x1, y1, b1[2], b1[3], b1[4], p1);
  rot_vertices(x2, y2, b2[2], b2[3], b2[4], p2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v1[i] = sub(p1[(i + 1) & 3], p1[i]);
    v2[i] = sub(p2[(i + 1) & 3], p2[i]);
  }
  P2 ip[24];
  int num = 0;
  const double EPS = 1e-5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float det = crs(v2[j], v1[i]);
      if (fabs((double)det) <= 1e-14) continue;
      P2 v12 = sub(p2[j], p1[i]);
      float t1 = crs(v2[j], v12) / det;
      float t2 = crs(v1[i], v12) / det;
      if ((double)t1 > -EPS && (double)t1 < (double)1.0f + EPS && (double)t2 > -EPS &&
          (double)t2 < (double)1.0f + EPS) {
        ip[num].x = p1[i].x + v1[i].x * t1;
        ip[num].y = p1[i].y + v1[i].y * t1;
        ++num;
      }
    }
  }
  {  // vertices of rect1 inside rect2
    const P2 AB = v2[0], DA = v2[3];
    const float ABdotAB = dt(AB, AB), ADdotAD = dt(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      P2 AP = sub(p1[i], p2[0]);
      float APdotAB = dt(AP, AB), APdotAD = -dt(AP, DA);
      if (((double)APdotAB > -EPS) && ((double)APdotAD > -EPS) && ((double)APdotAB < (double)ABdotAB + EPS) &&
          ((double)APdotAD < (double)ADdotAD + EPS))
        ip[num++] = p1[i];
    }
  }
  {  // vertices of rect2 inside rect1
    const P2 AB = v1[0], DA = v1[3];
    const float ABdotAB = dt(AB, AB), ADdotAD = dt(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      P2 AP = sub(p2[i], p1[0]);
      float APdotAB = dt(AP, AB), APdotAD = -dt(AP, DA);
      if (((double)APdotAB > -EPS) && ((double)APdotAD > -EPS) && ((double)APdotAB < (double)ABdotAB + EPS) &&
          ((double)APdotAD < (double)ADdotAD + EPS))
        ip[num++] = p2[i];
    }
  }
  float inter = 0.f;
  if (num > 2) {
    // Graham scan, shift_to_zero variant
    int t = 0;
    for (int i = 1; i < num; ++i)
      if (ip[i].y < ip[t].y || (ip[i].y == ip[t].y && ip[i].x < ip[t].x)) t = i;
    const P2 start = ip[t];
    P2 q[24];
    float dist[24];
    for (int i = 0; i < num; ++i) q[i] = sub(ip[i], start);
    {
      P2 tmp = q[0];
      q[0] = q[t];
      q[t] = tmp;
    }
    for (int i = 0; i < num; ++i) dist[i] = dt(q[i], q[i]);
    for (int i = 1; i < num - 1; ++i)
      for (int j = i + 1; j < num; ++j) {
        float cp = crs(q[i], q[j]);
        if (((double)cp < -1e-6) || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
          P2 qt = q[i];
          q[i] = q[j];
          q[j] = qt;
          float d = dist[i];
          dist[i] = dist[j];
          dist[j] = d;
        }
      }
    // the CPU reference recomputes dist after the sort; after the swaps above dist[] already travels with q[],
    // and dot(q,q) is a pure function of q, so the recomputed values are identical.
    int k;
    for (k = 1; k < num; ++k)
      if ((double)dist[k] > 1e-8) break;
    int m;
    if (k == num) {
      m = 1;
    } else {
      q[1] = q[k];
      m = 2;
      for (int i = k + 1; i < num; ++i) {
        while (m > 1) {
          P2 q1 = sub(q[i], q[m - 2]), q2 = sub(q[m - 1], q[m - 2]);
          if (q1.x * q2.y >= q2.x * q1.y) m--;
          else break;
        }
        q[m++] = q[i];
      }
    }
    if (m > 2) {
      float area = 0.f;
      for (int i = 1; i < m - 1; ++i) area += fabsf(crs(sub(q[i], q[0]), sub(q[i + 1], q[0])));
      inter = (float)((double)area / 2.0);
    }
  }
  return inter / (area1 + area2 - inter);
}

// box_iou_rotated_cuda.cu:14-63 equivalent: one thread per (i,j) pair, j fastest for coalesced output.
__global__ void __launch_bounds__(128) box_iou_rotated_kernel(const float* __restrict__ b1, long long N,
                                                              const float* __restrict__ b2, long long M,
                                                              float* __restrict__ out) {
  __shared__ float sb2[128 * 5];
  const long long j0 = (long long)blockIdx.x * 128;
  const int nj = (int)min((long long)128, M - j0);
  for (int t = threadIdx.x; t < nj * 5; t += 128) sb2[t] = b2[j0 * 5 + t];
  __syncthreads();
  if ((int)threadIdx.x >= nj) return;
  for (long long i = blockIdx.y; i < N; i += gridDim.y) {
    float a[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) a[c] = b1[i * 5 + c];
    out[i * M + j0 + threadIdx.x] = rotated_iou(a, sb2 + threadIdx.x * 5);
  }
}

// ------------------------------------------------------------------------------------------------
// NMS
// ------------------------------------------------------------------------------------------------
// v[i] = i (value array of the radix sorts) and keepflag[i] = 0 (output of the scans), one launch
__global__ void iota_kernel(int* __restrict__ v, unsigned char* __restrict__ keepflag, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    v[i] = i;
    keepflag[i] = 0;
  }
}

// coordinate range for the batched-NMS offset trick.  mm[0] = max, mm[1] = min.
//   axis-aligned: max over all 4 coordinates (torchvision _batched_nms_coordinate_trick: boxes.max())
//   rotated:      max(max(cx,cy) + max(w,h)/2), min(min(cx,cy) - max(w,h)/2)   (detectron2/layers/nms.py:137-143)
template <bool ROT>
__global__ void __launch_bounds__(1024) coord_range_kernel(const float* __restrict__ boxes, int M,
                                                           float* __restrict__ mm) {
  __shared__ float smax[32], smin[32];
  float mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < M; i += 1024) {
    if (ROT) {
      const float* b = boxes + (size_t)i * 5;
      float half = fmaxf(b[2], b[3]) / 2;
      mx = fmaxf(mx, fmaxf(b[0], b[1]) + half);
      mn = fminf(mn, fminf(b[0], b[1]) - half);
    } else {
      const float* b = boxes + (size_t)i * 4;
      mx = fmaxf(mx, fmaxf(fmaxf(b[0], b[1]), fmaxf(b[2], b[3])));
    }
  }
  for (int o = 16; o; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = mx;
    smin[threadIdx.x >> 5] = mn;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = smax[threadIdx.x];
    mn = smin[threadIdx.x];
    for (int o = 16; o; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (threadIdx.x == 0) {
      mm[0] = mx;
      mm[1] = mn;
    }
  }
}

// class id (as int32) of the box at score rank r
__global__ void class_of_rank_kernel(const int64_t* __restrict__ idxs, const int* __restrict__ order, int M,
                                     int* __restrict__ cls) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < M) cls[r] = (int)idxs[order[r]];
}

// sorted[p] = boxes[order[pos2 ? pos2[p] : p]]  (+ the batched-NMS coordinate offset of its class)
template <bool ROT>
__global__ void gather_boxes_kernel(const float* __restrict__ boxes, const int* __restrict__ order,
                                    const int* __restrict__ pos2, const int64_t* __restrict__ idxs,
                                    const float* __restrict__ mm, int M, float* __restrict__ sorted) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int src = order[pos2 ? pos2[r] : r];
  constexpr int D = ROT ? 5 : 4;
  float b[D];
#pragma unroll
  for (int c = 0; c < D; ++c) b[c] = boxes[(size_t)src * D + c];
  if (idxs) {
    if (ROT) {
      float off = (float)idxs[src] * (mm[0] - mm[1] + 1.0f);
      b[0] += off;
      b[1] += off;
    } else {
      float off = (float)idxs[src] * (mm[0] + 1.0f);
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] += off;
    }
  }
#pragma unroll
  for (int c = 0; c < D; ++c) sorted[(size_t)r * D + c] = b[c];
}

// 64x64 IoU tile -> one 64-bit word per row.  grid (col_block, row_block); only col_block >= row_block does work.
// maskT[(size_t)col_block * M + row]  (column-word-major).
// cls (optional): class of every position of the class-major order; only same-class pairs can suppress each other, and a
// tile whose row block and column block share no class is skipped altogether (never read by the scan).
// Thread layout: kSub threads per row, each testing 64/kSub columns, partial words OR-ed with warp shuffles.  The rotated
// IoU is ~50x the work of the axis-aligned one, so it gets 8 threads per row (512-thread CTAs), the cheap one gets 4.
template <bool ROT, int kSub>
__global__ void __launch_bounds__(64 * kSub) nms_mask_kernel(const float* __restrict__ sb, const int* __restrict__ cls,
                                                             int M, double thr, unsigned long long* __restrict__ maskT) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  if (cls && cb > rb && cls[min(rb * 64 + 63, M - 1)] != cls[cb * 64]) return;  // classes ascend with position
  constexpr int D = ROT ? 5 : 4;
  constexpr int kCols = 64 / kSub;
  __shared__ float cbox[64 * D];
  __shared__ int ccls[64];
  const int c0 = cb * 64, r0 = rb * 64;
  const int nc = min(64, M - c0);
  for (int t = threadIdx.x; t < nc * D; t += 64 * kSub) cbox[t] = sb[(size_t)c0 * D + t];
  if (cls && (int)threadIdx.x < nc) ccls[threadIdx.x] = cls[c0 + threadIdx.x];
  __syncthreads();
  const int lrow = threadIdx.x / kSub, sub = threadIdx.x % kSub;
  const int row = r0 + lrow;
  const bool row_ok = row < M;
  float a[D];
#pragma unroll
  for (int c = 0; c < D; ++c) a[c] = sb[(size_t)(row_ok ? row : M - 1) * D + c];
  unsigned long long bits = 0ull;
  const int start = (cb == rb) ? lrow + 1 : 0;
  const int my_cls = cls ? cls[row_ok ? row : M - 1] : 0;
  const int jbeg = max(start, sub * kCols), jend = min(nc, (sub + 1) * kCols);
  if (row_ok) {
    if (ROT) {
      for (int j = jbeg; j < jend; ++j) {
        if (cls && ccls[j] != my_cls) continue;
        float iou = rotated_iou(a, cbox + j * 5);
        if ((double)iou >= thr) bits |= 1ull << j;  // nms_rotated_cpu.cpp:54
      }
    } else {
      const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
      for (int j = jbeg; j < jend; ++j) {
        if (cls && ccls[j] != my_cls) continue;
        const float* b = cbox + j * 4;
        float xx1 = fmaxf(a[0], b[0]), yy1 = fmaxf(a[1], b[1]);
        float xx2 = fminf(a[2], b[2]), yy2 = fminf(a[3], b[3]);
        float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
        float inter = w * h;
        float area_b = (b[2] - b[0]) * (b[3] - b[1]);
        float ovr = inter / (area_a + area_b - inter);
        if ((double)ovr > thr) bits |= 1ull << j;  // torchvision nms: strict
      }
    }
  }
#pragma unroll
  for (int o = 1; o < kSub; o <<= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);  // the kSub lanes of a row are adjacent
  if (row_ok && sub == 0) maskT[(size_t)cb * M + row] = bits;
}

constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kMaxColsPerWarp = 10;  // register-prefetched column words per warp (covers segments <= 64*16*10 = 10240)

// Exclusive prefix sum of one int per thread over a 1024-thread CTA (shuffle scan inside the warps, one barrier);
// `total` receives the sum over the CTA.  warp_tot: 32 ints of shared memory.
__device__ __forceinline__ int block_excl_scan_1024(int v, int* __restrict__ warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
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

// segment table of the class-major order: seg_start[0..nseg], seg_start[nseg] = M.  Single CTA; every thread owns a
// contiguous strip of positions (count, one block-wide scan, write), so the table comes out in ascending order.
__global__ void __launch_bounds__(1024) nms_segments_kernel(const int* __restrict__ cls, int M, int* __restrict__ seg_start,
                                                            int* __restrict__ nseg) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  const int per = (M + 1023) / 1024;
  const int p0 = min(M, tid * per), p1 = min(M, p0 + per);
  int cnt = 0;
  for (int p = p0; p < p1; ++p) cnt += (cls == nullptr ? p == 0 : (p == 0 || cls[p] != cls[p - 1])) ? 1 : 0;
  int total;
  int idx = block_excl_scan_1024(cnt, warp_tot, total);
  for (int p = p0; p < p1; ++p)
    if (cls == nullptr ? p == 0 : (p == 0 || cls[p] != cls[p - 1])) seg_start[idx++] = p;
  if (tid == 0) {
    seg_start[total] = M;
    *nseg = total;
  }
}

// Greedy scan over the bitmask, one CTA per class segment (plain NMS = one segment covering everything).
// dynamic smem: removed[nb] (uint64).  Per 64-box block b of the segment: (B) thread 0 resolves the intra-block chain from
// removed[b] and the diagonal word of each row; (C) the warps OR the kept rows into the `removed` words of the later
// column blocks of the segment.  Everything the next block needs from global memory (its diagonal words, its rows of
// the later columns) is requested one full iteration ahead and parked in registers, so the serial chain never waits on
// L2.  Output: keepflag[rank] = 1 for every kept box, rank = its position in the global score order.
__global__ void __launch_bounds__(kScanThreads, 1) nms_scan_kernel(const unsigned long long* __restrict__ maskT,
                                                                   const int* __restrict__ pos2,
                                                                   const int* __restrict__ seg_start,
                                                                   const int* __restrict__ nseg_ptr, int M,
                                                                   unsigned char* __restrict__ keepflag) {
  extern __shared__ unsigned long long removed[];
  __shared__ __align__(16) unsigned long long s_diag[2][64];
  __shared__ unsigned long long s_kept;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nseg = *nseg_ptr;
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int p0 = seg_start[seg], p1 = seg_start[seg + 1];
    const int b0 = p0 >> 6, nb = ((p1 - 1) >> 6) + 1;  // blocks [b0, nb) of the global tiling touch this segment
    __syncthreads();                                     // previous segment done with removed[] / s_diag
    for (int i = b0 + tid; i < nb; i += kScanThreads) removed[i] = 0ull;

    // rows 2*lane, 2*lane+1 of block b, column words w = b + 1 + warp + kScanWarps*c
    auto fetch = [&](int b, ulonglong2 (&dst)[kMaxColsPerWarp]) {
      const int r = b * 64 + 2 * lane;
#pragma unroll
      for (int c = 0; c < kMaxColsPerWarp; ++c) {
        const int w = b + 1 + warp + kScanWarps * c;
        dst[c] = make_ulonglong2(0ull, 0ull);
        if (b < nb && w < nb) {
          const unsigned long long* p = maskT + (size_t)w * M + r;
          if (r + 1 < M) {
            if ((M & 1) == 0) dst[c] = *reinterpret_cast<const ulonglong2*>(p);  // 16 B aligned when M is even
            else dst[c] = make_ulonglong2(p[0], p[1]);
          } else if (r < M) {
            dst[c].x = p[0];
          }
        }
      }
    };
    auto diag_word = [&](int b) -> unsigned long long {
      const int r = b * 64 + tid;
      return (b < nb && tid < 64 && r >= p0 && r < p1) ? maskT[(size_t)b * M + r] : 0ull;
    };
    ulonglong2 bufA[kMaxColsPerWarp], bufB[kMaxColsPerWarp];
    fetch(b0, bufA);
    if (tid < 64) s_diag[b0 & 1][tid] = diag_word(b0);
    unsigned long long dnext = diag_word(b0 + 1);
    __syncthreads();

    // one block of 64 boxes; `cur` holds its rows (requested one iteration ago), `nxt` receives the next block's rows.
    // The two register buffers ping-pong (no copies: a copy would be a use and would expose the load latency).
    auto process = [&](int b, ulonglong2 (&cur)[kMaxColsPerWarp], ulonglong2 (&nxt)[kMaxColsPerWarp]) {
      // rows of this block that belong to the segment
      const int lo = max(p0 - b * 64, 0), hi = min(p1 - b * 64, 64);  // [lo, hi)
      const unsigned long long vmask =
          (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
      // ---- step B: intra-block chain.  One thread, branch-free: per row the dependent path is
      //      bit test -> mask -> and/or (about four ALU latencies); the diagonal words are pre-read from shared memory
      //      sixteen rows at a time so that no load sits on the chain.
      if (tid == 0) {
        const unsigned long long rem = removed[b];
        unsigned rlo = (unsigned)rem, rhi = (unsigned)(rem >> 32), klo = 0u, khi = 0u;
        const ulonglong2* dg = reinterpret_cast<const ulonglong2*>(s_diag[b & 1]);
        const unsigned vlo = (unsigned)vmask, vhi = (unsigned)(vmask >> 32);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          ulonglong2 d[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = dg[blk * 8 + q];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int i = blk * 16 + q;
            const unsigned long long dw = (q & 1) ? d[q >> 1].y : d[q >> 1].x;
            const unsigned m = 0u - ((~rlo & vlo) >> i & 1u);  // all-ones when row i is alive
            klo |= m & (1u << i);
            rlo |= m & (unsigned)dw;
            rhi |= m & (unsigned)(dw >> 32);
          }
        }
#pragma unroll
        for (int blk = 2; blk < 4; ++blk) {
          ulonglong2 d[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = dg[blk * 8 + q];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int i = (blk - 2) * 16 + q;
            const unsigned long long dw = (q & 1) ? d[q >> 1].y : d[q >> 1].x;
            const unsigned m = 0u - ((~rhi & vhi) >> i & 1u);
            khi |= m & (1u << i);
            rhi |= m & (unsigned)(dw >> 32);
          }
        }
        s_kept = ((unsigned long long)khi << 32) | klo;
      }
      __syncthreads();
      const unsigned long long kept = s_kept;
      // requests for block b+1 / b+2 go out now and are consumed one iteration later
      fetch(b + 1, nxt);
      if (tid < 64) s_diag[(b + 1) & 1][tid] = dnext;
      dnext = diag_word(b + 2);
      // ---- publish the kept boxes of this block at their score rank
      if (tid < 64 && ((kept >> tid) & 1ull)) {
        const int p = b * 64 + tid;
        keepflag[pos2 ? pos2[p] : p] = 1;
      }
      // ---- step C: OR kept rows into later column words
      const unsigned long long k0 = (kept >> (2 * lane)) & 1ull ? ~0ull : 0ull;
      const unsigned long long k1 = (kept >> (2 * lane + 1)) & 1ull ? ~0ull : 0ull;
#pragma unroll
      for (int c = 0; c < kMaxColsPerWarp; ++c) {
        const int w = b + 1 + warp + kScanWarps * c;
        if (w < nb) {  // warp-uniform
          unsigned long long v = (cur[c].x & k0) | (cur[c].y & k1);
          unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)v);
          unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
          if (lane == 0) removed[w] |= ((unsigned long long)hi32 << 32) | lo32;
        }
      }
      // columns beyond the register-prefetched window (very large segments): plain loads
      for (int w = b + 1 + warp + kScanWarps * kMaxColsPerWarp; w < nb; w += kScanWarps) {
        const int r = b * 64 + 2 * lane;
        unsigned long long v = 0ull;
        if (r < M) v |= maskT[(size_t)w * M + r] & k0;
        if (r + 1 < M) v |= maskT[(size_t)w * M + r + 1] & k1;
        unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)v);
        unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
        if (lane == 0) removed[w] |= ((unsigned long long)hi32 << 32) | lo32;
      }
      __syncthreads();  // removed[], s_diag[(b+1)&1] visible; s_kept consumed
    };
    for (int b = b0; b < nb; b += 2) {
      process(b, bufA, bufB);
      if (b + 1 < nb) process(b + 1, bufB, bufA);
    }
  }
}

// keep[] = original indices of the flagged ranks, in rank (= score) order; single CTA, contiguous strip per thread.
__global__ void __launch_bounds__(1024) nms_compact_kernel(const unsigned char* __restrict__ keepflag,
                                                           const int* __restrict__ order, int M,
                                                           long long* __restrict__ keep, long long* __restrict__ num_keep) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  const int per = (M + 1023) / 1024;
  const int r0 = min(M, tid * per), r1 = min(M, r0 + per);
  int cnt = 0;
  for (int r = r0; r < r1; ++r) cnt += keepflag[r] ? 1 : 0;
  int total;
  int idx = block_excl_scan_1024(cnt, warp_tot, total);
  for (int r = r0; r < r1; ++r)
    if (keepflag[r]) keep[idx++] = (long long)order[r];
  if (tid == 0) *num_keep = (long long)total;
}

struct NmsWorkspace {
  float* sorted_scores;
  int* iota;
  int* order;
  int* cls;
  int* cls_sorted;
  int* pos2;
  int* seg_start;
  int* nseg;
  unsigned char* keepflag;
  float* sorted_boxes;
  float* mm;
  unsigned long long* maskT;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

NmsWorkspace carve(void* base, int64_t M, int rotated) {
  NmsWorkspace w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align_up(bytes);
    return r;
  };
  const size_t m = (size_t)(M > 0 ? M : 1);
  const size_t nb = (m + 63) / 64;
  w.sorted_scores = (float*)take(m * 4);
  w.iota = (int*)take(m * 4);
  w.order = (int*)take(m * 4);
  w.cls = (int*)take(m * 4);
  w.cls_sorted = (int*)take(m * 4);
  w.pos2 = (int*)take(m * 4);
  w.seg_start = (int*)take((m + 1) * 4);
  w.nseg = (int*)take(16);
  w.keepflag = (unsigned char*)take(m);
  w.sorted_boxes = (float*)take(m * (rotated ? 5 : 4) * 4);
  w.mm = (float*)take(16);
  w.maskT = (unsigned long long*)take(nb * m * 8);
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, b1, (const float*)nullptr, (float*)nullptr, (const int*)nullptr,
                                            (int*)nullptr, (int)m);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr,
                                  (int)m);
  w.cub_bytes = b1 > b2 ? b1 : b2;
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace

D2B_API size_t d2b_nms_workspace_bytes(int64_t M, int flags) { return carve(nullptr, M, (flags & D2B_NMS_ROTATED) ? 1 : 0).total; }

D2B_API int d2b_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t M, double iou_threshold,
                    int flags, int64_t* keep, int64_t* num_keep, void* workspace, size_t workspace_bytes,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rotated = (flags & D2B_NMS_ROTATED) ? 1 : 0;
  const bool no_offset = (flags & D2B_NMS_NO_OFFSET) != 0;  // idxs only segment the boxes; coordinates are used as given
  if (!num_keep || M < 0) return D2B_EINVAL;
  if (M == 0) {
    D2B_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int64_t), stream));
    return D2B_OK;
  }
  if (!boxes || !scores || !keep || !workspace) return D2B_EINVAL;
  if (M > (1 << 30)) return D2B_EUNSUPPORTED;
  NmsWorkspace w = carve(workspace, M, rotated);
  if (workspace_bytes < w.total) return D2B_EWORKSPACE;
  const int m = (int)M, nb = (m + 63) / 64;
  const size_t smem = (size_t)nb * sizeof(unsigned long long);
  if (smem > 200 * 1024) return D2B_EUNSUPPORTED;
  // 1. global stable descending score order
  iota_kernel<<<d2b_cdiv(m, 256), 256, 0, stream>>>(w.iota, w.keepflag, m);
  D2B_CHECK_LAUNCH();
  size_t cub_bytes = w.cub_bytes;
  D2B_CUDA(cub::DeviceRadixSort::SortPairsDescending(w.cub_temp, cub_bytes, scores, w.sorted_scores, w.iota, w.order, m,
                                                     0, 32, stream));
  // 2. batched: class-major order (stable, so scores stay descending inside each class) + segment table
  const int* cls_sorted = nullptr;
  const int* pos2 = nullptr;
  if (idxs) {
    if (!no_offset) {
      if (rotated) coord_range_kernel<true><<<1, 1024, 0, stream>>>(boxes, m, w.mm);
      else coord_range_kernel<false><<<1, 1024, 0, stream>>>(boxes, m, w.mm);
      D2B_CHECK_LAUNCH();
    }
    class_of_rank_kernel<<<d2b_cdiv(m, 256), 256, 0, stream>>>(idxs, w.order, m, w.cls);
    D2B_CHECK_LAUNCH();
    cub_bytes = w.cub_bytes;
    D2B_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, cub_bytes, w.cls, w.cls_sorted, w.iota, w.pos2, m, 0, 32, stream));
    cls_sorted = w.cls_sorted;
    pos2 = w.pos2;
  }
  nms_segments_kernel<<<1, 1024, 0, stream>>>(cls_sorted, m, w.seg_start, w.nseg);
  D2B_CHECK_LAUNCH();
  // 3. boxes in that order (coordinate offsets of the reference's batched-NMS trick applied in fp32)
  const int64_t* off_idxs = no_offset ? nullptr : idxs;
  if (rotated) gather_boxes_kernel<true><<<d2b_cdiv(m, 256), 256, 0, stream>>>(boxes, w.order, pos2, off_idxs, w.mm, m, w.sorted_boxes);
  else gather_boxes_kernel<false><<<d2b_cdiv(m, 256), 256, 0, stream>>>(boxes, w.order, pos2, off_idxs, w.mm, m, w.sorted_boxes);
  D2B_CHECK_LAUNCH();
  // 4. IoU bitmask (same-class tiles only)
  dim3 grid(nb, nb);
  if (rotated) nms_mask_kernel<true, 8><<<grid, 512, 0, stream>>>(w.sorted_boxes, cls_sorted, m, iou_threshold, w.maskT);
  else nms_mask_kernel<false, 4><<<grid, 256, 0, stream>>>(w.sorted_boxes, cls_sorted, m, iou_threshold, w.maskT);
  D2B_CHECK_LAUNCH();
  // 5. per-segment greedy scans in parallel, then compaction in global score order
  D2B_ALLOW_BIG_SMEM(nms_scan_kernel);
  const int scan_grid = idxs ? (m < 2 * kNumSMs ? m : 2 * kNumSMs) : 1;
  nms_scan_kernel<<<scan_grid, kScanThreads, smem, stream>>>(w.maskT, pos2, w.seg_start, w.nseg, m, w.keepflag);
  D2B_CHECK_LAUNCH();
  nms_compact_kernel<<<1, 1024, 0, stream>>>(w.keepflag, w.order, m, (long long*)keep, (long long*)num_keep);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}

D2B_API int d2b_box_iou_rotated(const float* boxes1, int64_t N, const float* boxes2, int64_t M, float* ious,
                                void* stream) {
  if (N == 0 || M == 0) return D2B_OK;
  if (!boxes1 || !boxes2 || !ious || N < 0 || M < 0) return D2B_EINVAL;
  long long gx = (M + 127) / 128;
  if (gx > 2147483647LL) return D2B_EUNSUPPORTED;
  long long gy = N < 65535 ? N : 65535;
  // keep the grid a few waves deep; rows beyond gridDim.y are covered by the stride loop
  dim3 grid((unsigned)gx, (unsigned)gy);
  box_iou_rotated_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(boxes1, N, boxes2, M, ious);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}
