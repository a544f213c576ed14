// Deformable convolution forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The reference materialises columns[Cin*kh*kw, Ho*Wo] with a gather kernel and then calls cuBLAS per group
// (detectron2/layers/csrc/deformable/deform_conv_cuda.cu:382-431).  Here the gathered operand never exists in HBM:
//
//   D[pixel, oc] = sum_k' A[pixel, k'] * B[oc, k']        k' = kp * (Cin/G) + c   (kernel-point major: the bilinear taps of
//                                                          a (pixel, kernel point) are computed once and reused by 64 channels)
//
//   * MMA M = 128 output pixels (TMEM lanes), N = up to 128 output channels of the group (TMEM columns), K step 16 (bf16).
//   * A tile [128 x 64] bf16, K-major, 128-byte swizzle: written by the four GATHER warps (one thread per pixel row)
//     straight into the UMMA shared-memory layout -- bilinear gather -> registers -> st.shared.v4, then
//     fence.proxy.async + mbarrier arrive.
//   * B tile [N x 64] bf16, K-major, 128-byte swizzle: a loader warp copies it from a pre-converted bf16 weight copy.
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, fp32 accumulate in TMEM); tcgen05.commit releases the
//     smem stage to the producers (3-stage mbarrier ring) and finally signals the epilogue.
//   * epilogue: the gather warps read their 32 TMEM lanes with tcgen05.ld (32x32b.x16) and store NCHW output, one
//     coalesced 128-byte warp store per output channel (+ bias).
//
// precision 1 ("bf16x3"): operands are split x = hi + lo (two bf16) and three MMAs hi*hi + hi*lo + lo*hi are accumulated,
// which keeps ~16 mantissa bits per product -- fp32-class accuracy (<= 1e-4 rel) at 1/3 of the bf16 tensor peak.
// precision 2: plain bf16 operands (autocast-style), one MMA.
//
// Shapes taken: (Cin/G) % 64 == 0, (Cin/DG) % 64 == 0, (Cout/G) % 16 == 0.  Anything else returns D2B_EUNSUPPORTED
// (the fp32 FFMA kernel in deform_conv.cu covers it).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;      // pixels per tile  (UMMA M)
constexpr int BNMAX = 128;   // output channels per tile (UMMA N), multiple of 16
constexpr int BK = 64;       // k' per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kGatherThreads = 128;
constexpr int kThreads = 192;  // warps 0-3 gather + epilogue, warp 4 weight loader, warp 5 MMA issuer / TMEM owner
constexpr int kTileBytesA = BM * BK * 2;      // 16 KB
constexpr int kTileBytesB = BNMAX * BK * 2;   // 16 KB
constexpr int kStageBytes = 2 * kTileBytesA + 2 * kTileBytesB;  // hi + lo of A and B: 64 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kTmemCols = 128;

struct TcDims {
  int N, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, G, DG, Ho, Wo, cpg, opg, cpdg, KK, HoWo, K;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major tile, 128-byte swizzle, 8-row atoms 1024 B apart (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;         // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}

// D[tmem] (+)= A[smem] * B[smem];  issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------ weight pre-pass
// w fp32 [Cout][cpg][KK]  ->  bf16 hi / lo [Cout][KK][cpg]   (k' = kp * cpg + c)
__global__ void dcn_weight_split_kernel(const float* __restrict__ w, int Cout, int cpg, int KK,
                                        __nv_bfloat16* __restrict__ whi, __nv_bfloat16* __restrict__ wlo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Cout * cpg * KK;
  if (i >= total) return;
  const int c = i % cpg, kp = (i / cpg) % KK, oc = i / (cpg * KK);
  const float v = w[((size_t)oc * cpg + c) * KK + kp];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  whi[i] = h;
  wlo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------------ main kernel
// grid (pixel tiles, oc tiles, N*G)
__global__ void __launch_bounds__(kThreads, 1) dcn_fwd_tc_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ offset,
                                                                 const float* __restrict__ mask,
                                                                 const __nv_bfloat16* __restrict__ whi,
                                                                 const __nv_bfloat16* __restrict__ wlo,
                                                                 const float* __restrict__ bias, TcDims d, int split,
                                                                 float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]  producers -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages]  MMA -> producers
  uint64_t* accum_bar = bars + 2 * kStages;  // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BNMAX;
  const int b = blockIdx.z / d.G, g = blockIdx.z - b * d.G;
  const int bn = min(BNMAX, d.opg - n0);  // multiple of 16
  const int nchunk = d.K / BK;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], kGatherThreads + 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {  // TMEM allocation by one full warp; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =============================================================== GATHER producers (one pixel row per thread)
    const int row = tid;  // 0..127
    const int p = p0 + row;
    const bool pix_ok = p < d.HoWo;
    const int ho = pix_ok ? p / d.Wo : 0, wo = pix_ok ? p - (p / d.Wo) * d.Wo : 0;
    const size_t plane = (size_t)d.H * d.W;
    int cur_kp = -1, cur_dg = -1;
    int q0 = -1, q1 = -1, q2 = -1, q3 = -1;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
    const uint32_t sw = (uint32_t)(row & 7);
    for (int j = 0; j < nchunk; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (uint32_t)((j / kStages) & 1);
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int k0 = j * BK;
      const int kp = k0 / d.cpg, cl0 = k0 - kp * d.cpg;
      const int c_first = g * d.cpg + cl0;
      const int dg = c_first / d.cpdg;
      if (kp != cur_kp || dg != cur_dg) {  // taps of (pixel, kernel point): deform_conv_cuda_kernel.cu:263-282, :96-130
        cur_kp = kp;
        cur_dg = dg;
        q0 = q1 = q2 = q3 = -1;
        w0 = w1 = w2 = w3 = 0.f;
        if (pix_ok) {
          const int ki = kp / d.kw, kj = kp - ki * d.kw;
          const size_t ob = ((size_t)(b * d.DG + dg) * 2 * d.KK) * d.HoWo;
          const float oh = __ldg(offset + ob + (size_t)(2 * kp) * d.HoWo + p);
          const float ow = __ldg(offset + ob + (size_t)(2 * kp + 1) * d.HoWo + p);
          const float hf = (float)(ho * d.sh - d.ph + ki * d.dh) + oh;
          const float wf = (float)(wo * d.sw - d.pw + kj * d.dw) + ow;
          const float m = mask ? __ldg(mask + ((size_t)(b * d.DG + dg) * d.KK + kp) * d.HoWo + p) : 1.f;
          if (hf > -1.f && wf > -1.f && hf < (float)d.H && wf < (float)d.W) {
            const int hl = (int)floorf(hf), wl = (int)floorf(wf);
            const float lh = hf - (float)hl, lw = wf - (float)wl, hh = 1.f - lh, hw = 1.f - lw;
            const bool t0 = hl >= 0, t1 = hl + 1 <= d.H - 1, l0 = wl >= 0, l1 = wl + 1 <= d.W - 1;
            if (t0 && l0) { q0 = hl * d.W + wl; w0 = hh * hw * m; }
            if (t0 && l1) { q1 = hl * d.W + wl + 1; w1 = hh * lw * m; }
            if (t1 && l0) { q2 = (hl + 1) * d.W + wl; w2 = lh * hw * m; }
            if (t1 && l1) { q3 = (hl + 1) * d.W + wl + 1; w3 = lh * lw * m; }
          }
        }
      }
      uint8_t* a_hi = smem + s * kStageBytes;
      uint8_t* a_lo = a_hi + kTileBytesA;
      const float* __restrict__ xp = x + ((size_t)b * d.Cin + c_first) * plane;
#pragma unroll 2
      for (int c8 = 0; c8 < 8; ++c8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* __restrict__ pl = xp + (size_t)(c8 * 8 + i) * plane;
          float acc = 0.f;
          if (q0 >= 0) acc = fmaf(w0, __ldg(pl + q0), acc);
          if (q1 >= 0) acc = fmaf(w1, __ldg(pl + q1), acc);
          if (q2 >= 0) acc = fmaf(w2, __ldg(pl + q2), acc);
          if (q3 >= 0) acc = fmaf(w3, __ldg(pl + q3), acc);
          v[i] = acc;
        }
        uint4 hi, lo;
        {
          __nv_bfloat16 h[8];
          float r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = __float2bfloat16_rn(v[i]);
            r[i] = v[i] - __bfloat162float(h[i]);
          }
          hi.x = pack_bf16(__bfloat162float(h[0]), __bfloat162float(h[1]));
          hi.y = pack_bf16(__bfloat162float(h[2]), __bfloat162float(h[3]));
          hi.z = pack_bf16(__bfloat162float(h[4]), __bfloat162float(h[5]));
          hi.w = pack_bf16(__bfloat162float(h[6]), __bfloat162float(h[7]));
          lo.x = pack_bf16(r[0], r[1]);
          lo.y = pack_bf16(r[2], r[3]);
          lo.z = pack_bf16(r[4], r[5]);
          lo.w = pack_bf16(r[6], r[7]);
        }
        const uint32_t off = (uint32_t)row * 128u + (((uint32_t)c8 ^ sw) << 4);  // 128-byte swizzle: chunk ^= row % 8
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        if (split) *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[s]);
    }
    // =============================================================== EPILOGUE (TMEM -> registers -> NCHW global)
    mbar_wait(accum_bar, 0u);
    tc_fence_after();
    const uint32_t taddr_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col0 = 0; col0 < bn; col0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr_row + (uint32_t)col0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (pix_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int oc = g * d.opg + n0 + col0 + i;
          out[((size_t)b * d.Cout + oc) * d.HoWo + p] = __uint_as_float(r[i]) + (bias ? __ldg(bias + oc) : 0.f);
        }
      }
    }
  } else if (warp == 4) {
    // =============================================================== WEIGHT loader (B tiles)
    for (int j = 0; j < nchunk; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (uint32_t)((j / kStages) & 1);
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* b_hi = smem + s * kStageBytes + 2 * kTileBytesA;
      uint8_t* b_lo = b_hi + kTileBytesB;
      const size_t src0 = ((size_t)(g * d.opg + n0)) * d.K + (size_t)j * BK;
      const int c = lane & 7;
      for (int r = lane >> 3; r < bn; r += 4) {
        const size_t src = src0 + (size_t)r * d.K + c * 8;
        const uint32_t off = (uint32_t)r * 128u + (((uint32_t)c ^ (uint32_t)(r & 7)) << 4);
        *reinterpret_cast<uint4*>(b_hi + off) = __ldg(reinterpret_cast<const uint4*>(whi + src));
        if (split) *reinterpret_cast<uint4*>(b_lo + off) = __ldg(reinterpret_cast<const uint4*>(wlo + src));
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
    }
  } else {
    // =============================================================== MMA issuer (warp 5, one elected lane)
    // instruction descriptor: D = f32, A = B = bf16, both K-major, N = bn, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int j = 0; j < nchunk; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (uint32_t)((j / kStages) & 1);
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + s * kStageBytes);
        const uint32_t a_lo = a_hi + kTileBytesA;
        const uint32_t b_hi = a_hi + 2 * kTileBytesA;
        const uint32_t b_lo = b_hi + kTileBytesB;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t koff = (uint32_t)k * 32u;  // 16 bf16 = 32 bytes along K inside the swizzled row
          umma_bf16(tmem_base, umma_desc(a_hi + koff), umma_desc(b_hi + koff), idesc, (j > 0 || k > 0) ? 1u : 0u);
          if (split) {
            umma_bf16(tmem_base, umma_desc(a_hi + koff), umma_desc(b_lo + koff), idesc, 1u);
            umma_bf16(tmem_base, umma_desc(a_lo + koff), umma_desc(b_hi + koff), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);                    // frees the smem stage once these MMAs have read it
        if (j == nchunk - 1) umma_commit(accum_bar);   // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

bool make_tc_dims(const d2b_dcn_params* p, TcDims& d) {
  if (!p) return false;
  d.N = p->N; d.Cin = p->Cin; d.H = p->H; d.W = p->W; d.Cout = p->Cout; d.kh = p->kh; d.kw = p->kw;
  d.sh = p->stride_h; d.sw = p->stride_w; d.ph = p->pad_h; d.pw = p->pad_w; d.dh = p->dil_h; d.dw = p->dil_w;
  d.G = p->groups; d.DG = p->deformable_groups;
  if (d.N < 0 || d.Cin <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0 || d.kh <= 0 || d.kw <= 0 || d.sh <= 0 || d.sw <= 0 ||
      d.ph < 0 || d.pw < 0 || d.dh <= 0 || d.dw <= 0 || d.G <= 0 || d.DG <= 0)
    return false;
  if (d.Cin % d.G || d.Cout % d.G || d.Cin % d.DG) return false;
  d.Ho = (d.H + 2 * d.ph - (d.dh * (d.kh - 1) + 1)) / d.sh + 1;
  d.Wo = (d.W + 2 * d.pw - (d.dw * (d.kw - 1) + 1)) / d.sw + 1;
  if (d.Ho <= 0 || d.Wo <= 0) return false;
  d.cpg = d.Cin / d.G; d.opg = d.Cout / d.G; d.cpdg = d.Cin / d.DG; d.KK = d.kh * d.kw; d.HoWo = d.Ho * d.Wo;
  d.K = d.cpg * d.KK;
  return true;
}

}  // namespace

// shapes the tensor-core kernel takes
int d2b_deform_conv_tc_supported(const d2b_dcn_params* p) {
  TcDims d;
  if (!make_tc_dims(p, d)) return 0;
  return (d.cpg % BK == 0) && (d.cpdg % BK == 0) && (d.opg % 16 == 0) && ((size_t)d.HoWo * d.Cout < (1ull << 31));
}

size_t d2b_deform_conv_tc_workspace_bytes(const d2b_dcn_params* p) {
  TcDims d;
  if (!make_tc_dims(p, d)) return 0;
  return 2 * ((sizeof(__nv_bfloat16) * (size_t)d.Cout * d.K + 255) & ~(size_t)255);
}

int d2b_deform_conv_forward_tc(const float* x, const float* offset, const float* mask, const float* weight,
                               const float* bias, const d2b_dcn_params* p, int precision, float* out, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcDims d;
  if (!make_tc_dims(p, d)) return D2B_EINVAL;
  if (!d2b_deform_conv_tc_supported(p)) return D2B_EUNSUPPORTED;
  if (d.N == 0) return D2B_OK;
  if (!workspace || workspace_bytes < d2b_deform_conv_tc_workspace_bytes(p)) return D2B_EWORKSPACE;
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>((char*)workspace + d2b_deform_conv_tc_workspace_bytes(p) / 2);
  const int total = d.Cout * d.K;
  dcn_weight_split_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(weight, d.Cout, d.cpg, d.KK, whi, wlo);
  D2B_CHECK_LAUNCH();
  static bool attr_set = false;
  if (!attr_set) {
    D2B_CUDA(cudaFuncSetAttribute(dcn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  dim3 grid(d2b_cdiv(d.HoWo, BM), d2b_cdiv(d.opg, BNMAX), d.N * d.G);
  dcn_fwd_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(x, offset, mask, whi, wlo, bias, d, precision == 1 ? 1 : 0, out);
  D2B_CHECK_LAUNCH();
  return D2B_OK;
}
