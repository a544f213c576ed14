// Deformable convolution v1/v2 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a: forward, backward-data
// and backward-weight.  Replaces detectron2/layers/csrc/deformable/deform_conv_cuda.cu:272-444 (forward),
// :446-642 / :985-1221 (backward input + offset + mask) and :644-824 (backward filter).
//
// The reference materialises columns[Cin*kh*kw, Ho*Wo] in HBM with a gather kernel and calls cuBLAS per group; its
// backward materialises grad_columns, runs col2im / col2im_coord over them and re-runs im2col for the filter gradient.
// Here no column buffer exists: every kernel is an implicit GEMM on tcgen05 whose gathered operand is produced (or whose
// result is consumed) on the SM, with the operands that are plain matrices pre-tiled once into the UMMA shared-memory
// image so that ONE linear TMA copy (cp.async.bulk) brings a whole operand tile.
//
//   K1 forward      D[128 px, oc]   = col[128 px, k'] . W[oc, k']^T        A = gather (K-major), B = weight tiles (TMA)
//   K2 bwd data     gcol[128 px,k'] = gout[128 px, oc] . W[oc, k']          A = gout tiles (TMA), B = W^T tiles (TMA);
//                   epilogue: gcol -> red.global.add.v4 into grad_x (NHWC), grad_offset, grad_mask
//   K3 bwd weight   gW[k', oc]      = col^T[k', px] . gout[px, oc]          A = gather (MN-major), B = gout tiles (TMA)
//
// Layout decisions
//   * x is gathered from channels-last storage [N,H,W,C] fp32: the 64 channels of a tap are 256 contiguous bytes, a
//     half-warp reads them as 16 x LDG.128, a warp keeps 8 such loads in flight per lane.  NCHW inputs are re-laid out once
//     per call by nchw_to_nhwc_kernel (roi_align.cu); channels_last inputs are used in place.
//   * k' is ordered (kernel point, 64-channel block): one "unit" = 64 k' = one 128-byte swizzle row of bf16, so the
//     bilinear taps of a (pixel, kernel point) are computed once (tap table in shared memory) and reused by all channels.
//   * grouped convolutions with fewer than 64 channels per group are packed into "super-groups" of 64 input channels
//     with block-diagonal (zero-padded) weight tiles: the gather stays 256 bytes per tap and the wasted MMA flops are free.
//   * precision 1 ("bf16x3"): operands are split x = hi + lo (two bf16) and hi*hi + hi*lo + lo*hi is accumulated in fp32 --
//     fp32-class accuracy (<= 1e-4 rel) at a third of the bf16 tensor peak.  precision 2: plain bf16 operands.
//
// Warp roles (576 threads, one CTA per SM): warps 0-15 gather / scatter / epilogue, warp 16 TMA producer (one lane),
// warp 17 MMA issuer (one lane) and TMEM owner.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

using namespace d2b_tc;

namespace {

// 16 worker warps (112 registers each) measured faster than 8 warps with twice the loads in flight per lane (K1 res3 79 vs
// 94 us, K2 132 vs 141 us): the gather / scatter loops are bound by instruction issue and need the warps, not deeper queues.
constexpr int kWorkerWarps = 16;
constexpr int kWorkers = kWorkerWarps * 32;  // 512
constexpr int kThreads = kWorkers + 64;      // + producer warp + MMA warp
constexpr int kRowsPerHalf = 128 / (2 * kWorkerWarps);  // pixel rows of a 128-row tile owned by one half-warp: 4
constexpr int kTile = 16384;                 // [128 rows][128 B]
constexpr int kMaxSmem = 227 * 1024;
constexpr int kGcolPitch = 132;              // floats per pixel row of the drained gcol tile (128 + 4: conflict-free float4)

struct TC {
  int N, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, G, DG, Ho, Wo, cpg, opg, cpdg, KK, HoWo;
  int pack, SG, cps, ops, cbs, U;  // super-groups: cps input / ops output channels each, cbs 64-channel blocks, U units
  int tiles_img;                   // 128-pixel tiles per image
  int stages_img;                  // 64-pixel stages per image (K3)
  int MC;                          // macro-chunks (pairs of units) per super-group
  int nks;                         // 64-wide K stages over the super-group's output channels (K2)
  // offset / mask addressing: element strides between images, element offset of the mask block relative to `mask`
  // pointer (fused layout: offset and mask logits live in ONE [N, 3*DG*KK, Ho, Wo] tensor, backbone/resnet.py:307-312)
  long long off_bs, mask_bs;
  int mask_sigmoid;                // mask values are logits: sigmoid applied while the taps are built
};

// optional epilogue of the forward (and its transpose in the backward): y = relu(acc * scale[oc] + shift[oc]) --
// the FrozenBatchNorm / bias + ReLU that follow conv2 of a DeformBottleneckBlock (backbone/resnet.py:313-318)
struct Epi {
  const float* scale;  // may be null (1)
  const float* shift;  // may be null (0)
  int relu;
};

struct K1P { int BN, noct, gspan, nkp, ksplit, S, tap_bytes, stage_bytes, red; };  // red: k-split partial sums meet through red.add
struct K2P { int mper, msplit, tap_bytes; };
struct K3P { int BN, noct, sper, nsplit, S, stage_bytes; };

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

bool make_tc(const d2b_dcn_params* p, TC& d) {
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
  // ---- shapes the tensor-core kernels take
  if (d.cpg >= 64) {
    if (d.cpg % 64) return false;
    d.pack = 1;
  } else {
    if (d.cpg != 16 && d.cpg != 32) return false;
    d.pack = 64 / d.cpg;
    if (d.G % d.pack) return false;
  }
  if (d.cpdg % 64) return false;  // a 64-channel block never straddles deformable groups
  d.SG = d.G / d.pack; d.cps = d.cpg * d.pack; d.ops = d.opg * d.pack; d.cbs = d.cps / 64; d.U = d.KK * d.cbs;
  if (d.ops % 16) return false;
  if (d.KK > 49) return false;
  if ((long long)d.H * d.W * d.Cin >= (1LL << 29)) return false;  // 32-bit byte offsets inside one image
  if ((long long)d.HoWo * d.Cout >= (1LL << 31) || (long long)d.Cout * d.cpg * d.KK >= (1LL << 31)) return false;
  d.tiles_img = d2b_cdiv(d.HoWo, 128);
  d.stages_img = d2b_cdiv(d.HoWo, 64);
  d.MC = (d.U + 1) / 2;
  d.nks = d.ops / 64;
  if ((long long)d.N * d.tiles_img > 0x7fffffffLL) return false;
  d.off_bs = (long long)d.DG * 2 * d.KK * d.HoWo;
  d.mask_bs = (long long)d.DG * d.KK * d.HoWo;
  d.mask_sigmoid = 0;
  return true;
}

// fused offset+mask tensor: returns the mask pointer inside it and switches the strides / sigmoid on
const float* use_fused_offset_mask(TC& d, const float* offset_mask) {
  d.off_bs = d.mask_bs = (long long)d.DG * 3 * d.KK * d.HoWo;
  d.mask_sigmoid = 1;
  return offset_mask + (size_t)d.DG * 2 * d.KK * d.HoWo;
}

int largest_tile(int n) {  // largest of {256,...,16} dividing n
  for (int t = 256; t >= 16; t >>= 1)
    if (n % t == 0) return t;
  return 0;
}

// Split `work` units of a CTA `base` times replicated over up to `max_split` parts: the number of parts that minimises
// (waves over the 148 SMs) x (units per CTA + the CTA's fixed prologue / epilogue cost in units).  One CTA per SM is
// resident, so 153 CTAs cost two full waves.
int best_split(long long base, int work, int max_split, int overhead) {
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= max_split; ++s) {
    const int per = d2b_cdiv(work, s);
    const int parts = d2b_cdiv(work, per);
    const long long cost = (long long)d2b_cdiv(base * parts, kNumSMs) * (per + overhead);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = parts;
    }
  }
  return best;
}

int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

bool plan_k1(const TC& d, K1P& k) {
  if (d.ops <= 256) {
    k.BN = d.ops; k.noct = 1;
    k.gspan = std::max(1, std::min(d.SG, 256 / d.ops));
    while (d.SG % k.gspan) --k.gspan;
  } else {
    k.BN = largest_tile(d.ops); k.noct = d.ops / k.BN; k.gspan = 1;
  }
  if (k.BN < 16) return false;
  const int base = d.N * d.tiles_img * (d.SG / k.gspan) * k.noct;
  k.ksplit = best_split(base, d.KK, d.KK, 1);
  k.nkp = d2b_cdiv(d.KK, k.ksplit);
  k.ksplit = d2b_cdiv(d.KK, k.nkp);
  k.red = k.ksplit > 1;
  const int ndg = d.DG == 1 ? 1 : std::min(d.DG, d2b_cdiv(k.gspan * d.cps, d.cpdg) + 1);
  k.tap_bytes = ndg * k.nkp * 128 * 16;
  k.stage_bytes = 2 * kTile + 2 * k.BN * 128;
  k.S = 3;
  while (k.S > 1 && k.S * k.stage_bytes + k.tap_bytes + 1024 + 256 > kMaxSmem) --k.S;
  return k.S >= 2;
}

bool plan_k2(const TC& d, K2P& k) {
  if (d.ops % 64) return false;
  const int base = d.N * d.tiles_img * d.SG;
  k.msplit = best_split(base, d.MC, d.MC, 1);
  k.mper = d2b_cdiv(d.MC, k.msplit);
  k.msplit = d2b_cdiv(d.MC, k.mper);
  const int nkp = std::min(d.KK, (2 * k.mper + d.cbs - 1) / d.cbs + 1);
  const int ndg = d.DG == 1 ? 1 : std::min(d.DG, d2b_cdiv(d.cps, d.cpdg) + 1);
  k.tap_bytes = ndg * nkp * 128 * 16;
  return 2 * 4 * kTile + 128 * kGcolPitch * 4 + k.tap_bytes + 1024 + 256 <= kMaxSmem;
}

bool plan_k3(const TC& d, K3P& k) {
  k.BN = largest_tile(d.ops);
  if (k.BN < 16) return false;
  k.noct = d.ops / k.BN;
  const int units = d.MC * d.SG * k.noct;
  const int total = d.N * d.stages_img;
  k.nsplit = best_split(units, total, std::min(total, 4 * kNumSMs), 3);
  k.sper = d2b_cdiv(total, k.nsplit);
  k.nsplit = d2b_cdiv(total, k.sper);
  k.stage_bytes = 2 * kTile + 2 * k.BN * 128;
  k.S = 3;
  while (k.S > 1 && k.S * k.stage_bytes + 4096 + 1024 + 256 > kMaxSmem) --k.S;
  return k.S >= 2;
}

// ------------------------------------------------------------------------------------------------ sampling taps
// One 16-byte entry per (deformable group, kernel point, pixel): {code, lh, lw, mask} with
// code = ((pos0 + W + 1) << 4) | valid-corner bits, pos0 = floor(h)*W + floor(w) (may be "virtual": row/col -1).
// An entry of zeros means "sample outside (-1,H)x(-1,W)": contributes nothing (deform_conv_cuda_kernel.cu:273).
struct TapRaw {
  float oh, ow, m;
};

// the three global loads of a tap (issued early so that their latency overlaps other work)
__device__ __forceinline__ TapRaw tap_loads(const TC& d, const float* __restrict__ offset, const float* __restrict__ mask,
                                            int b, int dg, int kp, int p) {
  TapRaw r = {0.f, 0.f, 1.f};
  if (p < d.HoWo) {
    const size_t ob = (size_t)b * d.off_bs + ((size_t)dg * 2 * d.KK) * d.HoWo;
    r.oh = __ldg(offset + ob + (size_t)(2 * kp) * d.HoWo + p);       // deform_conv_cuda_kernel.cu:263-269
    r.ow = __ldg(offset + ob + (size_t)(2 * kp + 1) * d.HoWo + p);
    if (mask) r.m = __ldg(mask + (size_t)b * d.mask_bs + ((size_t)dg * d.KK + kp) * d.HoWo + p);
  }
  return r;
}

__device__ __forceinline__ int4 tap_finish(const TC& d, const TapRaw& r, bool has_mask, int kp, int p) {
  int4 t = make_int4(0, 0, 0, 0);
  if (p < d.HoWo) {
    const int ho = p / d.Wo, wo = p - ho * d.Wo;
    const int ki = kp / d.kw, kj = kp - ki * d.kw;
    const float hf = (float)(ho * d.sh - d.ph + ki * d.dh) + r.oh;
    const float wf = (float)(wo * d.sw - d.pw + kj * d.dw) + r.ow;
    float m = r.m;
    if (has_mask && d.mask_sigmoid) m = 1.f / (1.f + expf(-m));  // resnet.py:311 mask.sigmoid()
    if (hf > -1.f && wf > -1.f && hf < (float)d.H && wf < (float)d.W) {
      const float hfl = floorf(hf), wfl = floorf(wf);
      const int hl = (int)hfl, wl = (int)wfl;
      const bool t0 = hl >= 0, t1 = hl + 1 <= d.H - 1, l0 = wl >= 0, l1 = wl + 1 <= d.W - 1;
      const int flags = (t0 && l0 ? 1 : 0) | (t0 && l1 ? 2 : 0) | (t1 && l0 ? 4 : 0) | (t1 && l1 ? 8 : 0);
      t.x = ((hl * d.W + wl + d.W + 1) << 4) | flags;
      t.y = __float_as_int(hf - hfl);
      t.z = __float_as_int(wf - wfl);
      t.w = __float_as_int(m);
    }
  }
  return t;
}

__device__ __forceinline__ int4 make_tap(const TC& d, const float* __restrict__ offset, const float* __restrict__ mask,
                                         int b, int dg, int kp, int p) {
  return tap_finish(d, tap_loads(d, offset, mask, b, dg, kp, p), mask != nullptr, kp, p);
}

struct Taps4 {
  float4 v0, v1, v2, v3;
};

// the four corner pixels (4 channels each) of one tap entry; corners outside the image read as 0.
// `xcb` is the channel pointer biased by -(W + 1) pixels so that the tap's (pos0 + W + 1) indexes it directly; all offsets
// are 32-bit element counts (H * W * Cin < 2^30 is part of the shape gate), `rs` = W * Cin.
__device__ __forceinline__ void load_corners(const float* __restrict__ xcb, uint32_t Cin, uint32_t rs, int code, Taps4& t) {
  const float* __restrict__ p = xcb + (uint32_t)(code >> 4) * Cin;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  t.v0 = (code & 1) ? __ldg(reinterpret_cast<const float4*>(p)) : z;
  t.v1 = (code & 2) ? __ldg(reinterpret_cast<const float4*>(p + Cin)) : z;
  t.v2 = (code & 4) ? __ldg(reinterpret_cast<const float4*>(p + rs)) : z;
  t.v3 = (code & 8) ? __ldg(reinterpret_cast<const float4*>(p + (rs + Cin))) : z;
}

__device__ __forceinline__ void interp4(const Taps4& t, float w0, float w1, float w2, float w3, float (&v)[4]) {
  const F2 z = f2_pack(0.f, 0.f), p0 = f2_pack(w0, w0), p1 = f2_pack(w1, w1), p2 = f2_pack(w2, w2), p3 = f2_pack(w3, w3);
  F2 a = f2_fma(p0, f2_pack(t.v0.x, t.v0.y), z), b = f2_fma(p0, f2_pack(t.v0.z, t.v0.w), z);
  a = f2_fma(p1, f2_pack(t.v1.x, t.v1.y), a);
  b = f2_fma(p1, f2_pack(t.v1.z, t.v1.w), b);
  a = f2_fma(p2, f2_pack(t.v2.x, t.v2.y), a);
  b = f2_fma(p2, f2_pack(t.v2.z, t.v2.w), b);
  a = f2_fma(p3, f2_pack(t.v3.x, t.v3.y), a);
  b = f2_fma(p3, f2_pack(t.v3.z, t.v3.w), b);
  f2_unpack(a, v[0], v[1]);
  f2_unpack(b, v[2], v[3]);
}

// gather 4 channels of one (pixel, unit) and store them as bf16 hi / lo into a swizzled 128-byte row
__device__ __forceinline__ void gather_store(const Taps4& t, int4 tap, uint8_t* a_hi, uint8_t* a_lo, uint32_t off,
                                             bool split) {
  const float lh = __int_as_float(tap.y), lw = __int_as_float(tap.z), m = __int_as_float(tap.w);
  const float hh = 1.f - lh, hw = 1.f - lw;
  float v[4];
  interp4(t, hh * hw * m, hh * lw * m, lh * hw * m, lh * lw * m, v);
  uint2 hi, lo;
  split4(v, hi, lo);
  *reinterpret_cast<uint2*>(a_hi + off) = hi;
  if (split) *reinterpret_cast<uint2*>(a_lo + off) = lo;
}

// ================================================================================================ K1: forward
// grid (N * tiles_img, spans * oc tiles, k splits).  One more warp than K2 / K3: the SAVER, which (when the caller keeps the
// sampled columns for the weight gradient) copies every finished A stage -- already bf16 hi | lo in the tensor core's
// swizzled tile layout -- to global memory with one bulk store, so that the backward streams the tiles back instead of
// sampling x a second time (dcn_bwd_weight_cols_kernel).  Column tile of (image tile, super-group, unit): 16 KB hi [+ 16 KB lo].
constexpr int kThreadsK1 = kThreads + 32;
template <int kTmemCols>
__global__ void __launch_bounds__(kThreadsK1, 1) dcn_fwd_tc_kernel(const float* __restrict__ xh,
                                                                   const float* __restrict__ offset,
                                                                   const float* __restrict__ mask,
                                                                   const uint8_t* __restrict__ wt, const Epi ep, const TC d,
                                                                   const K1P k, const int split, float* __restrict__ out,
                                                                   uint8_t* __restrict__ cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset form keeps the shared address space
  int4* taps = reinterpret_cast<int4*>(smem + k.S * k.stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(taps) + k.tap_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + k.S;
  uint64_t* accum_bar = bars + 2 * k.S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k.S + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / d.tiles_img, p0 = (blockIdx.x - b * d.tiles_img) * 128;
  const int span = blockIdx.y / k.noct, oct = blockIdx.y - span * k.noct;
  const int sg0 = span * k.gspan;
  const int kp0 = blockIdx.z * k.nkp, nkp = min(d.KK, kp0 + k.nkp) - kp0;
  const int nt = k.gspan * nkp * d.cbs;
  const int dg0 = (sg0 * d.cps) / d.cpdg;
  const bool saving = cols != nullptr && oct == 0;  // CTAs of the other output-channel tiles build the same columns

  if (tid == 0) {
    for (int s = 0; s < k.S; ++s) {
      mbar_init(&full_bar[s], kWorkerWarps + 1);
      mbar_init(&empty_bar[s], saving ? 2 : 1);  // the tensor core is done with the stage [+ the saver has copied it out]
    }
    mbar_init(accum_bar, 1);
    mbar_init_fence();
  }
  if (warp == kWorkerWarps + 1) tmem_alloc<kTmemCols>(tmem_slot);
  if (tid < kWorkers) {
    const int ndg = ((sg0 + k.gspan) * d.cps - 1) / d.cpdg - dg0 + 1;
    for (int i = tid; i < ndg * nkp * 128; i += kWorkers) {
      const int row = i & 127, r = i >> 7;
      const int dgl = r / nkp, kpl = r - dgl * nkp;
      taps[i] = make_tap(d, offset, mask, b, dg0 + dgl, kp0 + kpl, p0 + row);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWorkerWarps) {
    // =============================================================== GATHER: A tile [128 px][64 k'] K-major, hi + lo
    const int half = lane >> 4, q = lane & 15;
    const float* __restrict__ ximg = xh + (size_t)b * d.H * d.W * d.Cin - (size_t)(d.W + 1) * d.Cin;  // biased: load_corners
    const uint32_t ucin = (uint32_t)d.Cin, urs = (uint32_t)(d.W * d.Cin);
    int sgl = 0, kpl = 0, cb = 0;
    for (int t = 0; t < nt; ++t) {
      const int s = t % k.S;
      const uint32_t par = (uint32_t)((t / k.S) & 1);
      const int cbase = (sg0 + sgl) * d.cps + cb * 64;
      const int4* __restrict__ tp = taps + ((cbase / d.cpdg - dg0) * nkp + kpl) * 128;
      const float* __restrict__ xc = ximg + cbase + q * 4;
      uint8_t* a_hi = smem + s * k.stage_bytes;
      uint8_t* a_lo = a_hi + kTile;
      mbar_wait(&empty_bar[s], par ^ 1u);
      {  // the half-warp's pixel rows in ONE batch: all their corner loads are in flight before the first use
        int4 tap[kRowsPerHalf];
        Taps4 c[kRowsPerHalf];
#pragma unroll
        for (int j = 0; j < kRowsPerHalf; ++j) tap[j] = tp[warp * 2 * kRowsPerHalf + j * 2 + half];
#pragma unroll
        for (int j = 0; j < kRowsPerHalf; ++j) load_corners(xc, ucin, urs, tap[j].x, c[j]);
#pragma unroll
        for (int j = 0; j < kRowsPerHalf; ++j) {
          const uint32_t r = (uint32_t)(warp * 2 * kRowsPerHalf + j * 2 + half);
          gather_store(c[j], tap[j], a_hi, a_lo, swz128(r, (uint32_t)(q >> 1)) + (uint32_t)(q & 1) * 8u, split != 0);
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (++cb == d.cbs) {
        cb = 0;
        if (++kpl == nkp) { kpl = 0; ++sgl; }
      }
    }
    // =============================================================== EPILOGUE: TMEM -> registers -> NCHW global
    mbar_wait(accum_bar, 0u);
    tc_fence_after();
    const int quad = warp & 3, cgrp = warp >> 2;
    const int row = quad * 32 + lane, p = p0 + row;
    const bool pix_ok = p < d.HoWo;
    const int ntot = k.gspan * k.BN;
    const int oc0 = sg0 * d.ops + oct * k.BN;
    const bool add_shift = ep.shift != nullptr && blockIdx.z == 0;
    for (int c16 = cgrp; c16 * 16 < ntot; c16 += kWorkerWarps / 4) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c16 * 16), r);
      tmem_ld_wait();
      if (pix_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int oc = oc0 + c16 * 16 + i;
          float* dst = out + ((size_t)b * d.Cout + oc) * d.HoWo + p;
          float v = __uint_as_float(r[i]);
          if (k.red) {  // k-split partial sums: scale / relu run in dcn_epilogue_kernel once all partials are in
            red_add(dst, v + ((add_shift && !ep.scale && !ep.relu) ? __ldg(ep.shift + oc) : 0.f));
          } else {
            if (ep.scale) v *= __ldg(ep.scale + oc);
            if (ep.shift) v += __ldg(ep.shift + oc);
            if (ep.relu) v = fmaxf(v, 0.f);
            *dst = v;
          }
        }
      }
    }
  } else if (warp == kWorkerWarps) {
    // =============================================================== TMA producer: weight tile of every chunk
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(split ? 2 : 1) * (uint32_t)k.BN * 128u;
      int sgl = 0, kpl = 0, cb = 0;
      for (int t = 0; t < nt; ++t) {
        const int s = t % k.S;
        const uint32_t par = (uint32_t)((t / k.S) & 1);
        mbar_wait(&empty_bar[s], par ^ 1u);
        const size_t tile = ((size_t)((sg0 + sgl) * k.noct + oct) * d.U + (size_t)(kp0 + kpl) * d.cbs + cb);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(smem + s * k.stage_bytes + 2 * kTile, wt + tile * (size_t)(2 * k.BN * 128), bytes, &full_bar[s]);
        if (++cb == d.cbs) {
          cb = 0;
          if (++kpl == nkp) { kpl = 0; ++sgl; }
        }
      }
    }
  } else if (warp == kWorkerWarps + 2) {
    // =============================================================== SAVER: finished A stages -> column tiles in global memory
    if (saving && lane == 0) {
      const uint32_t tile_bytes = (uint32_t)(split ? 2 : 1) * (uint32_t)kTile;
      int sgl = 0, kpl = 0, cb = 0;
      for (int t = 0; t < nt; ++t) {
        const int s = t % k.S;
        const uint32_t par = (uint32_t)((t / k.S) & 1);
        mbar_wait(&full_bar[s], par);  // the workers' writes are fenced to the async proxy before they arrive
        const size_t tile = ((size_t)blockIdx.x * d.SG + (sg0 + sgl)) * d.U + (size_t)(kp0 + kpl) * d.cbs + cb;
        bulk_s2g(cols + tile * tile_bytes, smem + s * k.stage_bytes, tile_bytes);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&empty_bar[s]);
        if (++cb == d.cbs) {
          cb = 0;
          if (++kpl == nkp) { kpl = 0; ++sgl; }
        }
      }
      bulk_wait0();
    }
  } else {
    // =============================================================== MMA issuer
    const uint32_t idesc = umma_idesc(k.BN, false, false);
    const int per_sg = nkp * d.cbs;
    for (int t = 0; t < nt; ++t) {
      const int s = t % k.S;
      const uint32_t par = (uint32_t)((t / k.S) & 1);
      mbar_wait(&full_bar[s], par);
      tc_fence_after();
      if (lane == 0) {
        const int sgl = t / per_sg;
        const bool first = (t - sgl * per_sg) == 0;
        const uint32_t a_hi = smem_u32(smem + s * k.stage_bytes), a_lo = a_hi + kTile;
        const uint32_t b_hi = a_hi + 2 * kTile, b_lo = b_hi + (uint32_t)k.BN * 128u;
        const uint32_t dcol = tmem_base + (uint32_t)(sgl * k.BN);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = (uint32_t)kk * 32u;  // 16 bf16 along K inside the swizzled row
          umma_bf16(dcol, umma_desc(a_hi + ko, 0, 1024), umma_desc(b_hi + ko, 0, 1024), idesc, (!first || kk > 0) ? 1u : 0u);
          if (split) {
            umma_bf16(dcol, umma_desc(a_hi + ko, 0, 1024), umma_desc(b_lo + ko, 0, 1024), idesc, 1u);
            umma_bf16(dcol, umma_desc(a_lo + ko, 0, 1024), umma_desc(b_hi + ko, 0, 1024), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (t == nt - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkerWarps + 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ================================================================================================ K2: backward data
// grid (N * tiles_img, SG, macro-chunk splits).  Per macro-chunk (2 units = 128 k'): gcol[128 px][128] = gout . W over the
// super-group's output channels (TMEM, double buffered), drained to shared memory, then scattered:
//   grad_x   += gcol * mask * bilinear weight      (deform_conv_cuda_kernel.cu:313-362, :923-975) red.global.add.v4 (NHWC)
//   grad_off += gcol * mask * d(sample)/d(h,w)     (:390-451, :977-1064)                           red.global.add
//   grad_msk += gcol * sample                      (:1053-1064)
__global__ void __launch_bounds__(kThreads, 1) dcn_bwd_data_tc_kernel(const float* __restrict__ xh,
                                                                      const float* __restrict__ offset,
                                                                      const float* __restrict__ mask,
                                                                      const uint8_t* __restrict__ gt,
                                                                      const uint8_t* __restrict__ wt, const TC d,
                                                                      const K2P k, const int split,
                                                                      float* __restrict__ gxh, float* __restrict__ goff,
                                                                      float* __restrict__ gmask) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset form keeps the shared address space
  constexpr int kStage = 4 * kTile;  // A hi | A lo | B hi | B lo
  float* gcol = reinterpret_cast<float*>(smem + 2 * kStage);
  int4* taps = reinterpret_cast<int4*>(smem + 2 * kStage + 128 * kGcolPitch * 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(taps) + k.tap_bytes);
  uint64_t* full_bar = bars;         // [2]
  uint64_t* empty_bar = bars + 2;    // [2]
  uint64_t* acc_full = bars + 4;     // [2]
  uint64_t* acc_empty = bars + 6;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / d.tiles_img, pt = blockIdx.x - b * d.tiles_img, p0 = pt * 128;
  const int sg = blockIdx.y;
  const int m0 = blockIdx.z * k.mper, m1 = min(d.MC, m0 + k.mper), nm = m1 - m0;
  const int u_begin = 2 * m0, u_end = min(d.U, 2 * m1);
  const int kp0 = u_begin / d.cbs, nkp = (u_end - 1) / d.cbs - kp0 + 1;
  const int dg0 = (sg * d.cps) / d.cpdg;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kWorkerWarps);
    }
    mbar_init_fence();
  }
  if (warp == kWorkerWarps + 1) tmem_alloc<256>(tmem_slot);
  if (tid < kWorkers) {
    const int ndg = ((sg + 1) * d.cps - 1) / d.cpdg - dg0 + 1;
    for (int i = tid; i < ndg * nkp * 128; i += kWorkers) {
      const int row = i & 127, r = i >> 7;
      const int dgl = r / nkp, kpl = r - dgl * nkp;
      taps[i] = make_tap(d, offset, mask, b, dg0 + dgl, kp0 + kpl, p0 + row);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWorkerWarps) {
    const int half = lane >> 4, q = lane & 15;
    const int quad = warp & 3, cq = warp >> 2;
    // both image pointers are biased by -(W + 1) pixels (see load_corners)
    const size_t img_off = (size_t)b * d.H * d.W * d.Cin - (size_t)(d.W + 1) * d.Cin;
    const float* __restrict__ ximg = xh + img_off;
    float* __restrict__ gimg = gxh ? gxh + img_off : nullptr;
    const uint32_t ucin = (uint32_t)d.Cin, urs = (uint32_t)(d.W * d.Cin);
    constexpr int R = kRowsPerHalf;  // pixel rows of the tile owned by this half-warp (fixed for the whole kernel)
    float sh[R], sw[R], sm[R];
#pragma unroll
    for (int j = 0; j < R; ++j) sh[j] = sw[j] = sm[j] = 0.f;
    int cur_kp = -1, cur_dg = -1;

    auto flush = [&]() {
      if (cur_kp < 0) return;
      // 12 partial sums (4 rows x {d/dh, d/dw, d/dmask}) per lane, to be summed over the 16 lanes of the half-warp.
      // Recursive halving: at each step a lane keeps one half of its values and hands the other half to its partner,
      // so 6 + 3 + 2 + 1 shuffles replace 12 butterflies of 4; lane q ends up owning value (row = q >> 2, which = q & 3).
      static_assert(R == 4, "the reduction below is written for 4 rows per half-warp");
      float v12[12];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        v12[3 * j] = sh[j];
        v12[3 * j + 1] = sw[j];
        v12[3 * j + 2] = sm[j];
      }
      const bool b8 = q & 8, b4 = q & 4, b2 = q & 2, b1 = q & 1;
      float v6[6], v3[3];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float keep = b8 ? v12[i + 6] : v12[i], send = b8 ? v12[i] : v12[i + 6];
        v6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float keep = b4 ? v6[i + 3] : v6[i], send = b4 ? v6[i] : v6[i + 3];
        v3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      // {v3[0], v3[1]} | {v3[2], -}
      const float k0 = b2 ? v3[2] : v3[0], s0 = b2 ? v3[0] : v3[2];
      const float k1 = b2 ? 0.f : v3[1], s1 = b2 ? v3[1] : 0.f;
      const float u0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
      const float u1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
      float v = (b1 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, b1 ? u0 : u1, 1);
      {
        const int j = q >> 2, which = q & 3;
        const int row = warp * 2 * R + j * 2 + half;
        const int p = p0 + row;
        if (which < 3 && p < d.HoWo) {
          if (which < 2) {
            if (goff) red_add(goff + (size_t)b * d.off_bs + ((size_t)cur_dg * 2 * d.KK + 2 * cur_kp + which) * d.HoWo + p, v);
          } else if (gmask) {
            if (d.mask_sigmoid) {  // gradient w.r.t. the logit: m (1 - m)
              const float m = __int_as_float(taps[((cur_dg - dg0) * nkp + (cur_kp - kp0)) * 128 + row].w);
              v *= m * (1.f - m);
            }
            red_add(gmask + (size_t)b * d.mask_bs + ((size_t)cur_dg * d.KK + cur_kp) * d.HoWo + p, v);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < R; ++j) sh[j] = sw[j] = sm[j] = 0.f;
    };

    for (int ml = 0; ml < nm; ++ml) {
      const int buf = ml & 1;
      const int m = m0 + ml;
      const int nu = (2 * m + 1 < d.U) ? 2 : 1;
      mbar_wait(&acc_full[buf], (uint32_t)((ml >> 1) & 1));
      tc_fence_after();
      // ---- drain: this warp's 32 TMEM lanes (pixels) x its share of the 128 columns -> gcol[px][col]
      constexpr int kDrainCols = 128 / (kWorkerWarps / 4);
      if (cq * kDrainCols < nu * 64) {
        float* dst = gcol + (quad * 32 + lane) * kGcolPitch + cq * kDrainCols;
#pragma unroll
        for (int h4 = 0; h4 < kDrainCols / 16; ++h4) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 128 + cq * kDrainCols + h4 * 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(dst + h4 * 16 + i * 4) =
                make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                            __uint_as_float(r[4 * i + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      named_bar_sync(1, kWorkers);
      // ---- scatter
      for (int ul = 0; ul < nu; ++ul) {
        const int u = 2 * m + ul;
        const int kp = u / d.cbs, cb = u - kp * d.cbs;
        const int cbase = sg * d.cps + cb * 64;
        const int dg = cbase / d.cpdg;
        if (kp != cur_kp || dg != cur_dg) {
          flush();
          cur_kp = kp;
          cur_dg = dg;
        }
        const int4* __restrict__ tp = taps + ((dg - dg0) * nkp + (kp - kp0)) * 128;
        const float* __restrict__ xc = ximg + cbase + q * 4;
        float* __restrict__ gc = gimg + cbase + q * 4;
        constexpr int kB2 = 2;  // pixel rows per batch: 4 * kB2 loads in flight per lane, then up to 4 * kB2 reductions
#pragma unroll
        for (int it = 0; it < R / kB2; ++it) {
          int4 tap[kB2];
          Taps4 c[kB2];
#pragma unroll
          for (int e = 0; e < kB2; ++e) tap[e] = tp[warp * 2 * R + (it * kB2 + e) * 2 + half];
#pragma unroll
          for (int e = 0; e < kB2; ++e) load_corners(xc, ucin, urs, tap[e].x, c[e]);
#pragma unroll
          for (int e = 0; e < kB2; ++e) {
            if ((tap[e].x & 15) == 0) continue;  // sample outside the image: no gradient anywhere
            const int j = it * kB2 + e;
            const int row = warp * 2 * R + j * 2 + half;
            const float4 g4 = *reinterpret_cast<const float4*>(gcol + row * kGcolPitch + ul * 64 + q * 4);
            const float lh = __int_as_float(tap[e].y), lw = __int_as_float(tap[e].z), mk = __int_as_float(tap[e].w);
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float w0 = hh * hw, w1 = hh * lw, w2 = lh * hw, w3 = lh * lw;
            const F2 LH = f2_pack(lh, lh), LW = f2_pack(lw, lw), HH = f2_pack(hh, hh), HW = f2_pack(hw, hw);
            const F2 W0 = f2_pack(w0, w0), W1 = f2_pack(w1, w1), W2 = f2_pack(w2, w2), W3 = f2_pack(w3, w3);
            const F2 MK = f2_pack(mk, mk);
            // two channels per instruction (packed fp32): pair 0 = channels 0,1; pair 1 = channels 2,3
            const F2 G[2] = {f2_pack(g4.x, g4.y), f2_pack(g4.z, g4.w)};
            const F2 A0[2] = {f2_pack(c[e].v0.x, c[e].v0.y), f2_pack(c[e].v0.z, c[e].v0.w)};
            const F2 A1[2] = {f2_pack(c[e].v1.x, c[e].v1.y), f2_pack(c[e].v1.z, c[e].v1.w)};
            const F2 A2[2] = {f2_pack(c[e].v2.x, c[e].v2.y), f2_pack(c[e].v2.z, c[e].v2.w)};
            const F2 A3[2] = {f2_pack(c[e].v3.x, c[e].v3.y), f2_pack(c[e].v3.z, c[e].v3.w)};
            F2 GM[2];
            F2 th2 = f2_pack(0.f, 0.f), tw2 = th2, tm2 = th2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              GM[i] = f2_mul(G[i], MK);
              // d val / d h = (1-lw)(v2-v0) + lw (v3-v1);  d val / d w = (1-lh)(v1-v0) + lh (v3-v2)
              const F2 dh = f2_fma(LW, f2_sub(A3[i], A1[i]), f2_mul(HW, f2_sub(A2[i], A0[i])));
              const F2 dw = f2_fma(LH, f2_sub(A3[i], A2[i]), f2_mul(HH, f2_sub(A1[i], A0[i])));
              const F2 val = f2_fma(W3, A3[i], f2_fma(W2, A2[i], f2_fma(W1, A1[i], f2_mul(W0, A0[i]))));
              th2 = f2_fma(GM[i], dh, th2);
              tw2 = f2_fma(GM[i], dw, tw2);
              tm2 = f2_fma(G[i], val, tm2);
            }
            {
              float e0, e1;
              f2_unpack(th2, e0, e1);
              sh[j] += e0 + e1;
              f2_unpack(tw2, e0, e1);
              sw[j] += e0 + e1;
              f2_unpack(tm2, e0, e1);
              sm[j] += e0 + e1;
            }
            if (gimg) {
              float* gp = gc + (uint32_t)(tap[e].x >> 4) * ucin;
              red_add_v4_if(tap[e].x & 1, gp, f2_mul(GM[0], W0), f2_mul(GM[1], W0));
              red_add_v4_if(tap[e].x & 2, gp + ucin, f2_mul(GM[0], W1), f2_mul(GM[1], W1));
              red_add_v4_if(tap[e].x & 4, gp + urs, f2_mul(GM[0], W2), f2_mul(GM[1], W2));
              red_add_v4_if(tap[e].x & 8, gp + (urs + ucin), f2_mul(GM[0], W3), f2_mul(GM[1], W3));
            }
          }
        }
      }
      named_bar_sync(1, kWorkers);  // everyone is done reading gcol before the next drain overwrites it
    }
    flush();
  } else if (warp == kWorkerWarps) {
    // =============================================================== TMA producer: gout tile + W^T tile per K stage
    if (lane == 0) {
      const uint32_t half_bytes = (uint32_t)(split ? 2 : 1) * (uint32_t)kTile;
      int c = 0;
      for (int ml = 0; ml < nm; ++ml) {
        const int m = m0 + ml;
        for (int ks = 0; ks < d.nks; ++ks, ++c) {
          const int s = c & 1;
          const uint32_t par = (uint32_t)((c >> 1) & 1);
          mbar_wait(&empty_bar[s], par ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], 2 * half_bytes);
          const size_t atile = (((size_t)b * d.tiles_img + pt) * d.SG + sg) * d.nks + ks;
          const size_t btile = ((size_t)sg * d.MC + m) * d.nks + ks;
          bulk_g2s(smem + s * kStage, gt + atile * (size_t)(2 * kTile), half_bytes, &full_bar[s]);
          bulk_g2s(smem + s * kStage + 2 * kTile, wt + btile * (size_t)(2 * kTile), half_bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // =============================================================== MMA issuer
    int c = 0;
    for (int ml = 0; ml < nm; ++ml) {
      const int buf = ml & 1;
      const int m = m0 + ml;
      const int ncols = (2 * m + 1 < d.U) ? 128 : 64;
      const uint32_t idesc = umma_idesc(ncols, false, false);
      mbar_wait(&acc_empty[buf], (uint32_t)(((ml >> 1) & 1) ^ 1));
      tc_fence_after();
      for (int ks = 0; ks < d.nks; ++ks, ++c) {
        const int s = c & 1;
        const uint32_t par = (uint32_t)((c >> 1) & 1);
        mbar_wait(&full_bar[s], par);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(smem + s * kStage), a_lo = a_hi + kTile, b_hi = a_hi + 2 * kTile, b_lo = b_hi + kTile;
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ko = (uint32_t)kk * 32u;
            umma_bf16(dcol, umma_desc(a_hi + ko, 0, 1024), umma_desc(b_hi + ko, 0, 1024), idesc, (ks > 0 || kk > 0) ? 1u : 0u);
            if (split) {
              umma_bf16(dcol, umma_desc(a_hi + ko, 0, 1024), umma_desc(b_lo + ko, 0, 1024), idesc, 1u);
              umma_bf16(dcol, umma_desc(a_lo + ko, 0, 1024), umma_desc(b_hi + ko, 0, 1024), idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (ks == d.nks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkerWarps + 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ================================================================================================ weight-gradient epilogue
// The accumulator tile of a K3 CTA is [128 k' rows][BN oc]; in the weight tensor [Cout][Cin/G][KK] one unit's 64 rows are
// 36 bytes apart (one kernel point of 64 channels), so writing it directly costs one 32-byte sector atomic per ELEMENT and
// per pixel split.  Instead every CTA adds its tile with 16-byte vector reductions (64-byte row segments) into the zero-filled
// staging matrix `part` [SG][MC*128][ops], and a small second kernel moves it into the weight tensor's layout.
__device__ __forceinline__ void gw_tile_store(uint32_t tmem_base, int quad, int cgrp, int lane, int BN,
                                              float* __restrict__ tile, int pitch) {
  float* __restrict__ prow = tile + (size_t)(quad * 32 + lane) * pitch;
  for (int c16 = cgrp; c16 * 16 < BN; c16 += kWorkerWarps / 4) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c16 * 16), r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i)
      red_add_v4(prow + c16 * 16 + i * 4, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                 __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  }
}

// one thread per (super-group, unit row, oc of the super-group); oc fastest: the reads of `part` are coalesced
__global__ void dcn_gw_reduce_kernel(const float* __restrict__ part, const TC d, float* __restrict__ gw) {
  const long long total = (long long)d.SG * d.U * 64 * d.ops;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ocl = (int)(idx % d.ops);
  const long long r = idx / d.ops;
  const int rowu = (int)(r % ((long long)d.U * 64)), sg = (int)(r / ((long long)d.U * 64));
  const int u = rowu >> 6, ch = rowu & 63;
  const int kp = u / d.cbs, cb = u - kp * d.cbs;
  const int ec = sg * d.cps + cb * 64 + ch, egrp = ec / d.cpg, ecin = ec - egrp * d.cpg;
  const int oc = sg * d.ops + ocl;
  if (oc / d.opg != egrp) return;  // block-diagonal padding of a packed super-group
  gw[((size_t)oc * d.cpg + ecin) * d.KK + kp] = __ldg(part + ((size_t)sg * d.MC * 128 + rowu) * d.ops + ocl);
}

// The same sum with coalesced writes for small kernels (KK <= 9): one block owns (super-group, 16 channels, 16 output
// channels), stages the KK x 16 x 16 sums in shared memory and writes, per output channel, the run of 16 * KK floats that the
// block's channels occupy in the weight tensor.
constexpr int kRedOc = 16, kRedCh = 16;
__global__ void __launch_bounds__(256) dcn_gw_reduce_tile_kernel(const float* __restrict__ part, const TC d,
                                                                  float* __restrict__ gw) {
  __shared__ float tile[9 * kRedCh][kRedOc + 1];
  const int per_sg = d.cps / kRedCh;
  const int sg = blockIdx.x / per_sg, c0 = (blockIdx.x - sg * per_sg) * kRedCh;  // first channel inside the super-group
  const int cb = c0 >> 6, ch0 = c0 & 63;
  const int ocl0 = blockIdx.y * kRedOc;
  const int tid = threadIdx.x;
  {
    const int j = tid & (kRedOc - 1), r0 = tid / kRedOc;  // 16 rows per pass
    for (int row = r0; row < d.KK * kRedCh; row += 256 / kRedOc) {
      const int kp = row / kRedCh, ch = row - kp * kRedCh;
      tile[row][j] = __ldg(part + ((size_t)sg * d.MC * 128 + (size_t)(kp * d.cbs + cb) * 64 + ch0 + ch) * d.ops + ocl0 + j);
    }
  }
  __syncthreads();
  const int ec0 = sg * d.cps + c0;
  for (int i = tid; i < kRedOc * kRedCh * d.KK; i += 256) {
    const int j = i / (kRedCh * d.KK), e = i - j * (kRedCh * d.KK);
    const int ch = e / d.KK, kp = e - ch * d.KK;
    const int oc = sg * d.ops + ocl0 + j, grp = oc / d.opg;
    const int ec = ec0 + ch;
    if (ec / d.cpg == grp) gw[((size_t)oc * d.cpg + (ec - grp * d.cpg)) * d.KK + kp] = tile[kp * kRedCh + ch][j];
  }
}

// ================================================================================================ K3: backward weight
// grid (M blocks = pairs of units, pixel splits, SG * oc tiles).  D[128 k'][BN oc] += col^T[128 k'][64 px] . gout[64 px][BN oc]
// per 64-pixel stage; the gathered tile is written [unit half][pixel][64 ch] = MN-major for the tensor core.
template <int kTmemCols>
__global__ void __launch_bounds__(kThreads, 1) dcn_bwd_weight_tc_kernel(const float* __restrict__ xh,
                                                                        const float* __restrict__ offset,
                                                                        const float* __restrict__ mask,
                                                                        const uint8_t* __restrict__ gt, const TC d,
                                                                        const K3P k, const int split,
                                                                        float* __restrict__ gw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset form keeps the shared address space
  int4* tap_s = reinterpret_cast<int4*>(smem + k.S * k.stage_bytes);  // [16 warps][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tap_s) + 4096);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + k.S;
  uint64_t* accum_bar = bars + 2 * k.S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k.S + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.x;
  const int sg = blockIdx.z / k.noct, oct = blockIdx.z - sg * k.noct;
  const int total = d.N * d.stages_img;
  const int gs0 = blockIdx.y * k.sper, gs1 = min(total, gs0 + k.sper), ns = gs1 - gs0;

  if (tid == 0) {
    for (int s = 0; s < k.S; ++s) {
      mbar_init(&full_bar[s], kWorkerWarps + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    mbar_init_fence();
  }
  if (warp == kWorkerWarps + 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWorkerWarps) {
    // =============================================================== GATHER: A tile [2 units][64 px][64 ch], hi + lo
    const int half = lane >> 4, q = lane & 15;
    const int u = 2 * mb + half;             // this half-warp's unit
    const bool u_ok = u < d.U;
    const int kp = u_ok ? u / d.cbs : 0, cb = u_ok ? u - kp * d.cbs : 0;
    const int cbase = sg * d.cps + cb * 64;
    const int dg = cbase / d.cpdg;
    // lanes 0..15 build the taps of this warp's 8 pixels x 2 units; the loads of the NEXT stage's offsets / mask are issued
    // before this stage's gather and consumed after it, so that their latency is hidden behind the gather
    constexpr int PX = 64 / kWorkerWarps;  // pixels of a stage per warp: 8
    const int tl_half = (lane / PX) & 1, tl_j = lane % PX;
    const int tl_u = 2 * mb + tl_half;
    const bool tl_ok = lane < 2 * PX && tl_u < d.U;
    const int tl_kp = tl_ok ? tl_u / d.cbs : 0;
    const int tl_dg = tl_ok ? (sg * d.cps + (tl_u - tl_kp * d.cbs) * 64) / d.cpdg : 0;
    int4* my_taps = tap_s + warp * 2 * (2 * PX);  // [2 buffers][2 units][PX]
    auto stage_of = [&](int i, int& b, int& pbase) {
      const int gs = gs0 + i;
      b = gs / d.stages_img;
      pbase = (gs - b * d.stages_img) * 64;
    };
    if (ns > 0) {
      int b0, pb0;
      stage_of(0, b0, pb0);
      if (lane < 2 * PX) my_taps[lane] = tl_ok ? make_tap(d, offset, mask, b0, tl_dg, tl_kp, pb0 + warp * PX + tl_j) : make_int4(0, 0, 0, 0);
    }
    __syncwarp();
    for (int i = 0; i < ns; ++i) {
      const int s = i % k.S;
      const uint32_t par = (uint32_t)((i / k.S) & 1);
      int b, pbase, bn = 0, pbn = 0;
      stage_of(i, b, pbase);
      const bool have_next = i + 1 < ns;
      TapRaw raw = {0.f, 0.f, 1.f};
      if (have_next) {
        stage_of(i + 1, bn, pbn);
        if (tl_ok) raw = tap_loads(d, offset, mask, bn, tl_dg, tl_kp, pbn + warp * PX + tl_j);
      }
      const float* __restrict__ xc = xh + ((size_t)b * d.H * d.W - (size_t)(d.W + 1)) * d.Cin + cbase + q * 4;  // biased
      const uint32_t ucin = (uint32_t)d.Cin, urs = (uint32_t)(d.W * d.Cin);
      uint8_t* a_hi = smem + s * k.stage_bytes + half * 8192;
      uint8_t* a_lo = a_hi + kTile;
      const int4* cur = my_taps + (i & 1) * (2 * PX) + half * PX;
      mbar_wait(&empty_bar[s], par ^ 1u);
      {
        int4 tap[PX];
        Taps4 c[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) tap[j] = cur[j];
#pragma unroll
        for (int j = 0; j < PX; ++j) load_corners(xc, ucin, urs, tap[j].x, c[j]);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const uint32_t r = (uint32_t)(warp * PX + j);
          gather_store(c[j], tap[j], a_hi, a_lo, swz128(r, (uint32_t)(q >> 1)) + (uint32_t)(q & 1) * 8u, split != 0);
        }
      }
      if (have_next && lane < 2 * PX)
        my_taps[((i + 1) & 1) * (2 * PX) + lane] =
            tl_ok ? tap_finish(d, raw, mask != nullptr, tl_kp, pbn + warp * PX + tl_j) : make_int4(0, 0, 0, 0);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // =============================================================== EPILOGUE: TMEM [k' row][oc col] -> this split's partial tile
    mbar_wait(accum_bar, 0u);
    tc_fence_after();
    const int quad = warp & 3, cgrp = warp >> 2;
    gw_tile_store(tmem_base, quad, cgrp, lane, k.BN,
                  gw + ((size_t)sg * d.MC * 128 + (size_t)mb * 128) * d.ops + (size_t)oct * k.BN,
                  d.ops);
  } else if (warp == kWorkerWarps) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(split ? 2 : 1) * (uint32_t)k.BN * 128u;
      for (int i = 0; i < ns; ++i) {
        const int s = i % k.S;
        const uint32_t par = (uint32_t)((i / k.S) & 1);
        mbar_wait(&empty_bar[s], par ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        const size_t tile = ((size_t)(gs0 + i) * d.SG + sg) * k.noct + oct;
        bulk_g2s(smem + s * k.stage_bytes + 2 * kTile, gt + tile * (size_t)(2 * k.BN * 128), bytes, &full_bar[s]);
      }
    }
  } else {
    const uint32_t idesc = umma_idesc(k.BN, true, false);
    for (int i = 0; i < ns; ++i) {
      const int s = i % k.S;
      const uint32_t par = (uint32_t)((i / k.S) & 1);
      mbar_wait(&full_bar[s], par);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + s * k.stage_bytes), a_lo = a_hi + kTile;
        const uint32_t b_hi = a_hi + 2 * kTile, b_lo = b_hi + (uint32_t)k.BN * 128u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // A: 16 pixels (K) = two 8-row atoms 1024 B apart; the two 64-channel M blocks are 8192 B apart
          const uint32_t ao = (uint32_t)kk * 2048u, bo = (uint32_t)kk * 32u;
          umma_bf16(tmem_base, umma_desc(a_hi + ao, 8192, 1024), umma_desc(b_hi + bo, 0, 1024), idesc, (i > 0 || kk > 0) ? 1u : 0u);
          if (split) {
            umma_bf16(tmem_base, umma_desc(a_hi + ao, 8192, 1024), umma_desc(b_lo + bo, 0, 1024), idesc, 1u);
            umma_bf16(tmem_base, umma_desc(a_lo + ao, 8192, 1024), umma_desc(b_hi + bo, 0, 1024), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (i == ns - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkerWarps + 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ================================================================================================ K3c: backward weight from saved columns
// Same grid, tiles and epilogue as K3, but the A operand is streamed back from the column tiles the forward saved
// (dcn_fwd_tc_kernel's saver warp) instead of being sampled from x again: a pure TMA -> tcgen05 pipeline.  A 64-pixel stage of
// a unit is one contiguous 8 KB half of its [128 px][64 ch] tile (the 128-byte swizzle repeats every 8 rows).
// (Measured: multicasting the shared grad_out tile over clusters of 2 along the macro-chunk axis changes nothing -- the kernel
// is bound by the three tcgen05 passes of the bf16x3 split and the column stream, not by L2 reads -- and clusters of 3 / 6 do
// not fit the one-wave grid, 135 / 132 resident CTAs; the kernel therefore runs without clusters.)
template <int kTmemCols>
__global__ void __launch_bounds__(kThreads, 1) dcn_bwd_weight_cols_kernel(const uint8_t* __restrict__ cols,
                                                                          const uint8_t* __restrict__ gt, const TC d,
                                                                          const K3P k, const int S, const int split,
                                                                          float* __restrict__ gw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * k.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + S;
  uint64_t* accum_bar = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.x;
  const int sg = blockIdx.z / k.noct, oct = blockIdx.z - sg * k.noct;
  const int total = d.N * d.stages_img;
  const int gs0 = blockIdx.y * k.sper, gs1 = min(total, gs0 + k.sper), ns = gs1 - gs0;
  const int nu = (2 * mb + 1 < d.U) ? 2 : 1;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    mbar_init_fence();
  }
  if (warp == kWorkerWarps + 1) tmem_alloc<kTmemCols>(tmem_slot);
  if (nu == 1 && tid < kWorkers) {  // odd unit count: the second 64 rows of A are never loaded; keep them finite
    for (int s = 0; s < S; ++s)
      for (int i = tid; i < 2 * 512; i += kWorkers) {
        const int part = i >> 9, w = i & 511;
        reinterpret_cast<uint4*>(smem + s * k.stage_bytes + part * kTile + 8192)[w] = make_uint4(0, 0, 0, 0);
      }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWorkerWarps) {
    // =============================================================== EPILOGUE: TMEM [k' row][oc col] -> this split's partial tile
    mbar_wait(accum_bar, 0u);
    tc_fence_after();
    const int quad = warp & 3, cgrp = warp >> 2;
    gw_tile_store(tmem_base, quad, cgrp, lane, k.BN,
                  gw + ((size_t)sg * d.MC * 128 + (size_t)mb * 128) * d.ops + (size_t)oct * k.BN,
                  d.ops);
  } else if (warp == kWorkerWarps) {
    // =============================================================== TMA producer: column halves + grad_out tile per stage
    if (lane == 0) {
      const uint32_t parts = split ? 2u : 1u;
      const uint32_t tile_bytes = parts * (uint32_t)kTile;
      const uint32_t gbytes = parts * (uint32_t)k.BN * 128u;
      for (int i = 0; i < ns; ++i) {
        const int s = i % S;
        const uint32_t par = (uint32_t)((i / S) & 1);
        const int gs = gs0 + i;
        const int b = gs / d.stages_img, st = gs - b * d.stages_img;
        mbar_wait(&empty_bar[s], par ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)nu * parts * 8192u + gbytes);
        uint8_t* stage = smem + s * k.stage_bytes;
        const uint8_t* src =
            cols + ((((size_t)b * d.tiles_img + (st >> 1)) * d.SG + sg) * d.U + 2 * mb) * tile_bytes + (size_t)(st & 1) * 8192;
        for (int ul = 0; ul < nu; ++ul)
          for (uint32_t part = 0; part < parts; ++part)
            bulk_g2s(stage + part * kTile + ul * 8192, src + (size_t)ul * tile_bytes + (size_t)part * kTile, 8192u, &full_bar[s]);
        const size_t tile = ((size_t)gs * d.SG + sg) * k.noct + oct;
        bulk_g2s(stage + 2 * kTile, gt + tile * (size_t)(2 * k.BN * 128), gbytes, &full_bar[s]);
      }
    }
  } else {
    // =============================================================== MMA issuer (as K3)
    const uint32_t idesc = umma_idesc(k.BN, true, false);
    for (int i = 0; i < ns; ++i) {
      const int s = i % S;
      const uint32_t par = (uint32_t)((i / S) & 1);
      mbar_wait(&full_bar[s], par);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + s * k.stage_bytes), a_lo = a_hi + kTile;
        const uint32_t b_hi = a_hi + 2 * kTile, b_lo = b_hi + (uint32_t)k.BN * 128u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ao = (uint32_t)kk * 2048u, bo = (uint32_t)kk * 32u;
          umma_bf16(tmem_base, umma_desc(a_hi + ao, 8192, 1024), umma_desc(b_hi + bo, 0, 1024), idesc, (i > 0 || kk > 0) ? 1u : 0u);
          if (split) {
            umma_bf16(tmem_base, umma_desc(a_hi + ao, 8192, 1024), umma_desc(b_lo + bo, 0, 1024), idesc, 1u);
            umma_bf16(tmem_base, umma_desc(a_lo + ao, 8192, 1024), umma_desc(b_hi + bo, 0, 1024), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (i == ns - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkerWarps + 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ================================================================================================ operand pre-tiling
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z);
  split2(v[6], v[7], hi.w, lo.w);
}

// weight element of (global output channel, global input channel, kernel point); 0 across original groups
__device__ __forceinline__ float w_elem(const float* __restrict__ w, const TC& d, int oc, int c, int kp) {
  const int g = c / d.cpg;
  if (oc / d.opg != g) return 0.f;
  return __ldg(w + ((size_t)oc * d.cpg + (c - g * d.cpg)) * d.KK + kp);
}

// K1 B tiles: [sg][oc tile][unit] -> [BN rows = oc][64 k' = channels of the unit], K-major swizzled, hi then lo
__global__ void dcn_wtile_fwd_kernel(const float* __restrict__ w, const TC d, int BN, int noct, uint8_t* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)d.SG * noct * d.U * BN * 8;
  if (idx >= total) return;
  const int c16 = (int)(idx & 7);
  const int r = (int)((idx >> 3) % BN);
  const long long tile = idx / (8LL * BN);
  const int u = (int)(tile % d.U);
  const int oct = (int)((tile / d.U) % noct);
  const int sg = (int)(tile / ((long long)d.U * noct));
  const int kp = u / d.cbs, cb = u - kp * d.cbs;
  const int oc = sg * d.ops + oct * BN + r;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = w_elem(w, d, oc, sg * d.cps + cb * 64 + c16 * 8 + e, kp);
  uint4 hi, lo;
  split8(v, hi, lo);
  uint8_t* base = dst + tile * (size_t)(2 * BN * 128);
  *reinterpret_cast<uint4*>(base + swz128((uint32_t)r, (uint32_t)c16)) = hi;
  *reinterpret_cast<uint4*>(base + BN * 128 + swz128((uint32_t)r, (uint32_t)c16)) = lo;
}

// K2 B tiles: [sg][macro-chunk][K stage] -> [128 rows = k' of the 2 units][64 oc of the stage], hi then lo
__global__ void dcn_wtile_bwd_kernel(const float* __restrict__ w, const TC d, uint8_t* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)d.SG * d.MC * d.nks * 128 * 8;
  if (idx >= total) return;
  const int c16 = (int)(idx & 7);
  const int r = (int)((idx >> 3) & 127);
  const long long tile = idx >> 10;
  const int ks = (int)(tile % d.nks);
  const int m = (int)((tile / d.nks) % d.MC);
  const int sg = (int)(tile / ((long long)d.nks * d.MC));
  const int u = 2 * m + (r >> 6);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (u < d.U) {
    const int kp = u / d.cbs, cb = u - kp * d.cbs;
    const int c = sg * d.cps + cb * 64 + (r & 63);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w_elem(w, d, sg * d.ops + ks * 64 + c16 * 8 + e, c, kp);
  }
  uint4 hi, lo;
  split8(v, hi, lo);
  uint8_t* base = dst + tile * (size_t)(2 * kTile);
  *reinterpret_cast<uint4*>(base + swz128((uint32_t)r, (uint32_t)c16)) = hi;
  *reinterpret_cast<uint4*>(base + kTile + swz128((uint32_t)r, (uint32_t)c16)) = lo;
}

// K2 A tiles: gout [N,Cout,HoWo] -> [b][pixel tile][sg][K stage] -> [128 rows = px][64 oc], hi then lo.  One CTA per tile.
// grad w.r.t. the convolution result when the forward ended in y = relu(acc * scale + shift)
__device__ __forceinline__ float gout_elem(const float* __restrict__ gout, const float* __restrict__ ysaved, const Epi& ep,
                                           size_t idx, int oc) {
  float g = __ldg(gout + idx);
  if (ep.relu && !(__ldg(ysaved + idx) > 0.f)) g = 0.f;
  if (ep.scale) g *= __ldg(ep.scale + oc);
  return g;
}

// Also writes the K3 B tiles (oc-row layout, see dcn_gout_oc_tiles_kernel) of the same elements when dst_oc != nullptr, so
// that a backward that needs both gradients reads grad_out once.
__global__ void __launch_bounds__(256) dcn_gout_px_tiles_kernel(const float* __restrict__ gout,
                                                                const float* __restrict__ ysaved, const Epi ep, const TC d,
                                                                uint8_t* __restrict__ dst, uint8_t* __restrict__ dst_oc,
                                                                int BN, int noct) {
  // grid (tiles, 4): a CTA converts 32 of the tile's 128 pixel rows -- 4x the CTAs of a tile-per-CTA layout, which matters for
  // the small maps (res5: 144 tiles on 148 SMs)
  __shared__ float t[64][33];
  const long long tile = blockIdx.x;
  const int ks = (int)(tile % d.nks);
  const int sg = (int)((tile / d.nks) % d.SG);
  const int pt = (int)((tile / ((long long)d.nks * d.SG)) % d.tiles_img);
  const int b = (int)(tile / ((long long)d.nks * d.SG * d.tiles_img));
  const int r0 = blockIdx.y * 32;
  const int p0 = pt * 128 + r0, oc0 = sg * d.ops + ks * 64;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // 8 independent loads per thread: 32 pixels (one 128-byte run) of 64 channels
    const int e = tid + i * 256;
    const int o = e >> 5, pp = e & 31;
    const int p = p0 + pp;
    t[o][pp] = p < d.HoWo ? gout_elem(gout, ysaved, ep, ((size_t)b * d.Cout + oc0 + o) * d.HoWo + p, oc0 + o) : 0.f;
  }
  __syncthreads();
  uint8_t* base = dst + tile * (size_t)(2 * kTile);
  {
    const int r = tid >> 3, c16 = tid & 7;  // 32 rows x 8 chunks = 256 threads
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[c16 * 8 + i][r];
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(base + swz128((uint32_t)(r0 + r), (uint32_t)c16)) = hi;
    *reinterpret_cast<uint4*>(base + kTile + swz128((uint32_t)(r0 + r), (uint32_t)c16)) = lo;
  }
  if (dst_oc) {  // rows = output channels, 64-pixel stages: this CTA's 32 pixels are chunks [c0, c0 + 4) of one stage
    const int o = tid >> 2, ch = tid & 3;  // 64 channel rows x 4 chunks of 8 pixels
    const int ocs = ks * 64 + o;           // channel inside the super-group
    const int oct = ocs / BN, row = ocs - oct * BN;
    const int pglob = pt * 128 + r0;       // first pixel of this CTA inside the image
    const int ps = pglob >> 6, c16 = ((pglob & 63) >> 3) + ch;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[o][ch * 8 + i];
    uint4 hi, lo;
    split8(v, hi, lo);
    if (ps < d.stages_img) {
      const size_t tile_oc = (((size_t)b * d.stages_img + ps) * d.SG + sg) * noct + oct;
      uint8_t* bo = dst_oc + tile_oc * (size_t)(2 * BN * 128);
      *reinterpret_cast<uint4*>(bo + swz128((uint32_t)row, (uint32_t)c16)) = hi;
      *reinterpret_cast<uint4*>(bo + BN * 128 + swz128((uint32_t)row, (uint32_t)c16)) = lo;
    }
  }
}

// K3 B tiles: gout -> [b][64-pixel stage][sg][oc tile] -> [BN rows = oc][64 px], K-major swizzled, hi then lo
__global__ void dcn_gout_oc_tiles_kernel(const float* __restrict__ gout, const float* __restrict__ ysaved, const Epi ep,
                                         const TC d, int BN, int noct, uint8_t* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)d.N * d.stages_img * d.SG * noct * BN * 8;
  if (idx >= total) return;
  const int c16 = (int)(idx & 7);
  const int r = (int)((idx >> 3) % BN);
  const long long tile = idx / (8LL * BN);
  const int oct = (int)(tile % noct);
  const int sg = (int)((tile / noct) % d.SG);
  const int ps = (int)((tile / ((long long)noct * d.SG)) % d.stages_img);
  const int b = (int)(tile / ((long long)noct * d.SG * d.stages_img));
  const int oc = sg * d.ops + oct * BN + r;
  const size_t src = ((size_t)b * d.Cout + oc) * d.HoWo;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int p = ps * 64 + c16 * 8 + e;
    v[e] = p < d.HoWo ? gout_elem(gout, ysaved, ep, src + p, oc) : 0.f;
  }
  uint4 hi, lo;
  split8(v, hi, lo);
  uint8_t* base = dst + tile * (size_t)(2 * BN * 128);
  *reinterpret_cast<uint4*>(base + swz128((uint32_t)r, (uint32_t)c16)) = hi;
  *reinterpret_cast<uint4*>(base + BN * 128 + swz128((uint32_t)r, (uint32_t)c16)) = lo;
}

// [N,HW,C] -> [N,C,HW]  (grad_x back to the reference's layout)
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ src, int C, int HW,
                                                           float* __restrict__ dst) {
  __shared__ float t[64][33];
  const int hw0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const float* __restrict__ s = src + (size_t)blockIdx.z * HW * C;
  float* __restrict__ o = dst + (size_t)blockIdx.z * C * HW;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * 32; e += 256) {
    const int pp = e >> 5, cc = e & 31;
    t[pp][cc] = (hw0 + pp < HW && c0 + cc < C) ? __ldg(s + (size_t)(hw0 + pp) * C + c0 + cc) : 0.f;
  }
  __syncthreads();
  for (int e = tid; e < 64 * 32; e += 256) {
    const int cc = e >> 6, pp = e & 63;
    if (hw0 + pp < HW && c0 + cc < C) o[(size_t)(c0 + cc) * HW + hw0 + pp] = t[pp][cc];
  }
}

// y = relu(y * scale + shift) in place: the forward's epilogue when the reduction was split over kernel points
__global__ void dcn_epilogue_kernel(float* __restrict__ y, long long total, int Cout, int HoWo, const Epi ep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int oc = (int)((i / HoWo) % Cout);
  float v = y[i];
  if (ep.scale) v *= __ldg(ep.scale + oc);
  if (ep.shift) v += __ldg(ep.shift + oc);
  if (ep.relu) v = fmaxf(v, 0.f);
  y[i] = v;
}

int to_nhwc(const float* x, const TC& d, float* dst, cudaStream_t stream) {
  d2b_pyramid P = {};
  P.num_levels = 1;
  P.feat[0] = x;
  P.H[0] = d.H;
  P.W[0] = d.W;
  P.scale[0] = 1.f;
  float* dsts[1] = {dst};
  return d2b_pyramid_nchw_to_nhwc(&P, d.N, d.Cin, dsts, (void*)stream);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host entry points
// (internal linkage across the library: declared again in deform_conv.cu)
int d2b_deform_conv_tc_supported(const d2b_dcn_params* p) {
  TC d;
  K1P k1;
  return make_tc(p, d) && plan_k1(d, k1);
}

int d2b_deform_conv_tc_bwd_supported(const d2b_dcn_params* p) {
  TC d;
  K2P k2;
  K3P k3;
  return make_tc(p, d) && plan_k2(d, k2) && plan_k3(d, k3);
}

size_t d2b_deform_conv_tc_fwd_workspace(const d2b_dcn_params* p, int x_nhwc) {
  TC d;
  K1P k;
  if (!make_tc(p, d) || !plan_k1(d, k)) return 0;
  size_t b = align256((size_t)d.SG * k.noct * d.U * 2 * k.BN * 128);
  if (!x_nhwc) b += align256(sizeof(float) * (size_t)d.N * d.H * d.W * d.Cin);
  return b;
}

// bytes of the column tiles one forward call saves for the backward (0: shape not taken by the tensor-core kernels)
size_t d2b_deform_conv_tc_cols_bytes(const d2b_dcn_params* p, int precision) {
  TC d;
  K1P k;
  if (precision == 0 || !make_tc(p, d) || !plan_k1(d, k)) return 0;
  return (size_t)d.N * d.tiles_img * d.SG * d.U * (size_t)(precision == 1 ? 2 : 1) * kTile;
}

// tcflags: bit 0 = x is NHWC, bit 1 = `offset` is the fused [N, 3*DG*KK, Ho, Wo] offset + mask-logit tensor (mask must be null)
// cols: optional, d2b_deform_conv_tc_cols_bytes() bytes, 16-byte aligned: receives the sampled columns (see the saver warp)
int d2b_deform_conv_forward_tc(const float* x, const float* offset, const float* mask, const float* weight,
                               const float* scale, const float* shift, int relu, const d2b_dcn_params* p, int precision,
                               int tcflags, float* out, void* cols, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int x_nhwc = tcflags & 1;
  TC d;
  K1P k;
  if (!make_tc(p, d) || !plan_k1(d, k)) return D2B_EUNSUPPORTED;  // argument validity was checked by the caller
  if (d.N == 0) return D2B_OK;
  if (tcflags & 2) {
    if (mask) return D2B_EINVAL;
    mask = use_fused_offset_mask(d, offset);
  }
  if (!workspace || workspace_bytes < d2b_deform_conv_tc_fwd_workspace(p, x_nhwc)) return D2B_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (x_nhwc && (reinterpret_cast<uintptr_t>(x) & 15)) ||
      (reinterpret_cast<uintptr_t>(cols) & 15))
    return D2B_EINVAL;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* wt = ws;
  ws += align256((size_t)d.SG * k.noct * d.U * 2 * k.BN * 128);
  const float* xh = x;
  if (!x_nhwc) {
    float* xn = reinterpret_cast<float*>(ws);
    int rc = to_nhwc(x, d, xn, stream);
    if (rc) return rc;
    xh = xn;
  }
  {
    const long long total = (long long)d.SG * k.noct * d.U * k.BN * 8;
    dcn_wtile_fwd_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(weight, d, k.BN, k.noct, wt);
    D2B_CHECK_LAUNCH();
  }
  if (k.red) D2B_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)d.N * d.Cout * d.HoWo, stream));
  const int smem_bytes = k.S * k.stage_bytes + k.tap_bytes + 1024 + 256;
  const int tcols = pow2_cols(k.gspan * k.BN);
  const int split = precision == 1 ? 1 : 0;
  const Epi ep = {scale, shift, relu};
  dim3 grid(d.N * d.tiles_img, (d.SG / k.gspan) * k.noct, k.ksplit);
#define D2B_LAUNCH_K1(COLS)                                                                                            \
  {                                                                                                                    \
    D2B_ALLOW_BIG_SMEM(dcn_fwd_tc_kernel<COLS>);                                                                       \
    dcn_fwd_tc_kernel<COLS><<<grid, kThreadsK1, smem_bytes, stream>>>(xh, offset, mask, wt, ep, d, k, split, out,      \
                                                                      reinterpret_cast<uint8_t*>(cols));               \
  }
  if (tcols <= 32) D2B_LAUNCH_K1(32)
  else if (tcols == 64) D2B_LAUNCH_K1(64)
  else if (tcols == 128) D2B_LAUNCH_K1(128)
  else D2B_LAUNCH_K1(256)
#undef D2B_LAUNCH_K1
  D2B_CHECK_LAUNCH();
  if (k.red && (scale || relu)) {
    const long long total = (long long)d.N * d.Cout * d.HoWo;
    dcn_epilogue_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(out, total, d.Cout, d.HoWo, ep);
    D2B_CHECK_LAUNCH();
  }
  return D2B_OK;
}

size_t d2b_deform_conv_tc_bwd_workspace(const d2b_dcn_params* p, int x_nhwc, int need_data, int need_weight) {
  TC d;
  K2P k2;
  K3P k3;
  if (!make_tc(p, d) || !plan_k2(d, k2) || !plan_k3(d, k3)) return 0;
  size_t b = 0;
  const size_t xbytes = align256(sizeof(float) * (size_t)d.N * d.H * d.W * d.Cin);
  if (!x_nhwc) b += xbytes;                     // x in NHWC
  if (need_data && !x_nhwc) b += xbytes;        // grad_x accumulated in NHWC
  if (need_data) {
    b += align256((size_t)d.N * d.tiles_img * d.SG * d.nks * 2 * kTile);  // gout, pixel-row tiles
    b += align256((size_t)d.SG * d.MC * d.nks * 2 * kTile);               // W^T tiles
  }
  if (need_weight) {
    b += align256((size_t)d.N * d.stages_img * d.SG * k3.noct * 2 * k3.BN * 128);  // gout, oc-row tiles
    b += align256(sizeof(float) * (size_t)d.SG * d.MC * 128 * d.ops);  // weight gradient in the accumulator tiles' layout
  }
  return b;
}

// grad_x: NCHW (or NHWC when x_nhwc) fully written; grad_offset / grad_mask / grad_weight fully written.
// With the fused offset + mask-logit tensor (tcflags bit 1) grad_offset is its [N, 3*DG*KK, Ho, Wo] gradient (mask part through
// the sigmoid) and grad_mask must be null.  scale / y_saved / relu: transpose of the forward's epilogue.
int d2b_deform_conv_backward_tc(const float* x, const float* offset, const float* mask, const float* weight,
                                const float* grad_out, const float* scale, const float* y_saved, int relu,
                                const d2b_dcn_params* p, int precision, int tcflags, const void* cols, float* grad_x,
                                float* grad_offset, float* grad_mask, float* grad_weight, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int x_nhwc = tcflags & 1;
  TC d;
  K2P k2;
  K3P k3;
  if (!make_tc(p, d) || !plan_k2(d, k2) || !plan_k3(d, k3)) return D2B_EUNSUPPORTED;
  const int need_data = (grad_x || grad_offset || grad_mask) ? 1 : 0, need_weight = grad_weight ? 1 : 0;
  const bool fused_om = (tcflags & 2) != 0;
  if (fused_om && (mask || grad_mask)) return D2B_EINVAL;
  if (relu && !y_saved) return D2B_EINVAL;
  const size_t nx = (size_t)d.N * d.H * d.W * d.Cin, noff = (size_t)d.N * d.DG * (fused_om ? 3 : 2) * d.KK * d.HoWo;
  const size_t nm = (size_t)d.N * d.DG * d.KK * d.HoWo, nw = (size_t)d.Cout * d.cpg * d.KK;
  if (fused_om) {
    mask = use_fused_offset_mask(d, offset);
    if (grad_offset) grad_mask = grad_offset + (size_t)d.DG * 2 * d.KK * d.HoWo;
  }
  const Epi ep = {scale, nullptr, relu};
  if (d.N == 0 || (!need_data && !need_weight)) {
    void* zp[3] = {grad_offset, fused_om ? nullptr : grad_mask, grad_weight};
    size_t zb[3] = {noff * 4, nm * 4, nw * 4};
    return d2b_zero_buffers(zp, zb, 3, stream);
  }
  if (!workspace || workspace_bytes < d2b_deform_conv_tc_bwd_workspace(p, x_nhwc, need_data, need_weight)) return D2B_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (x_nhwc && (reinterpret_cast<uintptr_t>(x) & 15)) ||
      (x_nhwc && grad_x && (reinterpret_cast<uintptr_t>(grad_x) & 15)) || (reinterpret_cast<uintptr_t>(cols) & 15))
    return D2B_EINVAL;
  const int split = precision == 1 ? 1 : 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const size_t xbytes = align256(sizeof(float) * nx);
  const float* xh = x;
  if (!x_nhwc) {
    float* xn = reinterpret_cast<float*>(ws);
    ws += xbytes;
    int rc = to_nhwc(x, d, xn, stream);
    if (rc) return rc;
    xh = xn;
  }
  float* gxh = nullptr;
  if (need_data && grad_x) gxh = x_nhwc ? grad_x : reinterpret_cast<float*>(ws);
  uint8_t* gt_px = nullptr;
  uint8_t* wt = nullptr;
  if (need_data) {
    if (!x_nhwc) ws += xbytes;
    gt_px = ws;
    ws += align256((size_t)d.N * d.tiles_img * d.SG * d.nks * 2 * kTile);
    wt = ws;
    ws += align256((size_t)d.SG * d.MC * d.nks * 2 * kTile);
  }
  uint8_t* gt_oc = need_weight ? ws : nullptr;
  float* gw_part = nullptr;
  if (need_weight) {
    ws += align256((size_t)d.N * d.stages_img * d.SG * k3.noct * 2 * k3.BN * 128);
    gw_part = reinterpret_cast<float*>(ws);
  }
  {  // every accumulated output of the call zero-filled by one launch (grad_mask of the fused layout lives inside grad_offset;
    // grad_weight is accumulated in the staging matrix and then written element by element)
    void* zp[4] = {grad_offset, fused_om ? nullptr : grad_mask, gxh, gw_part};
    size_t zb[4] = {noff * 4, nm * 4, nx * 4, sizeof(float) * (size_t)d.SG * d.MC * 128 * d.ops};
    int rc = d2b_zero_buffers(zp, zb, 4, stream);
    if (rc) return rc;
  }
  if (need_data) {
    // grad_out is read once: pixel-row tiles for K2 and (when the weight gradient is wanted too) channel-row tiles for K3
    dcn_gout_px_tiles_kernel<<<dim3((unsigned)((size_t)d.N * d.tiles_img * d.SG * d.nks), 4), 256, 0, stream>>>(
        grad_out, y_saved, ep, d, gt_px, gt_oc, k3.BN, k3.noct);
    D2B_CHECK_LAUNCH();
    {
      const long long total = (long long)d.SG * d.MC * d.nks * 128 * 8;
      dcn_wtile_bwd_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(weight, d, wt);
      D2B_CHECK_LAUNCH();
    }
    const int smem_bytes = 2 * 4 * kTile + 128 * kGcolPitch * 4 + k2.tap_bytes + 1024 + 256;
    D2B_ALLOW_BIG_SMEM(dcn_bwd_data_tc_kernel);
    dim3 grid(d.N * d.tiles_img, d.SG, k2.msplit);
    dcn_bwd_data_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(xh, offset, mask, gt_px, wt, d, k2, split, gxh, grad_offset,
                                                                  mask ? grad_mask : nullptr);
    D2B_CHECK_LAUNCH();
    if (grad_x && !x_nhwc) {
      dim3 g2(d2b_cdiv(d.H * d.W, 64), d2b_cdiv(d.Cin, 32), d.N);
      nhwc_to_nchw_kernel<<<g2, 256, 0, stream>>>(gxh, d.Cin, d.H * d.W, grad_x);
      D2B_CHECK_LAUNCH();
    }
  }
  if (need_weight) {
    if (!need_data) {
      const long long total = (long long)d.N * d.stages_img * d.SG * k3.noct * k3.BN * 8;
      dcn_gout_oc_tiles_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(grad_out, y_saved, ep, d, k3.BN, k3.noct, gt_oc);
      D2B_CHECK_LAUNCH();
    }
    const int tcols = pow2_cols(k3.BN);
    dim3 grid(d.MC, k3.nsplit, d.SG * k3.noct);
    if (cols) {  // the forward kept its sampled columns: stream them back (no second pass over x)
      const int S = std::min(6, (kMaxSmem - 2048) / k3.stage_bytes);
      const int smem_bytes = S * k3.stage_bytes + 1024 + 256;
      const uint8_t* cl = reinterpret_cast<const uint8_t*>(cols);
#define D2B_LAUNCH_K3C(COLS)                                                                                             \
  {                                                                                                                      \
    D2B_ALLOW_BIG_SMEM(dcn_bwd_weight_cols_kernel<COLS>);                                                                \
    dcn_bwd_weight_cols_kernel<COLS><<<grid, kThreads, smem_bytes, stream>>>(cl, gt_oc, d, k3, S, split, gw_part);       \
  }
      if (tcols <= 32) D2B_LAUNCH_K3C(32)
      else if (tcols == 64) D2B_LAUNCH_K3C(64)
      else if (tcols == 128) D2B_LAUNCH_K3C(128)
      else D2B_LAUNCH_K3C(256)
#undef D2B_LAUNCH_K3C
    } else {
      const int smem_bytes = k3.S * k3.stage_bytes + 4096 + 1024 + 256;
#define D2B_LAUNCH_K3(COLS)                                                                                              \
  {                                                                                                                      \
    D2B_ALLOW_BIG_SMEM(dcn_bwd_weight_tc_kernel<COLS>);                                                                  \
    dcn_bwd_weight_tc_kernel<COLS><<<grid, kThreads, smem_bytes, stream>>>(xh, offset, mask, gt_oc, d, k3, split, gw_part); \
  }
      if (tcols <= 32) D2B_LAUNCH_K3(32)
      else if (tcols == 64) D2B_LAUNCH_K3(64)
      else if (tcols == 128) D2B_LAUNCH_K3(128)
      else D2B_LAUNCH_K3(256)
#undef D2B_LAUNCH_K3
    }
    D2B_CHECK_LAUNCH();
    {
      if (d.KK <= 9) {  // d.ops is a multiple of 16 (shape gate)
        dcn_gw_reduce_tile_kernel<<<dim3(d.SG * (d.cps / kRedCh), d.ops / kRedOc), 256, 0, stream>>>(gw_part, d, grad_weight);
      } else {
        const long long total = (long long)d.SG * d.U * 64 * d.ops;
        dcn_gw_reduce_kernel<<<d2b_cdiv(total, 256), 256, 0, stream>>>(gw_part, d, grad_weight);
      }
      D2B_CHECK_LAUNCH();
    }
  }
  return D2B_OK;
}
