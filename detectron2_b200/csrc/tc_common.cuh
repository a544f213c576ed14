// PTX wrappers shared by the tcgen05 / TMEM / bulk-copy kernels (sm_100a only).
//
//   mbarrier            producer/consumer rings (generic, async-proxy and tensor-core arrivals)
//   cp.async.bulk       TMA engine, linear form: one instruction moves a whole pre-tiled operand block
//                       global -> shared and completes on an mbarrier (SASS: UBLKCP)
//   tcgen05.*           TMEM allocation, UMMA issue (SASS: UTCHMMA), commit (UTCBAR), TMEM loads (LDTM)
//   red.global.v4.f32   128-bit vector reduction to global memory (SASS: REDG.E.ADD.F32x4)
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace d2b_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure (trap -> cudaErrorLaunchFailure), never as a hung GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  for (uint32_t spin = 1;; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spin & 1023u) == 0 && global_timer_ns() - t0 > 2000000000ull) __trap();  // 2 s: far beyond any legitimate wait
  }
}

// ------------------------------------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------ bulk copy (TMA, linear)
// size % 16 == 0, both addresses 16-byte aligned; completes `bytes` on `bar` (pair with mbar_arrive_expect_tx)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the thread's bulk async-group (commit, then wait for the reads of the source to finish
// before the shared-memory tile is reused)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {  // the warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i <-> TMEM lane base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ UMMA
// Shared-memory matrix descriptor (sm_100 version 1), 128-byte swizzle.
//   K-major  tile [rows][64 bf16]: rows 128 B apart, 8-row atoms `sbo` bytes apart (1024 when rows are dense)
//   MN-major tile [k rows][64 bf16 of M/N]: 8-k-row atoms `sbo` bytes apart, 64-element M/N blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: D = f32, A = B = bf16, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by one thread for the CTA
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
// mbarrier arrive once every previously issued MMA of this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ misc
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {  // a -> low half
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x = hi + lo with hi, lo bf16 (lo = rn(x - hi)): three bf16 MMAs hi*hi + hi*lo + lo*hi keep ~16 mantissa bits per product
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);                                   // one F2FP for both values
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16(a - ha, b - hb);
}
__device__ __forceinline__ void split4(const float (&v)[4], uint2& hi, uint2& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
}
// Packed fp32 FMA (sm_100 FFMA2): two lanes of fp32 FMA in ONE issue slot -- the gather loops are issue-bound
struct F2 {
  unsigned long long v;
};
__device__ __forceinline__ F2 f2_pack(float a, float b) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(F2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ F2 f2_fma(F2 w, F2 v, F2 c) {
  F2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(w.v), "l"(v.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
  F2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
  F2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) {
  F2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ void red_add_v4(float* p, F2 ab, F2 cd) {
  float a, b, c, d;
  f2_unpack(ab, a, b);
  f2_unpack(cd, c, d);
  red_add_v4(p, a, b, c, d);
}
// predicated form: one instruction slot, no branch around it
__device__ __forceinline__ void red_add_v4_if(int pred, float* p, F2 ab, F2 cd) {
  float a, b, c, d;
  f2_unpack(ab, a, b);
  f2_unpack(cd, c, d);
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p),
      "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred)
      : "memory");
}

// byte offset of the 16-byte chunk `c16` of row `r` inside a 128-byte-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t swz128(uint32_t r, uint32_t c16) { return r * 128u + ((c16 ^ (r & 7u)) << 4); }

}  // namespace d2b_tc
