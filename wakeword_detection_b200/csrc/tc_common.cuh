// Hand-written sm_100a tensor-core primitives (inline PTX): tcgen05.mma with operands in
// shared memory and accumulators in TMEM, TMEM allocation, tcgen05.ld/st, mbarriers.
//
// Shared-memory operand layout used everywhere in this library: K-major, no swizzle
// ("interleave").  In the PTX matrix-descriptor model a K-major operand is a grid of
// 8-row x 16-byte core matrices; row r, k-chunk c (8 fp16 values) lives at
//      start + (r/8)*SBO + (r%8)*16 + c*LBO      bytes.
// We store every operand as *chunk panels*: panel c holds the 16-byte pieces of all rows
// back to back (row r at r*16), so SBO = 128 and LBO = rows_in_buffer*16.  With SBO = 128
// the address is linear in r, which lets an MMA read a row-shifted view of a tall buffer
// (dilated taps, strided windows) just by moving the start address.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace wwb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- descriptors -------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_NONE.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// 32-bit instruction descriptor for kind::f16: fp16 A/B (K-major), fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)      // a_format = b_format = F16
         | (0u << 15) | (0u << 16)     // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17)  // n_dim
         | ((uint32_t)(M >> 4) << 24); // m_dim
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- MMA -----------------------------------------------------------------------------
// NOTE on issue cost (measured with tools/tc_latency.cu): tcgen05.mma takes its operands from
// *uniform* registers.  If the descriptors are computed inside a divergent `if (lane == 0)` the
// compiler wraps every MMA in a waterfall loop (~180 clk per MMA); computed by the whole
// converged warp and issued under elect_one() a 128xNx16 MMA costs ~45 clk (N <= 64).
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// warp-collective forms: every lane passes the same (warp-uniform) operands, one lane issues
__device__ __forceinline__ void mma_f16_ss_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             bool accumulate) {
  if (elect_one()) mma_f16_ss(tmem_d, desc_a, desc_b, idesc, accumulate);
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------
// one full warp; writes the base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}

// warp-collective: lane l of the warp gets row (32*(warp%4) + l) of the accumulator,
// 16 consecutive fp32 columns starting at `taddr` (= base + (lane_quadrant*32 << 16) + col)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive that cannot be executed before `dep` is available: used to release a buffer only after the loads whose
// results feed `dep` have completed
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep) {
  asm volatile("{\n\t.reg .b32 t;\n\tmov.b32 t, %1;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(smem_u32(bar)), "r"(dep)
               : "memory");
}
// try_wait with a suspend-time hint: without the hint the instruction returns after a few tens of
// cycles and a waiting warp turns into a spin loop that competes for issue slots with the MMA
// issuer and the epilogue warps of its scheduler (measured: 35 % of all issued instructions of
// the CRNN front kernel were polls).  With the hint the warp sleeps in hardware until the phase
// completes or ~65 us pass.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x10000u)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-defined time when the
// phase is not complete yet; polling loops over several barriers must use test_wait)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  The retry loop is kept to try_wait + counter
// (waiting warps share issue slots with working ones; a clock64 comparison per retry made the wait
// loops ~30 % of all executed instructions of the WaveNet kernel).  try_wait sleeps ~100+ clk per
// call, so 2^24 retries are seconds; then trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
#pragma unroll 1
  while (!mbar_try_wait(bar, parity)) {
#ifdef WWB_WAIT_BACKOFF_NS
    __nanosleep(WWB_WAIT_BACKOFF_NS);
#endif
    if (++spins > (1u << 24)) __trap();   // surfaces on the host as a launch failure (no printf: its call ABI costs registers)
  }
}

// ---- fp16 hi/lo split ---------------------------------------------------------------------
// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits, so three fp16
// MMAs (hi*hi + lo*hi + hi*lo) reproduce an fp32 product to ~1e-6 relative.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// Cheap split of a PAIR of values: hi = x with the mantissa truncated to 10 bits (exactly an
// fp16 value for |x| in the normal fp16 range), lo = x - hi (exact in fp32), each pair packed
// by one cvt.rn.f16x2.  |x - (hi + lo)| <= 2^-21 |x| (plus 6e-8 absolute below 6e-5).
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  __half2 h = __floats2half2_rn(ah, bh);
  __half2 l = __floats2half2_rn(a - ah, b - bh);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tc
}  // namespace wwb

namespace wwb {
namespace tc {

// arrive + expect `bytes` of asynchronous (bulk-copy) traffic on the barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA-engine bulk copy global -> shared (contiguous, 16-byte multiples); completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


}  // namespace tc
}  // namespace wwb

namespace wwb {
namespace tc {

// ---- A operand in TENSOR MEMORY -------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T : A is 128 lanes x 8 columns per k-step (16 fp16 per lane; column j holds
// k = 2j (low half) and 2j+1).  Costs N/2 clk against 32 + N/4 with A in shared memory (tools/tc_rate_probe.cu).
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
// 16 consecutive columns of this thread's TMEM lane (e.g. an fp16 A chunk: hi 8 columns, lo 8 columns)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): one issue slot for two lanes of work ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fsub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// split a pair into fp16 hi (mantissa truncated to 10 bits) and fp16 lo (= x - hi, exact in fp32)
__device__ __forceinline__ void split2(u64 v, uint32_t& hi, uint32_t& lo) {
  float a, b;
  upk(v, a, b);
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  float la, lb;
  upk(fsub2(v, pk(ah, bh)), la, lb);
  __half2 h = __floats2half2_rn(ah, bh);
  __half2 l = __floats2half2_rn(la, lb);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tc
}  // namespace wwb
