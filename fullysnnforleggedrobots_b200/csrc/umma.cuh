// umma.cuh — thin inline-PTX wrappers for the sm_100a primitives the kernels use:
// mbarrier, bulk async copy (TMA engine, 1-D form), tcgen05 MMA / TMEM / commit / fences.
// No CUTLASS: descriptor formats follow the PTX ISA ("tcgen05 matrix / instruction descriptor").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace snn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Barrier waits of the pipelined kernels.  Release builds: a plain bounded wait.  -DSNN_ABLATE (tools only, a separate
// library file): the wait also accumulates its cycles into wcyc[slot] for tools/role_cycles.py.
#ifdef SNN_ABLATE
#define SNN_TWAIT(bar, par, slot)                                   \
  do {                                                              \
    const long long t0__ = clock64();                               \
    mbar_wait(bar, par);                                            \
    wcyc[slot] += static_cast<unsigned long long>(clock64() - t0__); \
  } while (0)
#define SNN_DBG_MASK(p) ((p).dbg)
#else
#define SNN_TWAIT(bar, par, slot) mbar_wait(bar, par)
#define SNN_DBG_MASK(p) 0          // the work-skipping ablation switches do not exist in the release binary
#endif

// ------------------------------------------------------------------ proxies / fences
// generic-proxy smem writes -> visible to the async proxy (MMA / bulk copy reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------ bulk async copy (UBLKCP)
// 1-D global -> shared copy executed by the TMA engine, completion signalled on an mbarrier.
// bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------ TMEM
// One full warp allocates `ncols` (power of two >= 32) columns; base address lands in *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 bit x 16 consecutive columns: thread i of the warp receives lane (base+i)'s columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptors are built from two 32-bit halves (smem_desc_lo / smem_desc_hi below): 128-byte
// swizzle, operand tiles made of 1024-byte atoms (8 rows x 128 B), version = 1 (sm_100), base_offset = 0.
// Instruction descriptor for kind::f16, A = B = bf16, D = fp32.
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4)      // D format: F32
         | (1u << 7)    // A format: BF16
         | (1u << 10)   // B format: BF16
         | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// mbarrier arrives once all tcgen05 ops issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// One lane of a converged warp (CUTLASS' elect_one_sync): keeps the enclosing control flow warp-uniform so
// that ptxas can keep tcgen05 / bulk-copy operands in uniform registers instead of emitting per-lane
// "waterfall" loops around every instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// descriptor = {hi = constant part, lo = start address >> 4 | LBO << 16}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);      // SBO, version 1, SWIZZLE_128B
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 8 spike bits (bit i = element i) -> eight bf16 {0, 1} (one 16-byte operand piece).  Pure ALU: a shared-memory
// look-up table costs one bank-conflicted LDS.128 per piece, and the kernels that expand spikes are bound by the
// shared-memory data pipe (profiles/r1g).  (t & 3) * 0x8001 & 0x10001 moves bit 0 -> bit 0 and bit 1 -> bit 16;
// * 0x3F80 turns each into bf16 1.0.
__device__ __forceinline__ uint4 expand_spike_byte(uint32_t b) {
  uint4 v;
  v.x = (((b & 3u) * 0x8001u) & 0x10001u) * 0x3F80u;
  v.y = ((((b >> 2) & 3u) * 0x8001u) & 0x10001u) * 0x3F80u;
  v.z = ((((b >> 4) & 3u) * 0x8001u) & 0x10001u) * 0x3F80u;
  v.w = ((((b >> 6) & 3u) * 0x8001u) & 0x10001u) * 0x3F80u;
  return v;
}
// 32 x 32 bit-matrix transpose across a warp: lane i passes row i (bit b = element (i, b)); lane b receives column b
// (bit i = element (i, b)).  Five block-swap stages (shuffle + two masked merges) instead of 32 ballots.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? ((x & ~m) | ((y & ~m) >> s)) : ((x & m) | ((y & m) << s));
  }
  return x;
}
// ------------------------------------------------------------------ packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2)
// Two independent IEEE fp32 operations per lane and instruction, each rounded exactly like its scalar form (so every
// result is bit-identical to the scalar code it replaces).  The time scans of the epilogues are bound by the fp32 /
// integer-multiply pipe (a three-register FFMA issues every second cycle per scheduler): pairing neighbouring
// sample-steps halves their floating-point instruction count.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_splat(float v) { return f2_pack(v, v); }
__device__ __forceinline__ float f2_lo(f32x2 v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float f2_hi(f32x2 v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_fma_ru(f32x2 a, f32x2 b, f32x2 c) {      // both lanes rounded towards +inf
  f32x2 d;
  asm("fma.rp.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ------------------------------------------------------------------ 16-bit voltage trace
// The backward pass uses a membrane voltage v in exactly three ways (actor_critic.py:154-167, 228-231): the spike flag
// v > 0.5, the surrogate window |v - 0.5| < 0.5, and — only inside that window — the value of v in the reset term.  The
// trace therefore keeps a 16-bit code per (neuron, sample-step) instead of the fp32 voltage:
//   k = ceil(65534 v) saturated to [0, 65535]
//   k == 0          v <= 0: below the window, no spike
//   k == 65535      v > 1: above the window, spike
//   1 <= k <= 65534 inside the window; k > 32767 <=> v > 0.5 EXACTLY (0.5 * 65534 is an integer and the rounding of the
//                   product is monotone); decoded as (k - 0.5) / 65534: |error| <= 2^-17.
// The spike flag is exact; the window flag is exact except for v == 1.0f and 0 < v < 2^-25 (where fp32 rounds v - 0.5 to
// -0.5): both sit on the window's edge, where the surrogate is discontinuous anyway.  Half the bytes of an fp32 trace.
constexpr float kVoltScale = 65534.f;
constexpr float kVoltInv = 1.f / 65534.f;
// both codes of a 32-bit trace word as floats {t(low half), t(high half)}, t in [0, 65535]: bytes {k.lo, k.hi, 0x00, 0x4B}
// are the float 2^23 + k (one byte permute each).  With u = t - 32767.5: spike <=> u > 0, window <=> |u| < 32767,
// 0.75 v = t * (0.75 / 65534) - 0.375 / 65534 inside the window.
__device__ __forceinline__ f32x2 volt_t2(uint32_t word) {
  return f2_add(f2_pack(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7610u)), __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7632u))),
                f2_splat(-8388608.f));
}
// two voltages -> one trace word, all on the fp32 pipe: clamp to [0, 65535 / 65534], then 2^23 + ceil(65534 v) by ONE fused
// multiply-add rounded towards +inf (exact product, no double rounding), the two low halves merged by a byte permute.
// (The saturating conversion cvt.rpi.u16.f32 gives the same codes but runs on the quarter-rate conversion pipe: forward
// layer 1 59.9 vs 57.8 us, profiles/r2k_encab.txt.)
__device__ __forceinline__ uint32_t volt_code2(float v0, float v1) {
  const float hi = 65535.f / 65534.f;
  const f32x2 k = f2_fma_ru(f2_pack(fminf(fmaxf(v0, 0.f), hi), fminf(fmaxf(v1, 0.f), hi)), f2_splat(kVoltScale), f2_splat(8388608.f));
  return __byte_perm(static_cast<uint32_t>(k), static_cast<uint32_t>(k >> 32), 0x5410u);
}

__device__ __forceinline__ float bf16_round(float x) {   // value of bf16(x) as fp32
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x));
  return __uint_as_float(r << 16);
}

}  // namespace snn
