// ikr_tc.cuh -- sm_100a tensor-core (tcgen05) building blocks: TMEM allocation, TMEM <-> register
// transfers, the single-thread MMA issue with the A operand in TMEM and the B operand in shared
// memory, and the descriptor encodings.  Raw PTX only (no CUTLASS types).
//
// Layout conventions used by every caller (validated on B200 by tests/tc_probe.cu):
//   * accumulator D (fp32, M = 128): D[m][n] lives in TMEM lane m, column d_col + n;
//   * operand A (bf16, M = 128, K-major) in TMEM: A[m][k] lives in lane m, column a_col + k / 2,
//     low half-word = even k;
//   * operand B (bf16, K-major, no swizzle) in shared memory: 8 x 8 core matrices of 128
//     contiguous bytes (row n % 8 at 16-byte pitch, k % 8 at 2-byte pitch); core matrices of one
//     K = 16 step are ordered (n / 8, k / 8): SBO (8-row group pitch) = 256 B, LBO (k-half pitch)
//     = 128 B, so one MMA reads N x 32 contiguous bytes.
#ifndef IKR_TC_CUH_
#define IKR_TC_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace ikr {
namespace tc {

constexpr uint32_t kTmemCols = 512;   // whole tensor memory of the SM (one CTA per SM)

// ---- TMEM allocation (one full warp executes these) -------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst_addr),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM <-> registers: 32 lanes x 32 bit, thread i of the warp <-> TMEM lane (warp % 4) * 32 + i
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                 "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a),
               "r"(b), "r"(c), "r"(d)
               : "memory");
}

#define IKR_TC_R4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
#define IKR_TC_W4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
// Wide variants: a TMEM store costs ~140 cycles per instruction plus ~1.5 cycles per column
// (measured, tests/tc_probe.cu "tmem"), so the epilogues write 16 / 32 columns per instruction.
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : IKR_TC_W4(v, 0), IKR_TC_W4(v, 4), IKR_TC_W4(v, 8), IKR_TC_W4(v, 12), IKR_TC_W4(v, 16),
        IKR_TC_W4(v, 20), IKR_TC_W4(v, 24), IKR_TC_W4(v, 28)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16};" ::"r"(taddr),
      IKR_TC_R4(v, 0), IKR_TC_R4(v, 4), IKR_TC_R4(v, 8), IKR_TC_R4(v, 12)
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(
          taddr),
      IKR_TC_R4(v, 0), IKR_TC_R4(v, 4), IKR_TC_R4(v, 8), IKR_TC_R4(v, 12), IKR_TC_R4(v, 16),
      IKR_TC_R4(v, 20), IKR_TC_R4(v, 24), IKR_TC_R4(v, 28)
      : "memory");
}

// one lane of a converged warp (the same lane every time)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- descriptors ----------------------------------------------------------------------------------
// instruction descriptor of tcgen05.mma.kind::f16: BF16 x BF16 -> FP32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4)                      // D format F32
         | (1u << 7)                    // A format BF16
         | (1u << 10)                   // B format BF16
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}
// FP16 x FP16 -> FP32, both operands K-major (format code 0 in both operand fields)
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with both operands MN-major (the reduction index is the slow one in shared memory)
__host__ __device__ constexpr uint32_t idesc_bf16_f32_mn(int M, int N) {
  return idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}
// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all / 16)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);   // descriptor version 1 (sm_100)
}

// D[tmem] (+)= A[tmem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t d_taddr, uint32_t a_taddr, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_taddr),
      "r"(a_taddr), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_ss(uint32_t d_taddr, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_taddr),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void commit(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr)
               : "memory");
}

// ---- fp32 -> three bf16 terms (x = h + m + l up to 2^-24 relative) ----------------------------------
// Packs two values: result words hold (even element in the low half, odd element in the high half).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  w1 = pack_bf16x2(x0, x1);
  float r0 = x0 - __uint_as_float(w1 << 16);
  float r1 = x1 - __uint_as_float(w1 & 0xFFFF0000u);
  w2 = pack_bf16x2(r0, r1);
  r0 -= __uint_as_float(w2 << 16);
  r1 -= __uint_as_float(w2 & 0xFFFF0000u);
  w3 = pack_bf16x2(r0, r1);
}


// ---- packed fp32 pairs (FADD2 / FMUL2 on sm_100): halve the issue slots of the epilogues -------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t p2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void u2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// LeakyReLU of a pair for 0 <= slope <= 1: max(z, slope z)
__device__ __forceinline__ f32x2_t leaky2(f32x2_t z, f32x2_t slope2) {
  const f32x2_t s = mul2(z, slope2);
  float z0, z1, s0, s1;
  u2(z, z0, z1);
  u2(s, s0, s1);
  return p2(fmaxf(z0, s0), fmaxf(z1, s1));
}
// bf16x3 split of a pair with packed subtractions (same values as split3)
__device__ __forceinline__ void split3p(f32x2_t h, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  float a, b;
  u2(h, a, b);
  w1 = pack_bf16x2(a, b);
  f32x2_t r = sub2(h, p2(__uint_as_float(w1 << 16), __uint_as_float(w1 & 0xFFFF0000u)));
  u2(r, a, b);
  w2 = pack_bf16x2(a, b);
  r = sub2(r, p2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xFFFF0000u)));
  u2(r, a, b);
  w3 = pack_bf16x2(a, b);
}

// Cheaper exact split for the forward epilogues: first term rounded (cvt.rn), second and third
// TRUNCATED to bf16 with a byte permute.  r1 = x - t1 has <= 16 significant bits, t2 takes its top 8,
// r2 = r1 - t2 has <= 8 and is a bf16 number itself, so x = t1 + t2 + t3 exactly and |t2|, |t3| obey
// the same 2^-8 / 2^-16 bounds as the rounded split: 11 instead of 14 pipe cycles per pair.  (The
// adjoint keeps the rounded split: its first two terms feed the weight-gradient GEMM, where a
// truncation bias would add up over the samples.)
__device__ __forceinline__ void split3t(f32x2_t h, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  float a, b;
  u2(h, a, b);
  w1 = pack_bf16x2(a, b);
  f32x2_t r = sub2(h, p2(__uint_as_float(w1 << 16), __uint_as_float(w1 & 0xFFFF0000u)));
  u2(r, a, b);
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  w2 = __byte_perm(ua, ub, 0x7632);
  r = sub2(r, p2(__uint_as_float(ua & 0xFFFF0000u), __uint_as_float(ub & 0xFFFF0000u)));
  u2(r, a, b);
  w3 = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
}

__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---- fp32 -> two fp16 terms ---------------------------------------------------------------------------
// x = t1 + t2 + e with t1 = rn_f16(x), t2 = rn_f16(x - t1): |e| <= 2^-24 |x| while t2 is a normal fp16
// number (|x| >= 2^-2 after the caller's power-of-two scaling), |e| <= 2^-25 absolute below that.  With
// both operands split this way the three products t1 u1 + t2 u1 + t1 u2 carry an fp32 product to
// ~3 x 2^-24 relative -- the same bound as the six-product bf16x3 scheme -- at half the MMAs.
// The caller scales its values by a power of two (exact) so that they sit inside the fp16 range.  A
// value beyond +-65504 converts to inf in the FIRST term and the MMA turns the products of that lane
// into inf / NaN: the whole evaluation of the lane comes out non-finite and the OWNER handles it
// (tc_range_filter in ikr_forward_tc.cuh) -- a wild trial stage of the adaptive solver is rejected
// like the reference rejects it, a range violation in the physical domain ends the lane with
// IKR_TC_RANGE.  No per-value check on the hot path.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ f32x2_t unpack_f16x2(uint32_t w) {
  float lo, hi;
  asm("{\n\t.reg .b16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "cvt.f32.f16 %0, l;\n\t"
      "cvt.f32.f16 %1, h;\n\t}"
      : "=f"(lo), "=f"(hi)
      : "r"(w));
  return p2(lo, hi);
}
__device__ __forceinline__ void split2h(f32x2_t h, uint32_t& w1, uint32_t& w2) {
  float a, b;
  u2(h, a, b);
  w1 = pack_f16x2(a, b);
  const f32x2_t r = sub2(h, unpack_f16x2(w1));
  u2(r, a, b);
  w2 = pack_f16x2(a, b);
}

}  // namespace tc
}  // namespace ikr
#endif  // IKR_TC_CUH_
