// ikr_backward_tc.cuh -- backward sweep (discrete adjoint) on the tensor cores.
//
//   ikr_adjoint_tc_kernel<S, G>   same lane adjoint machine as ikr_adjoint_kernel (ikr_math.h bdp_*),
//       128-trajectory tiles, thread <-> TMEM lane.  Per reversed stage the tile MLP runs forward
//       (L layer MMAs, LeakyReLU sign bits kept in shared memory) and backward (L layer MMAs with the
//       transposed weight image: dz_{l-1} = (dz_l W_l) * leaky'(H_{l-1})), all bf16x3 split like the
//       forward kernel.  What the weight gradients need goes to a global STASH, already in the
//       operand format of the weight-gradient GEMM: per evaluation slot the matrices dz_0..dz_L,
//       Hc_0..Hc_L (H with a constant-1 feature at index n, so bias gradients fall out of the same
//       GEMM), X = (nv, a, 1) and U = (up), each as TWO bf16 terms in MN-major core-matrix order
//       (8 samples x 8 features per 128-byte core matrix; a thread writes 16-byte pieces and 8
//       neighbouring lanes fill one 128-byte line).  The first two terms of the A operand ARE that
//       split, so the stash costs no extra arithmetic for the hidden layers.
//   ikr_wgrad_tc_kernel           dW_l = dz_l^T Hc_{l-1} (l = 1..L), [dw0 | db0] = dz_0^T X,
//       [dw_last; db_last] = Hc_L^T U over the stash: tcgen05.mma with both operands MN-major in
//       shared memory (K = samples), three MMAs per K = 16 step (d1 h1 + d1 h2 + d2 h1: products to
//       2^-16, unbiased), fp32 accumulation in TMEM over the CTA's share of the slots, fp64 partials.
//       HBM-bound by design (the stash is written once and read once).
//   ikr_grad_reduce_tc_kernel     partials -> flat state_dict-ordered fp64 gradient.
#ifndef IKR_BACKWARD_TC_CUH_
#define IKR_BACKWARD_TC_CUH_

#include "ikr_backward.cuh"
#include "ikr_forward_tc.cuh"

namespace ikr {

// ---- stash geometry ------------------------------------------------------------------------------------
struct TcStashGeom {
  int L, n, NP, NGb;            // NGb = NP / 8 feature groups of a big matrix
  long long big_term, big;      // bytes of one bf16 term / of both terms of a big matrix
  long long small_term, small_; // X and U: 2 feature groups
  long long off_x, off_u, slot; // byte offsets inside a slot; slot size
};
__host__ __device__ inline TcStashGeom tc_stash_geometry(const TcGeom& g) {
  TcStashGeom s;
  s.L = g.L; s.n = g.n; s.NP = g.NP; s.NGb = g.NP / 8;
  s.big_term = (long long)s.NGb * 2048;      // 8 K-steps x 2 k-groups x NGb x 128 B
  s.big = 2 * s.big_term;
  s.small_term = 2 * 2048;
  s.small_ = 2 * s.small_term;
  s.off_x = 2LL * (g.L + 1) * s.big;
  s.off_u = s.off_x + s.small_;
  s.slot = s.off_u + s.small_;
  return s;
}
__host__ __device__ inline long long tc_stash_dz(const TcStashGeom& s, int l) { return (long long)l * s.big; }
__host__ __device__ inline long long tc_stash_h(const TcStashGeom& s, int l) { return (long long)(s.L + 1 + l) * s.big; }
// byte offset of sample `lane` inside a term image with NG feature groups (add 128 * group)
__host__ __device__ inline long long tc_stash_sample(int lane, int NG) {
  return (long long)(lane >> 4) * (2LL * NG * 128) + (long long)((lane >> 3) & 1) * (NG * 128) + (lane & 7) * 16;
}
// the tensor-core backward needs a spare feature slot for the constant-1 column and the tail layout
__host__ __device__ inline bool tc_backward_ok(const TcGeom& g) { return tc_geometry_ok(g) && g.tail == 1; }

struct TcAdjParams {
  BwdParams b;           // b.M = trajectories per tile (<= 128); other tile geometry fields unused
  TcGeom g;
  TcStashGeom sg;
  const void* img;       // [2L] layer images: forward W_1..W_L, then backward W_L..W_1 (transposed)
  unsigned char* stash;  // [slots][sg.slot]
  int mask_words;        // sign-bit words per thread per layer = units per group + 1
  int timing;            // debug: block 0 prints its phase clocks (IKR_TC_TIMING=1)
};

template <typename S>
struct TcAdjSmemLayout {
  size_t off_bar, off_misc, off_lanes, off_xin, off_part, off_mask, off_sp, off_ring, total;
  __host__ __device__ TcAdjSmemLayout(const TcGeom& g, int stages, int G, int mask_words) {
    size_t o = 0;
    off_bar = o; o += (size_t)(2 * kTcMaxStages + 2 * 8 + 2) * 8;
    off_misc = o; o += 48;                                     // tmem base, stop, tile, slot, cmd
    off_lanes = o; o += (size_t)kTcM * sizeof(BLane<S>); o = (o + 15) & ~(size_t)15;
    off_xin = o; o += (size_t)kTcM * 4 * sizeof(float);
    off_part = o; o += (size_t)G * kTcM * sizeof(float);
    off_mask = o; o += (size_t)g.L * mask_words * 128 * G * sizeof(uint32_t);
    off_sp = o; o += (size_t)g.small_elems * sizeof(float); o = (o + 127) & ~(size_t)127;
    off_ring = o; o += (size_t)stages * g.ring_stride;
    total = o;
  }
};

// per-thread context of one adjoint evaluation
struct TcAdjLane {
  unsigned char* slot;     // stash slot base
  long long samp;          // tc_stash_sample(lane, NGb)
  uint32_t* mask;          // shared memory: word (l * mask_words + w) * nthreads + tid
  int mask_words, nthreads, tid;
};

__device__ __forceinline__ void tc_stash_words(unsigned char* dst, const uint32_t* w) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}
// both terms of NW consecutive feature groups (4 packed words each) of one big matrix
template <int NGRP>
__device__ __forceinline__ void tc_stash_groups(const TcStashGeom& sg, const TcAdjLane& al, long long mat_off,
                                                int group0, const uint32_t* t1, const uint32_t* t2) {
  unsigned char* base = al.slot + mat_off + al.samp + (long long)group0 * 128;
#pragma unroll
  for (int q = 0; q < NGRP; ++q) {
    tc_stash_words(base + q * 128, t1 + 4 * q);
    tc_stash_words(base + sg.big_term + q * 128, t2 + 4 * q);
  }
}

// ---- adjoint epilogue work units ---------------------------------------------------------------------
// Sign bookkeeping: LeakyReLU keeps the sign of its argument, and so does the bf16 rounding of the
// first split term, so the "negative" flags of a pair (features 2q, 2q + 1) are bits 15 and 31 of the
// packed term-1 word; word q of a unit contributes them as bits q and 16 + q of the unit's mask word.
__device__ __forceinline__ uint32_t tc_sign_bits(uint32_t bits, uint32_t w1, int q) {
  return bits | ((w1 >> (15 - q)) & (0x00010001u << q));
}
// (1 or slope) factors of pair q from a mask word
__device__ __forceinline__ tc::f32x2_t tc_leaky_grad2(uint32_t bits, int q, float slope) {
  return tc::p2((bits & (1u << q)) ? slope : 1.0f, (bits & (0x10000u << q)) ? slope : 1.0f);
}

// MODE 0: forward hidden/first layer: v + bias -> LeakyReLU -> sign bits, A operand, stash Hc_l
// MODE 1: forward LAST layer (H_L): stash Hc_L; dz_L = up w_last leaky'(H_L) -> A operand, stash dz_L
// MODE 2: backward layer: dz = v * leaky'(sign bits) -> A operand, stash dz
// MODE 3: backward FIRST layer (dz_0): stash dz_0, reduce dz_0 . w0b
// LITE: see tc_mma_warp -- units that feed a three-product pass carry no third term
template <int NK, int MODE, int LITE = 0>
__device__ __forceinline__ void tc_adj_unit(const TcGeom& g, const TcStashGeom& sg, const TcLane& tl,
                                            const TcAdjLane& al, int u, unsigned gi, int w_idx, int layer,
                                            const float* bias, uint32_t (&v)[16 * NK], float up,
                                            float& acc) {
  constexpr int NP2 = 8 * NK;       // pairs
  const int c0 = 32 * u;
  const float slope = tl.slope;
  const tc::f32x2_t slope2 = tc::p2(slope, slope);
  uint32_t* mword = al.mask + ((size_t)layer * al.mask_words + w_idx) * al.nthreads + al.tid;
  tc::f32x2_t h[NP2];
  uint32_t w12[2 * NP2], w3[NP2];
  uint32_t bits = 0;
  if (MODE == 0 || MODE == 1) {
#pragma unroll
    for (int q = 0; q < NP2 / 2; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
      h[2 * q] = tc::leaky2(tc::add2(tc::p2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                                     tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::add2(tc::p2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])),
                                         tc::p2(bb.z, bb.w)), slope2);
    }
#pragma unroll
    for (int q = 0; q < NP2; ++q) {
      tc::split3p(h[q], w12[q], w12[NP2 + q], w3[q]);
      bits = tc_sign_bits(bits, w12[q], q);
    }
    if (MODE == 0) {
      *mword = bits;
    } else {
      // H_L only feeds the output layer: stash it (two terms), then form dz_L = up w_last leaky'(H_L)
      tc_stash_groups<2 * NK>(sg, al, tc_stash_h(sg, g.L), 4 * u, w12, w12 + NP2);
      const float* wl = tl.sp + (size_t)(3 + g.L) * g.NP + c0;
      const tc::f32x2_t up2 = tc::p2(up, up);
#pragma unroll
      for (int q = 0; q < NP2 / 2; ++q) {
        const float4 ww = *reinterpret_cast<const float4*>(wl + 4 * q);
        h[2 * q] = tc::mul2(tc::mul2(tc::p2(ww.x, ww.y), up2), tc_leaky_grad2(bits, 2 * q, slope));
        h[2 * q + 1] = tc::mul2(tc::mul2(tc::p2(ww.z, ww.w), up2), tc_leaky_grad2(bits, 2 * q + 1, slope));
      }
#pragma unroll
      for (int q = 0; q < NP2; ++q) tc::split3p(h[q], w12[q], w12[NP2 + q], w3[q]);
    }
  } else {
    bits = *mword;
#pragma unroll
    for (int q = 0; q < NP2; ++q) {
      h[q] = tc::mul2(tc::p2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])),
                      tc_leaky_grad2(bits, q, slope));
      tc::split3p(h[q], w12[q], w12[NP2 + q], w3[q]);
    }
  }
  if (MODE != 3) {
    tc_unit_acquire<tc_ring_terms<3, LITE>()>(g, tl, u);
    const uint32_t dst = tl.taddr + tc_unit_slot_col<tc_ring_terms<3, LITE>()>(g, tl, u);
    constexpr bool third = LITE == 0 || (LITE == 1 && MODE == 0);
    if (NK == 2) {
      tc::st32(dst, reinterpret_cast<uint32_t(&)[32]>(w12));
      if (third) tc::st16(dst + 32, reinterpret_cast<uint32_t(&)[16]>(w3));
    } else {
      tc::st16(dst, reinterpret_cast<uint32_t(&)[16]>(w12));
      if (third) tc::st8(dst + 16, reinterpret_cast<uint32_t(&)[8]>(w3));
    }
    tc_unit_publish(g, tl, u);
  } else {
    const float* w0b = tl.sp + g.NP + c0;
#pragma unroll
    for (int q = 0; q < NP2; ++q) {
      float d0, d1;
      tc::u2(h[q], d0, d1);
      acc = __fmaf_rn(d0, w0b[2 * q], acc);
      acc = __fmaf_rn(d1, w0b[2 * q + 1], acc);
    }
  }
  // stash: MODE 0 -> Hc_layer, MODE 1 -> dz_L, MODE 2 -> dz_{layer}, MODE 3 -> dz_0
  const long long mat = MODE == 0 ? tc_stash_h(sg, layer)
                                  : tc_stash_dz(sg, MODE == 1 ? g.L : (MODE == 2 ? layer : 0));
  tc_stash_groups<2 * NK>(sg, al, mat, 4 * u, w12, w12 + NP2);
}

// the 8 tail features (feature group 2 KSf) -- same four modes; also writes the constant-1 feature of Hc
template <int MODE, int LITE = 0>
__device__ __forceinline__ void tc_adj_tail(const TcGeom& g, const TcStashGeom& sg, const TcLane& tl,
                                            const TcAdjLane& al, unsigned gi, int w_idx, int layer,
                                            const float* bias, uint32_t (&v)[8], float up, float& acc) {
  const int c0 = 16 * g.KSf;
  const int grp = 2 * g.KSf;
  const float slope = tl.slope;
  const tc::f32x2_t slope2 = tc::p2(slope, slope);
  uint32_t* mword = al.mask + ((size_t)layer * al.mask_words + w_idx) * al.nthreads + al.tid;
  tc::f32x2_t h[4];
  uint32_t t1[4], t2[4], t3[4];
  uint32_t bits = 0;
  const int ones = g.n - c0;   // position of the constant-1 feature: 1..8 (8 => next feature group)
  auto stash_h = [&](int l) {
    // Hc_l tail group with the constant-1 feature (index n) patched in
    uint32_t h1[4], h2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { h1[q] = t1[q]; h2[q] = t2[q]; }
    if (ones < 8) {
      const uint32_t one = 0x3F80u << (16 * (ones & 1));
      const uint32_t keep = (ones & 1) ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q == (ones >> 1)) { h1[q] = (h1[q] & keep) | one; h2[q] &= keep; }
    }
    tc_stash_groups<1>(sg, al, tc_stash_h(sg, l), grp, h1, h2);
    if (ones == 8) {
      const uint32_t o1[4] = {0x3F80u, 0u, 0u, 0u}, o2[4] = {0u, 0u, 0u, 0u};
      tc_stash_groups<1>(sg, al, tc_stash_h(sg, l), grp + 1, o1, o2);
    }
  };
  if (MODE == 0 || MODE == 1) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
      h[2 * q] = tc::leaky2(tc::add2(tc::p2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                                     tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::add2(tc::p2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])),
                                         tc::p2(bb.z, bb.w)), slope2);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      tc::split3p(h[q], t1[q], t2[q], t3[q]);
      bits = tc_sign_bits(bits, t1[q], q);
    }
    if (MODE == 0) {
      *mword = bits;
    } else {
      stash_h(g.L);
      const float* wl = tl.sp + (size_t)(3 + g.L) * g.NP + c0;
      const tc::f32x2_t up2 = tc::p2(up, up);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = tc::mul2(tc::mul2(tc::p2(wl[2 * q], wl[2 * q + 1]), up2), tc_leaky_grad2(bits, q, slope));
        tc::split3p(h[q], t1[q], t2[q], t3[q]);
      }
    }
  } else {
    bits = *mword;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h[q] = tc::mul2(tc::p2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])),
                      tc_leaky_grad2(bits, q, slope));
      tc::split3p(h[q], t1[q], t2[q], t3[q]);
    }
  }
  if (MODE != 3) {
    const uint32_t t[16] = {t1[0], t1[1], t1[2], t1[3], t2[0], t2[1], t2[2], t2[3],
                            t1[0], t1[1], t1[2], t1[3], t3[0], t3[1], t3[2], t3[3]};
    tc_unit_acquire<tc_ring_terms<3, LITE>()>(g, tl, g.units);
    tc::st16(tl.taddr + tc_unit_slot_col<tc_ring_terms<3, LITE>()>(g, tl, g.units), t);
    tc_unit_publish(g, tl, g.units);
  } else {
    const float* w0b = tl.sp + g.NP + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float d0, d1;
      tc::u2(h[q], d0, d1);
      acc = __fmaf_rn(d0, w0b[2 * q], acc);
      acc = __fmaf_rn(d1, w0b[2 * q + 1], acc);
    }
  }
  if (MODE == 0) stash_h(layer);
  else tc_stash_groups<1>(sg, al, tc_stash_dz(sg, MODE == 1 ? g.L : (MODE == 2 ? layer : 0)), grp, t1, t2);
}

// One forward + backward MLP evaluation of the tile (every lane thread of every group).
// xin[lane] = (nv, a, up, -); result: partial sums of up * d net / d a in tl.part.
template <int G, int LITE = 0, typename Hook = TcNoHook>
__device__ __forceinline__ void tc_adj_eval(const TcGeom& g, const TcStashGeom& sg, TcLane& tl,
                                            const TcAdjLane& al, Hook hook = Hook()) {
  const int NP = g.NP;
  const float4 in = *reinterpret_cast<const float4*>(tl.xin + 4 * tl.lane);
  const float nv = in.x, a = in.y, up = in.z;
  const int UT = g.units + g.tail;                 // tc_backward_ok: there is a tail unit
  const int w_tail = al.mask_words - 1;
  float acc = 0.0f;
  long long c0 = clock64();

  // X = (nv, a, 1) and U = (up): two feature groups each, only the first one carries data
  if (tl.group == 0) {
    uint32_t x1[4], x2[4], x3[4];
    tc::split3(nv, a, x1[0], x2[0], x3[0]);
    x1[1] = 0x3F80u; x2[1] = 0u;
    x1[2] = x1[3] = x2[2] = x2[3] = 0u;
    unsigned char* px = al.slot + sg.off_x + tc_stash_sample(tl.lane, 2);
    tc_stash_words(px, x1);
    tc_stash_words(px + sg.small_term, x2);
    uint32_t u1[4], u2[4];
    tc::split3(up, 0.0f, u1[0], u2[0], x3[0]);
    u1[1] = u1[2] = u1[3] = u2[1] = u2[2] = u2[3] = 0u;
    unsigned char* pu = al.slot + sg.off_u + tc_stash_sample(tl.lane, 2);
    tc_stash_words(pu, u1);
    tc_stash_words(pu + sg.small_term, u2);
  }

  // ---- layer 0 forward ------------------------------------------------------------------------------
  {
    const float* b0 = tl.sp + 2 * NP;
    for (int u = tl.group; u < UT; u += G) {
      const unsigned gi = tl.unit_idx + (unsigned)u;
      const int w = (u - tl.group) / G;
      if (u < g.units) {
        if (2 * u + 1 < g.KSf) {
          uint32_t v[32];
          tc_layer0_sums<32>(tl, NP, 32 * u, nv, a, v);
          tc_adj_unit<2, 0, LITE>(g, sg, tl, al, u, gi, w, 0, b0, v, up, acc);
        } else {
          uint32_t v[16];
          tc_layer0_sums<16>(tl, NP, 32 * u, nv, a, v);
          tc_adj_unit<1, 0, LITE>(g, sg, tl, al, u, gi, w, 0, b0, v, up, acc);
        }
      } else {
        uint32_t v[8];
        tc_layer0_sums<8>(tl, NP, 16 * g.KSf, nv, a, v);
        tc_adj_tail<0, LITE>(g, sg, tl, al, gi, w_tail, 0, b0, v, up, acc);
      }
    }
    tc_pass_advance<tc_ring_terms<3, LITE>()>(tl, UT);
    hook();      // owners: time-only terms of the next reversed stage, under the first layer's MMAs
  }
  // ---- hidden layers forward (l = 1..L; H_l has sign-bit row l, the last one is consumed at once) ----
  for (int l = 1; l <= g.L; ++l) {
    const float* bias = tl.sp + (size_t)(2 + l) * NP;
    const bool last = l == g.L;
    { const long long c1 = clock64(); tl.c_epi += c1 - c0; c0 = c1; }
    const uint32_t dcol = tc_wait_d(g, tl);
    { const long long c1 = clock64(); tl.c_wait += c1 - c0; c0 = c1; }
    for (int u = tl.group; u < UT; u += G) {
      const unsigned gi = tl.unit_idx + (unsigned)u;
      const int w = (u - tl.group) / G;
      if (u < g.units) {
        if (2 * u + 1 < g.KSf) {
          uint32_t v[32];
          tc::ld32(tl.taddr + dcol + 32 * u, v);
          tc::wait_ld();
          if (!last) tc_adj_unit<2, 0, LITE>(g, sg, tl, al, u, gi, w, l, bias, v, up, acc);
          else tc_adj_unit<2, 1, LITE>(g, sg, tl, al, u, gi, w, l, bias, v, up, acc);
        } else {
          uint32_t v[16];
          tc::ld16(tl.taddr + dcol + 32 * u, v);
          tc::wait_ld();
          if (!last) tc_adj_unit<1, 0, LITE>(g, sg, tl, al, u, gi, w, l, bias, v, up, acc);
          else tc_adj_unit<1, 1, LITE>(g, sg, tl, al, u, gi, w, l, bias, v, up, acc);
        }
      } else {
        uint32_t v[8];
        tc::ld8(tl.taddr + dcol + 16 * g.KSf, v);
        tc::wait_ld();
        if (!last) tc_adj_tail<0, LITE>(g, sg, tl, al, gi, w_tail, l, bias, v, up, acc);
        else tc_adj_tail<1, LITE>(g, sg, tl, al, gi, w_tail, l, bias, v, up, acc);
      }
    }
    tc_pass_advance<tc_ring_terms<3, LITE>()>(tl, UT);
  }
  // ---- backward: D = dz_l W_l  ->  dz_{l-1} = D * leaky'(H_{l-1}) ---------------------------------------
  for (int l = g.L; l >= 1; --l) {
    { const long long c1 = clock64(); tl.c_epi += c1 - c0; c0 = c1; }
    const bool first = l == 1;
    const uint32_t dcol = tc_wait_d(g, tl, !first);
    { const long long c1 = clock64(); tl.c_wait += c1 - c0; c0 = c1; }
    for (int u = tl.group; u < UT; u += G) {
      const unsigned gi = tl.unit_idx + (unsigned)u;
      const int w = (u - tl.group) / G;
      if (u < g.units) {
        if (2 * u + 1 < g.KSf) {
          uint32_t v[32];
          tc::ld32(tl.taddr + dcol + 32 * u, v);
          tc::wait_ld();
          if (!first) tc_adj_unit<2, 2, LITE>(g, sg, tl, al, u, gi, w, l - 1, nullptr, v, up, acc);
          else tc_adj_unit<2, 3, LITE>(g, sg, tl, al, u, gi, w, 0, nullptr, v, up, acc);
        } else {
          uint32_t v[16];
          tc::ld16(tl.taddr + dcol + 32 * u, v);
          tc::wait_ld();
          if (!first) tc_adj_unit<1, 2, LITE>(g, sg, tl, al, u, gi, w, l - 1, nullptr, v, up, acc);
          else tc_adj_unit<1, 3, LITE>(g, sg, tl, al, u, gi, w, 0, nullptr, v, up, acc);
        }
      } else {
        uint32_t v[8];
        tc::ld8(tl.taddr + dcol + 16 * g.KSf, v);
        tc::wait_ld();
        if (!first) tc_adj_tail<2, LITE>(g, sg, tl, al, gi, w_tail, l - 1, nullptr, v, up, acc);
        else tc_adj_tail<3, LITE>(g, sg, tl, al, gi, w_tail, 0, nullptr, v, up, acc);
      }
    }
    if (!first) tc_pass_advance<tc_ring_terms<3, LITE>()>(tl, UT);
  }
  tl.part[tl.group * kTcM + tl.lane] = acc;
  { const long long c1 = clock64(); tl.c_epi += c1 - c0; }
}

// Owner-side wrapper of one adjoint evaluation: publish (nv, a, up), take a stash slot, run, collect.
template <int G, int LITE = 0, typename Hook = TcNoHook>
__device__ __forceinline__ float tc_adj_owner_eval(const TcGeom& g, const TcStashGeom& sg, TcLane& tl,
                                                   TcAdjLane& al, unsigned char* stash,
                                                   volatile long long* stash_slot,
                                                   unsigned long long* counters, bool act, double nv,
                                                   double ain, double up, Hook hook = Hook()) {
  const int tid = tl.lane;
  *reinterpret_cast<float4*>(tl.xin + 4 * tid) =
      make_float4(act ? (float)nv : 0.0f, act ? (float)ain : 0.0f, act ? (float)up : 0.0f, 0.0f);
  if (tid == 0) *stash_slot = (long long)atomicAdd(&counters[1], 1ULL);
  if (G > 1) lanes_sync<G>(); else owners_sync();
  al.slot = stash + (size_t)(*stash_slot) * sg.slot;
  tc_adj_eval<G, LITE, Hook>(g, sg, tl, al, hook);
  if (G > 1) lanes_sync<G>(); else owners_sync();
  float da = tl.part[tid];
#pragma unroll
  for (int c = 1; c < G; ++c) da += tl.part[c * kTcM + tid];
  return da;
}

template <typename S, int G, int LITE = 0>
__global__ void __launch_bounds__(tc_threads(G), 1) ikr_adjoint_tc_kernel(const TcAdjParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef typename Vec2<S>::type V2;
  const BwdParams& p = tp.b;
  const TcGeom g = tp.g;
  const TcStashGeom sg = tp.sg;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kLaneThreads = 128 * G;
  constexpr int kMmaWarp = 4 * G, kLoadWarp = 4 * G + 1;
  const TcAdjSmemLayout<S> lay(g, g.stages, G, tp.mask_words);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  volatile long long* tile_slot = reinterpret_cast<volatile long long*>(smem_raw + lay.off_misc + 8);
  volatile long long* stash_slot = reinterpret_cast<volatile long long*>(smem_raw + lay.off_misc + 16);
  volatile int* cmd_exit = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 24);
  BLane<S>* lanes = reinterpret_cast<BLane<S>*>(smem_raw + lay.off_lanes);
  float* sp = reinterpret_cast<float*>(smem_raw + lay.off_sp);
  const TcEngineCtx eng = tc_engine_ctx(bars, stop_flag, smem_raw + lay.off_ring);

  if (tid == 0) {
    tc_engine_init(eng, g.stages);
    *cmd_exit = 0;
  }
  if (warp == kMmaWarp) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  {
    const float* P = reinterpret_cast<const float*>(p.mlp.base);
    const int NP = g.NP, npad = p.mlp.npad, n = g.n;
    for (int i = tid; i < g.small_elems; i += tc_threads(G)) {
      const int row = i / NP, c = i - row * NP;
      float v = 0.0f;
      if (row < 3) { if (c < n) v = P[p.mlp.off_w0 + (long long)row * npad + c]; }
      else if (row < 3 + g.L) { if (c < n) v = P[p.mlp.off_bh + (long long)(row - 3) * npad + c]; }
      else if (row == 3 + g.L) { if (c < n) v = P[p.mlp.off_wl + c]; }
      else if (i == (4 + g.L) * NP) v = P[p.mlp.off_wl + npad];
      sp[i] = v;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kMmaWarp) {
    tc_mma_warp<3, LITE>(g, eng, tbase, tp.timing && blockIdx.x == 0);
  } else if (warp == kLoadWarp) {
    if ((tid & 31) == 0)
      tc_producer_thread(g, eng, reinterpret_cast<const unsigned char*>(tp.img), (unsigned)(2 * g.L * g.KST),
                         LITE == 2 ? 2u * (unsigned)g.block_bytes : 0u);
  } else {
    TcLane tl;
    tl.group = warp >> 2;
    tl.lane = tid & 127;
    tl.taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    tc_lane_attach(tl, eng);
    tl.sp = sp;
    tl.xin = reinterpret_cast<float*>(smem_raw + lay.off_xin);
    tl.part = reinterpret_cast<float*>(smem_raw + lay.off_part);
    tl.slope = (float)p.mlp.slope;
    tl.c_l0 = tl.c_wait = tl.c_epi = 0;
    tl.c_sync_a = tl.c_sync_b = tl.c_last = 0;
    const long long c_begin = clock64();
    TcAdjLane al;
    al.samp = tc_stash_sample(tl.lane, sg.NGb);
    al.mask = reinterpret_cast<uint32_t*>(smem_raw + lay.off_mask);
    al.mask_words = tp.mask_words;
    al.nthreads = kLaneThreads;
    al.tid = tid;
    al.slot = nullptr;

    if (tl.group > 0) {
      while (true) {
        lanes_sync<G>();
        if (*cmd_exit) break;
        al.slot = tp.stash + (size_t)(*stash_slot) * sg.slot;
        tc_adj_eval<G, LITE>(g, sg, tl, al);
        lanes_sync<G>();
      }
    } else {
      const SolverCfg cfg = p.cfg;
      const bool rk4 = p.method == 1;
      const long long jB = p.B;
      const int T = p.T;
      const S* y0 = reinterpret_cast<const S*>(p.y0);
      const S* ckpt_y = reinterpret_cast<const S*>(p.ckpt_y);
      const V2* grad_y = reinterpret_cast<const V2*>(p.grad_y);
      const V2* y_out = reinterpret_cast<const V2*>(p.y_out);
      const S* gptr = reinterpret_cast<const S*>(p.g);
      const S* eptr = reinterpret_cast<const S*>(p.e_rev);
      const S* dptr = reinterpret_cast<const S*>(p.data);
      BLaneSave<S>* saved = reinterpret_cast<BLaneSave<S>*>(p.lane_state);

      while (true) {
        if (tid == 0) *tile_slot = (long long)atomicAdd(&p.counters[0], 1ULL);
        owners_sync();
        const long long tile = *tile_slot;
        if (tile >= p.n_tiles) break;
        const long long b = tile * p.M + tid;        // p.M <= 128 trajectories per tile
        const bool valid = tid < p.M && b < jB;
        S g_b = (S)1, e_b = (S)p.e_scalar;
        BLane<S>& L = lanes[tid];
        if (p.first_round) {
          const bool ok = valid && p.stats[4 * b + 3] == 0;
          blane_reset<S>(L, ok ? p.stats[4 * b] : 0, T, ok);
        } else {
          const BLaneSave<S> sv = saved[tile * kTcM + tid];
          blane_reset<S>(L, 0, T, false);
          L.lya = sv.lya; L.lyr = sv.lyr; L.lfa = sv.lfa; L.lfr = sv.lfr; L.gsum = sv.gsum;
          L.n_left = sv.n_left; L.out_idx = sv.out_idx; L.phase = sv.phase;
        }
        if (valid) {
          if (gptr) g_b = gptr[b];
          if (eptr) e_b = eptr[b];
        }

        auto grad = [&](int idx, S* ga, S* gr) {
          if (p.fused_loss == 0) {
            const V2 v = grad_y[(size_t)idx * jB + b];
            *ga = v.x; *gr = v.y;
          } else {
            const V2 y = y_out[(size_t)idx * jB + b];
            const double vm = p.v_out[idx] - (double)e_b;
            const double cur = (double)(g_b * y.x * y.y) * vm;
            const double d = (double)dptr[(size_t)idx * p.data_B + (p.data_B == 1 ? 0 : b)];
            const double diff = cur - d;
            const double w = p.fused_loss == 1 ? 2.0 * diff : (diff > 0 ? 1.0 : (diff < 0 ? -1.0 : 0.0));
            *ga = (S)(w * vm * (double)(g_b * y.y));
            *gr = (S)(w * vm * (double)(g_b * y.x));
            L.gsum = L.gsum + (S)(w * vm * (double)(y.x * y.y));
          }
        };
        TimeCache tcache;
        tcache.valid = 0;
        for (int r = 0; r < p.steps_per_round; ++r) {
          const bool act = L.phase == 0;
          if (!owners_or(act ? 1 : 0)) break;
          if (act) {
            const size_t o = (size_t)(L.n_left - 1) * jB + b;
            const double2 tt = *reinterpret_cast<const double2*>(p.ckpt_t + 2 * o);
            S ck[kCkptVals];
            const V2* src = reinterpret_cast<const V2*>(ckpt_y + (size_t)kCkptVals * o);
#pragma unroll
            for (int i = 0; i < kCkptVals / 2; ++i) {
              const V2 v = src[i];
              ck[2 * i] = v.x; ck[2 * i + 1] = v.y;
            }
            blane_load_step<S>(L, tt.x, tt.y, ck);
            if (rk4) brk4_seed_step<S>(L, p.t_out, p.time_f32 != 0, grad);
            else bdp_seed_step<S>(L, p.t_out, grad);
          }
#pragma unroll 1
          for (int s = (rk4 ? 3 : 5); s >= 0; --s) {
            double nv = 0, ain = 0, up = 0;
            if (act) {
              if (rk4) brk4_stage_inputs<S>(L, cfg, s, p.time_f32 != 0, p.rk4_perturb != 0, &nv, &ain, &up);
              else bdp_stage_inputs_cached<S>(L, cfg, s, &nv, &ain, &up, tcache);
            }
            const float da = tc_adj_owner_eval<G, LITE>(g, sg, tl, al, tp.stash, stash_slot, p.counters, act, nv,
                                                  ain, up, [&]() {
                                                    if (act && s > 0 && !rk4) bdp_prefetch_stage_time<S>(L, cfg, s - 1, tcache);
                                                  });
            if (act) {
              if (rk4) brk4_reverse_stage<S>(L, s, p.time_f32 != 0, (S)da);
              else bdp_reverse_stage<S>(L, s, (S)da);
            }
          }
          if (act) {
            if (rk4) brk4_finish_step<S>(L);
            else bdp_finish_step<S>(L);
          }
        }

        // f(t[0], y0): once no lane of the tile is still reversing steps
        {
          const int p0 = L.phase == 0 ? 1 : 0;
          const int p1 = L.phase == 1 ? 1 : 0;
          const int any0 = owners_or(p0);
          const int any1 = owners_or(p1);
          if (!any0 && any1 && rk4) {
            // fixed grid: no separate f(t[0], y0) evaluation, only the first output sample
            if (p1) {
              S g0a, g0r;
              grad(0, &g0a, &g0r);
              brk4_finish<S>(L, g0a, g0r);
              if (p.grad_y0) {
                V2 v;
                v.x = L.lya; v.y = L.lyr;
                reinterpret_cast<V2*>(p.grad_y0)[b] = v;
              }
              if (p.grad_g) reinterpret_cast<S*>(p.grad_g)[b] = L.gsum;
            }
          } else if (!any0 && any1) {
            double nv = 0, ain = 0, up = 0;
            S y0a = (S)0;
            if (p1) {
              y0a = y0[2 * b];
              bdp_f0_inputs<S>(L, cfg, p.t_out[0], y0a, &nv, &ain, &up);
            }
            const float da = tc_adj_owner_eval<G, LITE>(g, sg, tl, al, tp.stash, stash_slot, p.counters, p1 != 0,
                                                  nv, ain, up);
            if (p1) {
              S g0a, g0r;
              grad(0, &g0a, &g0r);
              bdp_f0_finish<S>(L, (S)da, g0a, g0r);
              if (p.grad_y0) {
                V2 v;
                v.x = L.lya; v.y = L.lyr;
                reinterpret_cast<V2*>(p.grad_y0)[b] = v;
              }
              if (p.grad_g) reinterpret_cast<S*>(p.grad_g)[b] = L.gsum;
            }
          }
        }
        {
          BLaneSave<S> sv;
          sv.lya = L.lya; sv.lyr = L.lyr; sv.lfa = L.lfa; sv.lfr = L.lfr; sv.gsum = L.gsum;
          sv.n_left = L.n_left; sv.out_idx = L.out_idx; sv.phase = L.phase;
          saved[tile * kTcM + tid] = sv;
        }
        owners_sync();
      }
      if (tp.timing && blockIdx.x == 0 && tid == 0) {
        const long long tot = clock64() - c_begin;
        printf("[tc timing] adjoint owner 0: total %lld cycles: wait_d %lld, epilogues+layer0+stash %lld, "
               "solver+other %lld\n", tot, tl.c_wait, tl.c_epi, tot - tl.c_wait - tl.c_epi);
      }
      if (tid == 0) { *cmd_exit = 1; *stop_flag = 1; }
      owners_sync();
      if (G > 1) lanes_sync<G>();
    }
    tc_release_engines(tl);
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// =============================================================================================
// Weight-gradient GEMM over the stash (tensor cores, both operands MN-major in shared memory)
// =============================================================================================
struct TcWgradParams {
  TcGeom g;
  TcStashGeom sg;
  const unsigned char* stash;
  const unsigned long long* counters;   // [1] = slots used this round (when slots_fixed < 0)
  long long slots_fixed;                // >= 0: number of slots, known to the host
  int S;                                // splits per pseudo-layer; grid = (L + 2) * S
  int stages;
  double* partial;                      // [(L + 2)][S][NP rows][NP cols] fp64, accumulated (+=)
};
constexpr int kWgTcThreads = 192;       // 4 epilogue warps + MMA warp + producer warp
constexpr int kWgTcMaxStages = 8;

__global__ void __launch_bounds__(kWgTcThreads, 1) ikr_wgrad_tc_kernel(const TcWgradParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* bar_empty = bar_full + kWgTcMaxStages;
  uint64_t* bar_done = bar_empty + kWgTcMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 8 * (2 * kWgTcMaxStages + 1));
  unsigned char* ring = smem_raw + 256;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const TcGeom& g = p.g;
  const TcStashGeom& sg = p.sg;
  const int L = g.L;
  const int j = blockIdx.x / p.S, split = blockIdx.x % p.S;

  // operands of this pseudo-layer: A (rows of the result) and B (columns)
  long long a_off, b_off, a_term, b_term;
  int NGa = sg.NGb, NGbm, Ncols;
  if (j == 0) { a_off = tc_stash_dz(sg, 0); b_off = sg.off_x; NGbm = 2; }
  else if (j <= L) { a_off = tc_stash_dz(sg, j); b_off = tc_stash_h(sg, j - 1); NGbm = sg.NGb; }
  else { a_off = tc_stash_h(sg, L); b_off = sg.off_u; NGbm = 2; }
  a_term = sg.big_term;
  b_term = NGbm == 2 ? sg.small_term : sg.big_term;
  Ncols = NGbm * 8;
  const unsigned a_step = 2u * NGa * 128u, b_step = 2u * NGbm * 128u;
  const unsigned stage_bytes = 2 * a_step + 2 * b_step;
  const int n_blocks = NGa > 16 ? 2 : 1;
  const int grp1 = NGa - 16;              // first feature group of the second 128-row block

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  const long long slots = p.slots_fixed >= 0 ? p.slots_fixed : (long long)p.counters[1];
  const long long my_slots = slots > split ? (slots - split + p.S - 1) / p.S : 0;
  const long long n_steps = my_slots * 8;      // K = 16 steps (8 per slot)

  if (warp == 5) {
    if ((tid & 31) == 0) {
      for (long long q = 0; q < n_steps; ++q) {
        const unsigned s = (unsigned)(q % p.stages);
        if (q >= p.stages) mbar_wait(&bar_empty[s], (unsigned)(((q / p.stages) - 1) & 1));
        const long long slot = split + (q >> 3) * p.S;
        const int ks = (int)(q & 7);
        const unsigned char* src = p.stash + (size_t)slot * sg.slot;
        unsigned char* dst = ring + (size_t)s * stage_bytes;
        mbar_expect_tx(&bar_full[s], stage_bytes);
        bulk_g2s(dst, src + a_off + (size_t)ks * a_step, a_step, &bar_full[s]);
        bulk_g2s(dst + a_step, src + a_off + a_term + (size_t)ks * a_step, a_step, &bar_full[s]);
        bulk_g2s(dst + 2 * a_step, src + b_off + (size_t)ks * b_step, b_step, &bar_full[s]);
        bulk_g2s(dst + 2 * a_step + b_step, src + b_off + b_term + (size_t)ks * b_step, b_step, &bar_full[s]);
      }
    }
  } else if (warp == 4) {
    const uint32_t idesc = tc::idesc_bf16_f32_mn(128, Ncols);
    const uint32_t lbo_a = (uint32_t)NGa * 128u, lbo_b = (uint32_t)NGbm * 128u;
    for (long long q = 0; q < n_steps; ++q) {
      const unsigned s = (unsigned)(q % p.stages);
      mbar_wait(&bar_full[s], (unsigned)((q / p.stages) & 1));
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t sb = smem_u32(ring + (size_t)s * stage_bytes);
        const uint64_t b1 = tc::smem_desc(sb + 2 * a_step, lbo_b, 128);
        const uint64_t b2 = tc::smem_desc(sb + 2 * a_step + b_step, lbo_b, 128);
        for (int blk = 0; blk < n_blocks; ++blk) {
          const uint32_t ga = blk ? (uint32_t)grp1 * 128u : 0u;
          const uint64_t a1 = tc::smem_desc(sb + ga, lbo_a, 128);
          const uint64_t a2 = tc::smem_desc(sb + a_step + ga, lbo_a, 128);
          const uint32_t d = tbase + (uint32_t)(blk * 256);
          tc::mma_ss(d, a1, b1, idesc, q > 0 ? 1u : 0u);
          tc::mma_ss(d, a1, b2, idesc, 1u);
          tc::mma_ss(d, a2, b1, idesc, 1u);
        }
        tc::commit(smem_u32(&bar_empty[s]));
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::commit(smem_u32(bar_done));
    __syncwarp();
  } else if (n_steps > 0) {
    // epilogue: TMEM -> fp64 partials (rows = A features, columns = B features)
    mbar_wait(bar_done, 0);
    tc::fence_after_sync();
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    double* out = p.partial + ((size_t)j * p.S + split) * (size_t)g.NP * g.NP;
    for (int blk = 0; blk < n_blocks; ++blk) {
      const int row = blk ? 8 * grp1 + tid : tid;
      const bool mine = blk ? tid >= 128 - 8 * grp1 : true;   // rows below are duplicates of block 0
      for (int c = 0; c < Ncols; c += 16) {
        uint32_t v[16];
        tc::ld16(taddr + (uint32_t)(blk * 256 + c), v);
        tc::wait_ld();
        if (mine && row < g.NP) {
          double* dst = out + (size_t)row * g.NP + c;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i] += (double)__uint_as_float(v[i]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

struct TcReduceParams {
  int L, n, NP, S;
  const double* partial;   // [(L + 2)][S][NP][NP]
  double* out;             // flat, state_dict order: w0 (n,2), b0, [W_l (n,n), b_l] x L, w_last (n), b_last
  long long n_params;
};

__global__ void ikr_grad_reduce_tc_kernel(const TcReduceParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_params) return;
  const long long n = p.n, per = n * n + n;
  const size_t mat = (size_t)p.NP * p.NP;
  auto at = [&](int j, long long row, long long col) {
    double s = 0.0;
    for (int sp = 0; sp < p.S; ++sp) s += p.partial[((size_t)j * p.S + sp) * mat + (size_t)row * p.NP + col];
    return s;
  };
  double s;
  if (i < 2 * n) s = at(0, i >> 1, i & 1);                        // w0[o][0|1]
  else if (i < 3 * n) s = at(0, i - 2 * n, 2);                    // b0
  else if (i < 3 * n + p.L * per) {
    const long long r = i - 3 * n;
    const long long l = r / per, q = r - l * per;
    if (q < n * n) s = at((int)l + 1, q / n, q % n);              // W_{l+1}[o][i]
    else s = at((int)l + 1, q - n * n, n);                        // b_{l+1}: the constant-1 column
  } else if (i < 3 * n + p.L * per + n) s = at(p.L + 1, i - 3 * n - p.L * per, 0);   // w_last
  else s = at(p.L + 1, n, 0);                                      // b_last: row of the constant-1 feature
  p.out[i] = s;
}

}  // namespace ikr
#endif  // IKR_BACKWARD_TC_CUH_
