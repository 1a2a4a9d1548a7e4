// ikr_forward_tc.cuh -- forward integration kernel with the hidden Linear(n, n) layers on the
// 5th-generation tensor cores (tcgen05, accumulators and activations in TMEM).
//
// Same per-trajectory solver as ikr_forward_kernel (ikr_math.h lane state machine, one owner
// thread per trajectory); the tile is fixed at M = 128 trajectories = the 128 TMEM lanes, and
// thread i owns TMEM lane i, so the whole MLP needs NO cross-thread exchange:
//
//   layer 0      owner thread: h = LeakyReLU(w0a nv + w0b a + b0)           -> A operand (TMEM)
//   layer 1..L   engine thread: D[128 x n] = A[128 x n] . W_l^T  (tcgen05.mma, B = W_l streamed
//                L2 -> shared memory by cp.async.bulk through a full/empty mbarrier ring)
//                owner thread: tcgen05.ld its D row, + bias, LeakyReLU      -> A operand (TMEM)
//   output       owner thread: dot(h, w_last) + b_last while reading the last D row
//
// FP32 accuracy on BF16 tensor cores: every fp32 value x is split into three bf16 terms
// x = x1 + x2 + x3 (exact to 2^-24 |x|), and each layer issues the six products whose weight is
// >= 2^-16: a1 b1, a2 b1, a3 b1, a1 b2, a2 b2, a1 b3 (dropped terms <= 2^-24 relative), all
// accumulated in fp32 in TMEM.  Measured on B200 (tests/tc_probe.cu): one 128 x 208 x 16 MMA
// = 104 cycles, i.e. 75 MMAs = 7.8 k cycles per layer against ~200 k cycles of FFMA2 work.
//
// TMEM budget (512 columns): D uses NP = roundup(n, 16) columns; each bf16 term of A uses n / 2
// columns.  For n = 200 (every shipped model): 208 + 3 x 96 (k < 192) + 16 (tail) = 512.  The last
// n % 16 <= 8 input features form a TAIL step: two A blocks [a1t | a2t], [a1t | a3t] against B
// blocks [b1t | b1t], [b2t | b2t], [b3t | b1t] give the same six products in 3 MMAs.
#ifndef IKR_FORWARD_TC_CUH_
#define IKR_FORWARD_TC_CUH_

#include "ikr_forward.cuh"
#include "ikr_tc.cuh"

namespace ikr {

constexpr int kTcM = 128;        // trajectories per tile = TMEM lanes
constexpr int kTcThreads = 160;  // 4 owner warps + 1 engine warp
constexpr int kTcMaxStages = 12;
constexpr int kTcRefillLag = 2;  // ring slot of k-step q - lag is refilled after issuing k-step q

struct TcGeom {
  int n, L, NP;      // NP = roundup(n, 16): D columns = rows of every B block
  int KSf;           // regular K = 16 steps per layer (incl. a zero-padded one when n % 16 > 8)
  int tail;          // 1: a final tail step covers the last n % 16 <= 8 features
  int KST;           // steps per layer = KSf + tail = ring stages consumed per layer
  int col_a[3];      // TMEM column of the three bf16 terms of A (relative to the allocation base)
  int col_t1, col_t2;
  int cols;          // TMEM columns used
  int block_bytes;   // one B block: NP x 16 bf16 = NP * 32 bytes
  int stage_bytes;   // one k-step: 3 blocks
  int stages;        // ring depth
  int small_elems;   // zero-padded small-parameter block: (4 + L) NP + 8 floats
};

__host__ __device__ inline TcGeom tc_geometry(int n, int L) {
  TcGeom g;
  g.n = n; g.L = L;
  g.NP = (n + 15) / 16 * 16;
  const int rem = n % 16;
  g.KSf = n / 16 + (rem > 8 ? 1 : 0);
  g.tail = (rem > 0 && rem <= 8) ? 1 : 0;
  g.KST = g.KSf + g.tail;
  int c = g.NP;
  for (int s = 0; s < 3; ++s) { g.col_a[s] = c; c += 8 * g.KSf; }
  g.col_t1 = c; g.col_t2 = c + 8;
  if (g.tail) c += 16;
  g.cols = c;
  g.block_bytes = g.NP * 32;
  g.stage_bytes = 3 * g.block_bytes;
  g.stages = 0;
  g.small_elems = (4 + L) * g.NP + 8;
  return g;
}
__host__ __device__ inline bool tc_geometry_ok(const TcGeom& g) {
  return g.n >= 16 && g.NP <= 256 && g.cols <= (int)tc::kTmemCols && g.KST >= 1;
}

// ---- weight image: [layer][k-step][block 0..2][NP x 16 bf16 in core-matrix order] -----------------
struct TcPackParams {
  const float* wn;   // [L][n][npad] rows = output features (packed-parameter section off_wn)
  int npad;
  TcGeom g;
  uint16_t* img;
};

__device__ __forceinline__ uint16_t bf16_term(float w, int term) {
  float r = w;
  uint32_t h = 0;
  for (int i = 0; i <= term; ++i) {
    h = tc::pack_bf16x2(r, 0.0f) & 0xFFFFu;
    r -= __uint_as_float(h << 16);
  }
  return (uint16_t)h;
}

__global__ void ikr_tc_pack_kernel(const TcPackParams p) {
  const TcGeom& g = p.g;
  const long long per_block = (long long)g.NP * 16;
  const long long total = (long long)g.L * g.KST * 3 * per_block;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % 16);
    const int o = (int)((i / 16) % g.NP);
    const long long blk = i / per_block;
    const int c = (int)(blk % 3);
    const int step = (int)((blk / 3) % g.KST);
    const int layer = (int)(blk / (3LL * g.KST));
    int k, term;
    if (step < g.KSf) {
      k = 16 * step + kk;
      term = c;
    } else {
      k = 16 * g.KSf + (kk & 7);
      term = c == 0 ? 0 : (c == 1 ? 1 : (kk < 8 ? 2 : 0));
    }
    float w = 0.0f;
    if (o < g.n && k < g.n) w = p.wn[((long long)layer * g.n + o) * p.npad + k];
    const long long byte = (long long)(o >> 3) * 256 + (kk >> 3) * 128 + (o & 7) * 16 + (kk & 7) * 2;
    p.img[(blk * g.block_bytes + byte) >> 1] = bf16_term(w, term);
  }
}

// ---- shared memory carve-up --------------------------------------------------------------------------
template <typename S>
struct TcSmemLayout {
  size_t off_bar, off_misc, off_job, off_lanes, off_obs, off_sp, off_ring, total;
  __host__ __device__ TcSmemLayout(const TcGeom& g, int stages) {
    size_t o = 0;
    off_bar = o; o += (size_t)(2 * kTcMaxStages + 2) * 8;          // full[], empty[], a_ready, d_ready
    off_misc = o; o += 32;                                          // tmem base, stop flag, tile slot
    off_job = o; o += (sizeof(FwdJob) + 15) & ~(size_t)15;
    off_lanes = o; o += (size_t)kTcM * sizeof(Lane<S>); o = (o + 15) & ~(size_t)15;
    off_obs = o; o += (size_t)kTcM * 2 * sizeof(double);
    off_sp = o; o += (size_t)g.small_elems * sizeof(float); o = (o + 127) & ~(size_t)127;
    off_ring = o; o += (size_t)stages * g.stage_bytes;
    total = o;
  }
};

struct TcFwdParams {
  FwdParams f;           // f.M == 128; MG / NG / n_worker_warps unused
  TcGeom g;
  const void* img;       // weight image written by ikr_tc_pack_kernel
};

// named barrier over the 128 owner threads (the engine warp never joins it)
__device__ __forceinline__ void owners_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ int owners_or(int pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, 1, 128, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(pred)
      : "memory");
  return (int)r;
}

// Per-owner view of the tensor-core MLP
struct TcLane {
  uint32_t taddr;        // TMEM address of this thread's lane, column 0 of the allocation
  uint32_t bar_a, bar_d; // shared-memory addresses of the a_ready / d_ready mbarriers
  uint64_t* bar_d_ptr;
  unsigned phase_d;      // parity of the next d_ready completion
  const float* sp;       // small parameters (stride NP): w0a | w0b | b0 | L x bias | w_last, b_last
  float slope;
};

__device__ __forceinline__ float tc_leaky(float x, float slope) { return x > 0.0f ? x : x * slope; }

// write 16 consecutive activations (features 16 j .. 16 j + 15) as the three bf16 terms of A
__device__ __forceinline__ void tc_store_step(const TcGeom& g, uint32_t taddr, int j, const float (&h)[16]) {
  uint32_t w1[8], w2[8], w3[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) tc::split3(h[2 * q], h[2 * q + 1], w1[q], w2[q], w3[q]);
  tc::st8(taddr + g.col_a[0] + 8 * j, w1);
  tc::st8(taddr + g.col_a[1] + 8 * j, w2);
  tc::st8(taddr + g.col_a[2] + 8 * j, w3);
}
// write the 8 tail activations as the blocks [a1t | a2t] and [a1t | a3t]
__device__ __forceinline__ void tc_store_tail(const TcGeom& g, uint32_t taddr, const float (&h)[8]) {
  uint32_t u1[4], u2[4], u3[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) tc::split3(h[2 * q], h[2 * q + 1], u1[q], u2[q], u3[q]);
  const uint32_t t1[8] = {u1[0], u1[1], u1[2], u1[3], u2[0], u2[1], u2[2], u2[3]};
  const uint32_t t2[8] = {u1[0], u1[1], u1[2], u1[3], u3[0], u3[1], u3[2], u3[3]};
  tc::st8(taddr + g.col_t1, t1);
  tc::st8(taddr + g.col_t2, t2);
}

__device__ __forceinline__ void tc_publish_a(const TcLane& tl) {
  tc::wait_st();
  tc::fence_before_sync();
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl.bar_a) : "memory");
}

// MLP of the trajectory owned by this thread; every owner thread of the CTA must call it
// (masked lanes included: the a_ready barrier counts 128 arrivals).
__device__ __forceinline__ float tc_mlp_eval(const TcGeom& g, TcLane& tl, float nv, float a) {
  const int NP = g.NP;
  const float slope = tl.slope;
  // ---- layer 0 ------------------------------------------------------------------------------
  {
    const float* w0a = tl.sp;
    const float* w0b = tl.sp + NP;
    const float* b0 = tl.sp + 2 * NP;
    for (int j = 0; j < g.KSf; ++j) {
      float h[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wa = *reinterpret_cast<const float4*>(w0a + 16 * j + 4 * q);
        const float4 wb = *reinterpret_cast<const float4*>(w0b + 16 * j + 4 * q);
        const float4 bb = *reinterpret_cast<const float4*>(b0 + 16 * j + 4 * q);
        h[4 * q + 0] = tc_leaky(__fmaf_rn(wb.x, a, __fmaf_rn(wa.x, nv, bb.x)), slope);
        h[4 * q + 1] = tc_leaky(__fmaf_rn(wb.y, a, __fmaf_rn(wa.y, nv, bb.y)), slope);
        h[4 * q + 2] = tc_leaky(__fmaf_rn(wb.z, a, __fmaf_rn(wa.z, nv, bb.z)), slope);
        h[4 * q + 3] = tc_leaky(__fmaf_rn(wb.w, a, __fmaf_rn(wa.w, nv, bb.w)), slope);
      }
      tc_store_step(g, tl.taddr, j, h);
    }
    if (g.tail) {
      float h[8];
      const int c0 = 16 * g.KSf;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 wa = *reinterpret_cast<const float4*>(w0a + c0 + 4 * q);
        const float4 wb = *reinterpret_cast<const float4*>(w0b + c0 + 4 * q);
        const float4 bb = *reinterpret_cast<const float4*>(b0 + c0 + 4 * q);
        h[4 * q + 0] = tc_leaky(__fmaf_rn(wb.x, a, __fmaf_rn(wa.x, nv, bb.x)), slope);
        h[4 * q + 1] = tc_leaky(__fmaf_rn(wb.y, a, __fmaf_rn(wa.y, nv, bb.y)), slope);
        h[4 * q + 2] = tc_leaky(__fmaf_rn(wb.z, a, __fmaf_rn(wa.z, nv, bb.z)), slope);
        h[4 * q + 3] = tc_leaky(__fmaf_rn(wb.w, a, __fmaf_rn(wa.w, nv, bb.w)), slope);
      }
      tc_store_tail(g, tl.taddr, h);
    }
    tc_publish_a(tl);
  }
  // ---- hidden layers: read D, bias + LeakyReLU, write the next A (or reduce the output) -----------
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  const float* wl = tl.sp + (size_t)(3 + g.L) * NP;
  for (int layer = 0; layer < g.L; ++layer) {
    const float* bias = tl.sp + (size_t)(3 + layer) * NP;
    const bool last = layer + 1 == g.L;
    mbar_wait(tl.bar_d_ptr, tl.phase_d);
    tl.phase_d ^= 1u;
    tc::fence_after_sync();
    for (int j = 0; j < g.KSf; ++j) {
      uint32_t v[16];
      tc::ld16(tl.taddr + 16 * j, v);
      tc::wait_ld();
      float h[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + 16 * j + 4 * q);
        h[4 * q + 0] = tc_leaky(__uint_as_float(v[4 * q + 0]) + bb.x, slope);
        h[4 * q + 1] = tc_leaky(__uint_as_float(v[4 * q + 1]) + bb.y, slope);
        h[4 * q + 2] = tc_leaky(__uint_as_float(v[4 * q + 2]) + bb.z, slope);
        h[4 * q + 3] = tc_leaky(__uint_as_float(v[4 * q + 3]) + bb.w, slope);
      }
      if (!last) {
        tc_store_step(g, tl.taddr, j, h);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = *reinterpret_cast<const float4*>(wl + 16 * j + 4 * q);
          s0 = __fmaf_rn(h[4 * q + 0], ww.x, s0);
          s1 = __fmaf_rn(h[4 * q + 1], ww.y, s1);
          s2 = __fmaf_rn(h[4 * q + 2], ww.z, s2);
          s3 = __fmaf_rn(h[4 * q + 3], ww.w, s3);
        }
      }
    }
    if (g.tail) {
      uint32_t v[8];
      const int c0 = 16 * g.KSf;
      tc::ld8(tl.taddr + c0, v);
      tc::wait_ld();
      float h[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
        h[4 * q + 0] = tc_leaky(__uint_as_float(v[4 * q + 0]) + bb.x, slope);
        h[4 * q + 1] = tc_leaky(__uint_as_float(v[4 * q + 1]) + bb.y, slope);
        h[4 * q + 2] = tc_leaky(__uint_as_float(v[4 * q + 2]) + bb.z, slope);
        h[4 * q + 3] = tc_leaky(__uint_as_float(v[4 * q + 3]) + bb.w, slope);
      }
      if (!last) {
        tc_store_tail(g, tl.taddr, h);
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 ww = *reinterpret_cast<const float4*>(wl + c0 + 4 * q);
          s0 = __fmaf_rn(h[4 * q + 0], ww.x, s0);
          s1 = __fmaf_rn(h[4 * q + 1], ww.y, s1);
          s2 = __fmaf_rn(h[4 * q + 2], ww.z, s2);
          s3 = __fmaf_rn(h[4 * q + 3], ww.w, s3);
        }
      }
    }
    if (!last) tc_publish_a(tl);
  }
  return ((s0 + s1) + (s2 + s3)) + wl[NP];
}

// ---- engine thread: weight ring + MMA issue -----------------------------------------------------------
struct TcEngine {
  uint64_t* full;
  uint64_t* empty;
  unsigned char* ring;
  const unsigned char* img;
  unsigned issued, consumed;   // k-steps (global sequence numbers)
  unsigned steps_per_cycle;    // L * KST
};

__device__ __forceinline__ void tc_engine_issue_copy(const TcGeom& g, TcEngine& e) {
  const unsigned q = e.issued;
  const unsigned s = q % (unsigned)g.stages;
  const unsigned src = q % e.steps_per_cycle;
  mbar_expect_tx(&e.full[s], (unsigned)g.stage_bytes);
  bulk_g2s(e.ring + (size_t)s * g.stage_bytes, e.img + (size_t)src * g.stage_bytes,
           (unsigned)g.stage_bytes, &e.full[s]);
  e.issued = q + 1;
}

// all MMAs of one hidden layer
__device__ __forceinline__ void tc_engine_layer(const TcGeom& g, TcEngine& e, uint32_t tbase, uint32_t idesc) {
  for (int j = 0; j < g.KST; ++j) {
    const unsigned q = e.consumed;
    const unsigned s = q % (unsigned)g.stages;
    mbar_wait(&e.full[s], (q / (unsigned)g.stages) & 1u);
    tc::fence_after_sync();
    const uint32_t sb = smem_u32(e.ring + (size_t)s * g.stage_bytes);
    const uint64_t b1 = tc::smem_desc(sb, 128, 256);
    const uint64_t b2 = tc::smem_desc(sb + g.block_bytes, 128, 256);
    const uint64_t b3 = tc::smem_desc(sb + 2 * g.block_bytes, 128, 256);
    if (j < g.KSf) {
      const uint32_t a1 = tbase + g.col_a[0] + 8 * j, a2 = tbase + g.col_a[1] + 8 * j,
                     a3 = tbase + g.col_a[2] + 8 * j;
      tc::mma_ts(tbase, a1, b1, idesc, j > 0 ? 1u : 0u);
      tc::mma_ts(tbase, a2, b1, idesc, 1u);
      tc::mma_ts(tbase, a3, b1, idesc, 1u);
      tc::mma_ts(tbase, a1, b2, idesc, 1u);
      tc::mma_ts(tbase, a2, b2, idesc, 1u);
      tc::mma_ts(tbase, a1, b3, idesc, 1u);
    } else {
      tc::mma_ts(tbase, tbase + g.col_t1, b1, idesc, j > 0 ? 1u : 0u);
      tc::mma_ts(tbase, tbase + g.col_t1, b2, idesc, 1u);
      tc::mma_ts(tbase, tbase + g.col_t2, b3, idesc, 1u);
    }
    tc::commit(smem_u32(&e.empty[s]));
    e.consumed = q + 1;
    if (q >= (unsigned)kTcRefillLag) {
      const unsigned qp = q - kTcRefillLag;      // its MMAs are (nearly) complete
      if (qp + g.stages == e.issued) {
        mbar_wait(&e.empty[qp % (unsigned)g.stages], (qp / (unsigned)g.stages) & 1u);
        tc_engine_issue_copy(g, e);
      }
    }
  }
}

template <typename S>
__global__ void __launch_bounds__(kTcThreads, 1) ikr_forward_tc_kernel(const TcFwdParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FwdParams& p = tp.f;
  const TcGeom g = tp.g;
  const int tid = threadIdx.x, warp = tid >> 5;
  const TcSmemLayout<S> lay(g, g.stages);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kTcMaxStages;
  uint64_t* bar_a = bars + 2 * kTcMaxStages;
  uint64_t* bar_d = bar_a + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  long long* tile_slot = reinterpret_cast<long long*>(smem_raw + lay.off_misc + 8);
  FwdJob* jobp = reinterpret_cast<FwdJob*>(smem_raw + lay.off_job);
  Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
  double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);
  float* sp = reinterpret_cast<float*>(smem_raw + lay.off_sp);
  unsigned char* ring = smem_raw + lay.off_ring;

  // ---- one-time setup ------------------------------------------------------------------------------
  if (tid == kTcM) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_a, kTcM);
    mbar_init(bar_d, 1);
    mbar_fence_init();
    *stop_flag = 0;
  }
  if (warp == 4) {
    __syncwarp();
    tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  }
  {
    // small parameters, zero-padded to stride NP
    const float* P = reinterpret_cast<const float*>(p.mlp.base);
    const int NP = g.NP, npad = p.mlp.npad, n = g.n;
    for (int i = tid; i < g.small_elems; i += kTcThreads) {
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

  if (warp == 4) {
    // ================================ engine ===========================================================
    if (tid == kTcM) {
      TcEngine e;
      e.full = bar_full; e.empty = bar_empty; e.ring = ring;
      e.img = reinterpret_cast<const unsigned char*>(tp.img);
      e.issued = 0; e.consumed = 0;
      e.steps_per_cycle = (unsigned)(g.L * g.KST);
      const uint32_t idesc = tc::idesc_bf16_f32(kTcM, g.NP);
      for (int s = 0; s < g.stages; ++s) tc_engine_issue_copy(g, e);
      unsigned phase_a = 0;
      while (true) {
        mbar_wait(bar_a, phase_a);
        phase_a ^= 1u;
        if (*stop_flag) break;
        tc::fence_after_sync();
        tc_engine_layer(g, e, tbase, idesc);
        tc::commit(smem_u32(bar_d));
      }
      // every MMA has completed (its D was consumed); wait for the copies still in flight
      for (unsigned q = e.consumed; q < e.issued; ++q)
        mbar_wait(&bar_full[q % (unsigned)g.stages], (q / (unsigned)g.stages) & 1u);
    }
  } else {
    // ================================ owners ===========================================================
    TcLane tl;
    tl.taddr = tbase + ((uint32_t)(warp * 32) << 16);
    tl.bar_a = smem_u32(bar_a);
    tl.bar_d = smem_u32(bar_d);
    tl.bar_d_ptr = bar_d;
    tl.phase_d = 0;
    tl.sp = sp;
    tl.slope = (float)p.mlp.slope;
    SolverCfg cfg = p.cfg;

    while (true) {
      if (tid == 0) {
        long long tile = (long long)atomicAdd(p.queue, 1ULL);
        *tile_slot = tile;
        if (tile < p.n_tiles) {
          int j = 0;
          while (j + 1 < p.n_jobs && p.jobs[j + 1].tile_begin <= tile) ++j;
          *jobp = p.jobs[j];
        }
      }
      owners_sync();
      const long long tile = *tile_slot;
      if (tile >= p.n_tiles) break;
      const FwdJob& job = *jobp;
      cfg.tab = job.tab;
      const S* y0 = reinterpret_cast<const S*>(job.y0);
      const S* gptr = reinterpret_cast<const S*>(job.g);
      const S* eptr = reinterpret_cast<const S*>(job.e_rev);
      const S* dptr = reinterpret_cast<const S*>(job.data);
      S* y_out = reinterpret_cast<S*>(job.y_out);
      S* i_out = reinterpret_cast<S*>(job.i_out);
      S* ckpt_y = reinterpret_cast<S*>(job.ckpt_y);
      const bool observe = (job.v_out != nullptr) && (job.i_out != nullptr || job.loss_out != nullptr);
      const long long jB = job.B;
      const int T = job.T;
      const long long b = (tile - job.tile_begin) * kTcM + tid;
      const bool valid = b < jB;
      S g_b = (S)1, e_b = (S)job.e_scalar;

      auto emit = [&](int idx, S a, S r) {
        if (y_out) {
          typename Vec2<S>::type v;
          v.x = a; v.y = r;
          *reinterpret_cast<typename Vec2<S>::type*>(y_out + ((size_t)idx * jB + b) * 2) = v;
        }
        if (observe) {
          double cur = (double)(g_b * a * r) * (job.v_out[idx] - (double)e_b);
          if (i_out) i_out[(size_t)idx * jB + b] = (S)cur;
          if (dptr) {
            double d = (double)dptr[(size_t)idx * job.data_B + (job.data_B == 1 ? 0 : b)];
            double diff = cur - d;
            obs[2 * tid] += diff * diff;
            obs[2 * tid + 1] += fabs(diff);
          }
        }
      };
      auto ckpt = [&](int step, const Lane<S>& lane) -> bool {
        if (!job.ckpt_t) return true;
        if (step >= job.ckpt_cap) return false;
        size_t o = (size_t)step * jB + b;
        double2 tt;
        tt.x = lane.t0; tt.y = lane.dt;
        *reinterpret_cast<double2*>(job.ckpt_t + 2 * o) = tt;
        S buf[kCkptVals];
        ckpt_pack<S>(lane, buf);
        typedef typename Vec2<S>::type V2;
        V2* dst = reinterpret_cast<V2*>(ckpt_y + (size_t)kCkptVals * o);
#pragma unroll
        for (int i = 0; i < kCkptVals / 2; ++i) {
          V2 v;
          v.x = buf[2 * i]; v.y = buf[2 * i + 1];
          dst[i] = v;
        }
        return true;
      };

      Lane<S>& L = lanes[tid];
      {
        S ya = (S)0, yr = (S)1;
        if (valid) {
          ya = y0[2 * b]; yr = y0[2 * b + 1];
          if (gptr) g_b = gptr[b];
          if (eptr) e_b = eptr[b];
        }
        lane_reset<S>(L, ya, yr, job.t_out[0], valid);
        obs[2 * tid] = 0.0; obs[2 * tid + 1] = 0.0;
        if (valid) emit(0, ya, yr);
      }

      double nv, ain;
      if (p.method == 0) {
        init_prepare_f0<S>(L, cfg, &nv, &ain);
        float out = tc_mlp_eval(g, tl, (float)nv, (float)ain);
        init_store_f0<S>(L, cfg, (double)out);
        if (cfg.first_step > 0) {
          L.dt = cfg.first_step;
        } else {
          init_prepare_f1<S>(L, cfg, &nv, &ain);
          out = tc_mlp_eval(g, tl, (float)nv, (float)ain);
          init_store_f1<S>(L, cfg, (double)out);
        }
        if (T <= 1 && lane_active(L)) L.status = LANE_DONE;

        while (true) {
          dp_check_before_step<S>(L, cfg);
          if (!owners_or(lane_active(L) ? 1 : 0)) break;
#pragma unroll 1
          for (int s = 0; s < 6; ++s) {
            dp_prepare_stage<S>(L, cfg, s, &nv, &ain);
            out = tc_mlp_eval(g, tl, (float)nv, (float)ain);
            dp_store_stage<S>(L, cfg, s, (double)out);
          }
          dp_finish_step<S>(L, cfg, job.t_out, T, emit, ckpt);
        }
      } else {
        if (T <= 1 && lane_active(L)) L.status = LANE_DONE;
        for (int gi = 0; gi + 1 < job.G; ++gi) {
          const double g0 = job.grid[gi], g1 = job.grid[gi + 1];
          if (!owners_or(lane_active(L) ? 1 : 0)) break;
#pragma unroll 1
          for (int s = 0; s < 4; ++s) {
            rk4_prepare_stage<S>(L, cfg, s, g0, g1, p.time_f32 != 0, p.rk4_perturb != 0, &nv, &ain);
            const float out = tc_mlp_eval(g, tl, (float)nv, (float)ain);
            rk4_store_stage<S>(L, cfg, s, (double)out);
          }
          rk4_finish_step<S>(L, g0, g1, p.time_f32 != 0, job.t_out, T, emit);
        }
      }

      if (valid) {
        job.stats_out[4 * b + 0] = L.n_acc;
        job.stats_out[4 * b + 1] = L.n_rej;
        job.stats_out[4 * b + 2] = L.nfe;
        job.stats_out[4 * b + 3] = L.status == LANE_DONE ? 0 : L.status;
        if (job.loss_out) {
          job.loss_out[2 * b] = obs[2 * tid];
          job.loss_out[2 * b + 1] = obs[2 * tid + 1];
        }
      }
      owners_sync();   // the job slot is rewritten by the next tile
    }
    // release the engine: stop flag, then one more a_ready phase
    if (tid == 0) *stop_flag = 1;
    owners_sync();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl.bar_a) : "memory");
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

}  // namespace ikr
#endif  // IKR_FORWARD_TC_CUH_
