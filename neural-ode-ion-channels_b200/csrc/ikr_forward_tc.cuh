// ikr_forward_tc.cuh -- forward integration kernel with the hidden Linear(n, n) layers on the
// 5th-generation tensor cores (tcgen05, accumulators and activations in TMEM).
//
// Same per-trajectory solver as ikr_forward_kernel (ikr_math.h lane state machine, one owner
// thread per trajectory); the tile is fixed at M = 128 trajectories = the 128 TMEM lanes, and
// thread i owns TMEM lane i, so the whole MLP needs NO cross-thread exchange:
//
//   layer 0      lane threads: h = LeakyReLU(w0a nv + w0b a + b0)           -> A operand (TMEM)
//   layer 1..L   MMA thread: D[128 x n] = A[128 x n] . W_l^T  (tcgen05.mma, B = W_l streamed
//                L2 -> shared memory by cp.async.bulk through a full/empty mbarrier ring that a
//                separate producer thread keeps full)
//                lane threads: tcgen05.ld the D row, + bias, LeakyReLU      -> A operand (TMEM)
//   output       lane threads: dot(h, w_last) while reading the last D row, owner adds b_last
// G threads share one TMEM lane (column groups, warps w and w + 4 c see the same lane quarter):
// group c produces the 32-feature UNITS u = c (mod G) of every layer; the owner broadcasts (nv, a)
// and collects the G partial output sums through shared memory.
//
// FP32 accuracy on BF16 tensor cores: every fp32 value x is split into three bf16 terms
// x = x1 + x2 + x3 (exact), and each layer issues the six products whose weight is >= 2^-16:
// a1 b1, a2 b1, a3 b1, a1 b2, a2 b2, a1 b3 (dropped terms <= 2^-24 relative), all accumulated in
// fp32 in TMEM.  Measured on B200 (tests/tc_probe.cu): one 128 x 208 x 16 MMA = 104 cycles, i.e.
// 75 MMAs = 7.8 k cycles per layer against ~200 k cycles of FFMA2 work.
//
// TMEM budget (512 columns), n = 200: TWO accumulators D (2 x 208 columns: the MMAs of layer l + 1
// write the one the epilogue of layer l is not reading) + a ring of TWO A-operand unit slots (2 x 48
// columns: [a1 16 | a2 16 | a3 16] = two K = 16 steps) = 512.  The activations never sit in TMEM as
// a whole: the epilogue of layer l hands them to the MMA warp unit by unit, and the MMAs of layer
// l + 1 start after the first unit.  The last n % 16 <= 8 input features form a TAIL unit: two A
// blocks [a1t | a2t], [a1t | a3t] against B blocks [b1t | b1t], [b2t | b2t], [b3t | b1t] give the
// same six products in 3 MMAs.
#ifndef IKR_FORWARD_TC_CUH_
#define IKR_FORWARD_TC_CUH_

#include <stdio.h>

#include "ikr_forward.cuh"
#include "ikr_tc.cuh"

namespace ikr {

constexpr int kTcM = 128;        // trajectories per tile = TMEM lanes
constexpr int kTcMaxStages = 12;
constexpr int kTcMinStages = 4;
// Threads: G column groups x 128 epilogue threads (group 0 = the owner threads of the trajectories;
// thread 128 c + i of group c works on TMEM lane i and on the k-steps j = c (mod G)), then one
// warp whose lane 0 issues the MMAs and one warp whose lane 0 streams the weight ring.
__host__ __device__ constexpr int tc_threads(int G) { return 128 * G + 64; }

struct TcGeom {
  int n, L, NP;      // NP = roundup(n, 16): D columns = rows of every B block
  int KSf;           // regular K = 16 steps per layer (incl. a zero-padded one when n % 16 > 8)
  int tail;          // 1: a final tail step covers the last n % 16 <= 8 features
  int KST;           // steps per layer = KSf + tail = ring stages consumed per layer
  int units;         // epilogue work units: pairs of k-steps (the last one may hold a single k-step)
  int terms;         // split terms per fp32 operand: 3 = bf16x3 (six products per fp32 product:
                     // a1b1 a2b1 a3b1 a1b2 a2b2 a1b3), 2 = fp16x2 (three: a1b1 a2b1 a1b2; operands
                     // scaled by powers of two into the fp16 range, see tc::split2h)
  int unit_cols;     // TMEM columns of one unit slot = 16 * terms
  int ring_slots;    // unit slots of the A-operand ring (2, or 3 when the columns allow it)
  int col_ring;      // TMEM column of the A-operand unit ring; a unit of two k-steps is laid out as
                     // [a1 | a2 (| a3)] of 16 columns each -- 8 each for a single k-step; the tail as
                     // [a1t | a2t | a1t | a3t] (bf16x3) or [a1t | a2t] (fp16x2) at the start of its slot.
                     // Columns [0, NP) and [NP, 2 NP) are the two D accumulators.
  int cols;          // TMEM columns used
  int block_bytes;   // one B block: NP x 16 16-bit values = NP * 32 bytes
  int stage_bytes;   // one k-step: `terms` blocks
  int ring_stride;   // bytes per slot of the shared-memory weight ring (= stage_bytes unless only the
                     // leading blocks of a stage are streamed: adjoint kernel, three-product passes)
  int stages;        // ring depth
  int small_elems;   // zero-padded small-parameter block: (4 + L) NP + 8 floats, then L per-layer
                     // accumulator scales (1 for bf16x3), padded to 8
};

__host__ __device__ inline TcGeom tc_geometry(int n, int L, int terms = 3) {
  TcGeom g;
  g.n = n; g.L = L;
  g.NP = (n + 15) / 16 * 16;
  const int rem = n % 16;
  g.KSf = n / 16 + (rem > 8 ? 1 : 0);
  g.tail = (rem > 0 && rem <= 8) ? 1 : 0;
  g.KST = g.KSf + g.tail;
  g.units = (g.KSf + 1) / 2;
  g.terms = terms;
  g.unit_cols = 16 * terms;
  g.col_ring = 2 * g.NP;
  g.ring_slots = (terms == 2 && g.col_ring + 3 * g.unit_cols <= 512) ? 3 : 2;
  g.cols = g.col_ring + g.ring_slots * g.unit_cols;
  g.block_bytes = g.NP * 32;
  g.stage_bytes = terms * g.block_bytes;
  g.ring_stride = g.stage_bytes;
  g.stages = 0;
  g.small_elems = (4 + L) * g.NP + 8 + (L + 7) / 8 * 8;
  return g;
}
// index of the accumulator scale of hidden layer l (0-based) in the small-parameter block
__host__ __device__ inline int tc_scale_index(const TcGeom& g, int l) { return (4 + g.L) * g.NP + 8 + l; }
// the A operand of the fp16x2 path carries 2^kTcActShift x the activation (exact), so that small
// LeakyReLU outputs keep their second term in the normal fp16 range; |h| must stay below 65504 / 16
constexpr int kTcActShift = 4;
__host__ __device__ inline bool tc_geometry_ok(const TcGeom& g) {
  return g.n >= 16 && g.NP <= 256 && g.cols <= (int)tc::kTmemCols && g.KST >= 1;
}
// ---- weight image: [layer][k-step][block 0..terms-1][NP x 16 16-bit values in core-matrix order] -----
struct TcPackParams {
  const float* wn;   // [L][n][npad] rows = output features (packed-parameter section off_wn)
  const float* wt;   // [L][n][npad] rows = input features  (section off_wt = W^T), backward images
  int npad;
  int n_seq;         // L: forward images W_1..W_L; 2L: followed by the transposed W_L..W_1;
                     // 3L (regression): forward images twice, then the transposed ones
  TcGeom g;
  uint16_t* img;
  // fp16x2 images (g.terms == 2, forward only): per-layer |W| maximum as float bits (filled by
  // ikr_tc_absmax_kernel, zeroed by the caller before) and the accumulator scales 2^-k_l the
  // forward kernel multiplies D with (written by this kernel, [n_seq] floats)
  unsigned* absmax;
  float* scales;
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
__device__ __forceinline__ uint16_t f16_term(float w, int term) {
  uint32_t h = tc::pack_f16x2(w, 0.0f);
  if (term == 0) return (uint16_t)(h & 0xFFFFu);
  float lo, hi;
  tc::u2(tc::unpack_f16x2(h), lo, hi);
  return (uint16_t)(tc::pack_f16x2(w - lo, 0.0f) & 0xFFFFu);
}
// exponent k of the power-of-two weight scale of a layer: max |W| 2^k in [2^13, 2^14)
__device__ __forceinline__ int tc_weight_shift(unsigned absmax_bits) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.0f) || !isfinite(m)) return 0;
  int e;
  frexpf(m, &e);            // m = f 2^e, f in [0.5, 1)
  int k = 14 - e;
  return k < -100 ? -100 : (k > 100 ? 100 : k);
}

// per-layer maximum of |W| (positive floats order like their bit patterns)
__global__ void ikr_tc_absmax_kernel(const TcPackParams p) {
  const TcGeom& g = p.g;
  const long long per = (long long)g.n * g.n;
  const long long total = (long long)g.L * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / per);
    const long long r = i - l * per;
    const int o = (int)(r / g.n), k = (int)(r - (long long)o * g.n);
    const float w = fabsf(p.wn[(long long)l * g.n * p.npad + (long long)o * p.npad + k]);
    unsigned m = __float_as_uint(w);
    const unsigned mask = __activemask();
    const int l0 = __shfl_sync(mask, l, __ffs(mask) - 1);
    if (__all_sync(mask, l == l0)) {
      m = __reduce_max_sync(mask, m);
      if ((int)(threadIdx.x & 31) == __ffs(mask) - 1) atomicMax(&p.absmax[l], m);
    } else {
      atomicMax(&p.absmax[l], m);          // a warp that straddles two layers
    }
  }
}

// B operand image of layer-MMA number `seq`: B[row][k] = src[row][k], src = W_l (forward: D = A W_l^T)
// or W_l^T (backward: D = A W_l)
__global__ void ikr_tc_pack_kernel(const TcPackParams p) {
  const TcGeom& g = p.g;
  const int T = g.terms;
  const long long per_block = (long long)g.NP * 16;
  const long long total = (long long)p.n_seq * g.KST * T * per_block;
  if (T == 2 && blockIdx.x == 0 && (int)threadIdx.x < p.n_seq)
    p.scales[threadIdx.x] = ldexpf(1.0f, -tc_weight_shift(p.absmax[threadIdx.x % g.L]));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % 16);
    const int o = (int)((i / 16) % g.NP);
    const long long blk = i / per_block;
    const int c = (int)(blk % T);
    const int step = (int)((blk / T) % g.KST);
    const int seq = (int)(blk / ((long long)T * g.KST));
    const int n_fwd = p.n_seq == 3 * g.L ? 2 * g.L : g.L;
    const float* src = seq < n_fwd ? p.wn + (long long)(seq % g.L) * g.n * p.npad
                                   : p.wt + (long long)(p.n_seq - 1 - seq) * g.n * p.npad;
    int k, term;
    bool zero = false;
    if (step < g.KSf) {
      k = 16 * step + kk;
      term = c;
    } else if (T == 3) {
      k = 16 * g.KSf + (kk & 7);
      term = c == 0 ? 0 : (c == 1 ? 1 : (kk < 8 ? 2 : 0));
    } else {
      // fp16x2 tail, A block [a1t | a2t]: block 0 = [b1t | b1t], block 1 = [b2t | 0]
      k = 16 * g.KSf + (kk & 7);
      term = c;
      zero = c == 1 && kk >= 8;
    }
    float w = 0.0f;
    if (!zero && o < g.n && k < g.n) w = src[(long long)o * p.npad + k];
    const long long byte = (long long)(o >> 3) * 256 + (kk >> 3) * 128 + (o & 7) * 16 + (kk & 7) * 2;
    uint16_t v;
    if (T == 3) v = bf16_term(w, term);
    else v = f16_term(ldexpf(w, tc_weight_shift(p.absmax[seq % g.L])), term);
    p.img[(blk * g.block_bytes + byte) >> 1] = v;
  }
}

// ---- shared memory carve-up --------------------------------------------------------------------------
template <typename S>
struct TcSmemLayout {
  size_t off_bar, off_misc, off_job, off_lanes, off_obs, off_aux, off_xin, off_part, off_sp, off_ring, total;
  __host__ __device__ TcSmemLayout(const TcGeom& g, int stages, int G) {
    size_t o = 0;
    off_bar = o; o += (size_t)(2 * kTcMaxStages + 2 * 8 + 2) * 8;  // full[], empty[], unit_ready[8], unit_done[8], d_ready
    off_misc = o; o += 32;                                          // tmem base, stop flag, tile slot
    off_job = o; o += (size_t)kInlineJobs * ((sizeof(FwdJob) + 15) & ~(size_t)15);   // pool: job table
    off_lanes = o; o += (size_t)kTcM * sizeof(Lane<S>); o = (o + 15) & ~(size_t)15;
    off_obs = o; o += (size_t)kTcM * 2 * sizeof(double);
    off_aux = o; o += (size_t)kTcM * 32;                               // LaneAux (lane-pool kernel)
    off_xin = o; o += (size_t)kTcM * 2 * sizeof(float);
    off_part = o; o += (size_t)G * kTcM * sizeof(float);
    off_sp = o; o += (size_t)g.small_elems * sizeof(float); o = (o + 127) & ~(size_t)127;
    off_ring = o; o += (size_t)stages * g.stage_bytes;
    total = o;
  }
};

struct TcFwdParams {
  FwdParams f;           // f.M = trajectories per tile (<= 128 TMEM lanes; fewer when the batch
                         // would otherwise leave SMs idle); MG / NG / n_worker_warps unused
  TcGeom g;
  const void* img;       // weight image written by ikr_tc_pack_kernel
  const float* scales;   // fp16x2: per-layer accumulator scales 2^-k_l written by ikr_tc_pack_kernel
  int timing;            // debug: block 0 prints its phase clocks (desc.reserved bit 3)
};

// named barriers: 1 = every lane thread (128 G), 2 = the 128 owner threads
template <int G>
__device__ __forceinline__ void lanes_sync() { asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory"); }
__device__ __forceinline__ void owners_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ int owners_or(int pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, 2, 128, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(pred)
      : "memory");
  return (int)r;
}

// Per-thread view of the tensor-core MLP
struct TcLane {
  uint32_t taddr;        // TMEM address of this thread's lane, column 0 of the allocation
  uint64_t* unit_ready;  // [kTcMaxUnits] unit u of the pass is in its slot (4 warp arrivals: its group)
  uint64_t* unit_done;   // [kTcMaxUnits] tcgen05.commit: the MMAs of unit u of the pass are done
  uint64_t* bar_d;       // tcgen05.commit: the layer's D is complete
  unsigned phase_d;      // parity of the next d_ready completion
  unsigned n_pass;       // layer passes whose D this thread has consumed: D of the next one is in buffer n_pass & 1
  unsigned unit_idx;     // global sequence number of the first unit of the pass being produced
  unsigned upass;        // passes produced so far (unit_idx = upass * units per pass)
  unsigned ring_base;    // unit_idx mod ring slots
  // two-tile kernel (ikr_forward_tc_pp.cuh) only, null / 0 elsewhere:
  uint64_t* last_d_bar;          // arrive (one per warp) once the D of an evaluation's LAST pass is there
  uint64_t* gate_bar;            // the first unit store of an evaluation waits for this barrier phase
  unsigned gate_parity;
  mutable int gate_pending;
  mutable long long c_gate;      // cycles spent waiting at the gate (timing runs)
  int group;             // column group 0..G-1
  int lane;              // TMEM lane 0..127
  const float* sp;       // small parameters (stride NP): w0a | w0b | b0 | L x bias | w_last, b_last
  float* xin;            // [128][2] (nv, a) broadcast by the owner ([128][4] in the adjoint kernel)
  float* part;           // [G][128] partial output sums
  float slope;
  long long c_l0, c_wait, c_epi;   // phase clocks (timing runs)
  long long c_sync_a, c_sync_b, c_last;   // owner: barrier A / barrier B waits, output-layer epilogue
#ifdef IKR_TC_TRACE
  int trace_eval;
#endif
};

__device__ __forceinline__ float tc_leaky(float x, float slope) { return x > 0.0f ? x : x * slope; }

// ---- A-operand unit ring ---------------------------------------------------------------------------
// The activations of the next layer do not wait in TMEM for the whole layer: they flow through a ring
// of TWO 48-column unit slots (two K-steps: [a1 16 | a2 16 | a3 16]; the tail uses 16 columns of a
// slot).  Units are numbered globally (all passes of all evaluations): unit i = pass i / UT, position
// i % UT, slot i & 1.  Barriers are per POSITION (every phase of a barrier is consumed by the same
// waiter, in order, so the parity never aliases): the producing group of unit i first waits for
// unit_done of unit i - 2 (the previous user of its slot), stores, and arrives on unit_ready (one
// arrival per warp); the MMA warp issues the unit's K-steps as soon as it lands and commits
// unit_done.  So the MMAs of layer l + 1 start after the FIRST unit of the epilogue of layer l and
// run under the rest of it, accumulating into the other D buffer.
#ifndef IKR_TC_STAGGER_NS
#define IKR_TC_STAGGER_NS 200
#endif
constexpr unsigned kTcStaggerNs = IKR_TC_STAGGER_NS;   // start delay per column group after d_ready
constexpr int kTcMaxUnits = 8;     // NP <= 208: at most 7 units of two K-steps + the tail
// ring geometry is a compile-time function of the operand split (a runtime modulo in the MMA warp's
// unit loop costs ~70 cycles per k-step: measured)
template <int TERMS>
__host__ __device__ constexpr unsigned tc_ring_slots() { return TERMS == 2 ? 3u : 2u; }
// ring geometry of a kernel: bf16x3 units without their third term (adjoint kernel, LITE = 2) are
// 32 columns wide like fp16x2 units, so three of them fit
template <int TERMS, int LITE>
__host__ __device__ constexpr int tc_ring_terms() { return (TERMS == 3 && LITE == 2) ? 2 : TERMS; }
// slot of unit u of the pass this thread is producing (tl.ring_base = first unit of the pass mod R)
template <int TERMS = 3>
__device__ __forceinline__ uint32_t tc_unit_slot_col(const TcGeom& g, const TcLane& tl, int u) {
  constexpr unsigned R = tc_ring_slots<TERMS>();
  return (uint32_t)g.col_ring + (uint32_t)(16 * TERMS) * ((tl.ring_base + (unsigned)u) % R);
}
// wait until the MMAs of the unit that used this slot before (R units earlier) are done
template <int TERMS = 3>
__device__ __forceinline__ void tc_unit_acquire(const TcGeom& g, const TcLane& tl, int u) {
  constexpr unsigned R = tc_ring_slots<TERMS>();
  const unsigned UT = (unsigned)(g.units + g.tail);
  if (tl.gate_pending) {
    // two-tile kernel: every pass of the other tile's evaluation must be complete before this
    // thread looks at the ring barriers again (their parities only tell neighbouring phases apart)
    const long long tg0 = clock64();
    mbar_wait(tl.gate_bar, tl.gate_parity);
    tl.gate_pending = 0;
    tl.c_gate += clock64() - tg0;
  }
  // UT < R: that unit belongs to a pass whose D this thread has already consumed (all done), and a
  // parity wait two phases behind the barrier would alias
  if (UT < R) return;
  if ((unsigned)u >= R) mbar_wait(&tl.unit_done[(unsigned)u - R], tl.upass & 1u);
  else if (tl.upass > 0u) mbar_wait(&tl.unit_done[(unsigned)u + UT - R], (tl.upass - 1u) & 1u);
  else return;
  tc::fence_after_sync();
}
__device__ __forceinline__ void tc_unit_publish(const TcGeom& g, const TcLane& tl, int u) {
  tc::wait_st();
  tc::fence_before_sync();
  __syncwarp();
  if ((tl.lane & 31) == 0) mbar_arrive(&tl.unit_ready[u]);
}
// this thread has produced its units of a pass
template <int TERMS = 3>
__device__ __forceinline__ void tc_pass_advance(TcLane& tl, int UT) {
  constexpr unsigned R = tc_ring_slots<TERMS>();
  tl.unit_idx += (unsigned)UT;
  tl.upass += 1u;
  tl.ring_base = (tl.ring_base + (unsigned)UT) % R;
}
// unit u of a pass belongs to column group u % G; the tail is unit number g.units
template <int G>
__device__ __forceinline__ int tc_units_total(const TcGeom& g) { return g.units + g.tail; }

// One epilogue work unit of NK = 1 or 2 k-steps (features 32 u ..): the pre-activations arrive in v
// (raw fp32 bits: D columns, or layer-0 sums), get bias + LeakyReLU, and either become a unit of the
// next A operand (three bf16 terms: one 16 NK-column store for [a1 | a2], one 8 NK-column store for
// a3, into ring slot `gi`) or are reduced against w_last.
template <int NK, int TERMS = 3>
__device__ __forceinline__ void tc_unit_finish(const TcGeom& g, const TcLane& tl, int u, unsigned gi,
                                               const float* bias, uint32_t (&v)[16 * NK], bool last,
                                               const float* wl, float (&s)[4], float dscale = 1.0f) {
  const int c0 = 32 * u;
  const tc::f32x2_t slope2 = tc::p2(tl.slope, tl.slope);
  const tc::f32x2_t ds2 = tc::p2(dscale, dscale);
  tc::f32x2_t h[8 * NK];
#pragma unroll
  for (int q = 0; q < 4 * NK; ++q) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
    const tc::f32x2_t d0 = tc::p2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
    const tc::f32x2_t d1 = tc::p2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    if (TERMS == 3) {
      h[2 * q] = tc::leaky2(tc::add2(d0, tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::add2(d1, tc::p2(bb.z, bb.w)), slope2);
    } else {
      // fp16x2: D carries 2^(k_l + 4) x the pre-activation sum, the bias is stored x 16:
      // z' = D 2^-k_l + 16 b = 16 z exactly (powers of two), h' = LeakyReLU(z') = 16 h
      h[2 * q] = tc::leaky2(tc::fma2(d0, ds2, tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::fma2(d1, ds2, tc::p2(bb.z, bb.w)), slope2);
    }
  }
  if (!last) {
    // the split arithmetic comes BEFORE the slot wait: it runs while the MMAs still read the slot
    uint32_t w12[16 * NK], w3[8 * NK];
#pragma unroll
    for (int q = 0; q < 8 * NK; ++q) {
      if (TERMS == 3) tc::split3t(h[q], w12[q], w12[8 * NK + q], w3[q]);
      else tc::split2h(h[q], w12[q], w12[8 * NK + q]);
    }
    tc_unit_acquire<TERMS>(g, tl, u);
    const uint32_t dst = tl.taddr + tc_unit_slot_col<TERMS>(g, tl, u);
    if (NK == 2) {
      tc::st32(dst, reinterpret_cast<uint32_t(&)[32]>(w12));
      if (TERMS == 3) tc::st16(dst + 32, reinterpret_cast<uint32_t(&)[16]>(w3));
    } else {
      tc::st16(dst, reinterpret_cast<uint32_t(&)[16]>(w12));
      if (TERMS == 3) tc::st8(dst + 16, reinterpret_cast<uint32_t(&)[8]>(w3));
    }
    tc_unit_publish(g, tl, u);
  } else {
#pragma unroll
    for (int q = 0; q < 4 * NK; ++q) {
      const float4 ww = *reinterpret_cast<const float4*>(wl + c0 + 4 * q);
      float h0, h1, h2, h3;
      tc::u2(h[2 * q], h0, h1);
      tc::u2(h[2 * q + 1], h2, h3);
      s[0] = __fmaf_rn(h0, ww.x, s[0]);
      s[1] = __fmaf_rn(h1, ww.y, s[1]);
      s[2] = __fmaf_rn(h2, ww.z, s[2]);
      s[3] = __fmaf_rn(h3, ww.w, s[3]);
    }
  }
}
// the 8 tail features (columns 16 KSf ..): blocks [a1t | a2t | a1t | a3t] in one 16-column store
// (bf16x3) or [a1t | a2t] in one 8-column store (fp16x2)
template <int TERMS = 3>
__device__ __forceinline__ void tc_tail_finish(const TcGeom& g, const TcLane& tl, unsigned gi,
                                               const float* bias, uint32_t (&v)[8], bool last,
                                               const float* wl, float (&s)[4], float dscale = 1.0f) {
  const int c0 = 16 * g.KSf;
  const tc::f32x2_t slope2 = tc::p2(tl.slope, tl.slope);
  const tc::f32x2_t ds2 = tc::p2(dscale, dscale);
  tc::f32x2_t h[4];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
    const tc::f32x2_t d0 = tc::p2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
    const tc::f32x2_t d1 = tc::p2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    if (TERMS == 3) {
      h[2 * q] = tc::leaky2(tc::add2(d0, tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::add2(d1, tc::p2(bb.z, bb.w)), slope2);
    } else {
      h[2 * q] = tc::leaky2(tc::fma2(d0, ds2, tc::p2(bb.x, bb.y)), slope2);
      h[2 * q + 1] = tc::leaky2(tc::fma2(d1, ds2, tc::p2(bb.z, bb.w)), slope2);
    }
  }
  if (!last) {
    uint32_t t[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (TERMS == 3) {
        tc::split3t(h[q], t[q], t[4 + q], t[12 + q]);
        t[8 + q] = t[q];
      } else {
        tc::split2h(h[q], t[q], t[4 + q]);
      }
    }
    tc_unit_acquire<TERMS>(g, tl, g.units);
    if (TERMS == 3) tc::st16(tl.taddr + tc_unit_slot_col<TERMS>(g, tl, g.units), t);
    else tc::st8(tl.taddr + tc_unit_slot_col<TERMS>(g, tl, g.units), reinterpret_cast<uint32_t(&)[8]>(t));
    tc_unit_publish(g, tl, g.units);
  } else {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 ww = *reinterpret_cast<const float4*>(wl + c0 + 4 * q);
      float h0, h1, h2, h3;
      tc::u2(h[2 * q], h0, h1);
      tc::u2(h[2 * q + 1], h2, h3);
      s[0] = __fmaf_rn(h0, ww.x, s[0]);
      s[1] = __fmaf_rn(h1, ww.y, s[1]);
      s[2] = __fmaf_rn(h2, ww.z, s[2]);
      s[3] = __fmaf_rn(h3, ww.w, s[3]);
    }
  }
}

// layer-0 pre-activations (without the bias, which tc_*_finish adds) of NV features from c0
template <int NV>
__device__ __forceinline__ void tc_layer0_sums(const TcLane& tl, int NP, int c0, float nv, float a,
                                               uint32_t (&v)[NV]) {
  const float* w0a = tl.sp + c0;
  const float* w0b = tl.sp + NP + c0;
#pragma unroll
  for (int q = 0; q < NV / 4; ++q) {
    const float4 wa = *reinterpret_cast<const float4*>(w0a + 4 * q);
    const float4 wb = *reinterpret_cast<const float4*>(w0b + 4 * q);
    v[4 * q + 0] = __float_as_uint(__fmaf_rn(wb.x, a, wa.x * nv));
    v[4 * q + 1] = __float_as_uint(__fmaf_rn(wb.y, a, wa.y * nv));
    v[4 * q + 2] = __float_as_uint(__fmaf_rn(wb.z, a, wa.z * nv));
    v[4 * q + 3] = __float_as_uint(__fmaf_rn(wb.w, a, wa.w * nv));
  }
}

struct TcNoHook {
  __device__ __forceinline__ void operator()() const {}
};
// wait for the D of the next layer pass; returns its TMEM column (the two D buffers alternate)
__device__ __forceinline__ uint32_t tc_wait_d(const TcGeom& g, TcLane& tl, bool stagger = true) {
#ifdef IKR_TC_BACKOFF
  mbar_wait_backoff(tl.bar_d, tl.phase_d, IKR_TC_BACKOFF);
#else
  mbar_wait(tl.bar_d, tl.phase_d);
#endif
  tl.phase_d ^= 1u;
  tc::fence_after_sync();
  const uint32_t col = (tl.n_pass & 1u) ? (uint32_t)g.NP : 0u;
  tl.n_pass += 1u;
  // Stagger the column groups: the MMAs of the next layer start with unit 0 (group 0), and a pass of
  // MMAs (~7.6 k cycles) is several times longer than a whole epilogue, so the later groups give the
  // first ones the issue slots (unit 0 lands in ~0.4 k instead of ~1.1 k cycles after d_ready).
  if (stagger && tl.group > 0) __nanosleep(kTcStaggerNs * (unsigned)tl.group);
  return col;
}

// ---- the three parts of one MLP evaluation (tc_mlp_eval = layer0 + hook + hidden + output) ----------
// layer 0: this thread's units (u = tl.group mod G) of the first pass, from (nv, a) in tl.xin
template <int G, int TERMS>
__device__ __forceinline__ void tc_eval_layer0(const TcGeom& g, TcLane& tl) {
  const int NP = g.NP;
  const float2 in = *reinterpret_cast<const float2*>(tl.xin + 2 * tl.lane);
  const float nv = in.x, a = in.y;
  const int UT = g.units + g.tail;
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const float* wl = tl.sp + (size_t)(3 + g.L) * NP;
  const float* b0 = tl.sp + 2 * NP;
  if (tl.group > 0) __nanosleep(kTcStaggerNs * (unsigned)tl.group);
  for (int u = tl.group; u < UT; u += G) {
    const unsigned gi = tl.unit_idx + (unsigned)u;
    if (u < g.units) {
      if (2 * u + 1 < g.KSf) {
        uint32_t v[32];
        tc_layer0_sums<32>(tl, NP, 32 * u, nv, a, v);
        tc_unit_finish<2, TERMS>(g, tl, u, gi, b0, v, false, wl, s);
      } else {
        uint32_t v[16];
        tc_layer0_sums<16>(tl, NP, 32 * u, nv, a, v);
        tc_unit_finish<1, TERMS>(g, tl, u, gi, b0, v, false, wl, s);
      }
    } else {
      uint32_t v[8];
      tc_layer0_sums<8>(tl, NP, 16 * g.KSf, nv, a, v);
      tc_tail_finish<TERMS>(g, tl, gi, b0, v, false, wl, s);
    }
  }
  tc_pass_advance<TERMS>(tl, UT);
}
// one epilogue over the D of a pass: `last` reduces against w_last into s, else produces the next units
template <int G, int TERMS>
__device__ __forceinline__ void tc_eval_epilogue(const TcGeom& g, TcLane& tl, int layer, uint32_t dcol,
                                                 bool last, float (&s)[4]) {
  const int NP = g.NP, UT = g.units + g.tail;
  const float* wl = tl.sp + (size_t)(3 + g.L) * NP;
  const float* bias = tl.sp + (size_t)(3 + layer) * NP;
  const float dscale = TERMS == 3 ? 1.0f : tl.sp[tc_scale_index(g, layer)];
  for (int u = tl.group; u < UT; u += G) {
    const unsigned gi = tl.unit_idx + (unsigned)u;
    if (u < g.units) {
      if (2 * u + 1 < g.KSf) {
        uint32_t v[32];
        tc::ld32(tl.taddr + dcol + 32 * u, v);
        tc::wait_ld();
        tc_unit_finish<2, TERMS>(g, tl, u, gi, bias, v, last, wl, s, dscale);
      } else {
        uint32_t v[16];
        tc::ld16(tl.taddr + dcol + 32 * u, v);
        tc::wait_ld();
        tc_unit_finish<1, TERMS>(g, tl, u, gi, bias, v, last, wl, s, dscale);
      }
    } else {
      uint32_t v[8];
      tc::ld8(tl.taddr + dcol + 16 * g.KSf, v);
      tc::wait_ld();
      tc_tail_finish<TERMS>(g, tl, gi, bias, v, last, wl, s, dscale);
    }
  }
  if (!last) tc_pass_advance<TERMS>(tl, UT);
}
// hidden layers 1 .. L-1 (their epilogues produce the units of the next pass), then the wait for the
// D of the last pass; returns its TMEM column
template <int G, int TERMS>
__device__ __forceinline__ uint32_t tc_eval_hidden(const TcGeom& g, TcLane& tl, long long& c0) {
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int layer = 0; layer + 1 < g.L; ++layer) {
    const uint32_t dcol = tc_wait_d(g, tl, true);
    { const long long c1 = clock64(); tl.c_wait += c1 - c0; c0 = c1; }
    tc_eval_epilogue<G, TERMS>(g, tl, layer, dcol, false, s);
    { const long long c1 = clock64(); tl.c_epi += c1 - c0; c0 = c1; }
  }
  const uint32_t dcol = tc_wait_d(g, tl, false);   // the output layer produces no units: no stagger
  if (tl.last_d_bar != nullptr) {
    __syncwarp();
    if ((tl.lane & 31) == 0) mbar_arrive(tl.last_d_bar);
  }
  { const long long c1 = clock64(); tl.c_wait += c1 - c0; c0 = c1; }
  return dcol;
}
// output layer: dot(h_L, w_last) over this thread's units -> tl.part
template <int G, int TERMS>
__device__ __forceinline__ void tc_eval_output(const TcGeom& g, TcLane& tl, uint32_t dcol) {
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  tc_eval_epilogue<G, TERMS>(g, tl, g.L - 1, dcol, true, s);
#ifdef IKR_TC_TRACE
  ++tl.trace_eval;
#endif
  tl.part[tl.group * kTcM + tl.lane] = (s[0] + s[1]) + (s[2] + s[3]);
}

// One MLP evaluation of the 128-trajectory tile, executed by EVERY lane thread (all G groups, masked
// lanes included).  The caller has published (nv, a) in tl.xin and passed the lanes barrier; the
// partial output sums land in tl.part (caller syncs).  Group c produces the units u = c (mod G) of
// every pass.  `hook` runs after this thread's layer-0 units are published, i.e. while the MMAs of
// layer 1 execute: the owners use it to compute time-only RHS terms of the next stage ahead.
template <int G, int TERMS = 3, typename Hook = TcNoHook>
__device__ __forceinline__ void tc_mlp_eval(const TcGeom& g, TcLane& tl, Hook hook = Hook()) {
  long long c0 = clock64();
  tc_eval_layer0<G, TERMS>(g, tl);
  { const long long c1 = clock64(); tl.c_l0 += c1 - c0; c0 = c1; }
  hook();
  const uint32_t dcol = tc_eval_hidden<G, TERMS>(g, tl, c0);
  tc_eval_output<G, TERMS>(g, tl, dcol);
  { const long long c1 = clock64(); tl.c_epi += c1 - c0; tl.c_last += c1 - c0; }
}

// Owner-side wrapper: publish the inputs, run the evaluation with the helper groups, collect.
template <int G, int TERMS = 3, typename Hook = TcNoHook>
__device__ __forceinline__ float tc_owner_eval(const TcGeom& g, TcLane& tl, float nv, float a,
                                               Hook hook = Hook()) {
  *reinterpret_cast<float2*>(tl.xin + 2 * tl.lane) = make_float2(nv, a);
  long long t0 = clock64();
  if (G > 1) lanes_sync<G>();        // inputs visible to the helper groups (cmd word = run)
  { const long long t1 = clock64(); tl.c_sync_a += t1 - t0; }
  tc_mlp_eval<G, TERMS, Hook>(g, tl, hook);
  t0 = clock64();
  if (G > 1) lanes_sync<G>();        // partial sums visible
  { const long long t1 = clock64(); tl.c_sync_b += t1 - t0; }
  float out = tl.part[tl.lane];
#pragma unroll
  for (int c = 1; c < G; ++c) out += tl.part[c * kTcM + tl.lane];
  return out + tl.sp[(size_t)(4 + g.L) * g.NP];
}

// ---- engine warps (shared by the forward, adjoint and regression kernels) ------------------------------
struct TcEngineCtx {
  uint64_t* bar_full;
  uint64_t* bar_empty;
  uint64_t* unit_ready;   // [kTcMaxUnits]
  uint64_t* unit_done;    // [kTcMaxUnits]
  uint64_t* bar_d;
  volatile int* stop_flag;
  unsigned char* ring;
};
__device__ __forceinline__ TcEngineCtx tc_engine_ctx(uint64_t* bars, volatile int* stop_flag, unsigned char* ring) {
  TcEngineCtx e;
  e.bar_full = bars;
  e.bar_empty = bars + kTcMaxStages;
  e.unit_ready = bars + 2 * kTcMaxStages;
  e.unit_done = e.unit_ready + kTcMaxUnits;
  e.bar_d = e.unit_done + kTcMaxUnits;
  e.stop_flag = stop_flag;
  e.ring = ring;
  return e;
}
// thread 0, before the first CTA barrier
__device__ __forceinline__ void tc_engine_init(const TcEngineCtx& e, int stages) {
  for (int s = 0; s < stages; ++s) {
    mbar_init(&e.bar_full[s], 1);
    mbar_init(&e.bar_empty[s], 1);
  }
  for (int s = 0; s < kTcMaxUnits; ++s) {
    mbar_init(&e.unit_ready[s], 4);     // the four warps of the producing column group
    mbar_init(&e.unit_done[s], 1);
  }
  mbar_init(e.bar_d, 1);
  mbar_fence_init();
  *e.stop_flag = 0;
}
__device__ __forceinline__ void tc_lane_attach(TcLane& tl, const TcEngineCtx& e) {
  tl.unit_ready = e.unit_ready;
  tl.unit_done = e.unit_done;
  tl.bar_d = e.bar_d;
  tl.phase_d = 0;
  tl.n_pass = 0;
  tl.unit_idx = 0;
  tl.upass = 0;
  tl.ring_base = 0;
  tl.last_d_bar = nullptr;
  tl.gate_bar = nullptr;
  tl.gate_parity = 0;
  tl.gate_pending = 0;
  tl.c_gate = 0;
}
// after the stop flag is set: the four warps of group 0 (the producers of unit 0) complete one more
// phase of unit_ready[0], on which the MMA warp is waiting between passes (it sees the flag and leaves)
__device__ __forceinline__ void tc_release_engines(const TcLane& tl) {
  __syncwarp();
  if (tl.group == 0 && (tl.lane & 31) == 0) mbar_arrive(&tl.unit_ready[0]);
}

// MMA issuer.  The whole warp runs the loop (warp-uniform control flow and operands => the descriptors
// live in uniform registers and each MMA is one UTCHMMA); one elected lane issues the MMAs and commits.
// A layer pass = the units 0 .. UT-1 in order; every unit is issued as soon as its slot is filled.
// LITE (bf16x3 only, adjoint kernel): three of the six products (a1 b1 + a2 b1 + a1 b2: products to
// ~2^-16, what the weight-gradient GEMM carries anyway) on the backward passes of an evaluation
// (LITE = 1: passes L .. 2L-1 of every 2L) or on all passes (LITE = 2).
template <int TERMS = 3, int LITE = 0>
__device__ __forceinline__ void tc_mma_warp(const TcGeom& g, const TcEngineCtx& c, uint32_t tbase, bool timing) {
  const unsigned stages = (unsigned)g.stages;
  constexpr bool f16 = TERMS == 2;      // compile time: the six / three MMAs stay back-to-back UTCHMMAs
  const uint32_t idesc = f16 ? tc::idesc_f16_f32(kTcM, g.NP) : tc::idesc_bf16_f32(kTcM, g.NP);
  const uint64_t desc0 = tc::smem_desc(smem_u32(c.ring), 128, 256);
  const uint32_t blk16 = (uint32_t)g.block_bytes >> 4, stage16 = (uint32_t)g.ring_stride >> 4;
  const int UT = g.units + g.tail;
  unsigned s = 0, round = 0, consumed = 0, gi = 0, pass = 0;
  long long e_wait = 0, e_wait0 = 0, e_wait00 = 0, e_issue = 0, ec = clock64();
  bool stop = false;
  while (!stop) {
    const uint32_t dcol = tbase + ((pass & 1u) ? (uint32_t)g.NP : 0u);
    const bool lite = LITE == 2 || (LITE == 1 && (pass % (2u * (unsigned)g.L)) >= (unsigned)g.L);
    bool first = true;
    for (int u = 0; u < UT; ++u, ++gi) {
      const unsigned slot = gi % tc_ring_slots<tc_ring_terms<TERMS, LITE>()>();
      { const long long c1 = clock64(); e_issue += c1 - ec; ec = c1; }
      mbar_wait(&c.unit_ready[u], pass & 1u);
      if (*c.stop_flag) { stop = true; break; }
      tc::fence_after_sync();
      { const long long c1 = clock64(); e_wait += c1 - ec; if (u == 0) { e_wait0 += c1 - ec; if (pass % (unsigned)g.L == 0) e_wait00 += c1 - ec; } ec = c1; }
      const uint32_t a_slot = tbase + (uint32_t)g.col_ring + (uint32_t)(16 * tc_ring_terms<TERMS, LITE>()) * slot;
      const bool is_tail = u >= g.units;
      const int nk = is_tail ? 1 : ((2 * u + 1 < g.KSf) ? 2 : 1);
#pragma unroll 1
      for (int jj = 0; jj < nk; ++jj) {
        mbar_wait(&c.bar_full[s], round);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t b1 = desc0 + (uint64_t)(s * stage16);
          const uint64_t b2 = b1 + blk16, b3 = b2 + blk16;
          if (f16) {
            // fp16x2: a1 b1 + a2 b1 + a1 b2; tail A block [a1t | a2t] against [b1t | b1t], [b2t | 0]
            if (!is_tail) {
              const uint32_t dt = nk == 2 ? 16u : 8u;
              const uint32_t a1 = a_slot + 8u * (uint32_t)jj;
              tc::mma_ts(dcol, a1, b1, idesc, first ? 0u : 1u);
              tc::mma_ts(dcol, a1 + dt, b1, idesc, 1u);
              tc::mma_ts(dcol, a1, b2, idesc, 1u);
            } else {
              tc::mma_ts(dcol, a_slot, b1, idesc, first ? 0u : 1u);
              tc::mma_ts(dcol, a_slot, b2, idesc, 1u);
            }
          } else if (!is_tail) {
            const uint32_t dt = nk == 2 ? 16u : 8u;
            const uint32_t a1 = a_slot + 8u * (uint32_t)jj;
            tc::mma_ts(dcol, a1, b1, idesc, first ? 0u : 1u);
            tc::mma_ts(dcol, a1 + dt, b1, idesc, 1u);
            if (LITE == 0 || !lite) tc::mma_ts(dcol, a1 + 2 * dt, b1, idesc, 1u);
            tc::mma_ts(dcol, a1, b2, idesc, 1u);
            if (LITE == 0 || !lite) {
              tc::mma_ts(dcol, a1 + dt, b2, idesc, 1u);
              tc::mma_ts(dcol, a1, b3, idesc, 1u);
            }
          } else {
            tc::mma_ts(dcol, a_slot, b1, idesc, first ? 0u : 1u);
            tc::mma_ts(dcol, a_slot, b2, idesc, 1u);
            if (LITE == 0 || !lite) tc::mma_ts(dcol, a_slot + 8u, b3, idesc, 1u);
          }
          tc::commit(smem_u32(&c.bar_empty[s]));   // weight slot free once these MMAs have read it
        }
        __syncwarp();
        first = false;
        ++consumed;
        if (++s == stages) { s = 0; round ^= 1u; }
      }
      if (tc::elect_one()) tc::commit(smem_u32(&c.unit_done[u]));      // its A slot may be refilled
      __syncwarp();
    }
    if (stop) break;
    if (tc::elect_one()) tc::commit(smem_u32(c.bar_d));
    __syncwarp();
    ++pass;
  }
  // stop: every MMA has completed (its D was consumed).  The producer can only be blocked on the
  // slot of the next k-step to consume: complete that phase by hand (it re-checks the flag).
  if (tc::elect_one()) {
    mbar_arrive(&c.bar_empty[s]);
    if (timing)
      printf("[tc timing] mma warp: wait_units %lld (first unit of a pass: %lld, of which first pass of an "
             "evaluation [forward kernels]: %lld) issue %lld cycles, %u k-steps, %u passes\n",
             e_wait, e_wait0, e_wait00, e_issue, consumed, pass);
  }
  __syncwarp();
}

// Weight producer (one thread): streams the image cyclically (`per_cycle` k-steps) through the ring.
// `copy_bytes` < g.stage_bytes: only the leading term blocks of every stage are needed (the adjoint
// kernel's three-product passes never read the third block)
__device__ __forceinline__ void tc_producer_thread(const TcGeom& g, const TcEngineCtx& c,
                                                   const unsigned char* img, unsigned per_cycle,
                                                   unsigned copy_bytes = 0u) {
  if (copy_bytes == 0u) copy_bytes = (unsigned)g.stage_bytes;
  const unsigned stages = (unsigned)g.stages;
  unsigned issued = 0;
  for (;; ++issued) {
    const unsigned q = issued, s = q % stages;
    if (q >= stages) mbar_wait(&c.bar_empty[s], ((q / stages) - 1u) & 1u);
    if (*c.stop_flag) break;
    mbar_expect_tx(&c.bar_full[s], copy_bytes);
    bulk_g2s(c.ring + (size_t)s * g.ring_stride, img + (size_t)(q % per_cycle) * g.stage_bytes,
             copy_bytes, &c.bar_full[s]);
  }
  // wait for the copies still in flight (the last `stages` issued chunks cover every slot once)
  for (unsigned q = issued > stages ? issued - stages : 0; q < issued; ++q)
    mbar_wait(&c.bar_full[q % stages], (q / stages) & 1u);
}

// fp16x2 split, owner side: a non-finite MLP output for finite inputs means that a hidden activation
// left the fp16 range (ikr_tc.cuh pack_f16x2).  The output is replaced by a huge finite value, so that
// the dopri5 step is REJECTED with the controller's smallest factor -- what the reference does with
// the finite, huge derivative it computes for such a wild trial stage -- and `hit` is raised.  A hit at
// the initial evaluations, on an rk4 step, or on kTcRangeRejects rejected attempts in a row is a
// range violation in the physical domain: the lane ends with LANE_RANGE (IKR_TC_RANGE).
constexpr int kTcRangeRejects = 12;
template <int TERMS>
__device__ __forceinline__ float tc_range_filter(float out, double nv, double a, bool& hit) {
  if (TERMS == 2 && !isfinite(out) && isfinite((float)nv) && isfinite((float)a)) {
    hit = true;
    return 1.0e30f;
  }
  return out;
}
template <int TERMS, typename S>
__device__ __forceinline__ void tc_range_after_step(Lane<S>& L, bool accepted, bool& hit, int& rejects) {
  if (TERMS != 2) return;
  if (hit) {
    rejects = accepted ? kTcRangeRejects : rejects + 1;
    if (rejects >= kTcRangeRejects && (L.status == LANE_OK || L.status == LANE_DONE)) L.status = LANE_RANGE;
  } else if (accepted) {
    rejects = 0;
  }
  hit = false;
}
template <int TERMS, typename S>
__device__ __forceinline__ void tc_range_fatal(Lane<S>& L, bool& hit) {
  if (TERMS == 2 && hit && (L.status == LANE_OK || L.status == LANE_DONE)) L.status = LANE_RANGE;
  hit = false;
}

// element i = (row, c) of the zero-padded small-parameter block kept in shared memory: rows w0a | w0b |
// b0 | L x hidden bias | w_last, then b_last and the per-layer accumulator scales.  The fp16x2 path
// works on 16 x the activations (kTcActShift): w0, b0 and the hidden biases are stored x 16, w_last
// x 1/16 (exact), and D of layer l is multiplied by scales[l] = 2^-k_l (the weight image holds
// W_l 2^k_l, see ikr_tc_pack_kernel).
template <int TERMS>
__device__ __forceinline__ float tc_small_param(const TcGeom& g, const MlpView& mlp, const float* P,
                                                const float* scales, int i, int row, int c, int NP,
                                                int npad, int n) {
  const float up = TERMS == 2 ? (float)(1 << kTcActShift) : 1.0f;
  float v = 0.0f;
  if (row < 3) { if (c < n) v = P[mlp.off_w0 + (long long)row * npad + c] * up; }
  else if (row < 3 + g.L) { if (c < n) v = P[mlp.off_bh + (long long)(row - 3) * npad + c] * up; }
  else if (row == 3 + g.L) { if (c < n) v = P[mlp.off_wl + c] / up; }
  else if (i == (4 + g.L) * NP) v = P[mlp.off_wl + npad];
  else if (i >= tc_scale_index(g, 0) && i < tc_scale_index(g, 0) + g.L)
    v = TERMS == 2 ? scales[i - tc_scale_index(g, 0)] : 1.0f;
  return v;
}

template <typename S, int G, int TERMS>
__global__ void __launch_bounds__(tc_threads(G), 1) ikr_forward_tc_kernel(const TcFwdParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FwdParams& p = tp.f;
  const TcGeom g = tp.g;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  constexpr int kLaneThreads = 128 * G;
  constexpr int kMmaWarp = 4 * G, kLoadWarp = 4 * G + 1;
  const TcSmemLayout<S> lay(g, g.stages, G);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  long long* tile_slot = reinterpret_cast<long long*>(smem_raw + lay.off_misc + 8);
  volatile int* cmd_exit = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 16);
  FwdJob* jobp = reinterpret_cast<FwdJob*>(smem_raw + lay.off_job);
  Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
  double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);
  float* sp = reinterpret_cast<float*>(smem_raw + lay.off_sp);
  const TcEngineCtx eng = tc_engine_ctx(bars, stop_flag, smem_raw + lay.off_ring);

  // ---- one-time setup ------------------------------------------------------------------------------
  if (tid == 0) {
    tc_engine_init(eng, g.stages);
    *cmd_exit = 0;
  }
  if (warp == kMmaWarp) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  {
    // small parameters, zero-padded to stride NP
    const float* P = reinterpret_cast<const float*>(p.mlp.base);
    const int NP = g.NP, npad = p.mlp.npad, n = g.n;
    for (int i = tid; i < g.small_elems; i += tc_threads(G)) {
      const int row = i / NP, c = i - row * NP;
      sp[i] = tc_small_param<TERMS>(g, p.mlp, P, tp.scales, i, row, c, NP, npad, n);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kMmaWarp) {
    tc_mma_warp<TERMS>(g, eng, tbase, tp.timing && blockIdx.x == 0);
  } else if (warp == kLoadWarp) {
    if ((tid & 31) == 0) tc_producer_thread(g, eng, reinterpret_cast<const unsigned char*>(tp.img),
                                            (unsigned)(g.L * g.KST));
  } else {
    // ================================ lane threads ======================================================
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
#ifdef IKR_TC_TRACE
    tl.trace_eval = 0;
#endif
    const long long c_begin = clock64();
    long long c_finish = 0;

    if (tl.group > 0) {
      // ---- helper groups: evaluate on command until the owners say exit -----------------------------
      while (true) {
        lanes_sync<G>();
        if (*cmd_exit) break;
        tc_mlp_eval<G, TERMS>(g, tl);
        lanes_sync<G>();
      }
    } else {
      // ---- owner threads: the solver ---------------------------------------------------------------
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
        const long long b = (tile - job.tile_begin) * p.M + tid;   // p.M <= 128 trajectories per tile
        const bool valid = tid < p.M && b < jB;
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
        bool range_hit = false;
        int range_rejects = 0;
        if (p.method == 0) {
          init_prepare_f0<S>(L, cfg, &nv, &ain);
          float out = tc_range_filter<TERMS>(tc_owner_eval<G, TERMS>(g, tl, (float)nv, (float)ain), nv, ain, range_hit);
          init_store_f0<S>(L, cfg, (double)out);
          if (cfg.first_step > 0) {
            L.dt = cfg.first_step;
          } else {
            init_prepare_f1<S>(L, cfg, &nv, &ain);
            out = tc_range_filter<TERMS>(tc_owner_eval<G, TERMS>(g, tl, (float)nv, (float)ain), nv, ain, range_hit);
            init_store_f1<S>(L, cfg, (double)out);
          }
          tc_range_fatal<TERMS, S>(L, range_hit);
          if (T <= 1 && lane_active(L)) L.status = LANE_DONE;

          TimeCache tcache;
          tcache.valid = 0;
          while (true) {
            dp_check_before_step<S>(L, cfg);
            if (!owners_or(lane_active(L) ? 1 : 0)) break;
#pragma unroll 1
            for (int s = 0; s < 6; ++s) {
              dp_prepare_stage_cached<S>(L, cfg, s, &nv, &ain, tcache);
              // while the MMAs of this stage run: V(t) and the HH rates of the next stage
              out = tc_owner_eval<G, TERMS>(g, tl, (float)nv, (float)ain, [&]() {
                if (s < 5 && lane_active(L)) dp_prefetch_stage_time<S>(L, cfg, s + 1, tcache);
              });
              out = tc_range_filter<TERMS>(out, nv, ain, range_hit);
              dp_store_stage<S>(L, cfg, s, (double)out);
            }
            const long long cf0 = clock64();
            const int acc0 = L.n_acc;
            dp_finish_step<S>(L, cfg, job.t_out, T, emit, ckpt);
            tc_range_after_step<TERMS, S>(L, L.n_acc != acc0, range_hit, range_rejects);
            c_finish += clock64() - cf0;
          }
        } else {
          if (T <= 1 && lane_active(L)) L.status = LANE_DONE;
          for (int gi = 0; gi + 1 < job.G; ++gi) {
            const double g0 = job.grid[gi], g1 = job.grid[gi + 1];
            if (!owners_or(lane_active(L) ? 1 : 0)) break;
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {
              rk4_prepare_stage<S>(L, cfg, s, g0, g1, p.time_f32 != 0, p.rk4_perturb != 0, &nv, &ain);
              const float out = tc_range_filter<TERMS>(tc_owner_eval<G, TERMS>(g, tl, (float)nv, (float)ain), nv, ain, range_hit);
              rk4_store_stage<S>(L, cfg, s, (double)out);
            }
            tc_range_fatal<TERMS, S>(L, range_hit);
            if (lane_active(L)) {
              // step checkpoint for the backward sweep: (g0, g1), y0, k1..k4
              L.t0 = g0; L.dt = g1;
              if (!ckpt(L.n_acc, L)) L.status = LANE_CKPT_OVERFLOW;
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
      if (tp.timing && blockIdx.x == 0 && tid == 0) {
        const long long tot = clock64() - c_begin;
        printf("[tc timing] owner 0: total %lld cycles: layer0 %lld, wait_d %lld, epilogue %lld (output layer %lld), "
               "barrier A %lld, barrier B %lld, solver %lld (of which dp_finish_step %lld)\n",
               tot, tl.c_l0, tl.c_wait, tl.c_epi, tl.c_last, tl.c_sync_a, tl.c_sync_b,
               tot - tl.c_l0 - tl.c_wait - tl.c_epi - tl.c_sync_a - tl.c_sync_b, c_finish);
      }
      // release the helper groups and the engine warps
      if (tid == 0) { *cmd_exit = 1; *stop_flag = 1; }
      owners_sync();
      if (G > 1) lanes_sync<G>();
    }
    tc_release_engines(tl);     // wakes the MMA warp, which sees the stop flag
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

// =============================================================================================
// Lane-pool variant (dopri5): every CTA owns 128 lane SLOTS; a slot whose trajectory has finished
// pulls the next trajectory -- of whatever job -- from one global queue (jobs longest first), so no
// TMEM lane idles while its neighbours finish and there is no tile / wave quantisation.  Lanes stay
// in lock-step by ROUND of six RHS evaluations: a lane attempts one dopri5 step, or, right after a
// refill, runs its two start-up evaluations.  Every lane's arithmetic is independent of its slot,
// so results are bit-identical to the tile-scheduled kernel.
// =============================================================================================
template <typename S, int G, int TERMS>
__global__ void __launch_bounds__(tc_threads(G), 1) ikr_forward_tc_pool_kernel(const TcFwdParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FwdParams& p = tp.f;
  const TcGeom g = tp.g;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kLaneThreads = 128 * G;
  constexpr int kMmaWarp = 4 * G, kLoadWarp = 4 * G + 1;
  const TcSmemLayout<S> lay(g, g.stages, G);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  volatile int* cmd_exit = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 16);
  Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
  double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);
  LaneAux<S>* aux = reinterpret_cast<LaneAux<S>*>(smem_raw + lay.off_aux);
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
      sp[i] = tc_small_param<TERMS>(g, p.mlp, P, tp.scales, i, row, c, NP, npad, n);
    }
  }
  // job descriptors: shared-memory copy of the inline table (every lane reads its job every round)
  constexpr size_t kJobStride = (sizeof(FwdJob) + 15) & ~(size_t)15;
  unsigned char* sjobs = smem_raw + lay.off_job;
  if (p.jobs_are_inline) {
    const int words = (int)(sizeof(FwdJob) / 4);
    for (int i = tid; i < p.n_jobs * words; i += tc_threads(G)) {
      const int j = i / words, w = i - j * words;
      reinterpret_cast<uint32_t*>(sjobs + (size_t)j * kJobStride)[w] =
          reinterpret_cast<const uint32_t*>(&p.jobs_inline[j])[w];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kMmaWarp) {
    tc_mma_warp<TERMS>(g, eng, tbase, tp.timing && blockIdx.x == 0);
  } else if (warp == kLoadWarp) {
    if ((tid & 31) == 0) tc_producer_thread(g, eng, reinterpret_cast<const unsigned char*>(tp.img),
                                            (unsigned)(g.L * g.KST));
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
#ifdef IKR_TC_TRACE
    tl.trace_eval = 0;
#endif

    if (tl.group > 0) {
      while (true) {
        lanes_sync<G>();
        if (*cmd_exit) break;
        tc_mlp_eval<G, TERMS>(g, tl);
        lanes_sync<G>();
      }
    } else {
      const bool heuristic = !(p.cfg.first_step > 0);
      Lane<S>& L = lanes[tid];
      LaneAux<S>& A = aux[tid];
      SolverCfg c = p.cfg;                      // per-thread copy; c.tab follows the lane's job
      auto job_of = [&](int j) -> const FwdJob* {
        return p.jobs_are_inline ? reinterpret_cast<const FwdJob*>(sjobs + (size_t)j * kJobStride) : p.jobs + j;
      };
      const FwdJob* jobp = job_of(0);
      TimeCache tcache;
      tcache.valid = 0;
      lane_reset<S>(L, (S)0, (S)1, 0.0, false);
      A.mode = POOL_EMPTY; A.job = 0; A.b = 0; A.g = (S)1; A.e = (S)0;
      bool queue_dry = false;
      bool range_hit = false;
      int range_rejects = 0;

      while (true) {
        // ---- round boundary: retire finished trajectories, refill free slots ----------------------
        if (A.mode == POOL_STEP) dp_check_before_step<S>(L, c);
        if (A.mode != POOL_EMPTY && !lane_active(L)) {
          const FwdJob& job = *jobp;
          int* st = job.stats_out + 4 * A.b;
          st[0] = L.n_acc; st[1] = L.n_rej; st[2] = L.nfe;
          st[3] = L.status == LANE_DONE ? 0 : L.status;
          if (job.loss_out) {
            job.loss_out[2 * A.b] = obs[2 * tid];
            job.loss_out[2 * A.b + 1] = obs[2 * tid + 1];
          }
          A.mode = POOL_EMPTY;
        }
        if (A.mode == POOL_EMPTY && !queue_dry) {
          const long long gidx = (long long)atomicAdd(p.queue, 1ULL);
          if (gidx < p.n_traj) {
            int j = 0;
            while (j + 1 < p.n_jobs && job_of(j + 1)->traj_begin <= gidx) ++j;
            jobp = job_of(j);
            const FwdJob& job = *jobp;
            c.tab = job.tab;
            const long long b = gidx - job.traj_begin;
            const S* y0 = reinterpret_cast<const S*>(job.y0);
            A.job = j; A.b = b; A.mode = POOL_INIT;
            A.g = job.g ? reinterpret_cast<const S*>(job.g)[b] : (S)1;
            A.e = job.e_rev ? reinterpret_cast<const S*>(job.e_rev)[b] : (S)job.e_scalar;
            lane_reset<S>(L, y0[2 * b], y0[2 * b + 1], job.t_out[0], true);
            range_hit = false; range_rejects = 0;
            obs[2 * tid] = 0.0; obs[2 * tid + 1] = 0.0;
          } else {
            queue_dry = true;
          }
        }
        if (!owners_or(A.mode != POOL_EMPTY ? 1 : 0)) break;

        // ---- one round = six RHS evaluations ---------------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          int what = 0;   // 0 masked, 1 dopri5 stage, 2 f0, 3 initial-step probe
          double nv = 0, ain = 0;
          if (A.mode == POOL_STEP) what = 1;
          else if (A.mode == POOL_INIT && s == 0) what = 2;
          else if (A.mode == POOL_INIT && s == 1 && heuristic) what = 3;
          if (what) {
            if (what == 1) dp_prepare_stage_cached<S>(L, c, s, &nv, &ain, tcache);
            else if (what == 2) init_prepare_f0<S>(L, c, &nv, &ain);
            else init_prepare_f1<S>(L, c, &nv, &ain);
          }
          float out = tc_owner_eval<G, TERMS>(g, tl, (float)nv, (float)ain, [&]() {
            if (what == 1 && s < 5) dp_prefetch_stage_time<S>(L, c, s + 1, tcache);
          });
          if (what) out = tc_range_filter<TERMS>(out, nv, ain, range_hit);
          if (what == 1) dp_store_stage<S>(L, c, s, (double)out);
          else if (what == 2) {
            init_store_f0<S>(L, c, (double)out);
            if (!heuristic) L.dt = c.first_step;
          } else if (what == 3) init_store_f1<S>(L, c, (double)out);
        }

        // ---- end of round: finish the attempted step / leave start-up ----------------------------------
        if (A.mode == POOL_STEP) {
          const FwdJob& job = *jobp;
          const long long jB = job.B, b = A.b;
          S* y_out = reinterpret_cast<S*>(job.y_out);
          S* i_out = reinterpret_cast<S*>(job.i_out);
          S* ckpt_y = reinterpret_cast<S*>(job.ckpt_y);
          const S* dptr = reinterpret_cast<const S*>(job.data);
          const bool observe = (job.v_out != nullptr) && (job.i_out != nullptr || job.loss_out != nullptr);
          const S g_b = A.g, e_b = A.e;
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
          const int acc0 = L.n_acc;
          dp_finish_step<S>(L, c, job.t_out, job.T, emit, ckpt);
          tc_range_after_step<TERMS, S>(L, L.n_acc != acc0, range_hit, range_rejects);
        } else if (A.mode == POOL_INIT) {
          tc_range_fatal<TERMS, S>(L, range_hit);      // f0 / the initial-step probe sit at y0
          // start-up done: emit y(t[0]) = y0 and start stepping (or finish if there is one output)
          const FwdJob& job = *jobp;
          const long long b = A.b;
          if (job.y_out) {
            typename Vec2<S>::type v;
            v.x = L.ya; v.y = L.yr;
            reinterpret_cast<typename Vec2<S>::type*>(job.y_out)[b] = v;
          }
          if (job.v_out && (job.i_out || job.loss_out)) {
            double cur = (double)(A.g * L.ya * L.yr) * (job.v_out[0] - (double)A.e);
            if (job.i_out) reinterpret_cast<S*>(job.i_out)[b] = (S)cur;
            if (job.data) {
              const S* dptr = reinterpret_cast<const S*>(job.data);
              double diff = cur - (double)dptr[job.data_B == 1 ? 0 : b];
              obs[2 * tid] += diff * diff;
              obs[2 * tid + 1] += fabs(diff);
            }
          }
          A.mode = POOL_STEP;
          if (job.T <= 1 && lane_active(L)) L.status = LANE_DONE;
        }
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

}  // namespace ikr
#endif  // IKR_FORWARD_TC_CUH_
