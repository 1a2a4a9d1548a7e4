// ikr_device.cuh -- device building blocks shared by the forward and backward kernels:
// mbarrier / bulk-copy (TMA) wrappers, the packed-parameter view, and the register-tiled
// trajectory-tile MLP (one CTA evaluates the MLP of M trajectories at once; hidden-layer weights
// are streamed L2 -> shared memory by cp.async.bulk into a ring of chunks, activations stay in
// shared memory feature-major, every thread owns an 8 x TN register tile of the output).
#ifndef IKR_DEVICE_CUH_
#define IKR_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#include "ikr_math.h"

namespace ikr {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
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
#ifdef IKR_MBAR_DEBUG
// bring-up build: a wait that does not end within ~2 s reports where it sits and aborts the kernel
__device__ __noinline__ void mbar_wait_dbg(uint64_t* bar, unsigned parity, int line) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("[mbar timeout] line %d block %d thread %d bar smem 0x%x parity %u\n", line, (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
#endif
// wait that yields the issue slots between polls (long waits of many threads on one barrier)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, unsigned parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float ikr_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// Packed FP32 pair FMA (sm_100 FFMA2): c.{lo,hi} = a.{lo,hi} * b.{lo,hi} + c.{lo,hi}, each half
// rounded exactly like a scalar fma.rn.  ptxas folds a {x, x} pack into a scalar-broadcast
// operand, so an outer-product update costs half the issue slots of scalar FFMA.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void f2_fma(f32x2& c, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ double ikr_fma(double a, double b, double c) { return __fma_rn(a, b, c); }

// ---------------------------------------------------------------------------------------------
// Packed parameters (element offsets into one buffer of the MLP dtype; see include/ikr.h)
// ---------------------------------------------------------------------------------------------
struct MlpView {
  const void* base;
  int L, n, npad;
  int kc;   // k-rows per streamed weight chunk
  int cpl;  // chunks per hidden layer
  long long off_w0;  // [3][npad]   w0[:,0] | w0[:,1] | b0
  long long off_wt;  // [L][n][npad] forward operand (k-major = W^T)
  long long off_bh;  // [L][npad]
  long long off_wl;  // [npad] + b_last
  long long off_wn;  // [L][n][npad] backward operand (rows = out features)
  double slope;
  int bwd_seq;       // 0: chunk sequence = forward layers only; 1: forward layers then backward
                     //    layers (adjoint kernel: Wt of layers 0..L-1, then Wn of layers L-1..0)
};

template <typename W>
struct MlpTileCfg;
template <>
struct MlpTileCfg<float> {
  static constexpr int TN = 8;  // output features per thread (wide tile; the narrow tile is TN / 2)
  static constexpr int V = 4;   // elements per 16-byte shared-memory vector
};
template <>
struct MlpTileCfg<double> {
  static constexpr int TN = 4;
  static constexpr int V = 2;
};
constexpr int kTM = 8;      // trajectories per thread
constexpr int kStages = 3;  // weight-chunk ring depth

// Shared-memory resources of the trajectory-tile MLP
template <typename W>
struct MlpSmem {
  W* xin;          // [2][M]   MLP inputs (nv, a)
  W* Hs;           // [npad][M] activations, feature-major
  W* Wr;           // [kStages][kc][npad] weight-chunk ring
  W* sp;           // small parameters staged once per CTA: w0a | w0b | b0 | L x b_hidden | w_last | b_last
  uint64_t* full;   // [kStages] chunk landed (TMA complete_tx)
  uint64_t* empty;  // [kStages] every worker warp is done reading the chunk
};

struct MlpPipe {
  unsigned q;       // next chunk (global sequence number) to consume
  unsigned issued;  // (thread 0) chunks issued so far
};

template <typename W>
__device__ __forceinline__ void mlp_issue_chunk(const MlpView& mv, const MlpSmem<W>& sm,
                                                unsigned q) {
  const unsigned stage = q % kStages;
  const unsigned c = q % (unsigned)mv.cpl;
  unsigned lq = q / (unsigned)mv.cpl;
  long long off = mv.off_wt;
  unsigned layer;
  if (mv.bwd_seq) {
    lq %= 2u * (unsigned)mv.L;
    if (lq < (unsigned)mv.L) {
      layer = lq;
    } else {
      layer = 2u * (unsigned)mv.L - 1u - lq;
      off = mv.off_wn;
    }
  } else {
    layer = lq % (unsigned)mv.L;
  }
  const int k0 = (int)c * mv.kc;
  const int rows = min(mv.kc, mv.n - k0);
  const unsigned bytes = (unsigned)(rows * mv.npad * (int)sizeof(W));
  const W* src = (const W*)mv.base + off + ((long long)layer * mv.n + k0) * mv.npad;
  mbar_expect_tx(&sm.full[stage], bytes);
  bulk_g2s(sm.Wr + (size_t)stage * mv.kc * mv.npad, src, bytes, &sm.full[stage]);
}

// Barrier initialisation (thread 0) + prologue filling the ring.  `n_worker_warps` warps arrive on
// the empty barriers.  Must be followed by __syncthreads() before first use.
template <typename W>
__device__ __forceinline__ void mlp_pipe_init(const MlpSmem<W>& sm, int n_worker_warps) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], n_worker_warps);
    }
    mbar_fence_init();
  }
}
template <typename W>
__device__ __forceinline__ void mlp_pipe_start(const MlpView& mv, const MlpSmem<W>& sm,
                                               MlpPipe& pp) {
  pp.q = 0;
  pp.issued = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mlp_issue_chunk<W>(mv, sm, pp.issued++);
  }
}
// epilogue: wait for every chunk still in flight before the CTA exits (call after a barrier)
template <typename W>
__device__ __forceinline__ void mlp_pipe_drain(const MlpView& mv, const MlpSmem<W>& sm,
                                               MlpPipe& pp) {
  if (threadIdx.x == 0) {
    for (unsigned q = pp.q; q < pp.issued; ++q) mbar_wait(&sm.full[q % kStages], (q / kStages) & 1);
  }
}

// number of elements of the small-parameter block ((3 + L + 1) npad + 8)
__host__ __device__ inline size_t mlp_small_elems(int L, int npad) { return (size_t)(4 + L) * npad + 8; }
// cooperative copy global -> shared (call once, follow with __syncthreads())
template <typename W>
__device__ __forceinline__ void mlp_stage_small(const MlpView& mv, const MlpSmem<W>& sm) {
  const W* P = (const W*)mv.base;
  const int n0 = 3 * mv.npad, n1 = (mv.L + 1) * mv.npad + 8;
  for (int i = threadIdx.x; i < n0; i += blockDim.x) sm.sp[i] = P[mv.off_w0 + i];
  for (int i = threadIdx.x; i < n1; i += blockDim.x) sm.sp[n0 + i] = P[mv.off_bh + i];
}
template <typename W>
__device__ __forceinline__ const W* sp_w0(const MlpView& mv, const MlpSmem<W>& sm) { return sm.sp; }
template <typename W>
__device__ __forceinline__ const W* sp_bh(const MlpView& mv, const MlpSmem<W>& sm, int layer) {
  return sm.sp + (size_t)(3 + layer) * mv.npad;
}
template <typename W>
__device__ __forceinline__ const W* sp_wl(const MlpView& mv, const MlpSmem<W>& sm) {
  return sm.sp + (size_t)(3 + mv.L) * mv.npad;
}

template <typename W>
__device__ __forceinline__ W leaky(W x, W slope) { return x > (W)0 ? x : x * slope; }

// Thread -> register-tile coordinates.  Threads are grouped in quarter-warps (8 lanes): the 8
// lanes of a quarter-warp own 8 consecutive trajectory groups gm of ONE feature group gn, so the
// activation vector load of a quarter-warp is one aligned 128-byte line and the weight vector
// load is a broadcast -- no shared-memory bank conflicts.  MG = 8a + rem: a remainder column of
// rem in {1, 2, 4} packs 8/rem feature groups per quarter-warp; other remainders use one partly
// idle quarter-warp per feature group.
struct TileCoord {
  int gm, gn;
  bool worker;
};
__host__ __device__ inline int tile_pack(int rem) { return (rem == 1 || rem == 2 || rem == 4) ? 8 / rem : 1; }
__host__ __device__ inline int tile_worker_threads(int MG, int NG) {
  const int a = MG / 8, rem = MG % 8;
  int qw = a * NG;
  if (rem != 0) {
    const int pk = tile_pack(rem);
    qw += (NG + pk - 1) / pk;
  }
  return qw * 8;
}
__device__ __forceinline__ TileCoord tile_coord(int tid, int MG, int NG) {
  const int a = MG / 8, rem = MG % 8;
  const int qw = tid >> 3, l8 = tid & 7;
  TileCoord c;
  const int n_full = a * NG;
  if (qw < n_full) {
    c.gn = qw / a;
    c.gm = (qw - c.gn * a) * 8 + l8;
    c.worker = true;
  } else if (rem != 0) {
    const int pk = tile_pack(rem);
    const int r = qw - n_full;
    if (pk > 1) {
      c.gn = pk * r + l8 / rem;
      c.gm = 8 * a + l8 % rem;
      c.worker = c.gn < NG;
    } else {
      c.gn = r;
      c.gm = 8 * a + l8;
      c.worker = c.gn < NG && l8 < rem;
    }
  } else {
    c.gn = 0; c.gm = 0; c.worker = false;
  }
  if (!c.worker) { c.gm = 0; c.gn = 0; }
  return c;
}

// row i (0..7) of a thread's register tile -> trajectory index inside the CTA tile.
// Rows are split in 8/V groups of V consecutive trajectories so that every 16-byte
// shared-memory vector access of a warp is contiguous (bank-conflict free).
template <int V>
__device__ __forceinline__ int tile_row(int i, int gm, int MG) {
  return (i / V) * (MG * V) + gm * V + (i % V);
}

// K-loop of one n x n layer over the weight-chunk ring: acc[i][j] += sum_k Hs[k][row_i] * Wr[k][col_j]
// for the thread's 8 x TN register tile.  No CTA barrier inside (full/empty mbarriers only).
template <typename W, int TN>
__device__ __forceinline__ void mlp_layer_kloop(const MlpView& mv, const MlpSmem<W>& sm,
                                                MlpPipe& pp, int M, int MG, const TileCoord& tc,
                                                bool warp_works, W (&acc)[kTM][TN]) {
  constexpr int V = MlpTileCfg<W>::V;
  constexpr int NP = TN / 2 > 0 ? TN / 2 : 1;   // packed fp32 pairs per row of the register tile
  const int tid = threadIdx.x;
  const int gm = tc.gm, gn = tc.gn;
  const bool worker = tc.worker;
  // fp32: accumulate in packed pairs (acc[i][2p], acc[i][2p+1]) with FFMA2
  f32x2 c2[kTM][NP];
  if (sizeof(W) == 4) {
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
      for (int pj = 0; pj < NP; ++pj) c2[i][pj] = f2_pack((float)acc[i][2 * pj], (float)acc[i][2 * pj + 1]);
  }
  for (int c = 0; c < mv.cpl; ++c) {
    const unsigned q = pp.q;
    const unsigned stage = q % kStages;
    const int k0 = c * mv.kc;
    const int rows = min(mv.kc, mv.n - k0);
    if (warp_works) {
      mbar_wait(&sm.full[stage], (q / kStages) & 1);
      if (worker) {
        const W* Hk = sm.Hs + (size_t)k0 * M + gm * V;
        const W* Wk = sm.Wr + (size_t)stage * mv.kc * mv.npad + gn * TN;
        if (sizeof(W) == 4) {
          // software-pipelined: the fragments of row kk+1 are loaded before row kk is consumed
          // (the ring is padded by one row so the last prefetch stays inside shared memory)
          const float* hp = reinterpret_cast<const float*>(Hk);
          const float* wp = reinterpret_cast<const float*>(Wk);
          const int a1_off = MG * V;
          float4 a0 = *reinterpret_cast<const float4*>(hp);
          float4 a1 = *reinterpret_cast<const float4*>(hp + a1_off);
          float4 b0 = *reinterpret_cast<const float4*>(wp);
          float4 b1 = b0;
          if (TN == 8) b1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll 2
          for (int kk = 0; kk < rows; ++kk) {
            hp += M;
            wp += mv.npad;
            const float4 na0 = *reinterpret_cast<const float4*>(hp);
            const float4 na1 = *reinterpret_cast<const float4*>(hp + a1_off);
            const float4 nb0 = *reinterpret_cast<const float4*>(wp);
            float4 nb1 = nb0;
            if (TN == 8) nb1 = *reinterpret_cast<const float4*>(wp + 4);
            const float a[kTM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const f32x2 b[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y),
                                f2_pack(b1.z, b1.w)};
#pragma unroll
            for (int i = 0; i < kTM; ++i) {
              const f32x2 aa = f2_pack(a[i], a[i]);
#pragma unroll
              for (int pj = 0; pj < NP; ++pj) f2_fma(c2[i][pj], aa, b[pj]);
            }
            a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
          }
        } else {
#pragma unroll 2
          for (int kk = 0; kk < rows; ++kk) {
            W a[kTM], b[TN];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              double2 av = *reinterpret_cast<const double2*>(Hk + (size_t)kk * M + g * MG * V);
              a[2 * g] = av.x; a[2 * g + 1] = av.y;
            }
#pragma unroll
            for (int g = 0; g < TN / 2; ++g) {
              double2 bv = *reinterpret_cast<const double2*>(Wk + (size_t)kk * mv.npad + 2 * g);
              b[2 * g] = bv.x; b[2 * g + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < kTM; ++i)
#pragma unroll
              for (int j = 0; j < TN; ++j) acc[i][j] = ikr_fma(a[i], b[j], acc[i][j]);
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&sm.empty[stage]);
      if (tid == 0 && q >= 1) {
        // refill the slot of the previous chunk once every worker warp has released it
        const unsigned qp = q - 1;
        mbar_wait(&sm.empty[qp % kStages], (qp / kStages) & 1);
        mlp_issue_chunk<W>(mv, sm, qp + kStages);
        pp.issued = qp + kStages + 1;
      }
    }
    pp.q = q + 1;
  }
  if (sizeof(W) == 4) {
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
      for (int pj = 0; pj < NP; ++pj) {
        float lo, hi;
        f2_unpack(c2[i][pj], lo, hi);
        acc[i][2 * pj] = (W)lo; acc[i][2 * pj + 1] = (W)hi;
      }
  }
}

// Hidden layers of the tile MLP: activations in sm.Hs (feature-major) are replaced layer by
// layer; weights arrive through the full/empty mbarrier ring (no CTA-wide barrier per chunk:
// warps drift up to kStages-1 chunks apart; thread 0 refills a ring slot one chunk late so that
// it rarely waits for the slowest warp).  Two CTA barriers per layer remain (all reads of the
// input activations before the in-place overwrite, all writes before the next layer reads).
template <typename W, int TN>
__device__ __forceinline__ void mlp_tile_hidden(const MlpView& mv, const MlpSmem<W>& sm,
                                                MlpPipe& pp, int M, int MG, const TileCoord& tc) {
  constexpr int V = MlpTileCfg<W>::V;
  const int tid = threadIdx.x;
  const W slope = (W)mv.slope;
  const W* P = (const W*)mv.base;
  const int gm = tc.gm, gn = tc.gn;
  const bool worker = tc.worker;
  const bool warp_works = __any_sync(0xffffffffu, worker);

  for (int layer = 0; layer < mv.L; ++layer) {
    W acc[kTM][TN];
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = (W)0;

    mlp_layer_kloop<W, TN>(mv, sm, pp, M, MG, tc, warp_works, acc);
    __syncthreads();  // all reads of this layer's input activations are done

    if (worker) {
      const W* bh = sp_bh<W>(mv, sm, layer) + gn * TN;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        W bb = bh[j];
        W* dst = sm.Hs + (size_t)(gn * TN + j) * M + gm * V;
        if (sizeof(W) == 4) {
          float4 o0, o1;
          o0.x = leaky(acc[0][j] + bb, slope); o0.y = leaky(acc[1][j] + bb, slope);
          o0.z = leaky(acc[2][j] + bb, slope); o0.w = leaky(acc[3][j] + bb, slope);
          o1.x = leaky(acc[4][j] + bb, slope); o1.y = leaky(acc[5][j] + bb, slope);
          o1.z = leaky(acc[6][j] + bb, slope); o1.w = leaky(acc[7][j] + bb, slope);
          *reinterpret_cast<float4*>(dst) = o0;
          *reinterpret_cast<float4*>(dst + MG * V) = o1;
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            double2 o;
            o.x = leaky(acc[2 * g][j] + bb, slope);
            o.y = leaky(acc[2 * g + 1][j] + bb, slope);
            *reinterpret_cast<double2*>(dst + g * MG * V) = o;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Evaluate the MLP for the M trajectories of the CTA tile.  Inputs in sm.xin, result (the
// network output before /netscale) returned to the owner threads tid < M.  Executed by every
// thread of the CTA (contains barriers).
template <typename W, int TN>
__device__ __forceinline__ W mlp_tile_forward(const MlpView& mv, const MlpSmem<W>& sm,
                                              MlpPipe& pp, int M, int MG, int NG) {
  constexpr int V = MlpTileCfg<W>::V;
  const int tid = threadIdx.x;
  const TileCoord tc = tile_coord(tid, MG, NG);
  const W slope = (W)mv.slope;
  const W* P = (const W*)mv.base;

  __syncthreads();  // xin written by the owners; Hs free (previous output layer finished)

  // ---- layer 0: Linear(2, n) + LeakyReLU ----------------------------------------------------
  if (tc.worker) {
    const W* w0 = sp_w0<W>(mv, sm);
    W nv[kTM], aa[kTM];
#pragma unroll
    for (int i = 0; i < kTM; ++i) {
      int m = tile_row<V>(i, tc.gm, MG);
      nv[i] = sm.xin[m];
      aa[i] = sm.xin[M + m];
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = tc.gn * TN + j;
      W wa = w0[col], wb = w0[mv.npad + col], bb = w0[2 * mv.npad + col];
#pragma unroll
      for (int i = 0; i < kTM; ++i) {
        W z = ikr_fma(wb, aa[i], ikr_fma(wa, nv[i], bb));
        sm.Hs[(size_t)col * M + tile_row<V>(i, tc.gm, MG)] = leaky(z, slope);
      }
    }
  }
  __syncthreads();

  mlp_tile_hidden<W, TN>(mv, sm, pp, M, MG, tc);

  // ---- output layer: Linear(n, 1), one owner thread per trajectory ---------------------------
  W out = (W)0;
  if (tid < M) {
    const W* wl = sp_wl<W>(mv, sm);
    W s0 = (W)0, s1 = (W)0, s2 = (W)0, s3 = (W)0;
    int k = 0;
    for (; k + 3 < mv.n; k += 4) {
      s0 = ikr_fma(sm.Hs[(size_t)(k + 0) * M + tid], wl[k + 0], s0);
      s1 = ikr_fma(sm.Hs[(size_t)(k + 1) * M + tid], wl[k + 1], s1);
      s2 = ikr_fma(sm.Hs[(size_t)(k + 2) * M + tid], wl[k + 2], s2);
      s3 = ikr_fma(sm.Hs[(size_t)(k + 3) * M + tid], wl[k + 3], s3);
    }
    for (; k < mv.n; ++k) s0 = ikr_fma(sm.Hs[(size_t)k * M + tid], wl[k], s0);
    out = ((s0 + s1) + (s2 + s3)) + wl[mv.npad];
  }
  return out;
}

}  // namespace ikr
#endif  // IKR_DEVICE_CUH_
