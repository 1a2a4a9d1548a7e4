// ikr_backward.cuh -- backward sweep through the solver (discrete adjoint of the accepted-step
// sequence == PyTorch autograd through torchdiffeq's non-adjoint odeint).
//
// Three kernels, run in ROUNDS of `steps_per_round` reversed steps (the round size is set by the
// stash the caller's workspace can hold):
//
//   ikr_adjoint_kernel<S,W>  persistent CTAs, one tile of M trajectories at a time.  For every
//       reversed step (last to first) and stage (5..0) the owner thread of a lane re-forms the
//       stage state from the forward checkpoint (t0, dt, y0, k_0..k_6), then the CTA runs the
//       tile MLP forward (keeping LeakyReLU sign masks in shared memory) and backward
//       (data-gradient GEMMs with the row-major weight operand).  The activations H_{l-1} and
//       the pre-activation gradients dz_l every weight gradient needs are written to a global
//       STASH ([slot][layer][lane][npad]); the first/last-layer gradients (rank-2 / rank-1) are
//       accumulated in shared memory.  Per-lane adjoint state survives between rounds in global
//       memory, so any CTA can continue any tile.
//   ikr_wgrad_kernel<W>      dW_l += dz_l^T H_{l-1} over the stash: a split-K FFMA GEMM with the
//       whole K range of a CTA accumulated in registers (8 x 8 tile per thread), operands
//       streamed by cp.async.bulk through a full/empty mbarrier ring; partial sums are kept per
//       (layer, split) in fp64.
//   ikr_grad_reduce_kernel   sums the partials into the flat state_dict-ordered gradient.
#ifndef IKR_BACKWARD_CUH_
#define IKR_BACKWARD_CUH_

#include "ikr_forward.cuh"

namespace ikr {

template <typename S>
struct BLaneSave {
  S lya, lyr, lfa, lfr, gsum;
  int n_left, out_idx, phase;
};

struct BwdParams {
  MlpView mlp;  // bwd_seq = 1
  SolverCfg cfg;
  int M, MG, NG, n_worker_warps;
  long long B;
  int T;
  long long n_tiles;
  const void* y0;
  const double* t_out;
  const int* stats;
  const double* ckpt_t;
  const void* ckpt_y;
  const void* grad_y;
  int fused_loss;
  const void* y_out;
  const double* v_out;
  const void* g;
  const void* e_rev;
  double e_scalar;
  const void* data;
  long long data_B;
  void* lane_state;
  int first_round;
  int steps_per_round;
  void* stash_h;
  void* stash_d;
  unsigned long long* counters;  // [0] tile queue, [1] stash slots used
  double* small_grad;            // [grid][small_stride]
  int small_stride;
  void* grad_y0;
  void* grad_g;
  int method;        // 0 dopri5, 1 rk4 (fixed grid, four stages per step)
  int time_f32;      // rk4: grid arithmetic in fp32
  int rk4_perturb;   // rk4 `perturb` option
};

template <typename S, typename W>
struct AdjSmemLayout {
  size_t off_bar, off_misc, off_lanes, off_xin, off_up, off_sg, off_sp, off_mask, off_hs, off_wr, total;
  __host__ __device__ AdjSmemLayout(int M, int npad, int kc, int L) {
    size_t o = 0;
    off_bar = o; o += 64;
    off_misc = o; o += 32;
    off_lanes = o; o += (size_t)M * sizeof(BLane<S>); o = (o + 15) & ~(size_t)15;
    off_xin = o; o += (size_t)2 * M * sizeof(W); o = (o + 15) & ~(size_t)15;
    off_up = o; o += (size_t)M * sizeof(W); o = (o + 15) & ~(size_t)15;
    off_sg = o; o += (size_t)(4 * npad + 8) * sizeof(double);
    off_sp = o; o += mlp_small_elems(L, npad) * sizeof(W); o = (o + 15) & ~(size_t)15;
    off_mask = o; o += (size_t)L * npad * (M / 8); o = (o + 127) & ~(size_t)127;
    off_hs = o; o += (size_t)npad * M * sizeof(W); o = (o + 127) & ~(size_t)127;
    off_wr = o; o += ((size_t)kStages * kc + 1) * npad * sizeof(W);   // +1 row: prefetch pad
    total = o;
  }
};

template <typename W>
struct AdjSmem {
  MlpSmem<W> mlp;
  W* up;                 // [M] upstream gradient of the MLP output per lane
  unsigned char* mask;   // [L][npad][MG] sign bits of H_0 .. H_{L-1} (8 lanes per byte)
  double* sg;            // [4 npad + 8] gradient accumulators: w0[:,0] | w0[:,1] | b0 | w_last | b_last
};

// ---- register tile <-> memory ------------------------------------------------------------------
template <typename W, int TN>
__device__ __forceinline__ void tile_store_smem(W* Hs, int M, int MG, const TileCoord& tc,
                                                const W (&v)[kTM][TN]) {
  constexpr int V = MlpTileCfg<W>::V;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    W* dst = Hs + (size_t)(tc.gn * TN + j) * M + tc.gm * V;
    if (sizeof(W) == 4) {
      float4 o0, o1;
      o0.x = v[0][j]; o0.y = v[1][j]; o0.z = v[2][j]; o0.w = v[3][j];
      o1.x = v[4][j]; o1.y = v[5][j]; o1.z = v[6][j]; o1.w = v[7][j];
      *reinterpret_cast<float4*>(dst) = o0;
      *reinterpret_cast<float4*>(dst + MG * V) = o1;
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        double2 o;
        o.x = v[2 * g][j]; o.y = v[2 * g + 1][j];
        *reinterpret_cast<double2*>(dst + g * MG * V) = o;
      }
    }
  }
}

// stash block layout [M][npad] (lane-major, so that the weight-gradient GEMM streams K-major rows)
template <typename W, int TN>
__device__ __forceinline__ void tile_store_stash(W* dst, int npad, int MG, const TileCoord& tc,
                                                 const W (&v)[kTM][TN]) {
  constexpr int V = MlpTileCfg<W>::V;
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    W* p = dst + (size_t)tile_row<V>(i, tc.gm, MG) * npad + tc.gn * TN;
#pragma unroll
    for (int g = 0; g < TN / V; ++g) {
      if (sizeof(W) == 4) {
        float4 o;
        o.x = v[i][4 * g]; o.y = v[i][4 * g + 1]; o.z = v[i][4 * g + 2]; o.w = v[i][4 * g + 3];
        *reinterpret_cast<float4*>(p + 4 * g) = o;
      } else {
        double2 o;
        o.x = v[i][2 * g]; o.y = v[i][2 * g + 1];
        *reinterpret_cast<double2*>(p + 2 * g) = o;
      }
    }
  }
}

template <typename W, int TN>
__device__ __forceinline__ void tile_store_mask(unsigned char* mask_l, int MG, const TileCoord& tc,
                                                const W (&v)[kTM][TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    unsigned bits = 0;
#pragma unroll
    for (int i = 0; i < kTM; ++i) bits |= (v[i][j] > (W)0 ? 1u : 0u) << i;
    mask_l[(size_t)(tc.gn * TN + j) * MG + tc.gm] = (unsigned char)bits;
  }
}

// One n x n layer: K-loop over the ring, then `epi(acc)` by the worker threads between the two
// CTA barriers (reads of the input activations done / writes of the outputs done).
template <typename W, int TN, typename Epi>
__device__ __forceinline__ void mlp_layer(const MlpView& mv, const MlpSmem<W>& sm, MlpPipe& pp,
                                          int M, int MG, const TileCoord& tc, bool warp_works,
                                          Epi epi) {
  W acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = (W)0;
  mlp_layer_kloop<W, TN>(mv, sm, pp, M, MG, tc, warp_works, acc);
  __syncthreads();
  if (tc.worker) epi(acc);
  __syncthreads();
}

// MLP forward + backward for the M lanes of the tile.  In: sm.xin (nv, a), as.up (dL/d net_out).
// Out (owner threads tid < M): up * d net_out / d a.  Side effects: stash blocks of this
// evaluation (`sh`, `sd`: [L][M][npad]) and the shared-memory gradient accumulators as.sg.
template <typename W, int TN>
__device__ __forceinline__ W mlp_tile_fwd_bwd(const MlpView& mv, const AdjSmem<W>& as, MlpPipe& pp,
                                              int M, int MG, int NG, W* stash_h, W* stash_d,
                                              const volatile long long* slot_ptr,
                                              size_t slot_elems) {
  constexpr int V = MlpTileCfg<W>::V;
  const MlpSmem<W>& sm = as.mlp;
  const int tid = threadIdx.x;
  const TileCoord tc = tile_coord(tid, MG, NG);
  const bool warp_works = __any_sync(0xffffffffu, tc.worker);
  const W slope = (W)mv.slope;
  const W* P = (const W*)mv.base;
  const int npad = mv.npad, L = mv.L;

  __syncthreads();  // xin / up / slot written; Hs free
  const long long slot = *slot_ptr;
  W* sh = stash_h + (size_t)slot * slot_elems;
  W* sd = stash_d + (size_t)slot * slot_elems;
  const size_t blk = (size_t)M * npad;

  // ---- layer 0 forward: H_0 = leaky(w0 [nv, a] + b0) -------------------------------------------
  if (tc.worker) {
    const W* w0 = sp_w0<W>(mv, sm);
    W nv[kTM], aa[kTM];
#pragma unroll
    for (int i = 0; i < kTM; ++i) {
      int m = tile_row<V>(i, tc.gm, MG);
      nv[i] = sm.xin[m];
      aa[i] = sm.xin[M + m];
    }
    W h[kTM][TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = tc.gn * TN + j;
      W wa = w0[col], wb = w0[npad + col], bb = w0[2 * npad + col];
#pragma unroll
      for (int i = 0; i < kTM; ++i) h[i][j] = leaky(ikr_fma(wb, aa[i], ikr_fma(wa, nv[i], bb)), slope);
    }
    tile_store_smem<W, TN>(sm.Hs, M, MG, tc, h);
    tile_store_mask<W, TN>(as.mask, MG, tc, h);
    tile_store_stash<W, TN>(sh, npad, MG, tc, h);
  }
  __syncthreads();

  // ---- hidden layers forward -------------------------------------------------------------------
  for (int l = 1; l <= L; ++l) {
    const W* bh = sp_bh<W>(mv, sm, l - 1) + tc.gn * TN;
    if (l < L) {
      mlp_layer<W, TN>(mv, sm, pp, M, MG, tc, warp_works, [&](W (&acc)[kTM][TN]) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          W bb = bh[j];
#pragma unroll
          for (int i = 0; i < kTM; ++i) acc[i][j] = leaky(acc[i][j] + bb, slope);
        }
        tile_store_smem<W, TN>(sm.Hs, M, MG, tc, acc);
        tile_store_mask<W, TN>(as.mask + (size_t)l * npad * MG, MG, tc, acc);
        tile_store_stash<W, TN>(sh + (size_t)l * blk, npad, MG, tc, acc);
      });
    } else {
      // last hidden layer: H_L feeds the output layer Linear(n, 1).  Its gradient is rank-1, so
      // it is consumed here: d w_last += sum_m up[m] H_L[:, m], and dz_L = w_last up leaky'(H_L)
      // replaces H_L in shared memory (and goes to the stash for dW_L).
      mlp_layer<W, TN>(mv, sm, pp, M, MG, tc, warp_works, [&](W (&acc)[kTM][TN]) {
        W upv[kTM];
#pragma unroll
        for (int i = 0; i < kTM; ++i) upv[i] = as.up[tile_row<V>(i, tc.gm, MG)];
        const W* wl = sp_wl<W>(mv, sm) + tc.gn * TN;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          W bb = bh[j], wj = wl[j];
          W s = (W)0;
#pragma unroll
          for (int i = 0; i < kTM; ++i) {
            W h = leaky(acc[i][j] + bb, slope);
            s = ikr_fma(upv[i], h, s);
            acc[i][j] = (wj * upv[i]) * (h > (W)0 ? (W)1 : slope);
          }
          atomicAdd(&as.sg[3 * npad + tc.gn * TN + j], (double)s);
        }
        tile_store_smem<W, TN>(sm.Hs, M, MG, tc, acc);
        tile_store_stash<W, TN>(sd + (size_t)(L - 1) * blk, npad, MG, tc, acc);
      });
    }
  }

  // ---- hidden layers backward: dz_{l-1} = (dz_l W_l) * leaky'(H_{l-1}) ---------------------------
  for (int l = L; l >= 1; --l) {
    mlp_layer<W, TN>(mv, sm, pp, M, MG, tc, warp_works, [&](W (&acc)[kTM][TN]) {
      const unsigned char* mk = as.mask + (size_t)(l - 1) * npad * MG;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        unsigned bits = mk[(size_t)(tc.gn * TN + j) * MG + tc.gm];
#pragma unroll
        for (int i = 0; i < kTM; ++i) acc[i][j] = acc[i][j] * (((bits >> i) & 1u) ? (W)1 : slope);
      }
      tile_store_smem<W, TN>(sm.Hs, M, MG, tc, acc);
      if (l >= 2) tile_store_stash<W, TN>(sd + (size_t)(l - 2) * blk, npad, MG, tc, acc);
    });
  }

  // ---- layer 0 backward (rank-2): Hs = dz_0 ------------------------------------------------------
  W da = (W)0;
  if (tid < M) {
    const W* w0b = sp_w0<W>(mv, sm) + npad;
    W s0 = (W)0, s1 = (W)0, s2 = (W)0, s3 = (W)0;
    int k = 0;
    for (; k + 3 < mv.n; k += 4) {
      s0 = ikr_fma(sm.Hs[(size_t)(k + 0) * M + tid], w0b[k + 0], s0);
      s1 = ikr_fma(sm.Hs[(size_t)(k + 1) * M + tid], w0b[k + 1], s1);
      s2 = ikr_fma(sm.Hs[(size_t)(k + 2) * M + tid], w0b[k + 2], s2);
      s3 = ikr_fma(sm.Hs[(size_t)(k + 3) * M + tid], w0b[k + 3], s3);
    }
    for (; k < mv.n; ++k) s0 = ikr_fma(sm.Hs[(size_t)k * M + tid], w0b[k], s0);
    da = (s0 + s1) + (s2 + s3);
    // d b_last = sum_m up[m]
    W u = as.up[tid];
    if (u != (W)0) atomicAdd(&as.sg[4 * npad], (double)u);
  }
  for (int k = tid; k < mv.n; k += blockDim.x) {
    // thread k owns feature k; lanes visited in rotated order (bank-conflict free for M % 32 == 0)
    const W* row = sm.Hs + (size_t)k * M;
    W g0 = (W)0, g1 = (W)0, g2 = (W)0;
    int m = k % M;
    for (int c = 0; c < M; ++c) {
      W dz = row[m];
      g0 = ikr_fma(dz, sm.xin[m], g0);
      g1 = ikr_fma(dz, sm.xin[M + m], g1);
      g2 += dz;
      m = (m + 1 == M) ? 0 : m + 1;
    }
    as.sg[k] += (double)g0;
    as.sg[npad + k] += (double)g1;
    as.sg[2 * npad + k] += (double)g2;
  }
  __syncthreads();  // xin / up / Hs may be rewritten by the owners for the next evaluation
  return da;
}

template <typename S, typename W, int TN>
__global__ void __launch_bounds__(512, 1) ikr_adjoint_kernel(const BwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef typename Vec2<S>::type V2;
  const int tid = threadIdx.x;
  const int M = p.M;
  const AdjSmemLayout<S, W> lay(M, p.mlp.npad, p.mlp.kc, p.mlp.L);
  BLane<S>* lanes = reinterpret_cast<BLane<S>*>(smem_raw + lay.off_lanes);
  long long* misc = reinterpret_cast<long long*>(smem_raw + lay.off_misc);  // [0] tile, [1] slot
  AdjSmem<W> as;
  as.mlp.full = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  as.mlp.empty = as.mlp.full + kStages;
  as.mlp.xin = reinterpret_cast<W*>(smem_raw + lay.off_xin);
  as.mlp.Hs = reinterpret_cast<W*>(smem_raw + lay.off_hs);
  as.mlp.Wr = reinterpret_cast<W*>(smem_raw + lay.off_wr);
  as.mlp.sp = reinterpret_cast<W*>(smem_raw + lay.off_sp);
  as.up = reinterpret_cast<W*>(smem_raw + lay.off_up);
  as.sg = reinterpret_cast<double*>(smem_raw + lay.off_sg);
  as.mask = smem_raw + lay.off_mask;

  mlp_pipe_init<W>(as.mlp, p.n_worker_warps);
  for (int k = tid; k < 4 * p.mlp.npad + 8; k += blockDim.x) as.sg[k] = 0.0;
  mlp_stage_small<W>(p.mlp, as.mlp);
  __syncthreads();
  MlpPipe pp;
  mlp_pipe_start<W>(p.mlp, as.mlp, pp);

  const SolverCfg cfg = p.cfg;
  const bool owner = tid < M;
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
  W* stash_h = reinterpret_cast<W*>(p.stash_h);
  W* stash_d = reinterpret_cast<W*>(p.stash_d);
  const size_t slot_elems = (size_t)p.mlp.L * M * p.mlp.npad;

  while (true) {
    if (tid == 0) misc[0] = (long long)atomicAdd(&p.counters[0], 1ULL);
    __syncthreads();
    const long long tile = misc[0];
    if (tile >= p.n_tiles) break;
    const long long b = tile * M + tid;
    const bool valid = owner && b < jB;
    S g_b = (S)1, e_b = (S)p.e_scalar;

    if (owner) {
      BLane<S>& L = lanes[tid];
      if (p.first_round) {
        const bool ok = valid && p.stats[4 * b + 3] == 0;
        blane_reset<S>(L, ok ? p.stats[4 * b] : 0, T, ok);
      } else {
        const BLaneSave<S> sv = saved[tile * M + tid];
        blane_reset<S>(L, 0, T, false);
        L.lya = sv.lya; L.lyr = sv.lyr; L.lfa = sv.lfa; L.lfr = sv.lfr; L.gsum = sv.gsum;
        L.n_left = sv.n_left; L.out_idx = sv.out_idx; L.phase = sv.phase;
      }
      if (valid) {
        if (gptr) g_b = gptr[b];
        if (eptr) e_b = eptr[b];
      }
    }

    // dL/dy_out[idx] of this lane: caller-provided, or derived from the fused loss on the current
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
        lanes[tid].gsum = lanes[tid].gsum + (S)(w * vm * (double)(y.x * y.y));
      }
    };
    auto evaluate = [&](bool act, double nv, double ain, double up) -> W {
      if (owner) {
        as.mlp.xin[tid] = act ? (W)nv : (W)0;
        as.mlp.xin[M + tid] = act ? (W)ain : (W)0;
        as.up[tid] = act ? (W)up : (W)0;
      }
      if (tid == 0) misc[1] = (long long)atomicAdd(&p.counters[1], 1ULL);
      return mlp_tile_fwd_bwd<W, TN>(p.mlp, as, pp, M, p.MG, p.NG, stash_h, stash_d, misc + 1,
                                 slot_elems);
    };

    for (int r = 0; r < p.steps_per_round; ++r) {
      const bool act = owner && lanes[tid].phase == 0;
      if (!__syncthreads_or(act ? 1 : 0)) break;
      if (act) {
        BLane<S>& L = lanes[tid];
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
          if (rk4) brk4_stage_inputs<S>(lanes[tid], cfg, s, p.time_f32 != 0, p.rk4_perturb != 0, &nv, &ain, &up);
          else bdp_stage_inputs<S>(lanes[tid], cfg, s, &nv, &ain, &up);
        }
        const W da = evaluate(act, nv, ain, up);
        if (act) {
          if (rk4) brk4_reverse_stage<S>(lanes[tid], s, p.time_f32 != 0, (S)da);
          else bdp_reverse_stage<S>(lanes[tid], s, (S)da);
        }
      }
      if (act) {
        if (rk4) brk4_finish_step<S>(lanes[tid]);
        else bdp_finish_step<S>(lanes[tid]);
      }
    }

    // f(t[0], y0): once no lane of the tile is still reversing steps
    {
      const int p0 = owner && lanes[tid].phase == 0 ? 1 : 0;
      const int p1 = owner && lanes[tid].phase == 1 ? 1 : 0;
      const int any0 = __syncthreads_or(p0);
      const int any1 = __syncthreads_or(p1);
      if (!any0 && any1 && rk4) {
        // fixed grid: there is no separate f(t[0], y0) evaluation, only the first output sample
        if (p1) {
          S g0a, g0r;
          grad(0, &g0a, &g0r);
          brk4_finish<S>(lanes[tid], g0a, g0r);
          if (p.grad_y0) {
            V2 v;
            v.x = lanes[tid].lya; v.y = lanes[tid].lyr;
            reinterpret_cast<V2*>(p.grad_y0)[b] = v;
          }
          if (p.grad_g) reinterpret_cast<S*>(p.grad_g)[b] = lanes[tid].gsum;
        }
      } else if (!any0 && any1) {
        double nv = 0, ain = 0, up = 0;
        S y0a = (S)0;
        if (p1) {
          y0a = y0[2 * b];
          bdp_f0_inputs<S>(lanes[tid], cfg, p.t_out[0], y0a, &nv, &ain, &up);
        }
        const W da = evaluate(p1 != 0, nv, ain, up);
        if (p1) {
          S g0a, g0r;
          grad(0, &g0a, &g0r);
          bdp_f0_finish<S>(lanes[tid], (S)da, g0a, g0r);
          if (p.grad_y0) {
            V2 v;
            v.x = lanes[tid].lya; v.y = lanes[tid].lyr;
            reinterpret_cast<V2*>(p.grad_y0)[b] = v;
          }
          if (p.grad_g) reinterpret_cast<S*>(p.grad_g)[b] = lanes[tid].gsum;
        }
      }
    }
    if (owner) {
      const BLane<S>& L = lanes[tid];
      BLaneSave<S> sv;
      sv.lya = L.lya; sv.lyr = L.lyr; sv.lfa = L.lfa; sv.lfr = L.lfr; sv.gsum = L.gsum;
      sv.n_left = L.n_left; sv.out_idx = L.out_idx; sv.phase = L.phase;
      saved[tile * M + tid] = sv;
    }
    __syncthreads();
  }
  mlp_pipe_drain<W>(p.mlp, as.mlp, pp);
  __syncthreads();
  double* sgo = p.small_grad + (size_t)blockIdx.x * p.small_stride;
  for (int k = tid; k < 4 * p.mlp.npad + 1; k += blockDim.x) sgo[k] += as.sg[k];
}

// =============================================================================================
// Weight-gradient GEMM over the stash
// =============================================================================================
struct WgradParams {
  int L, n, npad, M, KC;
  int n_ot, n_it, BO, BI, S;
  int IG;                 // thread columns (i groups) per CTA
  int stages;
  const void* stash_h;
  const void* stash_d;
  const unsigned long long* counters;
  double* partial_w;      // [L][S][npad][npad]
  double* partial_b;      // [L][S][npad]
};

template <typename W>
struct WgradCfg;
template <>
struct WgradCfg<float> { static constexpr int RI = 8, V = 4; };
template <>
struct WgradCfg<double> { static constexpr int RI = 4, V = 2; };
constexpr int kWgRO = 8;
constexpr int kWgMaxThreads = 384;

template <typename W>
__global__ void __launch_bounds__(kWgMaxThreads, 1) ikr_wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int RI = WgradCfg<W>::RI, V = WgradCfg<W>::V, RO = kWgRO;
  const int tid = threadIdx.x;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 8;
  W* ring = reinterpret_cast<W*>(smem_raw + 128);
  const int npad = p.npad;
  const size_t op_elems = (size_t)p.KC * npad;     // one operand chunk
  const size_t stage_elems = 2 * op_elems;

  int idx = blockIdx.x;
  const int split = idx % p.S; idx /= p.S;
  const int it = idx % p.n_it; idx /= p.n_it;
  const int ot = idx % p.n_ot;
  const int l = idx / p.n_ot;      // 0-based: hidden layer l+1 (dW_{l+1} = dz_{l+1}^T H_l)
  const int o0 = ot * p.BO, i0 = it * p.BI;
  const int BOt = min(p.BO, npad - o0), BIt = min(p.BI, npad - i0);
  const int half = BIt / 2;
  const int ig = tid % p.IG, og = tid / p.IG;
  const bool worker = (og * RO < BOt) && (ig * RI < BIt);
  const int n_warps = (blockDim.x + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], n_warps); }
    mbar_fence_init();
  }
  __syncthreads();

  const long long slots = (long long)p.counters[1];
  const int cps = p.M / p.KC;                         // chunks per slot
  const long long n_chunks = slots * cps;
  // this CTA's chunks: split, split + S, ...
  const long long my_chunks = n_chunks > split ? (n_chunks - split + p.S - 1) / p.S : 0;
  const W* SH = reinterpret_cast<const W*>(p.stash_h);
  const W* SD = reinterpret_cast<const W*>(p.stash_d);
  const unsigned bytes = (unsigned)(op_elems * sizeof(W));
  auto issue = [&](long long j) {
    const long long c = split + j * p.S;
    const long long slot = c / cps;
    const int sub = (int)(c - slot * cps);
    const size_t off = (((size_t)slot * p.L + l) * p.M + (size_t)sub * p.KC) * npad;
    const int st = (int)(j % p.stages);
    W* dst = ring + (size_t)st * stage_elems;
    mbar_expect_tx(&full[st], 2 * bytes);
    bulk_g2s(dst, SD + off, bytes, &full[st]);
    bulk_g2s(dst + op_elems, SH + off, bytes, &full[st]);
  };
  if (tid == 0) {
    for (int j = 0; j < p.stages && j < my_chunks; ++j) issue(j);
  }

  W acc[RO][RI];
  W bacc[RO];
#pragma unroll
  for (int r = 0; r < RO; ++r) {
    bacc[r] = (W)0;
#pragma unroll
    for (int c = 0; c < RI; ++c) acc[r][c] = (W)0;
  }

  f32x2 c2[RO][4];
  if (sizeof(W) == 4) {
#pragma unroll
    for (int r = 0; r < RO; ++r)
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) c2[r][pc] = f2_pack(0.f, 0.f);
  }
  for (long long j = 0; j < my_chunks; ++j) {
    const int st = (int)(j % p.stages);
    const unsigned par = (unsigned)((j / p.stages) & 1);
    mbar_wait(&full[st], par);
    if (worker) {
      const W* Dp = ring + (size_t)st * stage_elems + o0 + og * RO;
      const W* Hp = ring + (size_t)st * stage_elems + op_elems + i0 + ig * V;
      if (sizeof(W) == 4) {
        // software-pipelined fragment loads (the ring has one pad row behind the last stage)
        const float* dp = reinterpret_cast<const float*>(Dp);
        const float* hp = reinterpret_cast<const float*>(Hp);
        float4 d0 = *reinterpret_cast<const float4*>(dp);
        float4 d1 = *reinterpret_cast<const float4*>(dp + 4);
        float4 h0 = *reinterpret_cast<const float4*>(hp);
        float4 h1 = *reinterpret_cast<const float4*>(hp + half);
#pragma unroll 2
        for (int m = 0; m < p.KC; ++m) {
          dp += npad;
          hp += npad;
          const float4 nd0 = *reinterpret_cast<const float4*>(dp);
          const float4 nd1 = *reinterpret_cast<const float4*>(dp + 4);
          const float4 nh0 = *reinterpret_cast<const float4*>(hp);
          const float4 nh1 = *reinterpret_cast<const float4*>(hp + half);
          const float d[RO] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          const f32x2 h[4] = {f2_pack(h0.x, h0.y), f2_pack(h0.z, h0.w), f2_pack(h1.x, h1.y),
                              f2_pack(h1.z, h1.w)};
#pragma unroll
          for (int r = 0; r < RO; ++r) {
            const f32x2 dd = f2_pack(d[r], d[r]);
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) f2_fma(c2[r][pc], dd, h[pc]);
            bacc[r] += (W)d[r];
          }
          d0 = nd0; d1 = nd1; h0 = nh0; h1 = nh1;
        }
      } else {
#pragma unroll 2
        for (int m = 0; m < p.KC; ++m) {
          W d[RO], h[RI];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const double2 dv = *reinterpret_cast<const double2*>(Dp + (size_t)m * npad + 2 * g);
            d[2 * g] = dv.x; d[2 * g + 1] = dv.y;
          }
          const double2 h0 = *reinterpret_cast<const double2*>(Hp + (size_t)m * npad);
          const double2 h1 = *reinterpret_cast<const double2*>(Hp + (size_t)m * npad + half);
          h[0] = h0.x; h[1] = h0.y; h[RI - 2] = h1.x; h[RI - 1] = h1.y;
#pragma unroll
          for (int r = 0; r < RO; ++r) {
#pragma unroll
            for (int c = 0; c < RI; ++c) acc[r][c] = ikr_fma(d[r], h[c], acc[r][c]);
            bacc[r] += d[r];
          }
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[st]);
    if (tid == 0 && j >= 1 && j - 1 + p.stages < my_chunks) {
      // refill the slot of the previous chunk (one chunk late: rarely waits for the slowest warp)
      const long long jp = j - 1;
      mbar_wait(&empty[jp % p.stages], (unsigned)((jp / p.stages) & 1));
      issue(jp + p.stages);
    }
  }
  if (sizeof(W) == 4) {
#pragma unroll
    for (int r = 0; r < RO; ++r)
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        float lo, hi;
        f2_unpack(c2[r][pc], lo, hi);
        acc[r][(2 * pc) % RI] = (W)lo; acc[r][(2 * pc + 1) % RI] = (W)hi;
      }
  }

  if (worker) {
    double* pw = p.partial_w + ((size_t)l * p.S + split) * npad * npad;
#pragma unroll
    for (int r = 0; r < RO; ++r) {
      const int o = o0 + og * RO + r;
      if (o >= npad) continue;
      double* row = pw + (size_t)o * npad + i0;
#pragma unroll
      for (int c = 0; c < RI; ++c) {
        const int col = (c < RI / 2) ? ig * V + c : half + ig * V + (c - RI / 2);
        row[col] += (double)acc[r][c];
      }
      if (it == 0 && ig == 0) p.partial_b[((size_t)l * p.S + split) * npad + o] += (double)bacc[r];
    }
  }
}

struct ReduceParams {
  int L, n, npad, S, n_cta, small_stride;
  const double* partial_w;
  const double* partial_b;
  const double* small_grad;
  double* out;   // flat, state_dict order: w0 (n,2), b0, [W_l (n,n), b_l] x L, w_last (n), b_last
  long long n_params;
};

__global__ void ikr_grad_reduce_kernel(const ReduceParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_params) return;
  const long long n = p.n, per = n * n + n;
  double s = 0.0;
  auto small = [&](long long k) {
    double t = 0.0;
    for (int c = 0; c < p.n_cta; ++c) t += p.small_grad[(size_t)c * p.small_stride + k];
    return t;
  };
  if (i < 2 * n) {
    s = small((i & 1) * p.npad + (i >> 1));                  // w0[o][0|1]
  } else if (i < 3 * n) {
    s = small(2 * p.npad + (i - 2 * n));                     // b0
  } else if (i < 3 * n + p.L * per) {
    const long long r = i - 3 * n;
    const long long l = r / per, q = r - l * per;
    if (q < n * n) {
      const long long o = q / n, c = q - o * n;
      for (int sp = 0; sp < p.S; ++sp)
        s += p.partial_w[(((size_t)l * p.S + sp) * p.npad + o) * p.npad + c];
    } else {
      const long long o = q - n * n;
      for (int sp = 0; sp < p.S; ++sp) s += p.partial_b[((size_t)l * p.S + sp) * p.npad + o];
    }
  } else if (i < 3 * n + p.L * per + n) {
    s = small(3 * p.npad + (i - 3 * n - p.L * per));         // w_last
  } else {
    s = small(4 * p.npad);                                    // b_last
  }
  p.out[i] = s;
}

}  // namespace ikr
#endif  // IKR_BACKWARD_CUH_
