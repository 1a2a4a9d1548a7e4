// ikr_regress_tc.cuh -- the MLP regression stage of the reference's training scripts on the tensor
// cores: one full-batch iteration of
//     p = net(x_av) / netscale;  loss = MSELoss(reduction='sum')(p, y_dadt);  loss.backward()
// (train-s1.py:891-909, train-r1.py:917-925) over N = 69k..214k (V / vrange, a) points.
//
// The samples take the place of the trajectories of the ODE kernels: a tile is 128 samples (one per
// TMEM lane).  Per tile the CTA runs the forward MLP (tc_mlp_eval) to get the prediction, forms
// d loss / d net = 2 (p - y) / netscale, then the forward + backward evaluation of the adjoint kernel
// (tc_adj_eval) with that upstream gradient; the parameter gradients come out of the same stash and
// weight-gradient GEMM as the ODE backward (ikr_wgrad_tc_kernel, ikr_grad_reduce_tc_kernel).  The
// second forward is recomputation (4 L instead of 3 L layer MMAs per tile) that keeps the two
// evaluation routines untouched.
#ifndef IKR_REGRESS_TC_CUH_
#define IKR_REGRESS_TC_CUH_

#include "ikr_backward_tc.cuh"

namespace ikr {

struct TcRegParams {
  MlpView mlp;
  TcGeom g;
  TcStashGeom sg;
  const void* img;          // 3 L layer images: W_1..W_L, W_1..W_L, W_L^T..W_1^T
  unsigned char* stash;     // one slot per tile of this launch
  const float* x;           // [N][2] (V / vrange, a)
  const float* y;           // [N] target da/dt
  long long N;
  long long tile_begin, tile_end;   // tiles of this launch (slot = tile - tile_begin)
  float netscale;
  double* loss_out;         // [1] accumulated sum of squared errors
  int mask_words;
};

template <int G>
__global__ void __launch_bounds__(tc_threads(G), 1) ikr_regress_tc_kernel(const TcRegParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const TcGeom g = tp.g;
  const TcStashGeom sg = tp.sg;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kLaneThreads = 128 * G;
  constexpr int kMmaWarp = 4 * G, kLoadWarp = 4 * G + 1;
  const TcAdjSmemLayout<float> lay(g, g.stages, G, tp.mask_words);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  volatile long long* stash_slot = reinterpret_cast<volatile long long*>(smem_raw + lay.off_misc + 16);
  volatile int* cmd_exit = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 24);
  float* sp = reinterpret_cast<float*>(smem_raw + lay.off_sp);
  const TcEngineCtx eng = tc_engine_ctx(bars, stop_flag, smem_raw + lay.off_ring);

  if (tid == 0) {
    tc_engine_init(eng, g.stages);
    *cmd_exit = 0;
  }
  if (warp == kMmaWarp) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  {
    const float* P = reinterpret_cast<const float*>(tp.mlp.base);
    const int NP = g.NP, npad = tp.mlp.npad, n = g.n;
    for (int i = tid; i < g.small_elems; i += tc_threads(G)) {
      const int row = i / NP, c = i - row * NP;
      float v = 0.0f;
      if (row < 3) { if (c < n) v = P[tp.mlp.off_w0 + (long long)row * npad + c]; }
      else if (row < 3 + g.L) { if (c < n) v = P[tp.mlp.off_bh + (long long)(row - 3) * npad + c]; }
      else if (row == 3 + g.L) { if (c < n) v = P[tp.mlp.off_wl + c]; }
      else if (i == (4 + g.L) * NP) v = P[tp.mlp.off_wl + npad];
      sp[i] = v;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kMmaWarp) {
    tc_mma_warp(g, eng, tbase, false);
  } else if (warp == kLoadWarp) {
    if ((tid & 31) == 0)
      tc_producer_thread(g, eng, reinterpret_cast<const unsigned char*>(tp.img), (unsigned)(3 * g.L * g.KST));
  } else {
    TcLane tl;
    tl.group = warp >> 2;
    tl.lane = tid & 127;
    tl.taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    tc_lane_attach(tl, eng);
    tl.sp = sp;
    tl.xin = reinterpret_cast<float*>(smem_raw + lay.off_xin);
    tl.part = reinterpret_cast<float*>(smem_raw + lay.off_part);
    tl.slope = (float)tp.mlp.slope;
    tl.c_l0 = tl.c_wait = tl.c_epi = 0;
    tl.c_sync_a = tl.c_sync_b = tl.c_last = 0;
#ifdef IKR_TC_TRACE
    tl.trace_eval = 0;
#endif
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
        tc_mlp_eval<G>(g, tl);                 // prediction
        lanes_sync<G>();
        lanes_sync<G>();                       // upstream gradients and the slot are published
        al.slot = tp.stash + (size_t)(*stash_slot) * sg.slot;
        tc_adj_eval<G>(g, sg, tl, al);         // forward again + backward, stash
        lanes_sync<G>();
      }
    } else {
      double sse = 0.0;
      for (long long tile = tp.tile_begin + blockIdx.x; tile < tp.tile_end; tile += gridDim.x) {
        const long long i = tile * kTcM + tid;
        const bool valid = i < tp.N;
        float nv = 0.0f, a = 0.0f, yv = 0.0f;
        if (valid) {
          const float2 xv = reinterpret_cast<const float2*>(tp.x)[i];
          nv = xv.x; a = xv.y; yv = tp.y[i];
        }
        const float out = tc_owner_eval<G>(g, tl, nv, a);
        const float p = out / tp.netscale;                  // train-s1.py:903, fp32
        const float diff = p - yv;
        float up = 0.0f;
        if (valid) {
          sse += (double)diff * (double)diff;
          up = (2.0f * diff) / tp.netscale;
        }
        *reinterpret_cast<float4*>(tl.xin + 4 * tid) = make_float4(nv, a, up, 0.0f);
        if (tid == 0) *stash_slot = tile - tp.tile_begin;
        if (G > 1) lanes_sync<G>(); else owners_sync();
        al.slot = tp.stash + (size_t)(*stash_slot) * sg.slot;
        tc_adj_eval<G>(g, sg, tl, al);
        if (G > 1) lanes_sync<G>(); else owners_sync();
      }
      // block sum of the squared errors (fp64), one atomic per warp
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
      if ((tid & 31) == 0 && sse != 0.0) atomicAdd(tp.loss_out, sse);
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
#endif  // IKR_REGRESS_TC_CUH_
