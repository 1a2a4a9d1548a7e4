// ikr_hh.cuh -- Hodgkin-Huxley candidate model (no MLP): the forward model the reference fits with
// PINTS CMA-ES (train-d0.py:321-376 `ODEFunc`, :377-439 `Model.simulate`), batched over a
// POPULATION of parameter vectors.  RHS = the NN-d right-hand side with the network term absent:
//   da/dt = k1 (1 - a) - k2 a,  dr/dt = -k3 r + k4 (1 - r),  k_i = p exp(+-p V)
// One thread integrates one candidate (its own p1..p8, its own adaptive step size); the lane state
// machine is the same ikr_math.h code the MLP kernels run between network evaluations, so dopri5 /
// rk4 semantics (torchdiffeq 0.2.x) are shared, not re-implemented.  SURVEY.md 8f-1.
#ifndef IKR_HH_CUH_
#define IKR_HH_CUH_

#include "ikr_forward.cuh"

namespace ikr {

struct HhKernelParams {
  SolverCfg cfg;
  FwdJob job;
  const double* hh_params;   // [B][8] per-candidate p1..p8 (nullable: cfg.hp for everyone)
  int method, time_f32, rk4_perturb;
};

template <typename S>
__global__ void __launch_bounds__(128) ikr_hh_kernel(const HhKernelParams p) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const FwdJob& job = p.job;
  if (b >= job.B) return;
  SolverCfg c = p.cfg;
  c.tab = job.tab;
  c.nn_d = 1;
  if (p.hh_params) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c.hp.p[i] = p.hh_params[8 * b + i];
  }
  const long long jB = job.B;
  const int T = job.T;
  const S* y0 = reinterpret_cast<const S*>(job.y0);
  S* y_out = reinterpret_cast<S*>(job.y_out);
  S* i_out = reinterpret_cast<S*>(job.i_out);
  const S* dptr = reinterpret_cast<const S*>(job.data);
  const bool observe = (job.v_out != nullptr) && (job.i_out != nullptr || job.loss_out != nullptr);
  const S g_b = job.g ? reinterpret_cast<const S*>(job.g)[b] : (S)1;
  const S e_b = job.e_rev ? reinterpret_cast<const S*>(job.e_rev)[b] : (S)job.e_scalar;
  double sse = 0.0, sae = 0.0;
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
        sse += diff * diff;
        sae += fabs(diff);
      }
    }
  };
  auto no_ckpt = [](int, const Lane<S>&) { return true; };

  Lane<S> L;
  lane_reset<S>(L, y0[2 * b], y0[2 * b + 1], job.t_out[0], true);
  emit(0, L.ya, L.yr);
  double nv, ain;
  if (p.method == 0) {
    init_prepare_f0<S>(L, c, &nv, &ain);
    init_store_f0<S>(L, c, 0.0);
    if (c.first_step > 0) {
      L.dt = c.first_step;
    } else {
      init_prepare_f1<S>(L, c, &nv, &ain);
      init_store_f1<S>(L, c, 0.0);
    }
    if (T <= 1) L.status = LANE_DONE;
    while (true) {
      dp_check_before_step<S>(L, c);
      if (!lane_active(L)) break;
#pragma unroll 1
      for (int s = 0; s < 6; ++s) {
        dp_prepare_stage<S>(L, c, s, &nv, &ain);
        dp_store_stage<S>(L, c, s, 0.0);
      }
      dp_finish_step<S>(L, c, job.t_out, T, emit, no_ckpt);
    }
  } else {
    if (T <= 1) L.status = LANE_DONE;
    for (int gi = 0; gi + 1 < job.G && lane_active(L); ++gi) {
      const double g0 = job.grid[gi], g1 = job.grid[gi + 1];
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        rk4_prepare_stage<S>(L, c, s, g0, g1, p.time_f32 != 0, p.rk4_perturb != 0, &nv, &ain);
        rk4_store_stage<S>(L, c, s, 0.0);
      }
      rk4_finish_step<S>(L, g0, g1, p.time_f32 != 0, job.t_out, T, emit);
    }
  }
  job.stats_out[4 * b + 0] = L.n_acc;
  job.stats_out[4 * b + 1] = L.n_rej;
  job.stats_out[4 * b + 2] = L.nfe;
  job.stats_out[4 * b + 3] = L.status == LANE_DONE ? 0 : L.status;
  if (job.loss_out) {
    job.loss_out[2 * b] = sse;
    job.loss_out[2 * b + 1] = sae;
  }
}

}  // namespace ikr
#endif  // IKR_HH_CUH_
