// ikr_markov.cuh -- the 6-state Markov ground-truth model of the synthetic-data studies
// (train-d1.py:134-187 `Lambda`: states [c1, c2, i, ic1, ic2, o], twelve rate parameters) and its
// data production step (train-d1.py:539-569): integrate with dopri5, observe
// I(t_j) = g o(t_j) (V(t_j) - E) and add N(0, sigma^2) noise.  SURVEY.md 8f-2.
//
// One thread integrates one trajectory (own parameters, own initial state, own adaptive step size,
// own noise stream); the dopri5 / rk4 arithmetic is the same per-component code the MLP kernels use
// (dp_stage_state, dp_error, dp_dense_fit, dp_next_dt of ikr_math.h), generalised from 2 to 6 state
// components, with torchdiffeq's dtype rules: fp64 time, state-dtype stages, the RHS in fp64 inside
// the protocol table and in fp32 in the V = -80 fallback (train-d1.py:163-166).
#ifndef IKR_MARKOV_CUH_
#define IKR_MARKOV_CUH_

#include <curand_kernel.h>

#include "ikr_forward.cuh"

namespace ikr {

constexpr int kMkN = 6;

struct MarkovKernelParams {
  SolverCfg cfg;             // tab, rtol, atol, first_step, controller, max_num_steps
  double p[12];              // default parameters (p1..p12)
  const double* params;      // [B][12] per-trajectory parameters (nullable)
  long long B;
  int T, G;
  const void* y0;            // [B][6] state dtype
  const double* t_out;
  const double* grid;        // rk4
  const double* v_out;       // [T] V(t_out) (nullable unless i_out)
  void* y_out;               // [T][B][6] state dtype (nullable)
  double* i_out;             // [T][B] fp64 current (+ noise) (nullable)
  const void* g;             // [B] state dtype (nullable => 1)
  double e_rev;
  double noise_sigma;        // 0: no noise
  unsigned long long seed;
  int* stats_out;            // [B][4]
  int method, time_f32, rk4_perturb;
};

// dy/dt of the Markov model; returns true when the result is an fp64 quantity (in-table branch)
template <typename SS>
__device__ __forceinline__ bool markov_rhs(const double* p, const ProtocolTable& tab, double t, const SS* y,
                                           double* f) {
  double v;
  const bool in_table = table_voltage(tab, t, &v);
  if (in_table || sizeof(SS) == 8) {
    double a1, b1, bh, ah, a2, b2;
    if (in_table) {
      a1 = p[0] * exp(p[1] * v);  b1 = p[2] * exp(-p[3] * v);
      bh = p[4] * exp(p[5] * v);  ah = p[6] * exp(-p[7] * v);
      a2 = p[8] * exp(p[9] * v);  b2 = p[10] * exp(-p[11] * v);
    } else {
      const float vf = (float)v;
      a1 = (double)((float)p[0] * expf((float)p[1] * vf));  b1 = (double)((float)p[2] * expf(-(float)p[3] * vf));
      bh = (double)((float)p[4] * expf((float)p[5] * vf));  ah = (double)((float)p[6] * expf(-(float)p[7] * vf));
      a2 = (double)((float)p[8] * expf((float)p[9] * vf));  b2 = (double)((float)p[10] * expf(-(float)p[11] * vf));
    }
    const double c1 = (double)y[0], c2 = (double)y[1], i = (double)y[2], ic1 = (double)y[3],
                 ic2 = (double)y[4], o = (double)y[5];
    f[0] = a1 * c2 + ah * ic1 + b2 * o - (b1 + bh + a2) * c1;
    f[1] = b1 * c1 + ah * ic2 - (a1 + bh) * c2;
    f[2] = a2 * ic1 + bh * o - (b2 + ah) * i;
    f[3] = a1 * ic2 + bh * c1 + b2 * i - (b1 + ah + a2) * ic1;
    f[4] = b1 * ic1 + bh * c2 - (ah + a1) * ic2;
    f[5] = a2 * c1 + ah * i - (b2 + bh) * o;
    return true;
  }
  // fp32 state outside the table: everything in fp32
  const float vf = (float)v;
  const float a1 = (float)p[0] * expf((float)p[1] * vf), b1 = (float)p[2] * expf(-(float)p[3] * vf);
  const float bh = (float)p[4] * expf((float)p[5] * vf), ah = (float)p[6] * expf(-(float)p[7] * vf);
  const float a2 = (float)p[8] * expf((float)p[9] * vf), b2 = (float)p[10] * expf(-(float)p[11] * vf);
  const float c1 = (float)y[0], c2 = (float)y[1], i = (float)y[2], ic1 = (float)y[3], ic2 = (float)y[4],
              o = (float)y[5];
  f[0] = (double)(a1 * c2 + ah * ic1 + b2 * o - (b1 + bh + a2) * c1);
  f[1] = (double)(b1 * c1 + ah * ic2 - (a1 + bh) * c2);
  f[2] = (double)(a2 * ic1 + bh * o - (b2 + ah) * i);
  f[3] = (double)(a1 * ic2 + bh * c1 + b2 * i - (b1 + ah + a2) * ic1);
  f[4] = (double)(b1 * ic1 + bh * c2 - (ah + a1) * ic2);
  f[5] = (double)(a2 * c1 + ah * i - (b2 + bh) * o);
  return false;
}

template <typename S>
__global__ void __launch_bounds__(128) ikr_markov_kernel(const MarkovKernelParams mp) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mp.B) return;
  const SolverCfg& c = mp.cfg;
  double p[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) p[i] = mp.params ? mp.params[12 * b + i] : mp.p[i];
  const S* y0p = reinterpret_cast<const S*>(mp.y0) + (size_t)kMkN * b;
  S* y_out = reinterpret_cast<S*>(mp.y_out);
  const S g_b = mp.g ? reinterpret_cast<const S*>(mp.g)[b] : (S)1;
  const long long jB = mp.B;
  const int T = mp.T;
  curandStatePhilox4_32_10_t rng;
  const bool noisy = mp.noise_sigma > 0.0 && mp.i_out != nullptr;
  if (noisy) curand_init(mp.seed, (unsigned long long)b, 0ULL, &rng);

  auto emit = [&](int idx, const S* y) {
    if (y_out) {
      S* dst = y_out + ((size_t)idx * jB + b) * kMkN;
#pragma unroll
      for (int i = 0; i < kMkN; ++i) dst[i] = y[i];
    }
    if (mp.i_out) {
      // train-d1.py:545: true_y[:, 0, -1] * (V(t) + 86) + N(0, sigma^2): fp32 state times fp64 voltage
      double cur = (double)(g_b * y[kMkN - 1]) * (mp.v_out[idx] - mp.e_rev);
      if (noisy) cur += mp.noise_sigma * curand_normal_double(&rng);
      mp.i_out[(size_t)idx * jB + b] = cur;
    }
  };

  S y[kMkN], ys[kMkN], k[kMkN][7];
#pragma unroll
  for (int i = 0; i < kMkN; ++i) y[i] = y0p[i];
  double t0 = mp.t_out[0], dt = 0.0;
  int out_idx = 1, n_acc = 0, n_rej = 0, n_int = 0, nfe = 0, status = LANE_OK;
  emit(0, y);
  double f[kMkN];

  if (mp.method == 0) {
    // ---- f0 and the initial step (torchdiffeq _select_initial_step, order 4) -------------------
    double f0[kMkN];
    const bool f64 = markov_rhs<S>(p, c.tab, (double)(S)t0, y, f0);
    ++nfe;
#pragma unroll
    for (int i = 0; i < kMkN; ++i) k[i][0] = (S)f0[i];
    if (c.first_step > 0) {
      dt = c.first_step;
    } else {
      S scale[kMkN];
      S d0s = (S)0;
      double d1 = 0.0;
#pragma unroll
      for (int i = 0; i < kMkN; ++i) {
        scale[i] = (S)c.atol + ikr_abs(y[i]) * (S)c.rtol;
        const S q = y[i] / scale[i];
        d0s = d0s + q * q;
      }
      d0s = ikr_sqrt(d0s / (S)kMkN);
      double h0;
      if (f64) {
#pragma unroll
        for (int i = 0; i < kMkN; ++i) { const double q = f0[i] / (double)scale[i]; d1 += q * q; }
        d1 = sqrt(d1 / (double)kMkN);
        h0 = ((double)d0s < 1e-5 || d1 < 1e-5) ? (double)(S)1e-6 : 0.01 * (double)d0s / d1;
      } else {
        S d1s = (S)0;
#pragma unroll
        for (int i = 0; i < kMkN; ++i) { const S q = (S)f0[i] / scale[i]; d1s = d1s + q * q; }
        d1s = ikr_sqrt(d1s / (S)kMkN);
        d1 = (double)d1s;
        h0 = ((double)d0s < 1e-5 || d1 < 1e-5) ? (double)(S)1e-6 : (double)((S)0.01 * d0s / d1s);
      }
      double f1[kMkN], d2 = 0.0;
      if (f64) {
        double y1[kMkN];
#pragma unroll
        for (int i = 0; i < kMkN; ++i) y1[i] = (double)y[i] + h0 * f0[i];
        markov_rhs<double>(p, c.tab, (double)(S)t0 + h0, y1, f1);
#pragma unroll
        for (int i = 0; i < kMkN; ++i) { const double q = (f1[i] - f0[i]) / (double)scale[i]; d2 += q * q; }
        d2 = sqrt(d2 / (double)kMkN) / h0;
      } else {
        S y1[kMkN];
#pragma unroll
        for (int i = 0; i < kMkN; ++i) y1[i] = y[i] + (S)h0 * (S)f0[i];
        markov_rhs<S>(p, c.tab, (double)((S)t0 + (S)h0), y1, f1);
        S d2s = (S)0;
#pragma unroll
        for (int i = 0; i < kMkN; ++i) { const S q = ((S)f1[i] - (S)f0[i]) / scale[i]; d2s = d2s + q * q; }
        d2 = (double)(ikr_sqrt(d2s / (S)kMkN) / (S)h0);
      }
      ++nfe;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) {
        const double a = 1e-6, bb = h0 * 1e-3;
        h1 = a > bb ? a : bb;
      } else {
        const double m = d1 > d2 ? d1 : d2;
        h1 = pow(0.01 / m, 1.0 / 5.0);
      }
      dt = 100 * h0 < h1 ? 100 * h0 : h1;
      if (!f64) dt = (double)(S)dt;
    }
    if (T <= 1) status = LANE_DONE;

    while (status == LANE_OK) {
      if (!(t0 + dt > t0)) { status = LANE_DT_UNDERFLOW; break; }
      bool finite = true;
#pragma unroll
      for (int i = 0; i < kMkN; ++i) finite = finite && isfinite((double)y[i]);
      if (!finite) { status = LANE_NONFINITE; break; }
      if ((long long)n_int >= c.max_num_steps) { status = LANE_MAX_STEPS; break; }
#pragma unroll 1
      for (int s = 0; s < 6; ++s) {
        const S ti = dp_stage_time<S>(s, t0, dt);
#pragma unroll
        for (int i = 0; i < kMkN; ++i) ys[i] = dp_stage_state<S>(s, y[i], k[i], dt);
        markov_rhs<S>(p, c.tab, (double)ti, ys, f);
#pragma unroll
        for (int i = 0; i < kMkN; ++i) k[i][s + 1] = (S)f[i];
        ++nfe;
      }
      // error ratio over the six components (rms), accept / reject, dense output, controller
      S acc = (S)0;
#pragma unroll
      for (int i = 0; i < kMkN; ++i) {
        const S err = dp_error<S>(k[i], dt);
        const S tol = (S)c.atol + (S)c.rtol * ikr_max(ikr_abs(y[i]), ikr_abs(ys[i]));
        const S q = err / tol;
        acc = acc + q * q;
      }
      const S ratio_s = ikr_sqrt(acc / (S)kMkN);
      ++n_int;
      if (ratio_s <= (S)1) {
        const double t_hi = t0 + dt;
        if (out_idx < T && mp.t_out[out_idx] <= t_hi) {
          Dense<S> q[kMkN];
#pragma unroll
          for (int i = 0; i < kMkN; ++i) q[i] = dp_dense_fit<S>(y[i], ys[i], k[i], dt);
          while (out_idx < T && mp.t_out[out_idx] <= t_hi) {
            const S x = dp_dense_x<S>(t0, t_hi, mp.t_out[out_idx]);
            S yo[kMkN];
#pragma unroll
            for (int i = 0; i < kMkN; ++i) yo[i] = dp_dense_eval<S>(q[i], x);
            emit(out_idx, yo);
            ++out_idx;
            n_int = 0;
          }
        }
#pragma unroll
        for (int i = 0; i < kMkN; ++i) { y[i] = ys[i]; k[i][0] = k[i][6]; }
        t0 = t_hi;
        ++n_acc;
      } else {
        ++n_rej;
      }
      dt = dp_next_dt(dt, (double)ratio_s, c.ctl);
      if (out_idx >= T) status = LANE_DONE;
    }
  } else {
    // ---- rk4, 3/8 rule on the fixed grid ---------------------------------------------------------
    if (T <= 1) status = LANE_DONE;
    for (int gi = 0; gi + 1 < mp.G && status == LANE_OK; ++gi) {
      const double g0 = mp.grid[gi], g1 = mp.grid[gi + 1];
      const S h = rk4_dt<S>(g0, g1, mp.time_f32 != 0);
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const double ti = rk4_stage_time<S>(s, g0, g1, mp.time_f32 != 0, mp.rk4_perturb != 0);
#pragma unroll
        for (int i = 0; i < kMkN; ++i) ys[i] = rk4_stage_state<S>(s, y[i], k[i], h);
        markov_rhs<S>(p, c.tab, ti, ys, f);
#pragma unroll
        for (int i = 0; i < kMkN; ++i) k[i][s] = (S)f[i];
        ++nfe;
      }
      S y1[kMkN];
      bool finite = true;
#pragma unroll
      for (int i = 0; i < kMkN; ++i) { y1[i] = rk4_combine<S>(y[i], k[i], h); finite = finite && isfinite((double)y1[i]); }
      while (out_idx < T && g1 >= mp.t_out[out_idx]) {
        const double tj = mp.t_out[out_idx];
        S yo[kMkN];
        if (tj == g1) {
#pragma unroll
          for (int i = 0; i < kMkN; ++i) yo[i] = y1[i];
        } else {
          const S sl = mp.time_f32 ? (S)(((float)tj - (float)g0) / ((float)g1 - (float)g0))
                                   : (S)((tj - g0) / (g1 - g0));
#pragma unroll
          for (int i = 0; i < kMkN; ++i) yo[i] = tj == g0 ? y[i] : y[i] + sl * (y1[i] - y[i]);
        }
        emit(out_idx, yo);
        ++out_idx;
      }
#pragma unroll
      for (int i = 0; i < kMkN; ++i) y[i] = y1[i];
      ++n_acc;
      if (!finite) status = LANE_NONFINITE;
      else if (out_idx >= T) status = LANE_DONE;
    }
  }
  mp.stats_out[4 * b + 0] = n_acc;
  mp.stats_out[4 * b + 1] = n_rej;
  mp.stats_out[4 * b + 2] = nfe;
  mp.stats_out[4 * b + 3] = status == LANE_DONE ? 0 : status;
}

}  // namespace ikr
#endif  // IKR_MARKOV_CUH_
