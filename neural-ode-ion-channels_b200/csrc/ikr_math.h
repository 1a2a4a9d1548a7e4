// ikr_math.h -- per-trajectory solver arithmetic shared by every kernel (host+device inline).
//
// Everything a single trajectory ("lane") does between two MLP evaluations lives here: protocol
// interpolation, the Hodgkin-Huxley rate terms, Dormand-Prince stage assembly, error ratio, step
// size controller, initial-step heuristic, dense-output fit/evaluation and the 3/8-rule RK4
// combination.  The functions are __host__ __device__ so the exact same code is exercised by the
// host-side logic test (tests/host_harness.cpp, test infrastructure only) and by the kernels.
//
// dtype semantics follow torchdiffeq 0.2.x as used by the reference:
// S = state dtype (y, k, tableau, dense coefficients), times/step sizes in double.
#ifndef IKR_MATH_H_
#define IKR_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define IKR_HD __host__ __device__ __forceinline__
#else
#define IKR_HD inline
#endif

namespace ikr {

// ------------------------------------------------------------------------------------------
// Dormand-Prince 5(4) tableau (Shampine).  Stored as double expressions; the solver casts them
// to the state dtype exactly like torchdiffeq's `tableau.to(dtype=y0.dtype)`.
// ------------------------------------------------------------------------------------------
#define IKR_DP_A1 (1.0 / 5)
#define IKR_DP_A2 (3.0 / 10)
#define IKR_DP_A3 (4.0 / 5)
#define IKR_DP_A4 (8.0 / 9)

template <typename S>
IKR_HD S dp_alpha(int i) {
  switch (i) {
    case 0: return (S)(1.0 / 5);
    case 1: return (S)(3.0 / 10);
    case 2: return (S)(4.0 / 5);
    case 3: return (S)(8.0 / 9);
    default: return (S)1.0;
  }
}

template <typename S>
IKR_HD S dp_beta(int i, int j) {
  switch (i) {
    case 0: return (S)(1.0 / 5);
    case 1: return j == 0 ? (S)(3.0 / 40) : (S)(9.0 / 40);
    case 2: return j == 0 ? (S)(44.0 / 45) : j == 1 ? (S)(-56.0 / 15) : (S)(32.0 / 9);
    case 3:
      return j == 0 ? (S)(19372.0 / 6561)
                    : j == 1 ? (S)(-25360.0 / 2187)
                             : j == 2 ? (S)(64448.0 / 6561) : (S)(-212.0 / 729);
    case 4:
      return j == 0 ? (S)(9017.0 / 3168)
                    : j == 1 ? (S)(-355.0 / 33)
                             : j == 2 ? (S)(46732.0 / 5247)
                                      : j == 3 ? (S)(49.0 / 176) : (S)(-5103.0 / 18656);
    default:
      return j == 0 ? (S)(35.0 / 384)
                    : j == 1 ? (S)0.0
                             : j == 2 ? (S)(500.0 / 1113)
                                      : j == 3 ? (S)(125.0 / 192)
                                               : j == 4 ? (S)(-2187.0 / 6784) : (S)(11.0 / 84);
  }
}

template <typename S>
IKR_HD S dp_cerr(int j) {
  switch (j) {
    case 0: return (S)(35.0 / 384 - 1951.0 / 21600);
    case 1: return (S)0.0;
    case 2: return (S)(500.0 / 1113 - 22642.0 / 50085);
    case 3: return (S)(125.0 / 192 - 451.0 / 720);
    case 4: return (S)(-2187.0 / 6784 - -12231.0 / 42400);
    case 5: return (S)(11.0 / 84 - 649.0 / 6300);
    default: return (S)(-1.0 / 60.0);
  }
}

template <typename S>
IKR_HD S dp_cmid(int j) {
  switch (j) {
    case 0: return (S)(6025192743.0 / 30085553152.0 / 2);
    case 1: return (S)0.0;
    case 2: return (S)(51252292925.0 / 65400821598.0 / 2);
    case 3: return (S)(-2691868925.0 / 45128329728.0 / 2);
    case 4: return (S)(187940372067.0 / 1594534317056.0 / 2);
    case 5: return (S)(-1776094331.0 / 19743644256.0 / 2);
    default: return (S)(11237099.0 / 235043384.0 / 2);
  }
}

// nextafter(t, t-1): torchdiffeq's Perturb.PREV in the state dtype
IKR_HD float prev_representable(float t) { return nextafterf(t, t - 1.0f); }
IKR_HD double prev_representable(double t) { return nextafter(t, t - 1.0); }
IKR_HD float next_representable(float t) { return nextafterf(t, t + 1.0f); }
IKR_HD double next_representable(double t) { return nextafter(t, t + 1.0); }

IKR_HD float ikr_abs(float x) { return fabsf(x); }
IKR_HD double ikr_abs(double x) { return fabs(x); }
IKR_HD float ikr_max(float a, float b) { return fmaxf(a, b); }
IKR_HD double ikr_max(double a, double b) { return fmax(a, b); }
IKR_HD float ikr_sqrt(float x) { return sqrtf(x); }
IKR_HD double ikr_sqrt(double x) { return sqrt(x); }

// ------------------------------------------------------------------------------------------
// Protocol table: scipy.interpolate.interp1d(kind='linear') semantics (train-s1.py:218-225).
//   idx = searchsorted(x, t, side='left') clipped to [1, n-1]; lo = idx-1; hi = idx
//   v = (y_hi - y_lo) / (x_hi - x_lo) * (t - x_lo) + y_lo
// Out-of-table -> the callers' ValueError branch: V = -80 (train-s1.py:234-237).
// `hint` = (t0, inv_dt) of a uniform grid only seeds the search; the result is always verified
// against the table so it is exact for any monotone table.
// ------------------------------------------------------------------------------------------
struct ProtocolTable {
  const double* t;
  const double* v;
  int len;
  int uniform;
  double t0;
  double inv_dt;
};

IKR_HD int table_searchsorted_left(const ProtocolTable& tab, double x) {
  // first index i with tab.t[i] >= x  (in [0, len])
  int lo = 0, hi = tab.len;
  if (tab.uniform) {
    double guess = (x - tab.t0) * tab.inv_dt;
    int g = (int)guess;
    if (g < 0) g = 0;
    if (g > tab.len - 1) g = tab.len - 1;
    // establish a small bracket [lo, hi) around the guess, verified on real table values
    int a = g - 1 < 0 ? 0 : g - 1;
    int b = g + 2 > tab.len ? tab.len : g + 2;
    if ((a == 0 || tab.t[a - 1] < x) && (b == tab.len || tab.t[b] >= x)) {
      lo = a;
      hi = b;
    }
  }
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (tab.t[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// returns true when t is inside the table; *v receives the interpolated voltage
IKR_HD bool table_voltage(const ProtocolTable& tab, double t, double* v) {
  if (t < tab.t[0] || t > tab.t[tab.len - 1]) {
    *v = -80.0;
    return false;
  }
  int idx = table_searchsorted_left(tab, t);
  if (idx < 1) idx = 1;
  if (idx > tab.len - 1) idx = tab.len - 1;
  double x_lo = tab.t[idx - 1], x_hi = tab.t[idx];
  double y_lo = tab.v[idx - 1], y_hi = tab.v[idx];
  double slope = (y_hi - y_lo) / (x_hi - x_lo);
  *v = slope * (t - x_lo) + y_lo;
  return true;
}

// ------------------------------------------------------------------------------------------
// Hodgkin-Huxley terms of the RHS.  Mirrors the dtype promotions of the reference classes:
// inside the table V is fp64 and the rate terms are fp64; `(1 - y)` is formed in the state
// dtype; in the fallback branch V is an integer tensor and the rates are fp32.
// ------------------------------------------------------------------------------------------
struct HHParams {
  double p[8];
};

template <typename S>
IKR_HD double hh_rate_pair(double pa, double pb, double pc, double pd, double v, bool in_table,
                           S y) {
  // returns  ka * (1 - y) - kb * y  with ka = pa exp(pb v), kb = pc exp(-pd v)
  S one_minus = (S)1 - y;
  if (in_table) {
    double ka = pa * exp(pb * v);
    double kb = pc * exp(-pd * v);
    return ka * (double)one_minus - kb * (double)y;
  }
  float vf = (float)v;
  float ka = (float)pa * expf((float)pb * vf);
  float kb = (float)pc * expf(-(float)pd * vf);
  if (sizeof(S) == 4) return (double)(ka * (float)one_minus - kb * (float)y);
  return (double)ka * (double)one_minus - (double)kb * (double)y;
}

// dr/dt = -k3 r + k4 (1 - r)   (train-s1.py:239-242)
template <typename S>
IKR_HD double hh_drdt(const HHParams& hp, double v, bool in_table, S r) {
  S one_minus = (S)1 - r;
  if (in_table) {
    double k3 = hp.p[4] * exp(hp.p[5] * v);
    double k4 = hp.p[6] * exp(-hp.p[7] * v);
    return -k3 * (double)r + k4 * (double)one_minus;
  }
  float vf = (float)v;
  float k3 = (float)hp.p[4] * expf((float)hp.p[5] * vf);
  float k4 = (float)hp.p[6] * expf(-(float)hp.p[7] * vf);
  if (sizeof(S) == 4) return (double)(-k3 * (float)r + k4 * (float)one_minus);
  return -(double)k3 * (double)r + (double)k4 * (double)one_minus;
}

// d(dr/dt)/dr = -(k3 + k4)   (for the adjoint)
IKR_HD double hh_drdt_dr(const HHParams& hp, double v, bool in_table) {
  if (in_table) return -(hp.p[4] * exp(hp.p[5] * v) + hp.p[6] * exp(-hp.p[7] * v));
  float vf = (float)v;
  return -((double)((float)hp.p[4] * expf((float)hp.p[5] * vf)) +
           (double)((float)hp.p[6] * expf(-(float)hp.p[7] * vf)));
}

// NN-d: HH activation part  k1 (1 - a) - k2 a   (train-d2.py:247-250)
template <typename S>
IKR_HD double hh_dadt(const HHParams& hp, double v, bool in_table, S a) {
  return hh_rate_pair<S>(hp.p[0], hp.p[1], hp.p[2], hp.p[3], v, in_table, a);
}
IKR_HD double hh_dadt_da(const HHParams& hp, double v, bool in_table) {
  if (in_table) return -(hp.p[0] * exp(hp.p[1] * v) + hp.p[2] * exp(-hp.p[3] * v));
  float vf = (float)v;
  return -((double)((float)hp.p[0] * expf((float)hp.p[1] * vf)) +
           (double)((float)hp.p[2] * expf(-(float)hp.p[3] * vf)));
}

// MLP input nv = V / vrange as the reference forms it (fp64 division in the table, fp32 in the
// fallback), before the cast to the MLP dtype.
IKR_HD double mlp_input_nv(double v, bool in_table, double vrange, bool vrange_is_f64) {
  if (in_table || vrange_is_f64) return v / vrange;
  return (double)((float)v / (float)vrange);
}

// ------------------------------------------------------------------------------------------
// dopri5 pieces
// ------------------------------------------------------------------------------------------
// stage time in the state dtype: t0 + alpha_i dt, or prev(t1) for the alpha == 1 stages
template <typename S>
IKR_HD S dp_stage_time(int stage, double t0, double dt) {
  S t0s = (S)t0, dts = (S)dt;
  if (stage >= 4) {
    S t1s = (S)(t0 + dt);
    return prev_representable(t1s);
  }
  return t0s + dp_alpha<S>(stage) * dts;
}

// stage state  y0 + sum_j k_j (beta_ij dt)   for one state component; k[] = k_0..k_stage
template <typename S>
IKR_HD S dp_stage_state(int stage, S y0, const S* k, double dt) {
  S dts = (S)dt;
  S acc = k[0] * (dp_beta<S>(stage, 0) * dts);
  for (int j = 1; j <= stage; ++j) acc = acc + k[j] * (dp_beta<S>(stage, j) * dts);
  return y0 + acc;
}

template <typename S>
IKR_HD S dp_error(const S* k, double dt) {
  S dts = (S)dt;
  S acc = k[0] * (dts * dp_cerr<S>(0));
  for (int j = 1; j < 7; ++j) acc = acc + k[j] * (dts * dp_cerr<S>(j));
  return acc;
}

// rms( err / (atol + rtol max(|y0|,|y1|)) ) over the two state components, in S
template <typename S>
IKR_HD S dp_error_ratio(S err_a, S err_r, S y0a, S y0r, S y1a, S y1r, double rtol, double atol) {
  S tol_a = (S)atol + (S)rtol * ikr_max(ikr_abs(y0a), ikr_abs(y1a));
  S tol_r = (S)atol + (S)rtol * ikr_max(ikr_abs(y0r), ikr_abs(y1r));
  S qa = err_a / tol_a, qr = err_r / tol_r;
  return ikr_sqrt((qa * qa + qr * qr) / (S)2);
}

struct StepControl {
  double safety, ifactor, dfactor;
};

// torchdiffeq `_optimal_step_size` (order 5)
IKR_HD double dp_next_dt(double dt, double ratio, const StepControl& c) {
  if (ratio == 0) return dt * c.ifactor;
  double dfac = ratio < 1 ? 1.0 : c.dfactor;
  double f = c.safety / pow(ratio, 0.2);
  f = f > dfac ? f : (f != f ? f : dfac);   // torch.max propagates NaN
  f = c.ifactor < f ? c.ifactor : (f != f ? f : f);
  return dt * f;
}

// Dense output: quartic through (y0, y1, y_mid, f0, f1); coefficients in S
template <typename S>
struct Dense {
  S e, d, c, b, a;
};

template <typename S>
IKR_HD Dense<S> dp_dense_fit(S y0, S y1, const S* k, double dt) {
  S dts = (S)dt;
  S acc = k[0] * (dts * dp_cmid<S>(0));
  for (int j = 1; j < 7; ++j) acc = acc + k[j] * (dts * dp_cmid<S>(j));
  S ymid = y0 + acc;
  S f0 = k[0], f1 = k[6];
  Dense<S> q;
  q.a = (S)2 * dts * (f1 - f0) - (S)8 * (y1 + y0) + (S)16 * ymid;
  q.b = dts * ((S)5 * f0 - (S)3 * f1) + (S)18 * y0 + (S)14 * y1 - (S)32 * ymid;
  q.c = dts * (f1 - (S)4 * f0) - (S)11 * y0 - (S)5 * y1 + (S)16 * ymid;
  q.d = dts * f0;
  q.e = y0;
  return q;
}

template <typename S>
IKR_HD S dp_dense_x(double t_lo, double t_hi, double t) {
  return (S)((t - t_lo) / (t_hi - t_lo));
}

template <typename S>
IKR_HD S dp_dense_eval(const Dense<S>& q, S x) {
  S total = q.e + x * q.d;
  S xp = x * x;
  total = total + xp * q.c;
  xp = xp * x;
  total = total + xp * q.b;
  xp = xp * x;
  total = total + xp * q.a;
  return total;
}

// Initial step heuristic, part 1: h0 from d0 = rms(y0/scale), d1 = rms(f0/scale)
template <typename S>
struct InitStep {
  S scale_a, scale_r;
  double d1;   // torchdiffeq keeps d1 in the dtype of f0 (fp64 for the reference RHS)
  double h0;
};

// `f0a,f0r` are the RHS values in the precision the RHS returned them (double): the reference
// RHS returns an fp64 tensor, so d1 / h0 / y1 are fp64 even for an fp32 state.
template <typename S>
IKR_HD InitStep<S> init_step_h0(S ya, S yr, double f0a, double f0r, double rtol, double atol,
                                bool rhs_is_f64) {
  InitStep<S> r;
  r.scale_a = (S)atol + ikr_abs(ya) * (S)rtol;
  r.scale_r = (S)atol + ikr_abs(yr) * (S)rtol;
  S qa = ya / r.scale_a, qr = yr / r.scale_r;
  S d0 = ikr_sqrt((qa * qa + qr * qr) / (S)2);
  if (rhs_is_f64) {
    double fa = f0a / (double)r.scale_a, fr = f0r / (double)r.scale_r;
    r.d1 = sqrt((fa * fa + fr * fr) / 2.0);
    if ((double)d0 < 1e-5 || r.d1 < 1e-5) r.h0 = (double)(S)1e-6;
    else r.h0 = 0.01 * (double)d0 / r.d1;
  } else {
    S fa = (S)f0a / r.scale_a, fr = (S)f0r / r.scale_r;
    S d1 = ikr_sqrt((fa * fa + fr * fr) / (S)2);
    r.d1 = (double)d1;
    if ((double)d0 < 1e-5 || r.d1 < 1e-5) r.h0 = (double)(S)1e-6;
    else r.h0 = (double)((S)0.01 * d0 / d1);
  }
  return r;
}

// part 2: given f1 = f(t0 + h0, y0 + h0 f0) -> first dt
template <typename S>
IKR_HD double init_step_finish(const InitStep<S>& r, double f0a, double f0r, double f1a,
                               double f1r, bool rhs_is_f64) {
  double d2;
  if (rhs_is_f64) {
    double qa = (f1a - f0a) / (double)r.scale_a, qr = (f1r - f0r) / (double)r.scale_r;
    d2 = sqrt((qa * qa + qr * qr) / 2.0) / r.h0;
  } else {
    S qa = ((S)f1a - (S)f0a) / r.scale_a, qr = ((S)f1r - (S)f0r) / r.scale_r;
    d2 = (double)(ikr_sqrt((qa * qa + qr * qr) / (S)2) / (S)r.h0);
  }
  double h1;
  if (r.d1 <= 1e-15 && d2 <= 1e-15) {
    double a = 1e-6, b = r.h0 * 1e-3;
    h1 = a > b ? a : b;
  } else {
    double m = r.d1 > d2 ? r.d1 : d2;
    h1 = pow(0.01 / m, 1.0 / 5.0);
  }
  double h = 100 * r.h0 < h1 ? 100 * r.h0 : h1;
  if (!rhs_is_f64) h = (double)(S)h;
  return h;
}

// observation  I = g a r (V - E)   (train-s1.py:328)
template <typename S>
IKR_HD S observe_current(S g, S a, S r, double v, S e) {
  return (S)((double)(g * a * r) * (v - (double)e));
}


// ==========================================================================================
// Lane = the complete solver state of ONE trajectory.  Kernels keep an array of lanes in
// shared memory (one owner thread per lane); the host logic test keeps one on the stack.
// ==========================================================================================
enum LaneStatus { LANE_OK = 0, LANE_DT_UNDERFLOW = 1, LANE_MAX_STEPS = 2, LANE_NONFINITE = 3,
                  LANE_CKPT_OVERFLOW = 4, LANE_RANGE = 5, LANE_DONE = 100, LANE_EMPTY = 101 };

template <typename S>
struct Lane {
  double t0, dt;      // start / size of the step being attempted (fp64 time)
  double hh_r, hh_a;  // HH terms of the stage under evaluation, as the reference forms them
  double f0a, f0r;    // RHS-precision f(t0, y0) kept for the initial-step heuristic
  double h0, d1;      // initial-step scratch
  S scale_a, scale_r;
  S ya, yr;           // state at t0
  S sa, sr;           // stage state handed to the RHS
  S ka[7], kr[7];     // stage derivatives k_0..k_6 (k_0 = f0, FSAL)
  int out_idx;        // next output time to emit
  int status;
  int n_acc, n_rej, n_int, nfe;
};

struct SolverCfg {
  ProtocolTable tab;
  HHParams hp;
  StepControl ctl;
  double vrange, netscale;
  double rtol, atol, first_step;
  long long max_num_steps;
  int nn_d;
  int mlp_is_f64;
};

template <typename S>
IKR_HD bool lane_active(const Lane<S>& L) { return L.status == LANE_OK; }

template <typename S>
IKR_HD void lane_reset(Lane<S>& L, S ya, S yr, double t_start, bool valid) {
  L.t0 = t_start; L.dt = 0; L.hh_r = 0; L.hh_a = 0; L.f0a = 0; L.f0r = 0; L.h0 = 0; L.d1 = 0;
  L.scale_a = L.scale_r = (S)1;
  L.ya = ya; L.yr = yr; L.sa = ya; L.sr = yr;
  for (int j = 0; j < 7; ++j) { L.ka[j] = (S)0; L.kr[j] = (S)0; }
  L.out_idx = 1; L.n_acc = 0; L.n_rej = 0; L.n_int = 0; L.nfe = 0;
  L.status = valid ? LANE_OK : LANE_EMPTY;
}

// ---- RHS split around the MLP --------------------------------------------------------------
// rhs_prepare: everything before the MLP (V(t), HH terms, MLP inputs).  `SS` is the dtype the
// state has *for this evaluation* (double for the second evaluation of the initial-step
// heuristic whose y1 = y0 + h0 f0 is fp64 in torchdiffeq when the RHS returns fp64).
template <typename SS, typename S>
IKR_HD void rhs_prepare(Lane<S>& L, const SolverCfg& c, double t_eval, SS a, SS r, double* nv,
                        double* a_in) {
  double v;
  bool in_table = table_voltage(c.tab, t_eval, &v);
  L.hh_r = hh_drdt<SS>(c.hp, v, in_table, r);
  L.hh_a = c.nn_d ? hh_dadt<SS>(c.hp, v, in_table, a) : 0.0;
  *nv = mlp_input_nv(v, in_table, c.vrange, c.mlp_is_f64 != 0);
  *a_in = (double)a;
}

// rhs_finish: combine the MLP output (already in the MLP dtype, passed as double) with the HH
// terms.  Returns (da/dt, dr/dt) in the precision the reference RHS returns them (double).
template <typename S>
IKR_HD void rhs_finish(const Lane<S>& L, const SolverCfg& c, double net_out, double* fa,
                       double* fr) {
  double scaled = c.mlp_is_f64 ? net_out / c.netscale
                               : (double)((float)net_out / (float)c.netscale);
  *fa = c.nn_d ? L.hh_a + scaled : scaled;
  *fr = L.hh_r;
}

// ---- time-only part of the RHS, computed ahead --------------------------------------------------
// V(t) and the HH rate constants depend on the stage TIME only, and the times of stages 1..5 of a
// dopri5 step are known when the step starts.  The tensor-core kernels fill this cache for stage
// s + 1 while the MMAs of stage s run (the owner thread is idle then), which takes the table
// search (dependent global loads) and the fp64 `exp`s off the serial part of an evaluation.
// Same expressions as hh_drdt / hh_rate_pair (in-table branch) => bit-identical results; anything
// else (time mismatch, out-of-table fallback) takes the uncached path.
struct TimeCache {
  double t, v, nv, k1, k2, k3, k4;
  int valid, in_table;
};
IKR_HD void time_cache_fill(TimeCache& tcx, const SolverCfg& c, double t_eval) {
  tcx.t = t_eval;
  tcx.in_table = table_voltage(c.tab, t_eval, &tcx.v) ? 1 : 0;
  if (tcx.in_table) {
    tcx.nv = mlp_input_nv(tcx.v, true, c.vrange, c.mlp_is_f64 != 0);
    tcx.k3 = c.hp.p[4] * exp(c.hp.p[5] * tcx.v);
    tcx.k4 = c.hp.p[6] * exp(-c.hp.p[7] * tcx.v);
    if (c.nn_d) {
      tcx.k1 = c.hp.p[0] * exp(c.hp.p[1] * tcx.v);
      tcx.k2 = c.hp.p[2] * exp(-c.hp.p[3] * tcx.v);
    }
  }
  tcx.valid = 1;
}
template <typename SS, typename S>
IKR_HD void rhs_prepare_cached(Lane<S>& L, const SolverCfg& c, double t_eval, SS a, SS r, double* nv,
                               double* a_in, const TimeCache& tcx) {
  if (!(tcx.valid && tcx.in_table && tcx.t == t_eval)) {
    rhs_prepare<SS, S>(L, c, t_eval, a, r, nv, a_in);
    return;
  }
  const SS one_minus_r = (SS)1 - r;
  L.hh_r = -tcx.k3 * (double)r + tcx.k4 * (double)one_minus_r;
  if (c.nn_d) {
    const SS one_minus_a = (SS)1 - a;
    L.hh_a = tcx.k1 * (double)one_minus_a - tcx.k2 * (double)a;
  } else {
    L.hh_a = 0.0;
  }
  *nv = tcx.nv;
  *a_in = (double)a;
}

// ---- dopri5 --------------------------------------------------------------------------------
template <typename S>
IKR_HD void dp_prepare_stage_cached(Lane<S>& L, const SolverCfg& c, int stage, double* nv, double* a_in,
                                    const TimeCache& tcx) {
  S ti = dp_stage_time<S>(stage, L.t0, L.dt);
  L.sa = dp_stage_state<S>(stage, L.ya, L.ka, L.dt);
  L.sr = dp_stage_state<S>(stage, L.yr, L.kr, L.dt);
  rhs_prepare_cached<S, S>(L, c, (double)ti, L.sa, L.sr, nv, a_in, tcx);
}
// fill the cache for stage `stage` of the step the lane is attempting
template <typename S>
IKR_HD void dp_prefetch_stage_time(const Lane<S>& L, const SolverCfg& c, int stage, TimeCache& tcx) {
  time_cache_fill(tcx, c, (double)dp_stage_time<S>(stage, L.t0, L.dt));
}

template <typename S>
IKR_HD void dp_prepare_stage(Lane<S>& L, const SolverCfg& c, int stage, double* nv, double* a_in) {
  S ti = dp_stage_time<S>(stage, L.t0, L.dt);
  L.sa = dp_stage_state<S>(stage, L.ya, L.ka, L.dt);
  L.sr = dp_stage_state<S>(stage, L.yr, L.kr, L.dt);
  rhs_prepare<S, S>(L, c, (double)ti, L.sa, L.sr, nv, a_in);
}

template <typename S>
IKR_HD void dp_store_stage(Lane<S>& L, const SolverCfg& c, int stage, double net_out) {
  double fa, fr;
  rhs_finish(L, c, net_out, &fa, &fr);
  L.ka[stage + 1] = (S)fa;
  L.kr[stage + 1] = (S)fr;
  if (lane_active(L)) L.nfe += 1;
}

// f0 = f(t[0], y0)
template <typename S>
IKR_HD void init_prepare_f0(Lane<S>& L, const SolverCfg& c, double* nv, double* a_in) {
  rhs_prepare<S, S>(L, c, (double)(S)L.t0, L.ya, L.yr, nv, a_in);
}
template <typename S>
IKR_HD void init_store_f0(Lane<S>& L, const SolverCfg& c, double net_out) {
  rhs_finish(L, c, net_out, &L.f0a, &L.f0r);
  L.ka[0] = (S)L.f0a;
  L.kr[0] = (S)L.f0r;
  if (lane_active(L)) L.nfe += 1;
}
// f1 = f(t0 + h0, y0 + h0 f0) of the initial-step heuristic (fp64 state for this evaluation)
template <typename S>
IKR_HD void init_prepare_f1(Lane<S>& L, const SolverCfg& c, double* nv, double* a_in) {
  InitStep<S> is = init_step_h0<S>(L.ya, L.yr, L.f0a, L.f0r, c.rtol, c.atol, true);
  L.scale_a = is.scale_a; L.scale_r = is.scale_r; L.h0 = is.h0; L.d1 = is.d1;
  double y1a = (double)L.ya + is.h0 * L.f0a;
  double y1r = (double)L.yr + is.h0 * L.f0r;
  double t1 = (double)(S)L.t0 + is.h0;
  rhs_prepare<double, S>(L, c, t1, y1a, y1r, nv, a_in);
}
template <typename S>
IKR_HD void init_store_f1(Lane<S>& L, const SolverCfg& c, double net_out) {
  double f1a, f1r;
  rhs_finish(L, c, net_out, &f1a, &f1r);
  InitStep<S> is;
  is.scale_a = L.scale_a; is.scale_r = L.scale_r; is.h0 = L.h0; is.d1 = L.d1;
  L.dt = init_step_finish<S>(is, L.f0a, L.f0r, f1a, f1r, true);
  if (lane_active(L)) L.nfe += 1;
}

// Pre-attempt checks (torchdiffeq's asserts at the top of `_adaptive_step`)
template <typename S>
IKR_HD void dp_check_before_step(Lane<S>& L, const SolverCfg& c) {
  if (!lane_active(L)) return;
  if (!(L.t0 + L.dt > L.t0)) { L.status = LANE_DT_UNDERFLOW; return; }
  if (!(isfinite((double)L.ya) && isfinite((double)L.yr))) { L.status = LANE_NONFINITE; return; }
  if ((long long)L.n_int >= c.max_num_steps) { L.status = LANE_MAX_STEPS; return; }
}

// After the six stages: error ratio, accept/reject, dense outputs, controller.
//   emit(idx, a, r)            : write output sample idx
//   checkpoint(step, lane) -> bool : record an accepted step (t0, dt, y0, k_0..k_6)
template <typename S, typename Emit, typename Ckpt>
IKR_HD void dp_finish_step(Lane<S>& L, const SolverCfg& c, const double* t_out, int T, Emit emit,
                           Ckpt checkpoint) {
  if (!lane_active(L)) return;
  S y1a = L.sa, y1r = L.sr;          // y1 = stage state of the last stage (FSAL tableau)
  S ea = dp_error<S>(L.ka, L.dt), er = dp_error<S>(L.kr, L.dt);
  S ratio_s = dp_error_ratio<S>(ea, er, L.ya, L.yr, y1a, y1r, c.rtol, c.atol);
  double ratio = (double)ratio_s;
  bool accept = ratio_s <= (S)1;
  L.n_int += 1;
  if (accept) {
    if (!checkpoint(L.n_acc, L)) {
      L.status = LANE_CKPT_OVERFLOW;
      return;
    }
    Dense<S> qa = dp_dense_fit<S>(L.ya, y1a, L.ka, L.dt);
    Dense<S> qr = dp_dense_fit<S>(L.yr, y1r, L.kr, L.dt);
    double t_lo = L.t0, t_hi = L.t0 + L.dt;
    while (L.out_idx < T && t_out[L.out_idx] <= t_hi) {
      S x = dp_dense_x<S>(t_lo, t_hi, t_out[L.out_idx]);
      emit(L.out_idx, dp_dense_eval<S>(qa, x), dp_dense_eval<S>(qr, x));
      L.out_idx += 1;
      L.n_int = 0;
    }
    L.ya = y1a; L.yr = y1r;
    L.ka[0] = L.ka[6]; L.kr[0] = L.kr[6];
    L.t0 = t_hi;
    L.n_acc += 1;
  } else {
    L.n_rej += 1;
  }
  L.dt = dp_next_dt(L.dt, ratio, c.ctl);
  if (L.out_idx >= T) L.status = LANE_DONE;
}

// ---- rk4 (3/8 rule, fixed grid) --------------------------------------------------------------
// stage s in 0..3 of the step [g0, g1].  `tf32`: grid arithmetic in fp32 (caller's t was fp32).
template <typename S>
IKR_HD double rk4_stage_time(int s, double g0, double g1, bool tf32, bool perturb) {
  double t;
  if (tf32) {
    float a = (float)g0, b = (float)g1, h = b - a;
    float third = (float)(1.0 / 3.0), two_third = (float)(2.0 / 3.0);
    float tt = s == 0 ? a : s == 1 ? a + h * third : s == 2 ? a + h * two_third : b;
    t = (double)tt;
  } else {
    double h = g1 - g0;
    t = s == 0 ? g0 : s == 1 ? g0 + h * (1.0 / 3.0) : s == 2 ? g0 + h * (2.0 / 3.0) : g1;
  }
  S ts = (S)t;
  if (perturb && s == 0) ts = next_representable(ts);
  if (perturb && s == 3) ts = prev_representable(ts);
  return (double)ts;
}

template <typename S>
IKR_HD S rk4_dt(double g0, double g1, bool tf32) {
  if (tf32) return (S)((float)g1 - (float)g0);
  return (S)(g1 - g0);
}

// stage state for one component; k[] = k1..k4 stored at index 0..3
template <typename S>
IKR_HD S rk4_stage_state(int s, S y0, const S* k, S dt) {
  const S third = (S)(1.0 / 3.0);
  switch (s) {
    case 0: return y0;
    case 1: return y0 + dt * k[0] * third;
    case 2: return y0 + dt * (k[1] - k[0] * third);
    default: return y0 + dt * (k[0] - k[1] + k[2]);
  }
}

template <typename S>
IKR_HD S rk4_combine(S y0, const S* k, S dt) {
  return y0 + (k[0] + (S)3 * (k[1] + k[2]) + k[3]) * dt * (S)0.125;
}

template <typename S>
IKR_HD void rk4_prepare_stage(Lane<S>& L, const SolverCfg& c, int s, double g0, double g1,
                              bool tf32, bool perturb, double* nv, double* a_in) {
  double ti = rk4_stage_time<S>(s, g0, g1, tf32, perturb);
  S dt = rk4_dt<S>(g0, g1, tf32);
  L.sa = rk4_stage_state<S>(s, L.ya, L.ka, dt);
  L.sr = rk4_stage_state<S>(s, L.yr, L.kr, dt);
  rhs_prepare<S, S>(L, c, ti, L.sa, L.sr, nv, a_in);
}

template <typename S>
IKR_HD void rk4_store_stage(Lane<S>& L, const SolverCfg& c, int s, double net_out) {
  double fa, fr;
  rhs_finish(L, c, net_out, &fa, &fr);
  L.ka[s] = (S)fa;
  L.kr[s] = (S)fr;
  if (lane_active(L)) L.nfe += 1;
}

// finish the step [g0,g1]: y1, emit outputs t_out[j] <= g1 (linear interpolation off-grid)
template <typename S, typename Emit>
IKR_HD void rk4_finish_step(Lane<S>& L, double g0, double g1, bool tf32, const double* t_out,
                            int T, Emit emit) {
  if (!lane_active(L)) return;
  S dt = rk4_dt<S>(g0, g1, tf32);
  S y1a = rk4_combine<S>(L.ya, L.ka, dt), y1r = rk4_combine<S>(L.yr, L.kr, dt);
  while (L.out_idx < T && g1 >= t_out[L.out_idx]) {
    double tj = t_out[L.out_idx];
    if (tj == g0) emit(L.out_idx, L.ya, L.yr);
    else if (tj == g1) emit(L.out_idx, y1a, y1r);
    else {
      S slope = tf32 ? (S)(((float)tj - (float)g0) / ((float)g1 - (float)g0))
                     : (S)((tj - g0) / (g1 - g0));
      emit(L.out_idx, L.ya + slope * (y1a - L.ya), L.yr + slope * (y1r - L.yr));
    }
    L.out_idx += 1;
  }
  L.ya = y1a; L.yr = y1r;
  L.n_acc += 1;
  if (!(isfinite((double)L.ya) && isfinite((double)L.yr))) L.status = LANE_NONFINITE;
  else if (L.out_idx >= T) L.status = LANE_DONE;
}


// ==========================================================================================
// Discrete adjoint of one accepted step (backward sweep).  Semantics = PyTorch autograd through
// torchdiffeq's non-adjoint `odeint`: every accepted step's tensor ops are differentiated, step
// sizes, stage times and dense-output abscissae are constants, rejected steps contribute nothing.
//
// The forward pass checkpoints, per accepted step, (t0, dt, y0, k_0..k_6) -- see StepCkpt -- so
// the sweep never re-integrates: stage states are re-formed from the checkpoint (same arithmetic
// as the forward), every stage costs exactly one MLP forward (activations kept) + one MLP
// backward.  The sweep walks the steps last-to-first and, inside a step, the stages 5..0.
// ==========================================================================================
constexpr int kCkptVals = 16;   // per accepted step, state dtype: ya, yr, ka[0..6], kr[0..6]

template <typename S>
IKR_HD void ckpt_pack(const Lane<S>& L, S* dst) {
  dst[0] = L.ya; dst[1] = L.yr;
  for (int j = 0; j < 7; ++j) { dst[2 + j] = L.ka[j]; dst[9 + j] = L.kr[j]; }
}

template <typename S>
struct BLane {
  double t0, dt;         // the step being reversed
  S ya, yr;              // y0 of that step
  S ka[7], kr[7];        // k_0 .. k_6 of that step (checkpoint)
  S lka[7], lkr[7];      // adjoints of k_0 .. k_6
  S lya, lyr;            // in: adjoint of y1 (from the later steps); out: adjoint of y0
  S lfa, lfr;            // in: adjoint of f1 = k_6 (FSAL: it is k_0 of the next step)
  S la, lr;              // adjoint of y1 = Y_5 seeded by bdp_seed_step
  S ja, jr;              // HH Jacobian terms d(fa_hh)/da, d(fr)/dr of the stage in flight
  S gsum;                // accumulated dL/dg (fused-loss mode)
  int n_left;            // accepted steps still to reverse (step index = n_left - 1)
  int out_idx;           // largest output index not yet consumed
  int phase;             // 0: reversing steps, 1: f(t[0], y0) pending, 2: finished
};

template <typename S>
IKR_HD void blane_reset(BLane<S>& B, int n_acc, int T, bool valid) {
  B.t0 = 0; B.dt = 0; B.ya = (S)0; B.yr = (S)1;
  for (int j = 0; j < 7; ++j) { B.ka[j] = B.kr[j] = B.lka[j] = B.lkr[j] = (S)0; }
  B.lya = B.lyr = B.lfa = B.lfr = B.la = B.lr = B.ja = B.jr = B.gsum = (S)0;
  B.n_left = valid ? n_acc : 0;
  B.out_idx = T - 1;
  B.phase = valid ? (n_acc > 0 ? 0 : 1) : 2;
}

template <typename S>
IKR_HD void blane_load_step(BLane<S>& B, double t0, double dt, const S* ck) {
  B.t0 = t0; B.dt = dt; B.ya = ck[0]; B.yr = ck[1];
  for (int j = 0; j < 7; ++j) { B.ka[j] = ck[2 + j]; B.kr[j] = ck[9 + j]; }
}

template <typename S>
struct DenseAdj {
  S a, b, c, d, e;
};

// adjoint of  y(x) = e + x d + x^2 c + x^3 b + x^4 a   w.r.t. the coefficients
template <typename S>
IKR_HD void adj_dense_accumulate(DenseAdj<S>& q, S x, S g) {
  q.e = q.e + g;
  q.d = q.d + x * g;
  S xp = x * x;
  q.c = q.c + xp * g;
  xp = xp * x;
  q.b = q.b + xp * g;
  xp = xp * x;
  q.a = q.a + xp * g;
}

// coefficients -> (y0, y1, y_mid, f0, f1); see dp_dense_fit
template <typename S>
IKR_HD void adj_dense_to_inputs(const DenseAdj<S>& q, double dt, S* ly0, S* ly1, S* lymid, S* lf0,
                                S* lf1) {
  S dts = (S)dt;
  *ly0 = q.e - (S)8 * q.a + (S)18 * q.b - (S)11 * q.c;
  *ly1 = -(S)8 * q.a + (S)14 * q.b - (S)5 * q.c;
  *lymid = (S)16 * q.a - (S)32 * q.b + (S)16 * q.c;
  *lf0 = dts * (-(S)2 * q.a + (S)5 * q.b - (S)4 * q.c + q.d);
  *lf1 = dts * ((S)2 * q.a - (S)3 * q.b + q.c);
}

// Seed the adjoints of the loaded step from
//   - the incoming adjoints of y1 (B.lya/lyr) and f1 (B.lfa/lfr),
//   - the gradients of the outputs that fall in (t0, t0 + dt]   (grad(idx, &ga, &gr)).
// Leaves lka/lkr[0..6] seeded, B.la/lr = adjoint of Y_5 (= y1), and B.lya/lyr = the part of the
// adjoint of y0 that is already known.
template <typename S, typename Grad>
IKR_HD void bdp_seed_step(BLane<S>& B, const double* t_out, Grad grad) {
  DenseAdj<S> qa = {(S)0, (S)0, (S)0, (S)0, (S)0}, qr = {(S)0, (S)0, (S)0, (S)0, (S)0};
  const double t_lo = B.t0, t_hi = B.t0 + B.dt;
  bool any = false;
  while (B.out_idx >= 1 && t_out[B.out_idx] > t_lo) {
    S ga, gr;
    grad(B.out_idx, &ga, &gr);
    S x = dp_dense_x<S>(t_lo, t_hi, t_out[B.out_idx]);
    adj_dense_accumulate(qa, x, ga);
    adj_dense_accumulate(qr, x, gr);
    B.out_idx -= 1;
    any = true;
  }
  for (int j = 0; j < 7; ++j) { B.lka[j] = (S)0; B.lkr[j] = (S)0; }
  S ly1a = B.lya, ly1r = B.lyr;       // adjoint of y1 from the later steps
  S ly0a = (S)0, ly0r = (S)0;
  B.lka[6] = B.lfa; B.lkr[6] = B.lfr;  // f1 = k_6 is k_0 of the next step
  if (any) {
    S dts = (S)B.dt;
    S a0, a1, am, f0, f1;
    adj_dense_to_inputs(qa, B.dt, &a0, &a1, &am, &f0, &f1);
    ly0a = ly0a + a0 + am; ly1a = ly1a + a1;
    B.lka[0] = B.lka[0] + f0; B.lka[6] = B.lka[6] + f1;
    for (int j = 0; j < 7; ++j) B.lka[j] = B.lka[j] + (dts * dp_cmid<S>(j)) * am;
    adj_dense_to_inputs(qr, B.dt, &a0, &a1, &am, &f0, &f1);
    ly0r = ly0r + a0 + am; ly1r = ly1r + a1;
    B.lkr[0] = B.lkr[0] + f0; B.lkr[6] = B.lkr[6] + f1;
    for (int j = 0; j < 7; ++j) B.lkr[j] = B.lkr[j] + (dts * dp_cmid<S>(j)) * am;
  }
  B.la = ly1a; B.lr = ly1r;            // y1 = Y_5
  B.lya = ly0a; B.lyr = ly0r;
}

// d(fa)/d(net_out) = 1 / netscale applied to an adjoint of fa
IKR_HD double adj_mlp_upstream(double lk, const SolverCfg& c) {
  return c.mlp_is_f64 ? lk / c.netscale : (double)((float)lk / (float)c.netscale);
}

// Stage s (5..0) of the loaded step: MLP inputs (nv, a_in), the upstream gradient of the MLP
// output `up` = lambda_{k_{s+1}} / netscale, and the HH Jacobian terms (kept in B.ja/jr).
template <typename S>
IKR_HD void bdp_stage_inputs(BLane<S>& B, const SolverCfg& c, int s, double* nv, double* a_in,
                             double* up) {
  S ti = dp_stage_time<S>(s, B.t0, B.dt);
  S Ya = dp_stage_state<S>(s, B.ya, B.ka, B.dt);
  double v;
  bool in_table = table_voltage(c.tab, (double)ti, &v);
  B.jr = (S)hh_drdt_dr(c.hp, v, in_table);
  B.ja = c.nn_d ? (S)hh_dadt_da(c.hp, v, in_table) : (S)0;
  *nv = mlp_input_nv(v, in_table, c.vrange, c.mlp_is_f64 != 0);
  *a_in = (double)Ya;
  *up = adj_mlp_upstream((double)B.lka[s + 1], c);
}

// Same with the time-only terms (V, rate constants) taken from a TimeCache filled ahead
template <typename S>
IKR_HD void bdp_stage_inputs_cached(BLane<S>& B, const SolverCfg& c, int s, double* nv, double* a_in,
                                    double* up, const TimeCache& tcx) {
  S ti = dp_stage_time<S>(s, B.t0, B.dt);
  if (!(tcx.valid && tcx.in_table && tcx.t == (double)ti)) {
    bdp_stage_inputs<S>(B, c, s, nv, a_in, up);
    return;
  }
  S Ya = dp_stage_state<S>(s, B.ya, B.ka, B.dt);
  B.jr = (S)(-(tcx.k3 + tcx.k4));
  B.ja = c.nn_d ? (S)(-(tcx.k1 + tcx.k2)) : (S)0;
  *nv = tcx.nv;
  *a_in = (double)Ya;
  *up = adj_mlp_upstream((double)B.lka[s + 1], c);
}
template <typename S>
IKR_HD void bdp_prefetch_stage_time(const BLane<S>& B, const SolverCfg& c, int stage, TimeCache& tcx) {
  time_cache_fill(tcx, c, (double)dp_stage_time<S>(stage, B.t0, B.dt));
}

// Reverse stage s.  Before: B.lka/lkr[s+1] complete.  `mlp_grad_a` = up * d(net)/d(a_in) (the
// MLP backward result for this lane).
//   lambda_Y_s = J^T lambda_k_{s+1} (+ lambda_y1 for s = 5, carried in B.la/lr)
//   lambda_y0 += lambda_Y_s ; lambda_k_j += (beta_sj dt) lambda_Y_s  (j <= s)
template <typename S>
IKR_HD void bdp_reverse_stage(BLane<S>& B, int s, S mlp_grad_a) {
  S lYa = mlp_grad_a + B.lka[s + 1] * B.ja;
  S lYr = B.lkr[s + 1] * B.jr;
  if (s == 5) { lYa = lYa + B.la; lYr = lYr + B.lr; }
  S dts = (S)B.dt;
  B.lya = B.lya + lYa; B.lyr = B.lyr + lYr;
  for (int j = 0; j <= s; ++j) {
    S w = dp_beta<S>(s, j) * dts;
    B.lka[j] = B.lka[j] + w * lYa;
    B.lkr[j] = B.lkr[j] + w * lYr;
  }
}

// End of the step: the adjoints of y0 / k_0 become the incoming adjoints of the previous step.
template <typename S>
IKR_HD void bdp_finish_step(BLane<S>& B) {
  B.lfa = B.lka[0]; B.lfr = B.lkr[0];
  B.n_left -= 1;
  if (B.n_left <= 0) B.phase = 1;
}

// The first evaluation f0 = f(t[0], y0) (phase 1): its output adjoint is what the first step left
// in lfa/lfr.
template <typename S>
IKR_HD void bdp_f0_inputs(BLane<S>& B, const SolverCfg& c, double t_start, S y0a, double* nv,
                          double* a_in, double* up) {
  double v;
  bool in_table = table_voltage(c.tab, (double)(S)t_start, &v);
  B.jr = (S)hh_drdt_dr(c.hp, v, in_table);
  B.ja = c.nn_d ? (S)hh_dadt_da(c.hp, v, in_table) : (S)0;
  *nv = mlp_input_nv(v, in_table, c.vrange, c.mlp_is_f64 != 0);
  *a_in = (double)y0a;
  *up = adj_mlp_upstream((double)B.lfa, c);
}
// `g0a, g0r` = dL/dy_out[0] (the first output sample is y0 itself)
template <typename S>
IKR_HD void bdp_f0_finish(BLane<S>& B, S mlp_grad_a, S g0a, S g0r) {
  B.lya = B.lya + mlp_grad_a + B.lfa * B.ja + g0a;
  B.lyr = B.lyr + B.lfr * B.jr + g0r;
  B.phase = 2;
}

// ==========================================================================================
// Discrete adjoint of one rk4 (3/8 rule) step on the fixed grid -- same semantics as above (PyTorch
// autograd through torchdiffeq's fixed-grid solver: the grid and the interpolation abscissae are
// constants).  Checkpoint of a step: ckpt_t = (g0, g1), ckpt_y = (y0; k1..k4 in the first four k
// slots).  BLane.t0 / BLane.dt hold g0 / g1.  No FSAL: every step starts from y0 alone.
//   Y_0 = y0, Y_1 = y0 + dt k1 / 3, Y_2 = y0 + dt (k2 - k1 / 3), Y_3 = y0 + dt (k1 - k2 + k3),
//   y1 = y0 + (k1 + 3 (k2 + k3) + k4) dt / 8;  outputs inside (g0, g1) are y0 + slope (y1 - y0).
// ==========================================================================================
template <typename S, typename Grad>
IKR_HD void brk4_seed_step(BLane<S>& B, const double* t_out, bool tf32, Grad grad) {
  const double g0 = B.t0, g1 = B.dt;
  S ly1a = B.lya, ly1r = B.lyr;       // adjoint of y1 from the later steps
  S ly0a = (S)0, ly0r = (S)0;
  while (B.out_idx >= 1 && t_out[B.out_idx] > g0) {
    S ga, gr;
    grad(B.out_idx, &ga, &gr);
    const double tj = t_out[B.out_idx];
    if (tj == g1) {
      ly1a = ly1a + ga; ly1r = ly1r + gr;
    } else {
      const S slope = tf32 ? (S)(((float)tj - (float)g0) / ((float)g1 - (float)g0))
                           : (S)((tj - g0) / (g1 - g0));
      ly1a = ly1a + slope * ga; ly1r = ly1r + slope * gr;
      ly0a = ly0a + (ga - slope * ga); ly0r = ly0r + (gr - slope * gr);
    }
    B.out_idx -= 1;
  }
  const S dt = rk4_dt<S>(g0, g1, tf32);
  const S ga = (ly1a * (S)0.125) * dt, gr = (ly1r * (S)0.125) * dt;   // adjoint of the k sum
  for (int j = 0; j < 7; ++j) { B.lka[j] = (S)0; B.lkr[j] = (S)0; }
  B.lka[0] = ga; B.lka[3] = ga; B.lka[1] = (S)3 * ga; B.lka[2] = (S)3 * ga;
  B.lkr[0] = gr; B.lkr[3] = gr; B.lkr[1] = (S)3 * gr; B.lkr[2] = (S)3 * gr;
  B.lya = ly0a + ly1a; B.lyr = ly0r + ly1r;
  B.lfa = (S)0; B.lfr = (S)0;
}

// stage s (3..0): MLP inputs, upstream gradient of the MLP output (adjoint of k_{s+1} = ka[s]) and
// the HH Jacobian terms
template <typename S>
IKR_HD void brk4_stage_inputs(BLane<S>& B, const SolverCfg& c, int s, bool tf32, bool perturb,
                              double* nv, double* a_in, double* up) {
  const double ti = rk4_stage_time<S>(s, B.t0, B.dt, tf32, perturb);
  const S dt = rk4_dt<S>(B.t0, B.dt, tf32);
  const S Ya = rk4_stage_state<S>(s, B.ya, B.ka, dt);
  double v;
  bool in_table = table_voltage(c.tab, ti, &v);
  B.jr = (S)hh_drdt_dr(c.hp, v, in_table);
  B.ja = c.nn_d ? (S)hh_dadt_da(c.hp, v, in_table) : (S)0;
  *nv = mlp_input_nv(v, in_table, c.vrange, c.mlp_is_f64 != 0);
  *a_in = (double)Ya;
  *up = adj_mlp_upstream((double)B.lka[s], c);
}

template <typename S>
IKR_HD void brk4_reverse_stage(BLane<S>& B, int s, bool tf32, S mlp_grad_a) {
  const S lYa = mlp_grad_a + B.lka[s] * B.ja;
  const S lYr = B.lkr[s] * B.jr;
  B.lya = B.lya + lYa; B.lyr = B.lyr + lYr;
  const S dt = rk4_dt<S>(B.t0, B.dt, tf32);
  const S third = (S)(1.0 / 3.0);
  if (s == 1) {
    B.lka[0] = B.lka[0] + (dt * lYa) * third; B.lkr[0] = B.lkr[0] + (dt * lYr) * third;
  } else if (s == 2) {
    B.lka[1] = B.lka[1] + dt * lYa; B.lkr[1] = B.lkr[1] + dt * lYr;
    B.lka[0] = B.lka[0] - (dt * lYa) * third; B.lkr[0] = B.lkr[0] - (dt * lYr) * third;
  } else if (s == 3) {
    B.lka[0] = B.lka[0] + dt * lYa; B.lkr[0] = B.lkr[0] + dt * lYr;
    B.lka[1] = B.lka[1] - dt * lYa; B.lkr[1] = B.lkr[1] - dt * lYr;
    B.lka[2] = B.lka[2] + dt * lYa; B.lkr[2] = B.lkr[2] + dt * lYr;
  }
}

template <typename S>
IKR_HD void brk4_finish_step(BLane<S>& B) {
  B.n_left -= 1;
  if (B.n_left <= 0) B.phase = 1;
}
// after the first step has been reversed: the first output sample is y0 itself
template <typename S>
IKR_HD void brk4_finish(BLane<S>& B, S g0a, S g0r) {
  B.lya = B.lya + g0a; B.lyr = B.lyr + g0r;
  B.phase = 2;
}

}  // namespace ikr
#endif  // IKR_MATH_H_
