// host_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Drives the per-lane solver state machine of csrc/ikr_math.h (the exact code the CUDA kernels
// run between MLP evaluations) on the host with a scalar MLP, so that the solver logic can be
// checked against the CPU oracle in the GPU-less build container.  Built by
// tests/test_solver_logic_host.py with g++; the product path never loads it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ikr_math.h"

using namespace ikr;

namespace {

template <typename W>
struct ScalarMlp {
  int L, n;
  const W* w0;               // [n][2]
  const W* b0;               // [n]
  std::vector<const W*> wh;  // L x [n][n]  (out, in)
  std::vector<const W*> bh;  // L x [n]
  const W* wl;               // [n]
  W bl;
  W slope;
  std::vector<W> h, z;
  W operator()(W nv, W a) {
    h.assign(n, 0);
    z.assign(n, 0);
    for (int j = 0; j < n; ++j) {
      W s = std::fma(w0[2 * j + 1], a, std::fma(w0[2 * j], nv, (W)0)) + b0[j];
      h[j] = s > 0 ? s : s * slope;
    }
    for (int l = 0; l < L; ++l) {
      for (int j = 0; j < n; ++j) {
        W s = 0;
        for (int k = 0; k < n; ++k) s = std::fma(h[k], wh[l][(size_t)j * n + k], s);
        s += bh[l][j];
        z[j] = s > 0 ? s : s * slope;
      }
      h.swap(z);
    }
    W s = 0;
    for (int k = 0; k < n; ++k) s = std::fma(h[k], wl[k], s);
    return s + bl;
  }
  // forward keeping every layer's activations, then backward: accumulates dL/dparams (flat,
  // state_dict order) for upstream `up` = dL/d(net_out); returns dL/da (second input).
  W grad(W nv, W a, W up, std::vector<double>& g) {
    std::vector<std::vector<W>> H(L + 1, std::vector<W>(n));
    for (int j = 0; j < n; ++j) {
      W s = std::fma(w0[2 * j + 1], a, std::fma(w0[2 * j], nv, (W)0)) + b0[j];
      H[0][j] = s > 0 ? s : s * slope;
    }
    for (int l = 0; l < L; ++l)
      for (int j = 0; j < n; ++j) {
        W s = 0;
        for (int k = 0; k < n; ++k) s = std::fma(H[l][k], wh[l][(size_t)j * n + k], s);
        s += bh[l][j];
        H[l + 1][j] = s > 0 ? s : s * slope;
      }
    size_t off_w0 = 0, off_b0 = 2 * (size_t)n, off_h = 3 * (size_t)n;
    size_t per = (size_t)n * n + n;
    size_t off_wl = off_h + L * per, off_bl = off_wl + n;
    std::vector<W> dz(n), dh(n);
    for (int k = 0; k < n; ++k) {
      g[off_wl + k] += (double)(up * H[L][k]);
      dh[k] = wl[k] * up;
    }
    g[off_bl] += (double)up;
    for (int l = L; l >= 1; --l) {
      for (int o = 0; o < n; ++o) dz[o] = dh[o] * (H[l][o] > 0 ? (W)1 : slope);
      for (int o = 0; o < n; ++o) {
        for (int i = 0; i < n; ++i) g[off_h + (l - 1) * per + (size_t)o * n + i] += (double)(dz[o] * H[l - 1][i]);
        g[off_h + (l - 1) * per + (size_t)n * n + o] += (double)dz[o];
      }
      for (int i = 0; i < n; ++i) {
        W s = 0;
        for (int o = 0; o < n; ++o) s = std::fma(wh[l - 1][(size_t)o * n + i], dz[o], s);
        dh[i] = s;
      }
    }
    W da = 0;
    for (int o = 0; o < n; ++o) {
      W d0 = dh[o] * (H[0][o] > 0 ? (W)1 : slope);
      g[off_w0 + 2 * o] += (double)(d0 * nv);
      g[off_w0 + 2 * o + 1] += (double)(d0 * a);
      g[off_b0 + o] += (double)d0;
      da = std::fma(w0[2 * o + 1], d0, da);
    }
    return da;
  }
};

struct Args {
  int L, n, nn_d, method, time_f32, rk4_perturb;
  const void* params;  // flat: w0,b0,(wh,bh)xL,wl,bl in state_dict order
  const double* tab_t;
  const double* tab_v;
  int tab_len, tab_uniform;
  double tab_t0, tab_inv_dt;
  const double* p;     // 8
  double rtol, atol, first_step;
  const double* t_out;
  int T;
  const double* grid;
  int G;
  double y0a, y0r;
  double* y_out;       // [T][2] as double
  int* stats;          // n_acc, n_rej, nfe, status
  double* steps;       // optional [cap][2] accepted (t0, dt)
  int steps_cap;
};

template <typename S, typename W>
void run(const Args& a) {
  ScalarMlp<W> mlp;
  mlp.L = a.L; mlp.n = a.n; mlp.slope = (W)0.01;
  const W* p = (const W*)a.params;
  mlp.w0 = p; p += 2 * a.n;
  mlp.b0 = p; p += a.n;
  for (int l = 0; l < a.L; ++l) {
    mlp.wh.push_back(p); p += (size_t)a.n * a.n;
    mlp.bh.push_back(p); p += a.n;
  }
  mlp.wl = p; p += a.n;
  mlp.bl = *p;

  SolverCfg c;
  c.tab.t = a.tab_t; c.tab.v = a.tab_v; c.tab.len = a.tab_len; c.tab.uniform = a.tab_uniform;
  c.tab.t0 = a.tab_t0; c.tab.inv_dt = a.tab_inv_dt;
  for (int i = 0; i < 8; ++i) c.hp.p[i] = a.p[i];
  c.ctl.safety = 0.9; c.ctl.ifactor = 10.0; c.ctl.dfactor = 0.2;
  c.vrange = 100.0; c.netscale = 1000.0;
  c.rtol = a.rtol; c.atol = a.atol; c.first_step = a.first_step;
  c.max_num_steps = 2147483647LL;
  c.nn_d = a.nn_d;
  c.mlp_is_f64 = sizeof(W) == 8;

  Lane<S> L;
  lane_reset<S>(L, (S)a.y0a, (S)a.y0r, a.t_out[0], true);
  a.y_out[0] = (double)L.ya; a.y_out[1] = (double)L.yr;
  auto emit = [&](int idx, S ya, S yr) { a.y_out[2 * idx] = (double)ya; a.y_out[2 * idx + 1] = (double)yr; };
  double nv, ain;

  if (a.method == 0) {
    auto ckpt = [&](int step, const Lane<S>& lane) {
      if (a.steps && step < a.steps_cap) { a.steps[2 * step] = lane.t0; a.steps[2 * step + 1] = lane.dt; }
      return true;
    };
    init_prepare_f0<S>(L, c, &nv, &ain);
    init_store_f0<S>(L, c, (double)mlp((W)nv, (W)ain));
    if (a.first_step > 0) {
      L.dt = a.first_step;
    } else {
      init_prepare_f1<S>(L, c, &nv, &ain);
      init_store_f1<S>(L, c, (double)mlp((W)nv, (W)ain));
    }
    if (a.T <= 1) L.status = LANE_DONE;
    while (true) {
      dp_check_before_step<S>(L, c);
      if (!lane_active(L)) break;
      for (int s = 0; s < 6; ++s) {
        dp_prepare_stage<S>(L, c, s, &nv, &ain);
        dp_store_stage<S>(L, c, s, (double)mlp((W)nv, (W)ain));
      }
      dp_finish_step<S>(L, c, a.t_out, a.T, emit, ckpt);
    }
  } else {
    if (a.T <= 1) L.status = LANE_DONE;
    for (int gi = 0; gi + 1 < a.G && lane_active(L); ++gi) {
      double g0 = a.grid[gi], g1 = a.grid[gi + 1];
      for (int s = 0; s < 4; ++s) {
        rk4_prepare_stage<S>(L, c, s, g0, g1, a.time_f32 != 0, a.rk4_perturb != 0, &nv, &ain);
        rk4_store_stage<S>(L, c, s, (double)mlp((W)nv, (W)ain));
      }
      rk4_finish_step<S>(L, g0, g1, a.time_f32 != 0, a.t_out, a.T, emit);
    }
  }
  a.stats[0] = L.n_acc; a.stats[1] = L.n_rej; a.stats[2] = L.nfe;
  a.stats[3] = L.status == LANE_DONE ? 0 : L.status;
}

}  // namespace

extern "C" int harness_integrate(int state_f64, int mlp_f64, const Args* a) {
  if (state_f64 && mlp_f64) run<double, double>(*a);
  else if (state_f64 && !mlp_f64) run<double, float>(*a);
  else if (!state_f64 && !mlp_f64) run<float, float>(*a);
  else return -1;
  return 0;
}

namespace {
struct GradArgs {
  const double* grad_y;   // [T][2]
  double* grad_params;    // flat, state_dict order (accumulated)
  double* grad_y0;        // [2]
  int steps_cap;
};

template <typename S, typename W>
void run_grad(const Args& a, const GradArgs& ga) {
  // forward with step recording
  std::vector<double> rec_t;  // t0, dt
  std::vector<S> rec_y;       // ya, yr, fa, fr
  ScalarMlp<W> mlp;
  mlp.L = a.L; mlp.n = a.n; mlp.slope = (W)0.01;
  const W* p = (const W*)a.params;
  mlp.w0 = p; p += 2 * a.n;
  mlp.b0 = p; p += a.n;
  for (int l = 0; l < a.L; ++l) { mlp.wh.push_back(p); p += (size_t)a.n * a.n; mlp.bh.push_back(p); p += a.n; }
  mlp.wl = p; p += a.n;
  mlp.bl = *p;
  SolverCfg c;
  c.tab.t = a.tab_t; c.tab.v = a.tab_v; c.tab.len = a.tab_len; c.tab.uniform = a.tab_uniform;
  c.tab.t0 = a.tab_t0; c.tab.inv_dt = a.tab_inv_dt;
  for (int i = 0; i < 8; ++i) c.hp.p[i] = a.p[i];
  c.ctl.safety = 0.9; c.ctl.ifactor = 10.0; c.ctl.dfactor = 0.2;
  c.vrange = 100.0; c.netscale = 1000.0;
  c.rtol = a.rtol; c.atol = a.atol; c.first_step = a.first_step;
  c.max_num_steps = 2147483647LL; c.nn_d = a.nn_d; c.mlp_is_f64 = sizeof(W) == 8;

  Lane<S> L;
  lane_reset<S>(L, (S)a.y0a, (S)a.y0r, a.t_out[0], true);
  auto emit = [&](int idx, S ya, S yr) { a.y_out[2 * idx] = (double)ya; a.y_out[2 * idx + 1] = (double)yr; };
  auto ckpt = [&](int, const Lane<S>& lane) {
    rec_t.push_back(lane.t0); rec_t.push_back(lane.dt);
    S buf[kCkptVals];
    ckpt_pack<S>(lane, buf);
    for (int i = 0; i < kCkptVals; ++i) rec_y.push_back(buf[i]);
    return true;
  };
  double nv, ain, up;
  a.y_out[0] = (double)L.ya; a.y_out[1] = (double)L.yr;
  init_prepare_f0<S>(L, c, &nv, &ain);
  init_store_f0<S>(L, c, (double)mlp((W)nv, (W)ain));
  if (a.first_step > 0) L.dt = a.first_step;
  else { init_prepare_f1<S>(L, c, &nv, &ain); init_store_f1<S>(L, c, (double)mlp((W)nv, (W)ain)); }
  if (a.T <= 1) L.status = LANE_DONE;
  while (true) {
    dp_check_before_step<S>(L, c);
    if (!lane_active(L)) break;
    for (int s = 0; s < 6; ++s) {
      dp_prepare_stage<S>(L, c, s, &nv, &ain);
      dp_store_stage<S>(L, c, s, (double)mlp((W)nv, (W)ain));
    }
    dp_finish_step<S>(L, c, a.t_out, a.T, emit, ckpt);
  }
  a.stats[0] = L.n_acc; a.stats[1] = L.n_rej; a.stats[2] = L.nfe; a.stats[3] = L.status == LANE_DONE ? 0 : L.status;

  // reverse sweep (the same lane functions the backward kernel runs)
  size_t n_par = 3 * (size_t)a.n + (size_t)a.L * ((size_t)a.n * a.n + a.n) + a.n + 1;
  std::vector<double> g(n_par, 0.0);
  BLane<S> B;
  blane_reset<S>(B, L.n_acc, a.T, true);
  auto grad = [&](int idx, S* ga_, S* gr_) { *ga_ = (S)ga.grad_y[2 * idx]; *gr_ = (S)ga.grad_y[2 * idx + 1]; };
  while (B.phase == 0) {
    int j = B.n_left - 1;
    blane_load_step<S>(B, rec_t[2 * j], rec_t[2 * j + 1], &rec_y[(size_t)kCkptVals * j]);
    bdp_seed_step<S>(B, a.t_out, grad);
    for (int s = 5; s >= 0; --s) {
      bdp_stage_inputs<S>(B, c, s, &nv, &ain, &up);
      W da = mlp.grad((W)nv, (W)ain, (W)up, g);
      bdp_reverse_stage<S>(B, s, (S)da);
    }
    bdp_finish_step<S>(B);
  }
  bdp_f0_inputs<S>(B, c, a.t_out[0], (S)a.y0a, &nv, &ain, &up);
  W da = mlp.grad((W)nv, (W)ain, (W)up, g);
  bdp_f0_finish<S>(B, (S)da, (S)ga.grad_y[0], (S)ga.grad_y[1]);
  ga.grad_y0[0] = (double)B.lya;
  ga.grad_y0[1] = (double)B.lyr;
  for (size_t i = 0; i < n_par; ++i) ga.grad_params[i] = g[i];
}
}  // namespace

extern "C" int harness_grad(int state_f64, int mlp_f64, const Args* a, const GradArgs* g) {
  if (state_f64 && mlp_f64) run_grad<double, double>(*a, *g);
  else if (state_f64 && !mlp_f64) run_grad<double, float>(*a, *g);
  else if (!state_f64 && !mlp_f64) run_grad<float, float>(*a, *g);
  else return -1;
  return 0;
}

extern "C" double harness_table_voltage(const double* t, const double* v, int len, int uniform,
                                        double t0, double inv_dt, double x, int* in_table) {
  ProtocolTable tab{t, v, len, uniform, t0, inv_dt};
  double out;
  *in_table = table_voltage(tab, x, &out) ? 1 : 0;
  return out;
}
