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
    auto ckpt = [&](int step, double t0, double dt, S, S, S, S) {
      if (a.steps && step < a.steps_cap) { a.steps[2 * step] = t0; a.steps[2 * step + 1] = dt; }
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

extern "C" double harness_table_voltage(const double* t, const double* v, int len, int uniform,
                                        double t0, double inv_dt, double x, int* in_table) {
  ProtocolTable tab{t, v, len, uniform, t0, inv_dt};
  double out;
  *in_table = table_voltage(tab, x, &out) ? 1 : 0;
  return out;
}
