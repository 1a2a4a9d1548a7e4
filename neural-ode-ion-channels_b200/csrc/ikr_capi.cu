// ikr_capi.cu -- the C ABI (include/ikr.h) over the sm_100a kernels.  No torch types, no
// device allocation, no global state, stream-ordered, never synchronises (except ikr_fma_peak).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/ikr.h"
#include "ikr_backward.cuh"
#include "ikr_forward.cuh"

using namespace ikr;

namespace {

constexpr size_t kSmemLimit = 227 * 1024;
constexpr int kMaxThreads = 512;

struct Geometry {
  int npad, TN, NG, kc, cpl;
  int M, MG, threads, n_worker_warps;
  long long n_tiles;
  int grid;
  size_t smem;
  int sms;
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <typename S, typename W>
size_t fwd_smem(int M, int npad, int kc) { return FwdSmemLayout<S, W>(M, npad, kc).total; }

size_t fwd_smem_dyn(const ikr_desc* d, int M, int npad, int kc) {
  if (d->state_dtype == IKR_F32) return fwd_smem<float, float>(M, npad, kc);
  if (d->mlp_dtype == IKR_F32) return fwd_smem<double, float>(M, npad, kc);
  return fwd_smem<double, double>(M, npad, kc);
}

int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  return sms > 0 ? sms : 148;
}

bool valid_desc(const ikr_desc* d) {
  if (!d) return false;
  if (d->n_layers < 1 || d->n_nodes < 1 || d->n_nodes > 4096) return false;
  if (d->state_dtype != IKR_F32 && d->state_dtype != IKR_F64) return false;
  if (d->mlp_dtype != IKR_F32 && d->mlp_dtype != IKR_F64) return false;
  if (d->state_dtype == IKR_F32 && d->mlp_dtype == IKR_F64) return false;
  if (d->method != IKR_DOPRI5 && d->method != IKR_RK4) return false;
  return true;
}

void mlp_layout(const ikr_desc* d, int* npad, int* kc, int* cpl, long long off[5],
                long long* total) {
  const int n = d->n_nodes, L = d->n_layers;
  const int wsz = d->mlp_dtype == IKR_F32 ? 4 : 8;
  *npad = round_up(n, 8);
  int kc_max = (int)(20480 / ((long long)*npad * wsz));
  if (kc_max < 1) kc_max = 1;
  if (kc_max > n) kc_max = n;
  *cpl = (n + kc_max - 1) / kc_max;
  *kc = (n + *cpl - 1) / *cpl;
  long long o = 0;
  off[0] = o; o += 3LL * *npad;              // w0a | w0b | b0
  off[1] = o; o += (long long)L * n * *npad;  // Wt
  off[2] = o; o += (long long)L * *npad;      // bh
  off[3] = o; o += *npad + 8;                 // wl | bl (padded)
  off[4] = o; o += (long long)L * n * *npad;  // Wn (backward operand)
  *total = o;
}

// candidate trajectory-group counts MG (tile M = 8 MG): multiples of 4 keep the quarter-warp
// mapping bank-conflict free; 1..3 only serve tiny batches.
const int kMgCandidates[] = {1, 2, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64};

Geometry make_geometry(const ikr_desc* d, int n_jobs, const long long* B) {
  Geometry g;
  long long off[5], total;
  mlp_layout(d, &g.npad, &g.kc, &g.cpl, off, &total);
  g.TN = d->mlp_dtype == IKR_F32 ? 8 : 4;
  g.NG = g.npad / g.TN;
  g.sms = device_sms();
  long long b_total = 0;
  for (int j = 0; j < n_jobs; ++j) b_total += B[j];
  double best_score = -1.0;
  int best_mg = 1;
  const int forced = d->tile_m > 0 ? round_up(d->tile_m, 8) / 8 : 0;
  for (int mg : kMgCandidates) {
    const int M = 8 * mg;
    const int workers = tile_worker_threads(mg, g.NG);
    const int threads = round_up(workers > M ? workers : M, 32);
    if (threads > kMaxThreads) continue;
    if (fwd_smem_dyn(d, M, g.npad, g.kc) > kSmemLimit) continue;
    if (forced) {
      if (mg <= forced) best_mg = mg;       // largest feasible candidate not above the request
      continue;
    }
    long long tiles = 0;
    for (int j = 0; j < n_jobs; ++j) tiles += (B[j] + M - 1) / M;
    const double waves = (double)tiles / g.sms;
    const double wave_eff = waves / (double)((tiles + g.sms - 1) / g.sms);
    const double fill = (double)b_total / ((double)tiles * M);
    const int warps = (workers + 31) / 32;
    const double balance = (workers / 32.0) / (4.0 * ((warps + 3) / 4));
    const double amort = (double)M / (M + 6.0);   // per-evaluation owner-phase overhead
    const double score = wave_eff * fill * balance * amort;
    if (score > best_score) { best_score = score; best_mg = mg; }
  }
  g.MG = best_mg;
  g.M = 8 * best_mg;
  const int workers = tile_worker_threads(g.MG, g.NG);
  g.n_worker_warps = (workers + 31) / 32;
  g.threads = round_up(workers > g.M ? workers : g.M, 32);
  g.n_tiles = 0;
  for (int j = 0; j < n_jobs; ++j) g.n_tiles += (B[j] + g.M - 1) / g.M;
  g.grid = (int)(g.n_tiles < g.sms ? g.n_tiles : g.sms);
  if (g.grid < 1) g.grid = 1;
  g.smem = fwd_smem_dyn(d, g.M, g.npad, g.kc);
  return g;
}

MlpView make_view(const ikr_desc* d, const void* weights) {
  MlpView v;
  long long off[5], total;
  mlp_layout(d, &v.npad, &v.kc, &v.cpl, off, &total);
  v.base = weights;
  v.L = d->n_layers;
  v.n = d->n_nodes;
  v.off_w0 = off[0]; v.off_wt = off[1]; v.off_bh = off[2]; v.off_wl = off[3]; v.off_wn = off[4];
  v.slope = d->negative_slope;
  return v;
}

ProtocolTable make_table(const ikr_io* io) {
  ProtocolTable t;
  t.t = io->table_t; t.v = io->table_v; t.len = io->table_len; t.uniform = io->table_uniform;
  t.t0 = io->table_t0; t.inv_dt = io->table_inv_dt;
  return t;
}

SolverCfg make_cfg(const ikr_desc* d) {
  SolverCfg c;
  c.tab.t = nullptr; c.tab.v = nullptr; c.tab.len = 0; c.tab.uniform = 0;
  c.tab.t0 = 0; c.tab.inv_dt = 0;
  for (int i = 0; i < 8; ++i) c.hp.p[i] = d->p[i];
  c.ctl.safety = d->safety; c.ctl.ifactor = d->ifactor; c.ctl.dfactor = d->dfactor;
  c.vrange = d->vrange; c.netscale = d->netscale;
  c.rtol = d->rtol; c.atol = d->atol; c.first_step = d->first_step;
  c.max_num_steps = d->max_num_steps;
  c.nn_d = d->nn_d;
  c.mlp_is_f64 = d->mlp_dtype == IKR_F64;
  return c;
}

template <typename S, typename W>
int launch_forward(const FwdParams& p, const Geometry& g, cudaStream_t st) {
  auto kern = ikr_forward_kernel<S, W>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<g.grid, g.threads, g.smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

}  // namespace

extern "C" {

int ikr_abi_version(void) { return IKR_ABI_VERSION; }

const char* ikr_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case IKR_ERR_ARG: return "invalid argument";
    case IKR_ERR_UNSUPPORTED: return "unsupported configuration";
    case IKR_ERR_WORKSPACE: return "workspace too small";
    case IKR_ERR_LAUNCH: return "kernel launch failed";
    case IKR_ERR_DEVICE: return "CUDA device error";
    default: return "unknown error";
  }
}

int64_t ikr_packed_weight_elems(const ikr_desc* d) {
  if (!valid_desc(d)) return IKR_ERR_ARG;
  int npad, kc, cpl;
  long long off[5], total;
  mlp_layout(d, &npad, &kc, &cpl, off, &total);
  return total;
}

int ikr_packed_layout(const ikr_desc* d, int64_t out[8]) {
  if (!valid_desc(d) || !out) return IKR_ERR_ARG;
  int npad, kc, cpl;
  long long off[5], total;
  mlp_layout(d, &npad, &kc, &cpl, off, &total);
  out[0] = npad;
  for (int i = 0; i < 5; ++i) out[1 + i] = off[i];
  out[6] = total;
  out[7] = kc;
  return 0;
}

int64_t ikr_param_count(const ikr_desc* d) {
  if (!valid_desc(d)) return IKR_ERR_ARG;
  const long long n = d->n_nodes, L = d->n_layers;
  return 2 * n + n + L * (n * n + n) + n + 1;
}

int32_t ikr_tile_m(const ikr_desc* d, int32_t n_jobs, const int64_t* B) {
  if (!valid_desc(d) || n_jobs < 1 || !B) return IKR_ERR_ARG;
  for (int j = 0; j < n_jobs; ++j) if (B[j] < 1) return IKR_ERR_ARG;
  return make_geometry(d, n_jobs, (const long long*)B).M;
}

int ikr_launch_geometry(const ikr_desc* d, int32_t n_jobs, const int64_t* B, int64_t out[8]) {
  if (!valid_desc(d) || n_jobs < 1 || !B || !out) return IKR_ERR_ARG;
  for (int j = 0; j < n_jobs; ++j) if (B[j] < 1) return IKR_ERR_ARG;
  Geometry g = make_geometry(d, n_jobs, (const long long*)B);
  out[0] = g.M; out[1] = g.threads; out[2] = g.grid; out[3] = (int64_t)g.smem;
  out[4] = g.n_tiles; out[5] = g.kc; out[6] = g.cpl; out[7] = g.sms;
  return 0;
}

size_t ikr_workspace_bytes(const ikr_desc* d, int32_t n_jobs, int64_t B_total,
                           int32_t with_backward) {
  if (!valid_desc(d) || n_jobs < 1 || B_total < 1) return 0;
  size_t bytes = 256 + (size_t)n_jobs * sizeof(FwdJob);
  bytes = (bytes + 255) & ~(size_t)255;
  if (with_backward) bytes += bwd_workspace_bytes(d, B_total);
  return bytes;
}

int ikr_forward(const ikr_desc* d, const ikr_io* jobs, int32_t n_jobs, void* workspace,
                size_t workspace_bytes, void* cuda_stream) {
  if (!valid_desc(d) || !jobs || n_jobs < 1 || n_jobs > 4096) return IKR_ERR_ARG;
  for (int j = 0; j < n_jobs; ++j) {
    const ikr_io* io = &jobs[j];
    if (io->B < 1 || io->T < 1 || !io->weights || !io->table_t || !io->table_v || !io->y0 ||
        !io->t_out || !io->stats_out || io->table_len < 2)
      return IKR_ERR_ARG;
    if (d->method == IKR_RK4 && (!io->grid || io->G < 1)) return IKR_ERR_ARG;
    if ((io->i_out || io->loss_out) && !io->v_out) return IKR_ERR_ARG;
    if (io->data && io->data_B != 1 && io->data_B != io->B) return IKR_ERR_ARG;
    if (io->ckpt_t && (!io->ckpt_y || io->ckpt_cap < 1)) return IKR_ERR_ARG;
    if (io->T > 2147483647LL || io->G > 2147483647LL) return IKR_ERR_ARG;
    if (io->weights != jobs[0].weights) return IKR_ERR_ARG;
  }
  const size_t need = 256 + (size_t)n_jobs * sizeof(FwdJob);
  if (!workspace || workspace_bytes < need) return IKR_ERR_WORKSPACE;

  // longest jobs first (LPT) so that the dynamic tile queue balances the SMs
  std::vector<int> order(n_jobs);
  for (int j = 0; j < n_jobs; ++j) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const double ca = jobs[a].cost_hint > 0 ? jobs[a].cost_hint : (double)jobs[a].T;
    const double cb = jobs[b].cost_hint > 0 ? jobs[b].cost_hint : (double)jobs[b].T;
    return ca > cb;
  });
  std::vector<long long> Bs(n_jobs);
  for (int j = 0; j < n_jobs; ++j) Bs[j] = jobs[order[j]].B;
  Geometry g = make_geometry(d, n_jobs, Bs.data());
  if (g.smem > kSmemLimit || g.threads > kMaxThreads) return IKR_ERR_UNSUPPORTED;

  std::vector<FwdJob> table(n_jobs);
  long long tile = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const ikr_io* io = &jobs[order[j]];
    FwdJob& fj = table[j];
    fj.tab = make_table(io);
    fj.B = io->B; fj.T = (int)io->T; fj.G = (int)io->G;
    fj.tile_begin = tile;
    tile += (io->B + g.M - 1) / g.M;
    fj.y0 = io->y0; fj.t_out = io->t_out; fj.grid = io->grid; fj.v_out = io->v_out;
    fj.g = io->g; fj.e_rev = io->e_rev; fj.e_scalar = io->e_scalar;
    fj.data = io->data; fj.data_B = io->data_B;
    fj.y_out = io->y_out; fj.i_out = io->i_out; fj.loss_out = io->loss_out;
    fj.stats_out = io->stats_out;
    fj.ckpt_cap = io->ckpt_cap; fj.ckpt_t = io->ckpt_t; fj.ckpt_y = io->ckpt_y;
  }

  cudaStream_t st = (cudaStream_t)cuda_stream;
  unsigned char* ws = (unsigned char*)workspace;
  if (cudaMemsetAsync(ws, 0, 256, st) != cudaSuccess) return IKR_ERR_DEVICE;
  // pageable source: the runtime stages the bytes before returning, `table` may die afterwards
  if (cudaMemcpyAsync(ws + 256, table.data(), (size_t)n_jobs * sizeof(FwdJob),
                      cudaMemcpyHostToDevice, st) != cudaSuccess)
    return IKR_ERR_DEVICE;

  FwdParams p;
  p.mlp = make_view(d, jobs[0].weights);
  p.cfg = make_cfg(d);
  p.method = d->method;
  p.time_f32 = d->time_f32;
  p.rk4_perturb = d->rk4_perturb;
  p.M = g.M; p.MG = g.MG; p.NG = g.NG;
  p.n_worker_warps = g.n_worker_warps;
  p.n_jobs = n_jobs;
  p.n_tiles = g.n_tiles;
  p.jobs = reinterpret_cast<const FwdJob*>(ws + 256);
  p.queue = reinterpret_cast<unsigned long long*>(ws);

  if (d->state_dtype == IKR_F32) return launch_forward<float, float>(p, g, st);
  if (d->mlp_dtype == IKR_F32) return launch_forward<double, float>(p, g, st);
  return launch_forward<double, double>(p, g, st);
}

int ikr_backward(const ikr_desc* d, const ikr_io* io, const ikr_bwd_io* bio, void* workspace,
                 size_t workspace_bytes, void* cuda_stream) {
  if (!valid_desc(d) || !io || !bio) return IKR_ERR_ARG;
  return bwd_dispatch(d, io, bio, workspace, workspace_bytes, (cudaStream_t)cuda_stream);
}

int ikr_interp_protocol(const ikr_io* table, const double* t_query, int64_t T, double* v_out,
                        void* cuda_stream) {
  if (!table || !table->table_t || !table->table_v || !t_query || !v_out || T < 1 ||
      table->table_len < 2)
    return IKR_ERR_ARG;
  ProtocolTable tab = make_table(table);
  int threads = 256;
  long long blocks = (T + threads - 1) / threads;
  ikr_interp_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)cuda_stream>>>(tab, t_query, T,
                                                                               v_out);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

int ikr_fma_peak(int32_t dtype, int64_t iters, double* tflops_out, void* cuda_stream) {
  if (!tflops_out || iters < 1) return IKR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int sms = device_sms();
  void* sink = nullptr;
  // the sink is only written if an impossible value appears; use a static device symbol-free
  // approach: pass nullptr-guarded pointer from a tiny pinned host allocation is overkill, so
  // the caller-visible contract is: this entry point allocates nothing and the kernel never
  // stores (the comparison is against a value the recurrence cannot produce).
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
    return IKR_ERR_DEVICE;
  const int threads = 512, blocks = sms * 4;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0, st);
    if (dtype == IKR_F32)
      ikr_fma_peak_kernel<float><<<blocks, threads, 0, st>>>((float*)sink, iters);
    else
      ikr_fma_peak_kernel<double><<<blocks, threads, 0, st>>>((double*)sink, iters);
    cudaEventRecord(e1, st);
  }
  if (cudaEventSynchronize(e1) != cudaSuccess) {
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return IKR_ERR_DEVICE;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double flops = 2.0 * 16.0 * (double)iters * threads * (double)blocks;
  *tflops_out = flops / (ms * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
