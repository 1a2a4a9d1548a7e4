// ikr_capi.cu -- the C ABI (include/ikr.h) over the sm_100a kernels.  No torch types, no
// device allocation, no global state, stream-ordered, never synchronises (except ikr_fma_peak).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/ikr.h"
#include "ikr_backward.cuh"
#include "ikr_regress_tc.cuh"
#include "ikr_forward_tc_pp.cuh"
#include "ikr_hh.cuh"
#include "ikr_markov.cuh"

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
size_t fwd_smem(int M, int npad, int kc, int L) { return FwdSmemLayout<S, W>(M, npad, kc, L).total; }

size_t fwd_smem_dyn(const ikr_desc* d, int M, int npad, int kc) {
  if (d->state_dtype == IKR_F32) return fwd_smem<float, float>(M, npad, kc, d->n_layers);
  if (d->mlp_dtype == IKR_F32) return fwd_smem<double, float>(M, npad, kc, d->n_layers);
  return fwd_smem<double, double>(M, npad, kc, d->n_layers);
}

int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  return sms > 0 ? sms : 148;
}

// Dynamic shared memory limit of a kernel.  The attribute belongs to the kernel, i.e. to every host
// thread of the process: it is always raised to the device's opt-in maximum, never to the size of
// the launch at hand, so that concurrent callers with different geometries (one fit per thread and
// stream, bench.py `sweep`) cannot undercut each other between the attribute call and the launch.
template <typename K>
cudaError_t allow_max_smem(K kern, size_t needed) {
  int dev = 0, optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
    return cudaErrorUnknown;
  if (needed > (size_t)optin) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
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

// Relative SM throughput of the tile GEMM as a function of the worker warps per CTA (measured on
// B200: 4 warps = 1 per scheduler reach ~43 % of the 16-warp rate; latency hiding needs >= 3 per
// scheduler), times the balance of the warps over the four schedulers.
double warp_throughput(int workers) {
  const int warps = (workers + 31) / 32;
  const int per_sched = (warps + 3) / 4;
  static const double kRate[5] = {0.0, 0.43, 0.70, 0.88, 1.0};
  const double rate = kRate[per_sched > 4 ? 4 : per_sched];
  return rate * (workers / 32.0) / (4.0 * per_sched);
}

// candidate trajectory-group counts MG (tile M = 8 MG): multiples of 4 keep the quarter-warp
// mapping bank-conflict free; 1..3 only serve tiny batches.
const int kMgCandidates[] = {1, 2, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64};

// The register tile is 8 trajectories x TN features per thread.  The wide tile (TN = 8 fp32 / 4
// fp64) has the best FMA : LDS ratio; the narrow tile (TN = 4 / 2) doubles the worker threads of
// a small trajectory tile so that small batches still put >= 2 warps on every scheduler.
constexpr double kNarrowTilePenalty = 0.85;

template <typename SmemFn>
Geometry search_geometry(const ikr_desc* d, int n_jobs, const long long* B, bool pool,
                         SmemFn smem_of) {
  Geometry g;
  long long off[5], total;
  mlp_layout(d, &g.npad, &g.kc, &g.cpl, off, &total);
  const int tn_wide = d->mlp_dtype == IKR_F32 ? 8 : 4;
  g.sms = device_sms();
  long long b_total = 0;
  for (int j = 0; j < n_jobs; ++j) b_total += B[j];
  double best_score = -1.0;
  int best_mg = 1, best_tn = tn_wide;
  const int forced = d->tile_m > 0 ? round_up(d->tile_m, 8) / 8 : 0;
  for (int tn = tn_wide; tn >= tn_wide / 2; tn /= 2) {
    const int ng = g.npad / tn;
    for (int mg : kMgCandidates) {
      const int M = 8 * mg;
      const int workers = tile_worker_threads(mg, ng);
      const int threads = round_up(workers > M ? workers : M, 32);
      if (threads > kMaxThreads) continue;
      if (smem_of(M) > kSmemLimit) continue;
      if (forced) {
        // largest feasible candidate not above the request (wide tile only)
        if (tn == tn_wide && mg <= forced) { best_mg = mg; best_tn = tn; }
        continue;
      }
      long long tiles = 0;
      for (int j = 0; j < n_jobs; ++j) tiles += (B[j] + M - 1) / M;
      double wave_eff, fill;
      if (pool) {
        // lane-pool kernel: slots refill from one queue, so only an under-filled GPU costs
        tiles = (b_total + M - 1) / M;
        const long long ctas = tiles < g.sms ? tiles : g.sms;
        wave_eff = (double)ctas / g.sms;
        fill = b_total >= ctas * (long long)M ? 1.0 : (double)b_total / ((double)ctas * M);
      } else {
        const double waves = (double)tiles / g.sms;
        wave_eff = waves / (double)((tiles + g.sms - 1) / g.sms);
        fill = (double)b_total / ((double)tiles * M);
      }
      const double amort = (double)M / (M + 6.0);   // per-evaluation owner-phase overhead
      const double score = wave_eff * fill * warp_throughput(workers) * amort *
                           (tn == tn_wide ? 1.0 : kNarrowTilePenalty);
      if (score > best_score) { best_score = score; best_mg = mg; best_tn = tn; }
    }
  }
  g.TN = best_tn;
  g.NG = g.npad / g.TN;
  g.MG = best_mg;
  g.M = 8 * best_mg;
  const int workers = tile_worker_threads(g.MG, g.NG);
  g.n_worker_warps = (workers + 31) / 32;
  g.threads = round_up(workers > g.M ? workers : g.M, 32);
  g.n_tiles = 0;
  if (pool) g.n_tiles = (b_total + g.M - 1) / g.M;
  else for (int j = 0; j < n_jobs; ++j) g.n_tiles += (B[j] + g.M - 1) / g.M;
  g.grid = (int)(g.n_tiles < g.sms ? g.n_tiles : g.sms);
  if (g.grid < 1) g.grid = 1;
  g.smem = smem_of(g.M);
  return g;
}

Geometry make_geometry(const ikr_desc* d, int n_jobs, const long long* B, bool pool = false) {
  long long off[5], total;
  int npad, kc, cpl;
  mlp_layout(d, &npad, &kc, &cpl, off, &total);
  return search_geometry(d, n_jobs, B, pool,
                         [&](int M) { return fwd_smem_dyn(d, M, npad, kc); });
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
  v.bwd_seq = 0;
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

// dopri5 runs tile-scheduled by default; desc.reserved bit 0 selects the lane-pool kernel (slots
// refill from one trajectory queue).  Measured on B200 (profiles/r1_forward_v3_summary.md): the
// pool wins when one long job has many more trajectories than lane slots (+3.5 % at 65,536 x pr4)
// and loses ~3 % on the five-protocol bench mix, so it is opt-in.  rk4 is always tile-scheduled.
bool use_pool(const ikr_desc* d) { return d->method == IKR_DOPRI5 && (d->reserved & 1); }

// Tensor-core forward path (ikr_forward_tc.cuh): fp32 MLP whose hidden width fits the TMEM budget
// (n <= 200 covers every shipped model; s06-s08, n = 500, stay on the FFMA2 kernel).  desc.reserved
// bit 1 opts out (A/B measurements, FFMA-only parity runs).
struct TcPlan {
  bool ok;
  TcGeom g;
  int groups;            // threads sharing one TMEM lane (epilogue column groups)
  size_t smem, img_bytes;
};

constexpr int kTcDefaultGroups = 2;   // measured: 2 and 3 tie on full grids, 2 is ~3 % ahead on one-tile-per-SM launches (no register spills at 161-168 registers)

// Operand split of the tensor-core FORWARD kernels: fp16x2 (three MMAs per fp32 product) unless
// desc.reserved bit 8 asks for bf16x3 (six; no range restriction on the activations).  The adjoint /
// regression kernels always use bf16x3 (gradients span too many decades for fp16).
int tc_forward_terms(const ikr_desc* d) { return (d->reserved & 256) ? 3 : 2; }

TcPlan make_tc_plan(const ikr_desc* d, int terms = 3) {
  TcPlan t;
  t.ok = false;
  t.smem = 0; t.img_bytes = 0;
  t.g = tc_geometry(d->n_nodes, d->n_layers, terms);
  if (d->mlp_dtype != IKR_F32 || (d->reserved & 2) || d->tile_m > 0) return t;
  if (!(d->negative_slope >= 0.0 && d->negative_slope <= 1.0)) return t;   // epilogues use max(z, slope z)
  if (!tc_geometry_ok(t.g)) return t;
  t.groups = kTcDefaultGroups;
  {
    // desc.reserved bits 4-5: column-group override for tuning / A-B runs (0 = library default).
    // NOTE: the group count fixes the summation order of the output layer, i.e. results.
    const int v = (d->reserved >> 4) & 3;
    if (v >= 1 && v <= 3) t.groups = v;
  }
  // every column group must produce at least one unit per layer pass (it then consumes every phase
  // of the pass barriers in order)
  if (t.groups > t.g.units + t.g.tail) t.groups = t.g.units + t.g.tail;
  const size_t fixed = d->state_dtype == IKR_F32 ? TcSmemLayout<float>(t.g, 0, t.groups).total
                                                  : TcSmemLayout<double>(t.g, 0, t.groups).total;
  if (fixed + (size_t)kTcMinStages * t.g.stage_bytes > kSmemLimit) return t;
  int stages = (int)((kSmemLimit - fixed) / t.g.stage_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  t.g.stages = stages;
  t.smem = fixed + (size_t)stages * t.g.stage_bytes;
  // + 512 bytes behind the blocks: per-layer |W| maxima and accumulator scales of the fp16x2 image
  t.img_bytes = (((size_t)d->n_layers * t.g.KST * t.g.stage_bytes + 255) & ~(size_t)255) + 512;
  t.ok = d->n_layers <= 64;
  return t;
}

// Trajectories per tile of the tensor-core kernels.  Always the 128 TMEM lanes: spreading a small
// batch over more, thinner tiles was measured (B = 4,096 as 128 tiles of 32) and buys nothing -- a
// tile's run time is its chain of sequential RHS evaluations (~31 us each) whatever its lane count,
// and every tile already has an SM to itself -- while the backward stash grows with the tile count.
int tc_tile_lanes(long long /*b_total*/, int /*sms*/) { return kTcM; }

// Lane-pool scheduling of the tensor-core forward kernel.  desc.reserved bit 0 forces it, bit 2 forbids
// it; otherwise it is used when the launch holds more 128-trajectory tiles than SMs (measured on
// B200: +7 % at 65,536 x pr4, +2.4 % on the five-protocol bench mix, +5.7 / +4.2 / +2.9 % at 1.125 /
// 1.25 / 1.5 waves of pr4, -3 % when every SM gets exactly one tile).
bool use_pool_tc(const ikr_desc* d, long long b_total, int sms) {
  if (d->method != IKR_DOPRI5 || (d->reserved & 4)) return false;
  if (d->reserved & 1) return true;
  return b_total > (long long)kTcM * sms;   // more than one tile per SM
}

// Two-tile ping-pong lane pool (ikr_forward_tc_pp.cuh): dopri5 on the tensor cores, two tiles per CTA
// whose evaluations alternate on the tensor pipe.  EXPERIMENTAL in this round: bit-identical to the
// two-group tile kernel in the tests, but its first pass (layer 0 by one column group) does not
// keep up with the MMAs yet, so it is opt-in only (desc.reserved bit 7), never chosen automatically.
bool use_pp_tc(const ikr_desc* d, long long /*b_total*/, int /*sms*/, int /*n_jobs*/) {
  if (d->method != IKR_DOPRI5 || (d->reserved & (64 | 4))) return false;
  return (d->reserved & 128) != 0;
}

struct TcPpPlan {
  bool ok;
  TcGeom g;
  size_t smem;
};
TcPpPlan make_tc_pp_plan(const ikr_desc* d, const TcPlan& fw) {
  TcPpPlan t;
  t.ok = false;
  t.g = fw.g;
  t.smem = 0;
  if (!fw.ok || fw.g.units + fw.g.tail < 2) return t;   // two column groups need two units per pass
  const size_t fixed = d->state_dtype == IKR_F32 ? TcPpSmemLayout<float>(t.g, 0).total
                                                  : TcPpSmemLayout<double>(t.g, 0).total;
  if (fixed + (size_t)kTcMinStages * t.g.stage_bytes > kSmemLimit) return t;
  int stages = (int)((kSmemLimit - fixed) / t.g.stage_bytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  t.g.stages = stages;
  t.smem = fixed + (size_t)stages * t.g.stage_bytes;
  t.ok = true;
  return t;
}

template <typename S>
int launch_forward_tc_pp(const TcFwdParams& tp, const TcPpPlan& t, int grid, cudaStream_t st) {
  auto kern = tp.g.terms == 2 ? ikr_forward_tc_pp_kernel<S, 2> : ikr_forward_tc_pp_kernel<S, 3>;
  if (allow_max_smem(kern, t.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<grid, tc_threads(3), t.smem, st>>>(tp);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

size_t fwd_fixed_workspace(int n_jobs) {
  return (256 + (size_t)n_jobs * sizeof(FwdJob) + 255) & ~(size_t)255;
}

template <typename S, int G, int TERMS>
int launch_forward_tc_g(const TcFwdParams& tp, const TcPlan& t, int grid, cudaStream_t st, bool pool) {
  auto kern = pool ? ikr_forward_tc_pool_kernel<S, G, TERMS> : ikr_forward_tc_kernel<S, G, TERMS>;
  if (allow_max_smem(kern, t.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<grid, tc_threads(G), t.smem, st>>>(tp);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

template <typename S, int TERMS>
int launch_forward_tc_t(const TcFwdParams& tp, const TcPlan& t, int grid, cudaStream_t st, bool pool) {
  if (t.groups == 1) return launch_forward_tc_g<S, 1, TERMS>(tp, t, grid, st, pool);
  if (t.groups == 2) return launch_forward_tc_g<S, 2, TERMS>(tp, t, grid, st, pool);
  return launch_forward_tc_g<S, 3, TERMS>(tp, t, grid, st, pool);
}

template <typename S>
int launch_forward_tc(const TcFwdParams& tp, const TcPlan& t, int grid, cudaStream_t st, bool pool) {
  if (t.g.terms == 2) return launch_forward_tc_t<S, 2>(tp, t, grid, st, pool);
  return launch_forward_tc_t<S, 3>(tp, t, grid, st, pool);
}

template <typename S, typename W, int TN>
int launch_forward_tn(const FwdParams& p, const Geometry& g, cudaStream_t st, bool pool) {
  auto kern = pool ? ikr_forward_pool_kernel<S, W, TN> : ikr_forward_kernel<S, W, TN>;
  if (allow_max_smem(kern, g.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<g.grid, g.threads, g.smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

template <typename S, typename W>
int launch_forward(const FwdParams& p, const Geometry& g, cudaStream_t st, bool pool) {
  constexpr int wide = MlpTileCfg<W>::TN;
  if (g.TN == wide) return launch_forward_tn<S, W, wide>(p, g, st, pool);
  return launch_forward_tn<S, W, wide / 2>(p, g, st, pool);
}


// ---------------------------------------------------------------------------------------------
// backward: geometry, workspace plan, round loop
// ---------------------------------------------------------------------------------------------
template <typename S, typename W>
size_t adj_smem(int M, int npad, int kc, int L) { return AdjSmemLayout<S, W>(M, npad, kc, L).total; }

size_t adj_smem_dyn(const ikr_desc* d, int M, int npad, int kc) {
  if (d->state_dtype == IKR_F32) return adj_smem<float, float>(M, npad, kc, d->n_layers);
  if (d->mlp_dtype == IKR_F32) return adj_smem<double, float>(M, npad, kc, d->n_layers);
  return adj_smem<double, double>(M, npad, kc, d->n_layers);
}

Geometry make_geometry_bwd(const ikr_desc* d, long long B) {
  long long off[5], total;
  int npad, kc, cpl;
  mlp_layout(d, &npad, &kc, &cpl, off, &total);
  return search_geometry(d, 1, &B, false, [&](int M) { return adj_smem_dyn(d, M, npad, kc); });
}

int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

struct BwdPlan {
  Geometry g;
  // weight-gradient GEMM
  int KC, n_ot, n_it, BO, BI, S, IG, OG, wg_threads, wg_stages, wg_grid;
  size_t wg_smem;
  // workspace carve-up (byte offsets)
  size_t off_counters, off_lanes, off_small, off_pw, off_pb, off_stash_h, off_stash_d, fixed_bytes;
  size_t slot_bytes;   // one stash slot, ONE of the two stashes
  int small_stride;
  size_t zero_begin, zero_bytes;  // accumulators that must start at zero
};

BwdPlan make_bwd_plan(const ikr_desc* d, long long B) {
  BwdPlan pl;
  pl.g = make_geometry_bwd(d, B);
  const Geometry& g = pl.g;
  const int wsz = d->mlp_dtype == IKR_F32 ? 4 : 8;
  const int ssz = d->state_dtype == IKR_F32 ? 4 : 8;
  const int L = d->n_layers;
  const int RI = wsz == 4 ? 8 : 4;
  // K-chunk of the GEMM: a divisor of M (chunks never straddle stash slots)
  pl.KC = 8 * gcd_int(g.MG, 4);
  while (pl.KC > 8 && 2ULL * pl.KC * g.npad * wsz * 2 > 200 * 1024) pl.KC /= 2;
  pl.n_ot = (g.npad + 199) / 200;
  pl.BO = round_up((g.npad + pl.n_ot - 1) / pl.n_ot, 8);
  pl.OG = pl.BO / 8;
  const int ig_max = kWgMaxThreads / pl.OG;
  pl.n_it = (g.npad + ig_max * RI - 1) / (ig_max * RI);
  pl.BI = round_up((g.npad + pl.n_it - 1) / pl.n_it, 8);
  pl.IG = (pl.BI + RI - 1) / RI;
  pl.wg_threads = round_up(pl.OG * pl.IG, 32);
  if (pl.wg_threads > kWgMaxThreads) pl.wg_threads = kWgMaxThreads;
  pl.S = g.sms / (L * pl.n_ot * pl.n_it);
  if (pl.S < 1) pl.S = 1;
  pl.wg_grid = L * pl.n_ot * pl.n_it * pl.S;
  const size_t stage = 2ULL * pl.KC * g.npad * wsz;
  pl.wg_stages = (int)((200 * 1024) / stage);
  if (pl.wg_stages > 6) pl.wg_stages = 6;
  if (pl.wg_stages < 2) pl.wg_stages = 2;
  pl.wg_smem = 128 + (size_t)pl.wg_stages * stage + (size_t)g.npad * wsz;   // + prefetch pad row

  pl.small_stride = 4 * g.npad + 8;
  const size_t lane_save = ssz == 4 ? sizeof(BLaneSave<float>) : sizeof(BLaneSave<double>);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  pl.off_counters = take(256);
  pl.off_lanes = take((size_t)g.n_tiles * g.M * lane_save);
  pl.off_small = take((size_t)g.sms * pl.small_stride * sizeof(double));
  pl.off_pw = take((size_t)L * pl.S * g.npad * g.npad * sizeof(double));
  pl.off_pb = take((size_t)L * pl.S * g.npad * sizeof(double));
  pl.zero_begin = pl.off_small;
  pl.zero_bytes = o - pl.off_small;
  pl.fixed_bytes = o;
  pl.slot_bytes = (size_t)L * g.M * g.npad * wsz;
  return pl;
}

// steps per round a workspace of `bytes` can hold (0: too small)
int bwd_steps_per_round(const BwdPlan& pl, size_t bytes) {
  if (bytes <= pl.fixed_bytes + 512) return 0;
  const size_t avail = (bytes - pl.fixed_bytes - 512) / 2;       // per stash
  const long long slots = (long long)(avail / pl.slot_bytes);
  const long long per_tile = slots / pl.g.n_tiles;               // 6 R + 1 slots per tile
  if (per_tile < 7) return 0;
  long long R = (per_tile - 1) / 6;
  return (int)(R > 4096 ? 4096 : R);
}

size_t bwd_workspace_bytes(const ikr_desc* d, long long B, int gib = 4) {
  const BwdPlan pl = make_bwd_plan(d, B);
  // a stash of about `gib` GiB (both halves; default 4), at least one step per round
  const double target = (double)gib * 1024 * 1024 * 1024;
  long long R = (long long)((target / 2 / (double)pl.slot_bytes / (double)pl.g.n_tiles - 1) / 6);
  if (R < 1) R = 1;
  if (R > 64) R = 64;
  return pl.fixed_bytes + 512 + 2 * (size_t)pl.g.n_tiles * (6 * R + 1) * pl.slot_bytes;
}

template <typename S, typename W, int TN>
int launch_adjoint_tn(const BwdParams& p, const Geometry& g, cudaStream_t st) {
  auto kern = ikr_adjoint_kernel<S, W, TN>;
  if (allow_max_smem(kern, g.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<g.grid, g.threads, g.smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

template <typename S, typename W>
int launch_adjoint(const BwdParams& p, const Geometry& g, cudaStream_t st) {
  constexpr int wide = MlpTileCfg<W>::TN;
  if (g.TN == wide) return launch_adjoint_tn<S, W, wide>(p, g, st);
  return launch_adjoint_tn<S, W, wide / 2>(p, g, st);
}

template <typename W>
int launch_wgrad(const WgradParams& p, const BwdPlan& pl, cudaStream_t st) {
  auto kern = ikr_wgrad_kernel<W>;
  if (allow_max_smem(kern, pl.wg_smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<pl.wg_grid, pl.wg_threads, pl.wg_smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

// ---------------------------------------------------------------------------------------------
// tensor-core backward (ikr_backward_tc.cuh): plan, workspace, round loop
// ---------------------------------------------------------------------------------------------
struct TcBwdPlan {
  bool ok;
  TcGeom g;
  TcStashGeom sg;
  int groups, mask_words;
  size_t smem;
  long long n_tiles;
  int grid, sms, tile_lanes;
  int wg_S, wg_stages;
  size_t wg_smem;
  size_t off_counters, off_lanes, off_partial, off_img, fixed_bytes, partial_bytes, img_bytes;
};

// desc.reserved bit 10: products per fp32 product in the adjoint kernel's MMA passes.  Default
// (0): three of the six bf16x3 products (a1 b1 + a2 b1 + a1 b2, ~2^-16 per product, rounded splits:
// unbiased) -- the precision the weight-gradient GEMM has anyway, which is what bounds the gradient
// (measured on identical checkpoints: same deviation from the fp32-FMA backward in both modes,
// profiles/adjoint_products_accuracy.py).  1: all six.
int tc_adjoint_lite(const ikr_desc* d) { return (d->reserved & 1024u) ? 0 : 2; }

TcBwdPlan make_tc_bwd_plan(const ikr_desc* d, long long B) {
  TcBwdPlan pl;
  pl.ok = false;
  const TcPlan fw = make_tc_plan(d);
  if (!fw.ok || !tc_backward_ok(fw.g)) return pl;
  pl.g = fw.g;
  pl.sg = tc_stash_geometry(pl.g);
  pl.groups = fw.groups;
  pl.mask_words = (pl.g.units + pl.groups - 1) / pl.groups + 1;
  const size_t fixed = d->state_dtype == IKR_F32
                           ? TcAdjSmemLayout<float>(pl.g, 0, pl.groups, pl.mask_words).total
                           : TcAdjSmemLayout<double>(pl.g, 0, pl.groups, pl.mask_words).total;
  // three-product passes never read the third term block of a stage: it is neither streamed nor
  // given ring space, so more k-steps of weights are in flight
  if (tc_adjoint_lite(d) == 2) pl.g.ring_stride = 2 * pl.g.block_bytes;
  if (fixed + (size_t)kTcMinStages * pl.g.ring_stride > kSmemLimit) return pl;
  int stages = (int)((kSmemLimit - fixed) / pl.g.ring_stride);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  pl.g.stages = stages;
  pl.smem = fixed + (size_t)stages * pl.g.ring_stride;
  pl.sms = device_sms();
  pl.tile_lanes = tc_tile_lanes(B, pl.sms);
  pl.n_tiles = (B + pl.tile_lanes - 1) / pl.tile_lanes;
  pl.grid = (int)(pl.n_tiles < pl.sms ? pl.n_tiles : pl.sms);
  // weight-gradient GEMM: (L + 2) pseudo-layers x S splits
  pl.wg_S = pl.sms / (d->n_layers + 2);
  if (pl.wg_S < 1) pl.wg_S = 1;
  const size_t wg_stage = 4ull * pl.sg.NGb * 128 * 2;   // largest stage: two big operands, two terms each
  pl.wg_stages = (int)((kSmemLimit - 256) / wg_stage);
  if (pl.wg_stages > kWgTcMaxStages) pl.wg_stages = kWgTcMaxStages;
  if (pl.wg_stages < 2) return pl;
  pl.wg_smem = 256 + (size_t)pl.wg_stages * wg_stage;
  const size_t lane_save = d->state_dtype == IKR_F32 ? sizeof(BLaneSave<float>) : sizeof(BLaneSave<double>);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  pl.off_counters = take(256);
  pl.off_lanes = take((size_t)pl.n_tiles * kTcM * lane_save);
  pl.partial_bytes = (size_t)(d->n_layers + 2) * pl.wg_S * pl.g.NP * pl.g.NP * sizeof(double);
  pl.off_partial = take(pl.partial_bytes);
  pl.img_bytes = (size_t)2 * d->n_layers * pl.g.KST * pl.g.stage_bytes;
  pl.off_img = take(pl.img_bytes);
  pl.fixed_bytes = o;
  pl.ok = true;
  return pl;
}

int tc_bwd_steps_per_round(const TcBwdPlan& pl, size_t bytes) {
  if (bytes <= pl.fixed_bytes + 512) return 0;
  const long long slots = (long long)((bytes - pl.fixed_bytes - 512) / (size_t)pl.sg.slot);
  const long long per_tile = slots / pl.n_tiles;               // 6 R + 1 slots per tile
  if (per_tile < 7) return 0;
  const long long R = (per_tile - 1) / 6;
  return (int)(R > 4096 ? 4096 : R);
}

size_t tc_bwd_workspace_bytes(const TcBwdPlan& pl, int gib = 4) {
  // a stash of about `gib` GiB (default 4), at least one step per round
  const double target = (double)gib * 1024 * 1024 * 1024;
  long long R = (long long)((target / (double)pl.sg.slot / (double)pl.n_tiles - 1) / 6);
  if (R < 1) R = 1;
  if (R > 64) R = 64;
  return pl.fixed_bytes + 512 + (size_t)pl.n_tiles * (6 * R + 1) * (size_t)pl.sg.slot;
}

template <typename S, int G, int LITE>
int launch_adjoint_tc_gl(const TcAdjParams& tp, const TcBwdPlan& pl, cudaStream_t st) {
  auto kern = ikr_adjoint_tc_kernel<S, G, LITE>;
  if (allow_max_smem(kern, pl.smem) != cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<pl.grid, tc_threads(G), pl.smem, st>>>(tp);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}
template <typename S, int G>
int launch_adjoint_tc_g(const TcAdjParams& tp, const TcBwdPlan& pl, int lite, cudaStream_t st) {
  if (lite == 2) return launch_adjoint_tc_gl<S, G, 2>(tp, pl, st);
  return launch_adjoint_tc_gl<S, G, 0>(tp, pl, st);
}
template <typename S>
int launch_adjoint_tc(const TcAdjParams& tp, const TcBwdPlan& pl, int lite, cudaStream_t st) {
  if (pl.groups == 1) return launch_adjoint_tc_g<S, 1>(tp, pl, lite, st);
  if (pl.groups == 2) return launch_adjoint_tc_g<S, 2>(tp, pl, lite, st);
  return launch_adjoint_tc_g<S, 3>(tp, pl, lite, st);
}

int bwd_dispatch_tc(const ikr_desc* d, const ikr_io* io, const ikr_bwd_io* bio, const TcBwdPlan& pl,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int R = tc_bwd_steps_per_round(pl, workspace_bytes);
  if (R < 1) return IKR_ERR_WORKSPACE;
  unsigned char* ws = (unsigned char*)workspace;
  const size_t off_stash = (pl.fixed_bytes + 255) & ~(size_t)255;
  if (off_stash + (size_t)pl.n_tiles * (6 * (size_t)R + 1) * (size_t)pl.sg.slot > workspace_bytes)
    return IKR_ERR_WORKSPACE;
  if (cudaMemsetAsync(ws + pl.off_partial, 0, pl.partial_bytes, st) != cudaSuccess) return IKR_ERR_DEVICE;

  const MlpView mv = make_view(d, io->weights);
  TcPackParams pk;
  pk.wn = reinterpret_cast<const float*>(io->weights) + mv.off_wn;
  pk.wt = reinterpret_cast<const float*>(io->weights) + mv.off_wt;
  pk.npad = mv.npad;
  pk.n_seq = 2 * d->n_layers;
  pk.g = pl.g;
  pk.img = reinterpret_cast<uint16_t*>(ws + pl.off_img);
  pk.absmax = nullptr; pk.scales = nullptr;
  ikr_tc_pack_kernel<<<pl.sms, 256, 0, st>>>(pk);
  if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;

  TcAdjParams tp;
  BwdParams& p = tp.b;
  p.mlp = mv;
  p.cfg = make_cfg(d);
  p.cfg.tab = make_table(io);
  p.M = pl.tile_lanes; p.MG = 0; p.NG = 0; p.n_worker_warps = 0;
  p.B = io->B; p.T = (int)io->T; p.n_tiles = pl.n_tiles;
  p.y0 = io->y0; p.t_out = io->t_out; p.stats = io->stats_out;
  p.ckpt_t = io->ckpt_t; p.ckpt_y = io->ckpt_y;
  p.grad_y = bio->grad_y; p.fused_loss = bio->fused_loss;
  p.y_out = io->y_out; p.v_out = io->v_out; p.g = io->g; p.e_rev = io->e_rev;
  p.e_scalar = io->e_scalar; p.data = io->data; p.data_B = io->data_B;
  p.lane_state = ws + pl.off_lanes;
  p.steps_per_round = R;
  p.stash_h = nullptr; p.stash_d = nullptr;
  p.counters = reinterpret_cast<unsigned long long*>(ws + pl.off_counters);
  p.small_grad = nullptr; p.small_stride = 0;
  p.grad_y0 = bio->grad_y0; p.grad_g = bio->grad_g;
  p.method = d->method; p.time_f32 = d->time_f32; p.rk4_perturb = d->rk4_perturb;
  tp.g = pl.g;
  tp.sg = pl.sg;
  tp.img = ws + pl.off_img;
  tp.stash = ws + off_stash;
  tp.mask_words = pl.mask_words;
  tp.timing = (d->reserved & 8) ? 1 : 0;

  TcWgradParams wp;
  wp.g = pl.g; wp.sg = pl.sg;
  wp.stash = tp.stash;
  wp.counters = p.counters;
  wp.slots_fixed = -1;
  wp.S = pl.wg_S;
  wp.stages = pl.wg_stages;
  wp.partial = reinterpret_cast<double*>(ws + pl.off_partial);
  if (allow_max_smem(ikr_wgrad_tc_kernel, pl.wg_smem) != cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }

  long long max_steps = bio->max_accepted_steps > 0 ? bio->max_accepted_steps : io->ckpt_cap;
  if (max_steps > io->ckpt_cap) max_steps = io->ckpt_cap;
  // When the adjoint kernel leaves SMs idle (fewer tiles than SMs: e.g. 4,096 datasets = 32 tiles), the
  // weight-gradient GEMM of round r runs on a second stream UNDER the adjoint kernel of round r + 1:
  // the stash and the round counters are double-buffered (half the steps per round each), events
  // order producer and consumer.  The stream and the events live for this call only.
  const bool overlap = pl.n_tiles * 2 <= pl.sms && R >= 3 && !(d->reserved & 512);
  if (!overlap) {
    const long long rounds = max_steps > 0 ? (max_steps + R - 1) / R : 1;
    for (long long r = 0; r < rounds; ++r) {
      if (cudaMemsetAsync(ws + pl.off_counters, 0, 256, st) != cudaSuccess) return IKR_ERR_DEVICE;
      p.first_round = r == 0 ? 1 : 0;
      const int rc = d->state_dtype == IKR_F32 ? launch_adjoint_tc<float>(tp, pl, tc_adjoint_lite(d), st)
                                               : launch_adjoint_tc<double>(tp, pl, tc_adjoint_lite(d), st);
      if (rc != 0) return rc;
      ikr_wgrad_tc_kernel<<<(d->n_layers + 2) * pl.wg_S, kWgTcThreads, pl.wg_smem, st>>>(wp);
      if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;
    }
  } else {
    const int R2 = (R - 1) / 2;      // 2 x (6 R2 + 1) slots per tile fit where (6 R + 1) did
    const size_t half = (size_t)pl.n_tiles * (6 * (size_t)R2 + 1) * (size_t)pl.sg.slot;
    const long long rounds = max_steps > 0 ? (max_steps + R2 - 1) / R2 : 1;
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_adj[2] = {nullptr, nullptr}, ev_wg[2] = {nullptr, nullptr};
    bool ok = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&ev_adj[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&ev_wg[i], cudaEventDisableTiming) == cudaSuccess;
    int rc = ok ? 0 : IKR_ERR_DEVICE;
    p.steps_per_round = R2;
    for (long long r = 0; r < rounds && rc == 0; ++r) {
      const int b = (int)(r & 1);
      // buffer b was last read by the weight-gradient GEMM of round r - 2
      if (r >= 2 && cudaStreamWaitEvent(st, ev_wg[b], 0) != cudaSuccess) { rc = IKR_ERR_DEVICE; break; }
      unsigned char* cnt = ws + pl.off_counters + 128 * b;
      if (cudaMemsetAsync(cnt, 0, 128, st) != cudaSuccess) { rc = IKR_ERR_DEVICE; break; }
      p.first_round = r == 0 ? 1 : 0;
      p.counters = reinterpret_cast<unsigned long long*>(cnt);
      tp.stash = ws + off_stash + (size_t)b * half;
      rc = d->state_dtype == IKR_F32 ? launch_adjoint_tc<float>(tp, pl, tc_adjoint_lite(d), st)
                                     : launch_adjoint_tc<double>(tp, pl, tc_adjoint_lite(d), st);
      if (rc != 0) break;
      if (cudaEventRecord(ev_adj[b], st) != cudaSuccess || cudaStreamWaitEvent(s2, ev_adj[b], 0) != cudaSuccess) {
        rc = IKR_ERR_DEVICE;
        break;
      }
      wp.stash = tp.stash;
      wp.counters = p.counters;
      ikr_wgrad_tc_kernel<<<(d->n_layers + 2) * pl.wg_S, kWgTcThreads, pl.wg_smem, s2>>>(wp);
      if (cudaGetLastError() != cudaSuccess) { rc = IKR_ERR_LAUNCH; break; }
      if (cudaEventRecord(ev_wg[b], s2) != cudaSuccess) { rc = IKR_ERR_DEVICE; break; }
    }
    // the caller's stream continues after every weight-gradient launch
    if (ok) {
      for (int i = 0; i < 2; ++i)
        if (cudaStreamWaitEvent(st, ev_wg[i], 0) != cudaSuccess && rc == 0) rc = IKR_ERR_DEVICE;
    }
    for (int i = 0; i < 2; ++i) {
      if (ev_adj[i]) cudaEventDestroy(ev_adj[i]);
      if (ev_wg[i]) cudaEventDestroy(ev_wg[i]);
    }
    if (s2) cudaStreamDestroy(s2);
    if (rc != 0) return rc;
  }

  TcReduceParams rp;
  rp.L = d->n_layers; rp.n = d->n_nodes; rp.NP = pl.g.NP; rp.S = pl.wg_S;
  rp.partial = wp.partial;
  rp.out = bio->grad_weights;
  const long long n = d->n_nodes, Ln = d->n_layers;
  rp.n_params = 3 * n + Ln * (n * n + n) + n + 1;
  const int threads = 256;
  const unsigned blocks = (unsigned)((rp.n_params + threads - 1) / threads);
  ikr_grad_reduce_tc_kernel<<<blocks, threads, 0, st>>>(rp);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

int bwd_dispatch(const ikr_desc* d, const ikr_io* io, const ikr_bwd_io* bio, void* workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  if (io->B < 1 || io->T < 1 || !io->weights || !io->table_t || !io->table_v || !io->y0 ||
      !io->t_out || !io->stats_out || !io->ckpt_t || !io->ckpt_y || io->ckpt_cap < 1 ||
      !bio->grad_weights)
    return IKR_ERR_ARG;
  if (bio->fused_loss == 0 && !bio->grad_y) return IKR_ERR_ARG;
  if (bio->fused_loss != 0 && (!io->y_out || !io->v_out || !io->data)) return IKR_ERR_ARG;
  if (bio->fused_loss < 0 || bio->fused_loss > 2) return IKR_ERR_ARG;
  if (!workspace) return IKR_ERR_WORKSPACE;
  {
    const TcBwdPlan tpl = make_tc_bwd_plan(d, io->B);
    if (tpl.ok) return bwd_dispatch_tc(d, io, bio, tpl, workspace, workspace_bytes, st);
  }
  const BwdPlan pl = make_bwd_plan(d, io->B);
  const Geometry& g = pl.g;
  if (g.smem > kSmemLimit || g.threads > kMaxThreads || pl.wg_smem > kSmemLimit)
    return IKR_ERR_UNSUPPORTED;
  const int R = bwd_steps_per_round(pl, workspace_bytes);
  if (R < 1) return IKR_ERR_WORKSPACE;
  unsigned char* ws = (unsigned char*)workspace;
  const size_t stash_each = (size_t)g.n_tiles * (6 * (size_t)R + 1) * pl.slot_bytes;
  const size_t off_sh = (pl.fixed_bytes + 255) & ~(size_t)255;
  const size_t off_sd = (off_sh + stash_each + 255) & ~(size_t)255;
  if (off_sd + stash_each > workspace_bytes) return IKR_ERR_WORKSPACE;

  if (cudaMemsetAsync(ws + pl.zero_begin, 0, pl.zero_bytes, st) != cudaSuccess)
    return IKR_ERR_DEVICE;

  BwdParams p;
  p.mlp = make_view(d, io->weights);
  p.mlp.bwd_seq = 1;
  p.cfg = make_cfg(d);
  p.cfg.tab = make_table(io);
  p.M = g.M; p.MG = g.MG; p.NG = g.NG; p.n_worker_warps = g.n_worker_warps;
  p.B = io->B; p.T = (int)io->T; p.n_tiles = g.n_tiles;
  p.y0 = io->y0; p.t_out = io->t_out; p.stats = io->stats_out;
  p.ckpt_t = io->ckpt_t; p.ckpt_y = io->ckpt_y;
  p.grad_y = bio->grad_y; p.fused_loss = bio->fused_loss;
  p.y_out = io->y_out; p.v_out = io->v_out; p.g = io->g; p.e_rev = io->e_rev;
  p.e_scalar = io->e_scalar; p.data = io->data; p.data_B = io->data_B;
  p.lane_state = ws + pl.off_lanes;
  p.steps_per_round = R;
  p.stash_h = ws + off_sh; p.stash_d = ws + off_sd;
  p.counters = reinterpret_cast<unsigned long long*>(ws + pl.off_counters);
  p.small_grad = reinterpret_cast<double*>(ws + pl.off_small);
  p.small_stride = pl.small_stride;
  p.grad_y0 = bio->grad_y0; p.grad_g = bio->grad_g;
  p.method = d->method; p.time_f32 = d->time_f32; p.rk4_perturb = d->rk4_perturb;

  WgradParams wp;
  wp.L = d->n_layers; wp.n = d->n_nodes; wp.npad = g.npad; wp.M = g.M; wp.KC = pl.KC;
  wp.n_ot = pl.n_ot; wp.n_it = pl.n_it; wp.BO = pl.BO; wp.BI = pl.BI; wp.S = pl.S; wp.IG = pl.IG;
  wp.stages = pl.wg_stages;
  wp.stash_h = p.stash_h; wp.stash_d = p.stash_d; wp.counters = p.counters;
  wp.partial_w = reinterpret_cast<double*>(ws + pl.off_pw);
  wp.partial_b = reinterpret_cast<double*>(ws + pl.off_pb);

  long long max_steps = bio->max_accepted_steps > 0 ? bio->max_accepted_steps : io->ckpt_cap;
  if (max_steps > io->ckpt_cap) max_steps = io->ckpt_cap;
  const long long rounds = max_steps > 0 ? (max_steps + R - 1) / R : 1;
  for (long long r = 0; r < rounds; ++r) {
    if (cudaMemsetAsync(ws + pl.off_counters, 0, 256, st) != cudaSuccess) return IKR_ERR_DEVICE;
    p.first_round = r == 0 ? 1 : 0;
    int rc;
    if (d->state_dtype == IKR_F32) rc = launch_adjoint<float, float>(p, g, st);
    else if (d->mlp_dtype == IKR_F32) rc = launch_adjoint<double, float>(p, g, st);
    else rc = launch_adjoint<double, double>(p, g, st);
    if (rc != 0) return rc;
    rc = d->mlp_dtype == IKR_F32 ? launch_wgrad<float>(wp, pl, st) : launch_wgrad<double>(wp, pl, st);
    if (rc != 0) return rc;
  }

  ReduceParams rp;
  rp.L = d->n_layers; rp.n = d->n_nodes; rp.npad = g.npad; rp.S = pl.S; rp.n_cta = g.sms;
  rp.small_stride = pl.small_stride;
  rp.partial_w = wp.partial_w; rp.partial_b = wp.partial_b; rp.small_grad = p.small_grad;
  rp.out = bio->grad_weights;
  const long long n = d->n_nodes, Ln = d->n_layers;
  rp.n_params = 3 * n + Ln * (n * n + n) + n + 1;
  const int threads = 256;
  const unsigned blocks = (unsigned)((rp.n_params + threads - 1) / threads);
  ikr_grad_reduce_kernel<<<blocks, threads, 0, st>>>(rp);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

// ---------------------------------------------------------------------------------------------
// MLP regression stage (ikr_regress_tc.cuh): plan + launch
// ---------------------------------------------------------------------------------------------
struct TcRegPlan {
  bool ok;
  TcBwdPlan b;          // geometry / smem / wgrad configuration shared with the ODE backward
  long long n_tiles;
  size_t off_loss, off_partial, off_img, fixed_bytes, img_bytes;
};

TcRegPlan make_tc_reg_plan(const ikr_desc* d, long long N) {
  TcRegPlan r;
  r.ok = false;
  ikr_desc dd = *d;
  dd.method = IKR_DOPRI5;
  dd.state_dtype = IKR_F32;
  dd.reserved |= 1024;          // the regression kernel issues all six products: full-width weight ring
  r.b = make_tc_bwd_plan(&dd, N);
  if (!r.b.ok) return r;
  r.n_tiles = (N + kTcM - 1) / kTcM;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
  r.off_loss = take(256);
  r.off_partial = take(r.b.partial_bytes);
  r.img_bytes = (size_t)3 * d->n_layers * r.b.g.KST * r.b.g.stage_bytes;
  r.off_img = take(r.img_bytes);
  r.fixed_bytes = o;
  r.ok = true;
  return r;
}

template <int G>
int launch_regress_g(const TcRegParams& tp, const TcBwdPlan& pl, int grid, cudaStream_t st) {
  auto kern = ikr_regress_tc_kernel<G>;
  if (allow_max_smem(kern, pl.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  kern<<<grid, tc_threads(G), pl.smem, st>>>(tp);
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

int32_t ikr_uses_tensor_cores(const ikr_desc* d) {
  if (!valid_desc(d)) return IKR_ERR_ARG;
  return make_tc_plan(d).ok ? 1 : 0;
}

int32_t ikr_tile_m(const ikr_desc* d, int32_t n_jobs, const int64_t* B) {
  if (!valid_desc(d) || n_jobs < 1 || !B) return IKR_ERR_ARG;
  for (int j = 0; j < n_jobs; ++j) if (B[j] < 1) return IKR_ERR_ARG;
  if (make_tc_plan(d).ok) {
    long long b_total = 0;
    for (int j = 0; j < n_jobs; ++j) b_total += B[j];
    return tc_tile_lanes(b_total, device_sms());
  }
  return make_geometry(d, n_jobs, (const long long*)B, use_pool(d)).M;
}

int ikr_launch_geometry(const ikr_desc* d, int32_t n_jobs, const int64_t* B, int64_t out[16]) {
  if (!valid_desc(d) || n_jobs < 1 || !B || !out) return IKR_ERR_ARG;
  for (int j = 0; j < n_jobs; ++j) if (B[j] < 1) return IKR_ERR_ARG;
  for (int i = 0; i < 16; ++i) out[i] = 0;
  const TcPlan tcp = make_tc_plan(d, tc_forward_terms(d));
  if (tcp.ok) {
    const int sms = device_sms();
    long long tiles = 0, b_total = 0;
    for (int j = 0; j < n_jobs; ++j) b_total += B[j];
    const int tl = tc_tile_lanes(b_total, sms);
    const bool pool = use_pool_tc(d, b_total, sms);
    if (pool) tiles = (b_total + tl - 1) / tl;
    else for (int j = 0; j < n_jobs; ++j) tiles += (B[j] + tl - 1) / tl;
    out[0] = tl; out[1] = tc_threads(tcp.groups); out[2] = tiles < sms ? tiles : sms; out[3] = (int64_t)tcp.smem;
    out[4] = tiles; out[5] = 16; out[6] = tcp.g.KST; out[7] = sms;
    out[8] = pool ? 1 : 0;
    if (use_pp_tc(d, b_total, sms, n_jobs) && make_tc_pp_plan(d, tcp).ok) {
      const TcPpPlan ppl = make_tc_pp_plan(d, tcp);
      const long long ctas = (b_total + 2 * kTcM - 1) / (2 * kTcM);
      out[8] = 2;
      out[1] = tc_threads(3);
      out[2] = ctas < sms ? ctas : sms;
      out[3] = (int64_t)ppl.smem;
      out[4] = (b_total + tl - 1) / tl;
      out[11] = 2;
    }
    out[9] = tcp.g.terms == 2 ? 3 : 2;   // (ikr_tc_absmax_kernel +) ikr_tc_pack_kernel + the forward kernel
    out[10] = 1;
    if (out[8] != 2) out[11] = tcp.groups;
    out[12] = tcp.g.terms == 2 ? 3 : 6;  // 16-bit MMAs per fp32 product
    return 0;
  }
  const bool pool = use_pool(d);
  Geometry g = make_geometry(d, n_jobs, (const long long*)B, pool);
  out[0] = g.M; out[1] = g.threads; out[2] = g.grid; out[3] = (int64_t)g.smem;
  out[4] = g.n_tiles; out[5] = g.kc; out[6] = g.cpl; out[7] = g.sms;
  out[8] = pool ? 1 : 0;
  out[9] = 1;
  out[10] = 0;
  out[11] = 0;
  return 0;
}

size_t ikr_workspace_bytes(const ikr_desc* d, int32_t n_jobs, int64_t B_total,
                           int32_t with_backward) {
  if (!valid_desc(d) || n_jobs < 1 || B_total < 1) return 0;
  size_t bytes = fwd_fixed_workspace(n_jobs);
  const TcPlan tcp = make_tc_plan(d, tc_forward_terms(d));
  if (tcp.ok) bytes += (tcp.img_bytes + 255) & ~(size_t)255;   // weight image of the tcgen05 path
  if (with_backward) {
    const TcBwdPlan tpl = make_tc_bwd_plan(d, B_total);
    const int gib = with_backward > 1 ? (with_backward > 128 ? 128 : with_backward) : 4;
    bytes += tpl.ok ? tc_bwd_workspace_bytes(tpl, gib) : bwd_workspace_bytes(d, B_total, gib);
  }
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
  const TcPlan tcp = make_tc_plan(d, tc_forward_terms(d));
  const size_t need = fwd_fixed_workspace(n_jobs) + (tcp.ok ? tcp.img_bytes : 0);
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
  bool pool = use_pool(d);
  Geometry g = make_geometry(d, n_jobs, Bs.data(), pool);
  if (tcp.ok) {
    // tensor-core kernel: tiles of up to 128 trajectories (one per TMEM lane)
    long long b_total = 0;
    for (int j = 0; j < n_jobs; ++j) b_total += Bs[j];
    pool = use_pool_tc(d, b_total, g.sms);
    g.M = tc_tile_lanes(b_total, g.sms);
    g.n_tiles = 0;
    if (pool) g.n_tiles = (b_total + g.M - 1) / g.M;      // lane slots refill from one queue
    else for (int j = 0; j < n_jobs; ++j) g.n_tiles += (Bs[j] + g.M - 1) / g.M;
    g.grid = (int)(g.n_tiles < g.sms ? g.n_tiles : g.sms);
    g.threads = tc_threads(tcp.groups);
    g.smem = tcp.smem;
  }
  if (g.smem > kSmemLimit || g.threads > kMaxThreads) return IKR_ERR_UNSUPPORTED;

  std::vector<FwdJob> table(n_jobs);
  long long tile = 0, traj = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const ikr_io* io = &jobs[order[j]];
    FwdJob& fj = table[j];
    fj.traj_begin = traj;
    traj += io->B;
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
  p.n_traj = traj;
  p.jobs_are_inline = n_jobs <= kInlineJobs ? 1 : 0;
  for (int j = 0; j < kInlineJobs; ++j) p.jobs_inline[j] = table[j < n_jobs ? j : 0];
  p.jobs = reinterpret_cast<const FwdJob*>(ws + 256);
  p.queue = reinterpret_cast<unsigned long long*>(ws);

  if (tcp.ok) {
    TcFwdParams tp;
    tp.f = p;
    tp.g = tcp.g;
    unsigned char* img = ws + fwd_fixed_workspace(n_jobs);
    tp.img = img;
    tp.timing = (d->reserved & 8) ? 1 : 0;
    TcPackParams pk;
    pk.wn = reinterpret_cast<const float*>(jobs[0].weights) + p.mlp.off_wn;
    pk.wt = reinterpret_cast<const float*>(jobs[0].weights) + p.mlp.off_wt;
    pk.n_seq = d->n_layers;
    pk.npad = p.mlp.npad;
    pk.g = tcp.g;
    pk.img = reinterpret_cast<uint16_t*>(img);
    unsigned char* tail = img + tcp.img_bytes - 512;
    pk.absmax = reinterpret_cast<unsigned*>(tail);
    pk.scales = reinterpret_cast<float*>(tail + 256);
    tp.scales = pk.scales;
    if (tcp.g.terms == 2) {
      if (cudaMemsetAsync(tail, 0, 256, st) != cudaSuccess) return IKR_ERR_DEVICE;
      ikr_tc_absmax_kernel<<<g.sms, 256, 0, st>>>(pk);
      if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;
    }
    ikr_tc_pack_kernel<<<g.sms, 256, 0, st>>>(pk);
    if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;
    if (use_pp_tc(d, traj, g.sms, n_jobs)) {
      const TcPpPlan ppl = make_tc_pp_plan(d, tcp);
      if (ppl.ok) {
        tp.g = ppl.g;
        // 256 lane slots per CTA; the queue hands out trajectories, so the grid only needs to cover them
        const long long ctas = (traj + 2 * kTcM - 1) / (2 * kTcM);
        const int grid = (int)(ctas < g.sms ? ctas : g.sms);
        tp.f.n_tiles = (traj + kTcM - 1) / kTcM;
        if (d->state_dtype == IKR_F32) return launch_forward_tc_pp<float>(tp, ppl, grid, st);
        return launch_forward_tc_pp<double>(tp, ppl, grid, st);
      }
    }
    if (d->state_dtype == IKR_F32) return launch_forward_tc<float>(tp, tcp, g.grid, st, pool);
    return launch_forward_tc<double>(tp, tcp, g.grid, st, pool);
  }
  if (d->state_dtype == IKR_F32) return launch_forward<float, float>(p, g, st, pool);
  if (d->mlp_dtype == IKR_F32) return launch_forward<double, float>(p, g, st, pool);
  return launch_forward<double, double>(p, g, st, pool);
}

int ikr_backward(const ikr_desc* d, const ikr_io* io, const ikr_bwd_io* bio, void* workspace,
                 size_t workspace_bytes, void* cuda_stream) {
  if (!valid_desc(d) || !io || !bio) return IKR_ERR_ARG;
  return bwd_dispatch(d, io, bio, workspace, workspace_bytes, (cudaStream_t)cuda_stream);
}

int ikr_forward_hh(const ikr_desc* d, const ikr_io* io, const double* hh_params, void* cuda_stream) {
  if (!d || !io) return IKR_ERR_ARG;
  if (d->state_dtype != IKR_F32 && d->state_dtype != IKR_F64) return IKR_ERR_ARG;
  if (d->method != IKR_DOPRI5 && d->method != IKR_RK4) return IKR_ERR_ARG;
  if (io->B < 1 || io->T < 1 || !io->table_t || !io->table_v || !io->y0 || !io->t_out ||
      !io->stats_out || io->table_len < 2)
    return IKR_ERR_ARG;
  if (d->method == IKR_RK4 && (!io->grid || io->G < 1)) return IKR_ERR_ARG;
  if ((io->i_out || io->loss_out) && !io->v_out) return IKR_ERR_ARG;
  if (io->data && io->data_B != 1 && io->data_B != io->B) return IKR_ERR_ARG;
  if (io->T > 2147483647LL || io->G > 2147483647LL) return IKR_ERR_ARG;
  HhKernelParams p;
  p.cfg = make_cfg(d);
  p.cfg.mlp_is_f64 = 1;   // no network: the (absent) MLP term is an exact zero in any precision
  FwdJob& fj = p.job;
  fj.tab = make_table(io);
  fj.B = io->B; fj.T = (int)io->T; fj.G = (int)io->G;
  fj.tile_begin = 0; fj.traj_begin = 0;
  fj.y0 = io->y0; fj.t_out = io->t_out; fj.grid = io->grid; fj.v_out = io->v_out;
  fj.g = io->g; fj.e_rev = io->e_rev; fj.e_scalar = io->e_scalar;
  fj.data = io->data; fj.data_B = io->data_B;
  fj.y_out = io->y_out; fj.i_out = io->i_out; fj.loss_out = io->loss_out;
  fj.stats_out = io->stats_out;
  fj.ckpt_cap = 0; fj.ckpt_t = nullptr; fj.ckpt_y = nullptr;
  p.hh_params = hh_params;
  p.method = d->method; p.time_f32 = d->time_f32; p.rk4_perturb = d->rk4_perturb;
  const int threads = 128;
  const unsigned blocks = (unsigned)((io->B + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (d->state_dtype == IKR_F32) ikr_hh_kernel<float><<<blocks, threads, 0, st>>>(p);
  else ikr_hh_kernel<double><<<blocks, threads, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
}

size_t ikr_regression_workspace_bytes(const ikr_desc* d, int64_t N) {
  if (!valid_desc(d) || N < 1) return 0;
  const TcRegPlan r = make_tc_reg_plan(d, N);
  if (!r.ok) return 0;
  // stash: every tile of the batch when that stays below ~4 GiB, else rounds of tiles
  long long tiles = r.n_tiles;
  const long long cap = (long long)(4.0 * 1024 * 1024 * 1024 / (double)r.b.sg.slot);
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  return r.fixed_bytes + 512 + (size_t)tiles * (size_t)r.b.sg.slot;
}

int ikr_regression_loss_grad(const ikr_desc* d, const void* weights, const void* x, const void* y,
                             int64_t N, double* loss_out, double* grad_weights, void* workspace,
                             size_t workspace_bytes, void* cuda_stream) {
  if (!valid_desc(d) || !weights || !x || !y || N < 1 || !loss_out || !grad_weights) return IKR_ERR_ARG;
  const TcRegPlan r = make_tc_reg_plan(d, N);
  if (!r.ok) return IKR_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes <= r.fixed_bytes + 512) return IKR_ERR_WORKSPACE;
  const TcBwdPlan& pl = r.b;
  unsigned char* ws = (unsigned char*)workspace;
  const size_t off_stash = (r.fixed_bytes + 255) & ~(size_t)255;
  const long long slots_cap = (long long)((workspace_bytes - off_stash) / (size_t)pl.sg.slot);
  if (slots_cap < 1) return IKR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (cudaMemsetAsync(ws + r.off_loss, 0, 256, st) != cudaSuccess) return IKR_ERR_DEVICE;
  if (cudaMemsetAsync(ws + r.off_partial, 0, pl.partial_bytes, st) != cudaSuccess) return IKR_ERR_DEVICE;

  const MlpView mv = make_view(d, weights);
  TcPackParams pk;
  pk.wn = reinterpret_cast<const float*>(weights) + mv.off_wn;
  pk.wt = reinterpret_cast<const float*>(weights) + mv.off_wt;
  pk.npad = mv.npad;
  pk.n_seq = 3 * d->n_layers;
  pk.g = pl.g;
  pk.img = reinterpret_cast<uint16_t*>(ws + r.off_img);
  pk.absmax = nullptr; pk.scales = nullptr;
  ikr_tc_pack_kernel<<<pl.sms, 256, 0, st>>>(pk);
  if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;

  TcRegParams tp;
  tp.mlp = mv;
  tp.g = pl.g;
  tp.sg = pl.sg;
  tp.img = ws + r.off_img;
  tp.stash = ws + off_stash;
  tp.x = reinterpret_cast<const float*>(x);
  tp.y = reinterpret_cast<const float*>(y);
  tp.N = N;
  tp.netscale = (float)d->netscale;
  tp.loss_out = reinterpret_cast<double*>(ws + r.off_loss);
  tp.mask_words = pl.mask_words;

  TcWgradParams wp;
  wp.g = pl.g; wp.sg = pl.sg;
  wp.stash = tp.stash;
  wp.counters = nullptr;
  wp.S = pl.wg_S;
  wp.stages = pl.wg_stages;
  wp.partial = reinterpret_cast<double*>(ws + r.off_partial);
  if (allow_max_smem(ikr_wgrad_tc_kernel, pl.wg_smem) != cudaSuccess) {
    cudaGetLastError();
    return IKR_ERR_LAUNCH;
  }
  for (long long t0 = 0; t0 < r.n_tiles; t0 += slots_cap) {
    const long long t1 = t0 + slots_cap < r.n_tiles ? t0 + slots_cap : r.n_tiles;
    tp.tile_begin = t0; tp.tile_end = t1;
    const int grid = (int)((t1 - t0) < pl.sms ? (t1 - t0) : pl.sms);
    int rc;
    if (pl.groups == 1) rc = launch_regress_g<1>(tp, pl, grid, st);
    else if (pl.groups == 2) rc = launch_regress_g<2>(tp, pl, grid, st);
    else rc = launch_regress_g<3>(tp, pl, grid, st);
    if (rc != 0) return rc;
    wp.slots_fixed = t1 - t0;
    ikr_wgrad_tc_kernel<<<(d->n_layers + 2) * pl.wg_S, kWgTcThreads, pl.wg_smem, st>>>(wp);
    if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;
  }
  TcReduceParams rp;
  rp.L = d->n_layers; rp.n = d->n_nodes; rp.NP = pl.g.NP; rp.S = pl.wg_S;
  rp.partial = wp.partial;
  rp.out = grad_weights;
  const long long n = d->n_nodes, Ln = d->n_layers;
  rp.n_params = 3 * n + Ln * (n * n + n) + n + 1;
  const int threads = 256;
  const unsigned blocks = (unsigned)((rp.n_params + threads - 1) / threads);
  ikr_grad_reduce_tc_kernel<<<blocks, threads, 0, st>>>(rp);
  if (cudaGetLastError() != cudaSuccess) return IKR_ERR_LAUNCH;
  if (cudaMemcpyAsync(loss_out, ws + r.off_loss, sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return IKR_ERR_DEVICE;
  return 0;
}

int ikr_forward_markov(const ikr_desc* d, const ikr_markov_io* io, void* cuda_stream) {
  if (!d || !io) return IKR_ERR_ARG;
  if (d->state_dtype != IKR_F32 && d->state_dtype != IKR_F64) return IKR_ERR_ARG;
  if (d->method != IKR_DOPRI5 && d->method != IKR_RK4) return IKR_ERR_ARG;
  if (io->B < 1 || io->T < 1 || !io->table_t || !io->table_v || io->table_len < 2 || !io->y0 ||
      !io->t_out || !io->stats_out)
    return IKR_ERR_ARG;
  if (d->method == IKR_RK4 && (!io->grid || io->G < 1)) return IKR_ERR_ARG;
  if (io->i_out && !io->v_out) return IKR_ERR_ARG;
  if (io->T > 2147483647LL || io->G > 2147483647LL || !(io->noise_sigma >= 0.0)) return IKR_ERR_ARG;
  MarkovKernelParams p;
  p.cfg = make_cfg(d);
  p.cfg.tab.t = io->table_t; p.cfg.tab.v = io->table_v; p.cfg.tab.len = io->table_len;
  p.cfg.tab.uniform = io->table_uniform; p.cfg.tab.t0 = io->table_t0; p.cfg.tab.inv_dt = io->table_inv_dt;
  for (int i = 0; i < 12; ++i) p.p[i] = io->p[i];
  p.params = io->params;
  p.B = io->B; p.T = (int)io->T; p.G = (int)io->G;
  p.y0 = io->y0; p.t_out = io->t_out; p.grid = io->grid; p.v_out = io->v_out;
  p.y_out = io->y_out; p.i_out = io->i_out; p.g = io->g; p.e_rev = io->e_rev;
  p.noise_sigma = io->noise_sigma; p.seed = io->seed;
  p.stats_out = io->stats_out;
  p.method = d->method; p.time_f32 = d->time_f32; p.rk4_perturb = d->rk4_perturb;
  const int threads = 128;
  const unsigned blocks = (unsigned)((io->B + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (d->state_dtype == IKR_F32) ikr_markov_kernel<float><<<blocks, threads, 0, st>>>(p);
  else ikr_markov_kernel<double><<<blocks, threads, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : IKR_ERR_LAUNCH;
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
    else if (dtype == 2)
      ikr_fma2_peak_kernel<<<blocks, threads, 0, st>>>((float*)sink, iters);
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
  double flops = 2.0 * 16.0 * (double)iters * threads * (double)blocks * (dtype == 2 ? 2.0 : 1.0);
  *tflops_out = flops / (ms * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
