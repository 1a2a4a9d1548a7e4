// ikr_forward.cuh -- the forward integration kernel.
//
// One persistent CTA integrates a tile of M trajectories from t[0] to t[-1] inside ONE launch:
// per-trajectory adaptive dopri5 (own dt, own error control, own dense output) or fixed-grid
// RK4 (3/8 rule).  All lanes of a tile advance in lock-step *by RHS evaluation*, not by time: a
// lane whose step is rejected simply retries while its neighbours move on; finished lanes are
// masked.  The RHS evaluation is fused: V(t) interpolation + HH rates (owner thread of the lane,
// ikr_math.h) and the MLP as a register-tiled tile GEMM (ikr_device.cuh).
#ifndef IKR_FORWARD_CUH_
#define IKR_FORWARD_CUH_

#include "ikr_device.cuh"

namespace ikr {

// One job = one protocol table + one batch of trajectories (one reference `odeint` call shape).
// A launch integrates any number of jobs that share the MLP weights: tiles of all jobs go
// through one dynamic queue (longest jobs first) so that every SM stays busy.
struct FwdJob {
  ProtocolTable tab;
  long long B;
  int T;
  int G;
  long long tile_begin;  // first global tile index of this job (tile-scheduled kernel)
  long long traj_begin;  // first global trajectory index of this job (lane-pool kernel)
  const void* y0;
  const double* t_out;
  const double* grid;
  const double* v_out;
  const void* g;
  const void* e_rev;
  double e_scalar;
  const void* data;
  long long data_B;
  void* y_out;
  void* i_out;
  double* loss_out;
  int* stats_out;
  long long ckpt_cap;
  double* ckpt_t;
  void* ckpt_y;
};

constexpr int kInlineJobs = 8;   // job descriptors carried in the kernel parameters (constant bank)

struct FwdParams {
  MlpView mlp;
  SolverCfg cfg;     // cfg.tab is overwritten per job
  int method;        // 0 dopri5, 1 rk4
  int time_f32;      // rk4
  int rk4_perturb;   // rk4
  int M, MG, NG;     // tile geometry
  int n_worker_warps;
  int n_jobs;
  long long n_tiles;
  long long n_traj;          // trajectories of all jobs (lane-pool kernel)
  const FwdJob* jobs;        // device array [n_jobs]
  int jobs_are_inline;       // n_jobs <= kInlineJobs: read jobs_inline (no L2 hot spot)
  FwdJob jobs_inline[kInlineJobs];
  unsigned long long* queue; // device tile counter (zeroed by the host before launch)
};

template <typename S>
struct Vec2;
template <>
struct Vec2<float> { typedef float2 type; };
template <>
struct Vec2<double> { typedef double2 type; };

// shared memory carve-up (host and device agree through this one function)
template <typename S, typename W>
struct FwdSmemLayout {
  size_t off_lanes, off_obs, off_aux, off_xin, off_sp, off_hs, off_wr, off_bar, off_job, off_tile, total;
  __host__ __device__ FwdSmemLayout(int M, int npad, int kc, int L) {
    size_t o = 0;
    off_bar = o; o += 64;
    off_job = o; o += (sizeof(FwdJob) + 15) & ~(size_t)15;
    off_tile = o; o += 16;
    off_lanes = o; o += (size_t)M * sizeof(Lane<S>); o = (o + 15) & ~(size_t)15;
    off_obs = o; o += (size_t)M * 2 * sizeof(double); o = (o + 15) & ~(size_t)15;
    off_aux = o; o += (size_t)M * 32; o = (o + 15) & ~(size_t)15;   // LaneAux (lane-pool kernel)
    off_xin = o; o += (size_t)2 * M * sizeof(W); o = (o + 15) & ~(size_t)15;
    off_sp = o; o += mlp_small_elems(L, npad) * sizeof(W); o = (o + 127) & ~(size_t)127;
    off_hs = o; o += (size_t)npad * M * sizeof(W); o = (o + 127) & ~(size_t)127;
    off_wr = o; o += ((size_t)kStages * kc + 1) * npad * sizeof(W);   // +1 row: prefetch pad
    total = o;
  }
};

template <typename S, typename W, int TN>
__global__ void __launch_bounds__(512, 1) ikr_forward_kernel(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int M = p.M;
  const FwdSmemLayout<S, W> lay(M, p.mlp.npad, p.mlp.kc, p.mlp.L);
  Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
  double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);  // [M][2] sse, sae
  FwdJob* jobp = reinterpret_cast<FwdJob*>(smem_raw + lay.off_job);
  long long* tile_slot = reinterpret_cast<long long*>(smem_raw + lay.off_tile);
  MlpSmem<W> sm;
  sm.full = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  sm.empty = sm.full + kStages;
  sm.xin = reinterpret_cast<W*>(smem_raw + lay.off_xin);
  sm.Hs = reinterpret_cast<W*>(smem_raw + lay.off_hs);
  sm.Wr = reinterpret_cast<W*>(smem_raw + lay.off_wr);
  sm.sp = reinterpret_cast<W*>(smem_raw + lay.off_sp);

  mlp_pipe_init<W>(sm, p.n_worker_warps);
  mlp_stage_small<W>(p.mlp, sm);
  __syncthreads();
  MlpPipe pp;
  mlp_pipe_start<W>(p.mlp, sm, pp);

  SolverCfg cfg = p.cfg;
  const bool owner = tid < M;

  while (true) {
    // ---- fetch the next tile from the queue and stage its job descriptor --------------------
    if (tid == 0) {
      long long tile = (long long)atomicAdd(p.queue, 1ULL);
      *tile_slot = tile;
      if (tile < p.n_tiles) {
        int j = 0;
        while (j + 1 < p.n_jobs && p.jobs[j + 1].tile_begin <= tile) ++j;
        *jobp = p.jobs[j];
      }
    }
    __syncthreads();
    const long long tile = *tile_slot;
    if (tile >= p.n_tiles) break;
    const FwdJob& job = *jobp;
    cfg.tab = job.tab;
    const S* y0 = reinterpret_cast<const S*>(job.y0);
    const S* gptr = reinterpret_cast<const S*>(job.g);
    const S* eptr = reinterpret_cast<const S*>(job.e_rev);
    const S* dptr = reinterpret_cast<const S*>(job.data);
    S* y_out = reinterpret_cast<S*>(job.y_out);
    S* i_out = reinterpret_cast<S*>(job.i_out);
    S* ckpt_y = reinterpret_cast<S*>(job.ckpt_y);
    const bool observe = (job.v_out != nullptr) && (job.i_out != nullptr || job.loss_out != nullptr);
    const long long jB = job.B;
    const int T = job.T;
    const long long b = (tile - job.tile_begin) * M + tid;  // trajectory owned by this thread
    const bool valid = owner && b < jB;
    S g_b = (S)1, e_b = (S)job.e_scalar;

    // ---- output sample writer -----------------------------------------------------------
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
          obs[2 * tid] += diff * diff;
          obs[2 * tid + 1] += fabs(diff);
        }
      }
    };
    auto ckpt = [&](int step, const Lane<S>& lane) -> bool {
      if (!job.ckpt_t) return true;
      if (step >= job.ckpt_cap) return false;
      size_t o = (size_t)step * jB + b;
      double2 tt;
      tt.x = lane.t0; tt.y = lane.dt;
      *reinterpret_cast<double2*>(job.ckpt_t + 2 * o) = tt;
      S buf[kCkptVals];
      ckpt_pack<S>(lane, buf);
      typedef typename Vec2<S>::type V2;
      V2* dst = reinterpret_cast<V2*>(ckpt_y + (size_t)kCkptVals * o);
#pragma unroll
      for (int i = 0; i < kCkptVals / 2; ++i) {
        V2 v;
        v.x = buf[2 * i]; v.y = buf[2 * i + 1];
        dst[i] = v;
      }
      return true;
    };

    if (owner) {
      Lane<S>& L = lanes[tid];
      S ya = (S)0, yr = (S)1;
      if (valid) {
        ya = y0[2 * b]; yr = y0[2 * b + 1];
        if (gptr) g_b = gptr[b];
        if (eptr) e_b = eptr[b];
      }
      lane_reset<S>(L, ya, yr, job.t_out[0], valid);
      obs[2 * tid] = 0.0; obs[2 * tid + 1] = 0.0;
      if (valid) emit(0, ya, yr);
    }

    double nv, ain;
    if (p.method == 0) {
      // ================================ dopri5 ===========================================
      if (owner) {
        init_prepare_f0<S>(lanes[tid], cfg, &nv, &ain);
        sm.xin[tid] = (W)nv; sm.xin[M + tid] = (W)ain;
      }
      W out = mlp_tile_forward<W, TN>(p.mlp, sm, pp, M, p.MG, p.NG);
      if (owner) {
        Lane<S>& L = lanes[tid];
        init_store_f0<S>(L, cfg, (double)out);
        if (cfg.first_step > 0) {
          L.dt = cfg.first_step;
        } else {
          init_prepare_f1<S>(L, cfg, &nv, &ain);
          sm.xin[tid] = (W)nv; sm.xin[M + tid] = (W)ain;
        }
      }
      if (!(cfg.first_step > 0)) {
        out = mlp_tile_forward<W, TN>(p.mlp, sm, pp, M, p.MG, p.NG);
        if (owner) init_store_f1<S>(lanes[tid], cfg, (double)out);
      }
      if (owner && T <= 1 && lane_active(lanes[tid])) lanes[tid].status = LANE_DONE;

      while (true) {
        int act = 0;
        if (owner) {
          dp_check_before_step<S>(lanes[tid], cfg);
          act = lane_active(lanes[tid]) ? 1 : 0;
        }
        if (!__syncthreads_or(act)) break;
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          if (owner) {
            dp_prepare_stage<S>(lanes[tid], cfg, s, &nv, &ain);
            sm.xin[tid] = (W)nv; sm.xin[M + tid] = (W)ain;
          }
          out = mlp_tile_forward<W, TN>(p.mlp, sm, pp, M, p.MG, p.NG);
          if (owner) dp_store_stage<S>(lanes[tid], cfg, s, (double)out);
        }
        if (owner) dp_finish_step<S>(lanes[tid], cfg, job.t_out, T, emit, ckpt);
      }
    } else {
      // ================================ rk4 (3/8 rule) ====================================
      if (owner && T <= 1 && lane_active(lanes[tid])) lanes[tid].status = LANE_DONE;
      for (int gi = 0; gi + 1 < job.G; ++gi) {
        const double g0 = job.grid[gi], g1 = job.grid[gi + 1];
        int act = owner && lane_active(lanes[tid]) ? 1 : 0;
        if (!__syncthreads_or(act)) break;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          if (owner) {
            rk4_prepare_stage<S>(lanes[tid], cfg, s, g0, g1, p.time_f32 != 0, p.rk4_perturb != 0,
                                 &nv, &ain);
            sm.xin[tid] = (W)nv; sm.xin[M + tid] = (W)ain;
          }
          W out = mlp_tile_forward<W, TN>(p.mlp, sm, pp, M, p.MG, p.NG);
          if (owner) rk4_store_stage<S>(lanes[tid], cfg, s, (double)out);
        }
        if (owner) {
          Lane<S>& L = lanes[tid];
          if (lane_active(L)) {
            // step checkpoint for the backward sweep: (g0, g1), y0, k1..k4
            L.t0 = g0; L.dt = g1;
            if (!ckpt(L.n_acc, L)) L.status = LANE_CKPT_OVERFLOW;
          }
          rk4_finish_step<S>(L, g0, g1, p.time_f32 != 0, job.t_out, T, emit);
        }
      }
    }

    if (valid) {
      const Lane<S>& L = lanes[tid];
      job.stats_out[4 * b + 0] = L.n_acc;
      job.stats_out[4 * b + 1] = L.n_rej;
      job.stats_out[4 * b + 2] = L.nfe;
      job.stats_out[4 * b + 3] = L.status == LANE_DONE ? 0 : L.status;
      if (job.loss_out) {
        job.loss_out[2 * b] = obs[2 * tid];
        job.loss_out[2 * b + 1] = obs[2 * tid + 1];
      }
    }
    __syncthreads();  // lanes[] / job slot are re-initialised by the next tile
  }
  mlp_pipe_drain<W>(p.mlp, sm, pp);
}


// =============================================================================================
// Lane-pool forward kernel (dopri5).  Every CTA owns M lane SLOTS; a slot whose trajectory has
// finished pulls the next trajectory -- of whatever job -- from one global queue (jobs longest
// first), so no slot idles while its neighbours finish and there is no tile/wave quantisation:
// the only tail is the very end of the launch.  Lanes stay in lock-step by ROUND of six RHS
// evaluations: a lane either attempts one dopri5 step (6 stages) or, right after a refill, runs
// its two start-up evaluations f(t0, y0) and the initial-step probe (then idles 4 slots of ~10^3).
// Results are identical to the tile-scheduled kernel: every lane's arithmetic is independent of
// its slot and of its neighbours.
// =============================================================================================
// Job descriptor j.  Every lane reads its job every evaluation: from the kernel parameters
// (constant cache) when they fit, else from global memory (all SMs then hammer the same L2 lines).
__device__ __forceinline__ FwdJob load_job(const FwdParams& p, int j) {
  if (p.jobs_are_inline) return p.jobs_inline[j];
  return p.jobs[j];
}
__device__ __forceinline__ long long job_traj_begin(const FwdParams& p, int j) {
  return p.jobs_are_inline ? p.jobs_inline[j].traj_begin : p.jobs[j].traj_begin;
}

template <typename S>
struct LaneAux {
  long long b;    // trajectory index inside its job
  int job;
  int mode;       // 0 empty, 1 start-up pending, 2 stepping
  S g, e;
};
enum { POOL_EMPTY = 0, POOL_INIT = 1, POOL_STEP = 2 };

template <typename S, typename W, int TN>
__global__ void __launch_bounds__(512, 1) ikr_forward_pool_kernel(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int M = p.M;
  const FwdSmemLayout<S, W> lay(M, p.mlp.npad, p.mlp.kc, p.mlp.L);
  Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
  double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);  // [M][2] sse, sae
  LaneAux<S>* aux = reinterpret_cast<LaneAux<S>*>(smem_raw + lay.off_aux);
  MlpSmem<W> sm;
  sm.full = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  sm.empty = sm.full + kStages;
  sm.xin = reinterpret_cast<W*>(smem_raw + lay.off_xin);
  sm.Hs = reinterpret_cast<W*>(smem_raw + lay.off_hs);
  sm.Wr = reinterpret_cast<W*>(smem_raw + lay.off_wr);
  sm.sp = reinterpret_cast<W*>(smem_raw + lay.off_sp);

  mlp_pipe_init<W>(sm, p.n_worker_warps);
  mlp_stage_small<W>(p.mlp, sm);
  __syncthreads();
  MlpPipe pp;
  mlp_pipe_start<W>(p.mlp, sm, pp);

  const bool owner = tid < M;
  const bool heuristic = !(p.cfg.first_step > 0);
  if (owner) {
    lane_reset<S>(lanes[tid], (S)0, (S)1, 0.0, false);
    aux[tid].mode = POOL_EMPTY; aux[tid].job = 0; aux[tid].b = 0;
    aux[tid].g = (S)1; aux[tid].e = (S)0;
  }
  bool queue_dry = false;

  while (true) {
    // ---- round boundary: retire finished trajectories, refill free slots -----------------------
    int act = 0;
    if (owner) {
      Lane<S>& L = lanes[tid];
      LaneAux<S>& A = aux[tid];
      if (A.mode == POOL_STEP) {
        SolverCfg c = p.cfg;
        dp_check_before_step<S>(L, c);
      }
      if (A.mode != POOL_EMPTY && !lane_active(L)) {
        const FwdJob job = load_job(p, A.job);
        int* st = job.stats_out + 4 * A.b;
        st[0] = L.n_acc; st[1] = L.n_rej; st[2] = L.nfe;
        st[3] = L.status == LANE_DONE ? 0 : L.status;
        if (job.loss_out) {
          job.loss_out[2 * A.b] = obs[2 * tid];
          job.loss_out[2 * A.b + 1] = obs[2 * tid + 1];
        }
        A.mode = POOL_EMPTY;
      }
      if (A.mode == POOL_EMPTY && !queue_dry) {
        const long long gidx = (long long)atomicAdd(p.queue, 1ULL);
        if (gidx < p.n_traj) {
          int j = 0;
          while (j + 1 < p.n_jobs && job_traj_begin(p, j + 1) <= gidx) ++j;
          const FwdJob job = load_job(p, j);
          const long long b = gidx - job.traj_begin;
          const S* y0 = reinterpret_cast<const S*>(job.y0);
          A.job = j; A.b = b; A.mode = POOL_INIT;
          A.g = job.g ? reinterpret_cast<const S*>(job.g)[b] : (S)1;
          A.e = job.e_rev ? reinterpret_cast<const S*>(job.e_rev)[b] : (S)job.e_scalar;
          lane_reset<S>(L, y0[2 * b], y0[2 * b + 1], job.t_out[0], true);
          obs[2 * tid] = 0.0; obs[2 * tid + 1] = 0.0;
        } else {
          queue_dry = true;
        }
      }
      act = A.mode != POOL_EMPTY ? 1 : 0;
    }
    if (!__syncthreads_or(act)) break;

    // ---- one round = six RHS evaluations --------------------------------------------------------
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
      int what = 0;   // 0 masked, 1 dopri5 stage, 2 f0, 3 initial-step probe
      if (owner) {
        Lane<S>& L = lanes[tid];
        const LaneAux<S>& A = aux[tid];
        double nv = 0, ain = 0;
        if (A.mode == POOL_STEP) what = 1;
        else if (A.mode == POOL_INIT && s == 0) what = 2;
        else if (A.mode == POOL_INIT && s == 1 && heuristic) what = 3;
        if (what) {
          SolverCfg c = p.cfg;
          c.tab = load_job(p, A.job).tab;
          if (what == 1) dp_prepare_stage<S>(L, c, s, &nv, &ain);
          else if (what == 2) init_prepare_f0<S>(L, c, &nv, &ain);
          else init_prepare_f1<S>(L, c, &nv, &ain);
        }
        sm.xin[tid] = (W)nv; sm.xin[M + tid] = (W)ain;
      }
      const W out = mlp_tile_forward<W, TN>(p.mlp, sm, pp, M, p.MG, p.NG);
      if (what) {
        Lane<S>& L = lanes[tid];
        SolverCfg c = p.cfg;
        if (what == 1) dp_store_stage<S>(L, c, s, (double)out);
        else if (what == 2) {
          init_store_f0<S>(L, c, (double)out);
          if (!heuristic) L.dt = c.first_step;
        } else init_store_f1<S>(L, c, (double)out);
      }
    }

    // ---- end of round: finish the attempted step / leave start-up ---------------------------------
    if (owner) {
      Lane<S>& L = lanes[tid];
      LaneAux<S>& A = aux[tid];
      if (A.mode == POOL_STEP) {
        const FwdJob job = load_job(p, A.job);
        const long long jB = job.B, b = A.b;
        S* y_out = reinterpret_cast<S*>(job.y_out);
        S* i_out = reinterpret_cast<S*>(job.i_out);
        S* ckpt_y = reinterpret_cast<S*>(job.ckpt_y);
        const S* dptr = reinterpret_cast<const S*>(job.data);
        const bool observe = (job.v_out != nullptr) && (job.i_out != nullptr || job.loss_out != nullptr);
        const S g_b = A.g, e_b = A.e;
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
              obs[2 * tid] += diff * diff;
              obs[2 * tid + 1] += fabs(diff);
            }
          }
        };
        auto ckpt = [&](int step, const Lane<S>& lane) -> bool {
          if (!job.ckpt_t) return true;
          if (step >= job.ckpt_cap) return false;
          size_t o = (size_t)step * jB + b;
          double2 tt;
          tt.x = lane.t0; tt.y = lane.dt;
          *reinterpret_cast<double2*>(job.ckpt_t + 2 * o) = tt;
          S buf[kCkptVals];
          ckpt_pack<S>(lane, buf);
          typedef typename Vec2<S>::type V2;
          V2* dst = reinterpret_cast<V2*>(ckpt_y + (size_t)kCkptVals * o);
#pragma unroll
          for (int i = 0; i < kCkptVals / 2; ++i) {
            V2 v;
            v.x = buf[2 * i]; v.y = buf[2 * i + 1];
            dst[i] = v;
          }
          return true;
        };
        SolverCfg c = p.cfg;
        dp_finish_step<S>(L, c, job.t_out, job.T, emit, ckpt);
      } else if (A.mode == POOL_INIT) {
        // start-up done: emit y(t[0]) = y0 and start stepping (or finish if there is one output)
        const FwdJob job = load_job(p, A.job);
        const long long b = A.b;
        if (job.y_out) {
          typename Vec2<S>::type v;
          v.x = L.ya; v.y = L.yr;
          reinterpret_cast<typename Vec2<S>::type*>(job.y_out)[b] = v;
        }
        if (job.v_out && (job.i_out || job.loss_out)) {
          double cur = (double)(A.g * L.ya * L.yr) * (job.v_out[0] - (double)A.e);
          if (job.i_out) reinterpret_cast<S*>(job.i_out)[b] = (S)cur;
          if (job.data) {
            const S* dptr = reinterpret_cast<const S*>(job.data);
            double diff = cur - (double)dptr[job.data_B == 1 ? 0 : b];
            obs[2 * tid] += diff * diff;
            obs[2 * tid + 1] += fabs(diff);
          }
        }
        A.mode = POOL_STEP;
        if (job.T <= 1 && lane_active(L)) L.status = LANE_DONE;
      }
    }
  }
  mlp_pipe_drain<W>(p.mlp, sm, pp);
}

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
__global__ void ikr_interp_kernel(ProtocolTable tab, const double* tq, long long T, double* v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) {
    double out;
    table_voltage(tab, tq[i], &out);
    v[i] = out;
  }
}

template <typename W>
__global__ void ikr_fma_peak_kernel(W* sink, long long iters) {
  // 16 independent accumulator chains per thread: enough ILP to saturate the FMA pipe
  W a[16];
  W x = (W)(1.0 + 1e-7 * threadIdx.x), y = (W)(1e-9 * blockIdx.x);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (W)i;
  for (long long it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = ikr_fma(a[i], x, y);
  }
  W s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == (W)123.456) sink[0] = s;
}

// FFMA2 (packed fp32 pair) variant: counts 2 FMA per instruction
__global__ void ikr_fma2_peak_kernel(float* sink, long long iters) {
  f32x2 a[16];
  const f32x2 x = f2_pack(1.0f + 1e-7f * threadIdx.x, 1.0f - 1e-7f * threadIdx.x);
  const f32x2 y = f2_pack(1e-9f * blockIdx.x, 2e-9f * blockIdx.x);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = f2_pack((float)i, (float)-i);
  for (long long it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float lo, hi; f2_unpack(a[i], lo, hi); s += lo + hi; }
  if (s == 123.456f) sink[0] = s;
}

}  // namespace ikr
#endif  // IKR_FORWARD_CUH_
