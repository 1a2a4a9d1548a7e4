// ikr_forward_tc_pp.cuh -- two-tile ("ping-pong") lane-pool forward kernel on the tensor cores.
//
// One RHS evaluation of a 128-trajectory tile is a serial chain: solver stage -> layer 0 -> L layer
// passes on the tensor pipe -> output reduction -> solver stage ...  With one tile per CTA the
// tensor pipe idles through everything that is not a layer pass (measured: ~8 k of ~45 k cycles per
// evaluation for the evaluation boundary alone).  This kernel gives every CTA TWO tiles, X and Y,
// whose evaluations strictly alternate on the tensor pipe (X1..XL Y1..YL X1..XL ...): while the MMA
// warp runs the passes of Y, the owners of X collect their outputs, run their solver stage and
// publish the next inputs, and vice versa.  TMEM does not need a second tile's worth of columns:
// between its evaluations a tile holds NOTHING in TMEM (its last D is reduced to one scalar per
// lane), so the two D accumulators and the A-operand unit ring are time-shared, exactly as within
// one tile.
//
// Threads: 128 owners of X (group 0), 128 owners of Y (group 1), 128 helper threads (group 2), the MMA
// warp and the weight-producer warp of ikr_forward_tc.cuh -- unchanged: they only see a stream of
// passes.  An evaluation of tile Z is produced by TWO column groups, owner(Z) and the helper (units
// u = 0, 2, 4, ... / 1, 3, 5, ...), i.e. it IS the G = 2 evaluation of tc_mlp_eval: results are
// bit-identical to the single-tile kernels run with two column groups.  At a tile switch the helper
// produces its share of the next evaluation's first pass BEFORE it reduces the finished tile's output.  The other owner group takes
// no part in it -- that is when it runs its solver.
//
// Hand-shakes (mbarriers, all phases consumed in order by every waiter):
//   xin_ready[Z]   4 arrivals (owner warps): inputs (nv, a) of Z's next evaluation are in shared
//                  memory -- or Z is finished (dead[Z] set before the last arrival).  Waited by the
//                  helper (every evaluation) and by the OTHER owner (once per evaluation of Z, to keep
//                  its pass / ring counters in step with the global sequence).
//   d_last[Z]      8 arrivals (owner + helper warps, right after the D of Z's LAST pass is there): every
//                  pass of Z's evaluation is complete.  The other owner's first unit store waits for
//                  it -- the ring barriers tell only neighbouring phases apart, so a thread that sat
//                  out a whole evaluation may not touch them before the sequence has caught up.
//   part_ready[Z]  8 arrivals (owner + helper warps): partial output sums of Z's evaluation are in
//                  shared memory and nobody reads Z's last D any more.  Waited by owner(Z), and by the
//                  other owner before it produces the units of ITS second pass (the MMAs of that pass
//                  overwrite the D buffer Z's output reduction read).
// Lane slots refill from the global trajectory queue as in ikr_forward_tc_pool_kernel; a tile whose
// slots are all empty with the queue dry marks itself dead and leaves the alternation.
#ifndef IKR_FORWARD_TC_PP_CUH_
#define IKR_FORWARD_TC_PP_CUH_

#include "ikr_forward_tc.cuh"

namespace ikr {

template <typename S>
struct TcPpSmemLayout {
  size_t off_bar, off_pp, off_misc, off_job, off_lanes, off_obs, off_aux, off_xin, off_part, off_sp, off_ring, total;
  __host__ __device__ TcPpSmemLayout(const TcGeom& g, int stages) {
    size_t o = 0;
    off_bar = o; o += (size_t)(2 * kTcMaxStages + 2 * 8 + 2) * 8;    // engine barriers (tc_engine_ctx)
    off_pp = o; o += 6 * 8;                                            // xin_ready[2], part_ready[2], d_last[2]
    off_misc = o; o += 32;                                             // tmem base, stop flag, dead[2]
    off_job = o; o += (size_t)kInlineJobs * ((sizeof(FwdJob) + 15) & ~(size_t)15);
    off_lanes = o; o += (size_t)2 * kTcM * sizeof(Lane<S>); o = (o + 15) & ~(size_t)15;
    off_obs = o; o += (size_t)2 * kTcM * 2 * sizeof(double);
    off_aux = o; o += (size_t)2 * kTcM * 32;
    off_xin = o; o += (size_t)2 * kTcM * 2 * sizeof(float);
    off_part = o; o += (size_t)2 * 2 * kTcM * sizeof(float);
    off_sp = o; o += (size_t)g.small_elems * sizeof(float); o = (o + 127) & ~(size_t)127;
    off_ring = o; o += (size_t)stages * g.stage_bytes;
    total = o;
  }
};

// a whole evaluation of the OTHER tile went by: keep this thread's view of the global pass sequence
template <int TERMS>
__device__ __forceinline__ void tc_pp_skip_eval(const TcGeom& g, TcLane& tl) {
  const int UT = g.units + g.tail;
  for (int l = 0; l < g.L; ++l) tc_pass_advance<TERMS>(tl, UT);
  tl.n_pass += (unsigned)g.L;
  tl.phase_d ^= (unsigned)(g.L & 1);
}

__device__ __forceinline__ int pp_owners_or(int pred, int tile) {
  uint32_t r;
  if (tile == 0) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.or.pred p, 2, 128, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r) : "r"(pred) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.or.pred p, 3, 128, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r) : "r"(pred) : "memory");
  }
  return (int)r;
}

template <typename S, int TERMS>
__global__ void __launch_bounds__(tc_threads(3), 1) ikr_forward_tc_pp_kernel(const TcFwdParams tp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FwdParams& p = tp.f;
  const TcGeom g = tp.g;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int kMmaWarp = 12, kLoadWarp = 13;
  const TcPpSmemLayout<S> lay(g, g.stages);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bar);
  uint64_t* xin_ready = reinterpret_cast<uint64_t*>(smem_raw + lay.off_pp);      // [2]
  uint64_t* part_ready = xin_ready + 2;                                            // [2]
  uint64_t* d_last = xin_ready + 4;                                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + lay.off_misc);
  volatile int* stop_flag = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 4);
  volatile int* dead = reinterpret_cast<volatile int*>(smem_raw + lay.off_misc + 8);   // [2]
  float* sp = reinterpret_cast<float*>(smem_raw + lay.off_sp);
  float* xin_all = reinterpret_cast<float*>(smem_raw + lay.off_xin);              // [2][128][2]
  float* part_all = reinterpret_cast<float*>(smem_raw + lay.off_part);            // [2][2][128]
  const TcEngineCtx eng = tc_engine_ctx(bars, stop_flag, smem_raw + lay.off_ring);

  if (tid == 0) {
    tc_engine_init(eng, g.stages);
    for (int z = 0; z < 2; ++z) {
      mbar_init(&xin_ready[z], 4);
      mbar_init(&part_ready[z], 8);
      mbar_init(&d_last[z], 8);
      dead[z] = 0;
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  {
    const float* P = reinterpret_cast<const float*>(p.mlp.base);
    const int NP = g.NP, npad = p.mlp.npad, n = g.n;
    for (int i = tid; i < g.small_elems; i += tc_threads(3)) {
      const int row = i / NP, c = i - row * NP;
      sp[i] = tc_small_param<TERMS>(g, p.mlp, P, tp.scales, i, row, c, NP, npad, n);
    }
  }
  constexpr size_t kJobStride = (sizeof(FwdJob) + 15) & ~(size_t)15;
  unsigned char* sjobs = smem_raw + lay.off_job;
  if (p.jobs_are_inline) {
    const int words = (int)(sizeof(FwdJob) / 4);
    for (int i = tid; i < p.n_jobs * words; i += tc_threads(3)) {
      const int j = i / words, w = i - j * words;
      reinterpret_cast<uint32_t*>(sjobs + (size_t)j * kJobStride)[w] =
          reinterpret_cast<const uint32_t*>(&p.jobs_inline[j])[w];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == kMmaWarp) {
    tc_mma_warp<TERMS>(g, eng, tbase, tp.timing && blockIdx.x == 0);
  } else if (warp == kLoadWarp) {
    if ((tid & 31) == 0) tc_producer_thread(g, eng, reinterpret_cast<const unsigned char*>(tp.img),
                                            (unsigned)(g.L * g.KST));
  } else {
    const int grp = warp >> 2;          // 0: owners of X, 1: owners of Y, 2: helpers
    TcLane tl;
    tl.lane = tid & 127;
    tl.taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    tc_lane_attach(tl, eng);
    tl.sp = sp;
    tl.slope = (float)p.mlp.slope;
    tl.c_l0 = tl.c_wait = tl.c_epi = 0;
    tl.c_sync_a = tl.c_sync_b = tl.c_last = 0;
#ifdef IKR_TC_TRACE
    tl.trace_eval = 0;
#endif

    if (grp == 2) {
      // ============================ helper group: column group 1 of EVERY evaluation ================
      // Software-pipelined across the tile switch: after the D of tile Z's last pass is there, the
      // helper first produces its units of the NEXT evaluation's first pass (layer 0 of the other
      // tile, whose inputs were published long ago) and only then reduces Z's output -- the MMAs of
      // the next pass must not wait for an output reduction.  When the other tile is gone the next
      // evaluation is Z's own, whose inputs need Z's output: classic order.
      tl.group = 1;
      unsigned ph[2] = {0u, 0u};
      bool gone[2] = {false, false};
      long long c0 = clock64();
      // next_alive(Z): wait for tile Z's turn; false when Z has left the alternation
      auto turn = [&](int Z) -> bool {
        if (gone[Z]) return false;
        mbar_wait_backoff(&xin_ready[Z], ph[Z], 32);
        ph[Z] ^= 1u;
        if (dead[Z]) { gone[Z] = true; return false; }
        return true;
      };
      auto bind = [&](int Z) {
        tl.xin = xin_all + (size_t)Z * 2 * kTcM;
        tl.part = part_all + (size_t)Z * 2 * kTcM;
        tl.last_d_bar = &d_last[Z];
      };
      int Z = 0;
      bool running = turn(0);
      if (!running) { Z = 1; running = turn(1); }
      if (running) {
        bind(Z);
        tc_eval_layer0<2, TERMS>(g, tl);
      }
      while (running) {
        const uint32_t dcol = tc_eval_hidden<2, TERMS>(g, tl, c0);     // evaluation of tile Z
        float* part_z = tl.part;
        const int O = Z ^ 1;
        if (turn(O)) {
          // pipelined switch: first pass of O's evaluation, then Z's output
          bind(O);
          tc_eval_layer0<2, TERMS>(g, tl);
          tl.part = part_z;
          tc_eval_output<2, TERMS>(g, tl, dcol);
          __syncwarp();
          if ((tl.lane & 31) == 0) mbar_arrive(&part_ready[Z]);
          tl.part = part_all + (size_t)O * 2 * kTcM;
          Z = O;
        } else {
          tc_eval_output<2, TERMS>(g, tl, dcol);
          __syncwarp();
          if ((tl.lane & 31) == 0) mbar_arrive(&part_ready[Z]);
          running = turn(Z);
          if (running) {
            bind(Z);
            tc_eval_layer0<2, TERMS>(g, tl);
          }
        }
      }
      // both tiles are finished: release the engine warps
      if (tl.lane == 0) *stop_flag = 1;
      asm volatile("bar.sync 4, 128;" ::: "memory");
      __syncwarp();
      if ((tl.lane & 31) == 0) mbar_arrive(&tl.unit_ready[0]);   // wakes the MMA warp, which sees the flag
    } else {
      // ============================ owner group of tile Z = grp ======================================
      const int Z = grp, O = grp ^ 1;
      tl.group = 0;
      tl.xin = xin_all + (size_t)Z * 2 * kTcM;
      tl.part = part_all + (size_t)Z * 2 * kTcM;
      const int slot = Z * kTcM + tl.lane;      // index into the per-slot shared arrays
      Lane<S>* lanes = reinterpret_cast<Lane<S>*>(smem_raw + lay.off_lanes);
      double* obs = reinterpret_cast<double*>(smem_raw + lay.off_obs);
      LaneAux<S>* aux = reinterpret_cast<LaneAux<S>*>(smem_raw + lay.off_aux);
      long long c_acct = 0, c_eval = 0, c_collect = 0, c_opart = 0;
      const long long c_begin = clock64();
      unsigned ph_other_xin = 0u, ph_other_part = 0u, ph_other_last = 0u, ph_my_part = 0u;
      bool other_gone = false, first_eval = true;
      tl.last_d_bar = &d_last[Z];

      // One evaluation of this tile, interleaved with the other tile's (see the file header)
      auto owner_eval = [&](float nv, float a, auto hook) -> float {
        bool wait_other_part = false;
        long long ct = clock64();
        if (!other_gone && !(first_eval && Z == 0)) {
          // the other tile's evaluation that precedes this one in the global sequence
          mbar_wait_backoff(&xin_ready[O], ph_other_xin, 32);
          ph_other_xin ^= 1u;
          if (dead[O]) other_gone = true;
          else {
            tc_pp_skip_eval<TERMS>(g, tl);
            wait_other_part = true;
            // first unit store of this evaluation: only after the other tile's last pass is complete
            tl.gate_bar = &d_last[O];
            tl.gate_parity = ph_other_last;
            tl.gate_pending = 1;
            ph_other_last ^= 1u;
          }
        }
        first_eval = false;
        { const long long c1 = clock64(); c_acct += c1 - ct; ct = c1; }
        *reinterpret_cast<float2*>(tl.xin + 2 * tl.lane) = make_float2(nv, a);
        __syncwarp();
        if ((tl.lane & 31) == 0) mbar_arrive(&xin_ready[Z]);
        tc_mlp_eval<2, TERMS>(g, tl, [&]() {
          // before the units of pass 2 (its MMAs overwrite the D the other tile's output reduction read)
          if (wait_other_part) {
            const long long c0 = clock64();
            mbar_wait(&part_ready[O], ph_other_part);
            ph_other_part ^= 1u;
            c_opart += clock64() - c0;
          }
          hook();
        });
        { const long long c1 = clock64(); c_eval += c1 - ct; ct = c1; }
        __syncwarp();
        if ((tl.lane & 31) == 0) mbar_arrive(&part_ready[Z]);
        mbar_wait(&part_ready[Z], ph_my_part);
        ph_my_part ^= 1u;
        c_collect += clock64() - ct;
        return tl.part[tl.lane] + tl.part[kTcM + tl.lane] + tl.sp[(size_t)(4 + g.L) * g.NP];
      };

      const bool heuristic = !(p.cfg.first_step > 0);
      Lane<S>& L = lanes[slot];
      LaneAux<S>& A = aux[slot];
      SolverCfg c = p.cfg;
      auto job_of = [&](int j) -> const FwdJob* {
        return p.jobs_are_inline ? reinterpret_cast<const FwdJob*>(sjobs + (size_t)j * kJobStride) : p.jobs + j;
      };
      const FwdJob* jobp = job_of(0);
      TimeCache tcache;
      tcache.valid = 0;
      lane_reset<S>(L, (S)0, (S)1, 0.0, false);
      A.mode = POOL_EMPTY; A.job = 0; A.b = 0; A.g = (S)1; A.e = (S)0;
      bool queue_dry = false;
      bool range_hit = false;
      int range_rejects = 0;

      while (true) {
        // ---- round boundary: retire finished trajectories, refill free slots ----------------------
        if (A.mode == POOL_STEP) dp_check_before_step<S>(L, c);
        if (A.mode != POOL_EMPTY && !lane_active(L)) {
          const FwdJob& job = *jobp;
          int* st = job.stats_out + 4 * A.b;
          st[0] = L.n_acc; st[1] = L.n_rej; st[2] = L.nfe;
          st[3] = L.status == LANE_DONE ? 0 : L.status;
          if (job.loss_out) {
            job.loss_out[2 * A.b] = obs[2 * slot];
            job.loss_out[2 * A.b + 1] = obs[2 * slot + 1];
          }
          A.mode = POOL_EMPTY;
        }
        if (A.mode == POOL_EMPTY && !queue_dry) {
          const long long gidx = (long long)atomicAdd(p.queue, 1ULL);
          if (gidx < p.n_traj) {
            int j = 0;
            while (j + 1 < p.n_jobs && job_of(j + 1)->traj_begin <= gidx) ++j;
            jobp = job_of(j);
            const FwdJob& job = *jobp;
            c.tab = job.tab;
            const long long b = gidx - job.traj_begin;
            const S* y0 = reinterpret_cast<const S*>(job.y0);
            A.job = j; A.b = b; A.mode = POOL_INIT;
            A.g = job.g ? reinterpret_cast<const S*>(job.g)[b] : (S)1;
            A.e = job.e_rev ? reinterpret_cast<const S*>(job.e_rev)[b] : (S)job.e_scalar;
            lane_reset<S>(L, y0[2 * b], y0[2 * b + 1], job.t_out[0], true);
            range_hit = false; range_rejects = 0;
            obs[2 * slot] = 0.0; obs[2 * slot + 1] = 0.0;
          } else {
            queue_dry = true;
          }
        }
        if (!pp_owners_or(A.mode != POOL_EMPTY ? 1 : 0, Z)) break;

        // ---- one round = six RHS evaluations ---------------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          int what = 0;   // 0 masked, 1 dopri5 stage, 2 f0, 3 initial-step probe
          double nv = 0, ain = 0;
          if (A.mode == POOL_STEP) what = 1;
          else if (A.mode == POOL_INIT && s == 0) what = 2;
          else if (A.mode == POOL_INIT && s == 1 && heuristic) what = 3;
          if (what) {
            if (what == 1) dp_prepare_stage_cached<S>(L, c, s, &nv, &ain, tcache);
            else if (what == 2) init_prepare_f0<S>(L, c, &nv, &ain);
            else init_prepare_f1<S>(L, c, &nv, &ain);
          }
          float out = owner_eval((float)nv, (float)ain, [&]() {
            if (what == 1 && s < 5) dp_prefetch_stage_time<S>(L, c, s + 1, tcache);
          });
          if (what) out = tc_range_filter<TERMS>(out, nv, ain, range_hit);
          if (what == 1) dp_store_stage<S>(L, c, s, (double)out);
          else if (what == 2) {
            init_store_f0<S>(L, c, (double)out);
            if (!heuristic) L.dt = c.first_step;
          } else if (what == 3) init_store_f1<S>(L, c, (double)out);
        }

        // ---- end of round: finish the attempted step / leave start-up ----------------------------------
        if (A.mode == POOL_STEP) {
          const FwdJob& job = *jobp;
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
                obs[2 * slot] += diff * diff;
                obs[2 * slot + 1] += fabs(diff);
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
          const int acc0 = L.n_acc;
          dp_finish_step<S>(L, c, job.t_out, job.T, emit, ckpt);
          tc_range_after_step<TERMS, S>(L, L.n_acc != acc0, range_hit, range_rejects);
        } else if (A.mode == POOL_INIT) {
          tc_range_fatal<TERMS, S>(L, range_hit);      // f0 / the initial-step probe sit at y0
          const FwdJob& job = *jobp;
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
              obs[2 * slot] += diff * diff;
              obs[2 * slot + 1] += fabs(diff);
            }
          }
          A.mode = POOL_STEP;
          if (job.T <= 1 && lane_active(L)) L.status = LANE_DONE;
        }
      }
      if (tp.timing && blockIdx.x == 0 && tl.lane == 0) {
        const long long tot = clock64() - c_begin;
        printf("[pp timing] owner of tile %d: total %lld cycles: accounting wait %lld, evaluation %lld (gate wait %lld, "
               "other-part wait %lld, wait_d %lld, layer0 %lld, epilogue %lld), collect %lld, solver+boundary %lld\n",
               Z, tot, c_acct, c_eval, tl.c_gate, c_opart, tl.c_wait, tl.c_l0, tl.c_epi, c_collect,
               tot - c_acct - c_eval - c_collect);
      }
      // this tile is finished: leave the alternation (the helper and the other owner see dead[Z] at
      // this tile's next turn)
      if (tl.lane == 0) dead[Z] = 1;
      if (Z == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
      else asm volatile("bar.sync 3, 128;" ::: "memory");
      __syncwarp();
      if ((tl.lane & 31) == 0) mbar_arrive(&xin_ready[Z]);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

}  // namespace ikr
#endif  // IKR_FORWARD_TC_PP_CUH_
