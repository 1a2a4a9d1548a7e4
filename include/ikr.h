/*
 * ikr.h -- C ABI of the B200-native IKr neural-ODE integrator (libikr_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of chonlei/neural-ode-ion-channels:
 * batched `odeint(func, y0, t, method, rtol, atol)` of the hERG NN-f / NN-d models, forward and
 * backward.  Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only; no torch types; every buffer (including
 * the workspace) is owned by the caller; the library keeps no device allocation and no global
 * state; calls are stream-ordered and never synchronise.  Thread safety: any number of host
 * threads may call into the library at the same time, each on its own stream with its own
 * buffers (independent fits of an architecture sweep share one GPU that way).
 *
 * Replaces, on the reference side:
 *   - `torchdiffeq.odeint(func, y0, t[, method='dopri5'])`     train-s1.py:322,327; table-1.py:404,413;
 *                                                               train-r1.py:934,941; train-d0.py:428,436
 *   - `ODEFunc.forward` NN-f / NN-d (the RHS the solver calls)  train-s1.py:231-247; train-d2.py:257-272
 *   - `ODEFunc._v` (protocol interpolation, host scipy)         train-s1.py:218-229
 *   - observation `I = g a r (V - E)` and the loss reductions   train-s1.py:328-329; table-1.py:414-415;
 *                                                               train-d0.py:509 (sum of squares)
 *   - backward through the solver (torchdiffeq autograd)        absent from the reference
 */
#ifndef IKR_H_
#define IKR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IKR_ABI_VERSION 2

/* dtype codes */
#define IKR_F32 0
#define IKR_F64 1
/* method codes */
#define IKR_DOPRI5 0
#define IKR_RK4 1
/* per-trajectory status codes (status_out) -- torchdiffeq's host asserts become lane codes */
#define IKR_OK 0
#define IKR_DT_UNDERFLOW 1  /* "underflow in dt"                     */
#define IKR_MAX_STEPS 2     /* "max_num_steps exceeded"              */
#define IKR_NONFINITE 3     /* "non-finite values in state `y`"      */
#define IKR_CKPT_OVERFLOW 4 /* step-checkpoint capacity too small    */
#define IKR_TC_RANGE 5      /* fp16x2 tensor-core forward: a hidden activation of an ACCEPTED step left
                               the fp16 range (|h| >= 4094) in the physical domain; rerun with reserved bit 8
                               (bf16x3 split) or bit 1 (FFMA2 kernel)                                 */
/* return codes (0 ok, negative = argument / launch error) */
#define IKR_ERR_ARG (-1)
#define IKR_ERR_UNSUPPORTED (-2)
#define IKR_ERR_WORKSPACE (-3)
#define IKR_ERR_LAUNCH (-4)
#define IKR_ERR_DEVICE (-5)

/* Model + solver descriptor (POD, host memory).  Mirrors the attributes of the reference
 * `ODEFunc` (train-s1.py:181-216 / train-d2.py:191-232) and the `odeint` keyword arguments. */
typedef struct ikr_desc {
  int32_t n_layers;       /* hidden Linear(n,n) count L (architectures/sNN.py: n_layers)         */
  int32_t n_nodes;        /* hidden width n             (architectures/sNN.py: n_nodes)          */
  int32_t nn_d;           /* 0: NN-f  da/dt = net/netscale; 1: NN-d  da/dt = HH + net/netscale   */
  int32_t method;         /* IKR_DOPRI5 | IKR_RK4                                                */
  int32_t state_dtype;    /* dtype of y0 / y_out / k (torchdiffeq: y0.dtype)                     */
  int32_t mlp_dtype;      /* dtype of the MLP weights and arithmetic                             */
  int32_t time_f32;       /* rk4: grid arithmetic in fp32 (caller passed an fp32 `t`)            */
  int32_t rk4_perturb;    /* rk4 `perturb` option (default 0)                                    */
  double p[8];            /* p1..p8 (p1..p4 used when nn_d)                                       */
  double vrange;          /* 100   */
  double netscale;        /* 1000  */
  double negative_slope;  /* LeakyReLU slope 0.01 */
  double rtol, atol;      /* dopri5 tolerances (torchdiffeq defaults 1e-7 / 1e-9)                */
  double first_step;      /* <= 0: Hairer initial-step heuristic                                  */
  double safety, ifactor, dfactor; /* 0.9, 10, 0.2                                               */
  int64_t max_num_steps;  /* per output interval, torchdiffeq semantics                          */
  int32_t tile_m;         /* 0: library picks the trajectories-per-CTA tile                      */
  int32_t reserved;       /* option bits (0 = library defaults; every knob lives HERE, the library
                             reads no environment variable and keeps no state between calls):
                             bit 0: dopri5 on the lane-pool kernel (slots refill from a queue);
                             bit 1: keep an fp32 MLP on the FFMA2 kernel (no tensor cores);
                             bit 2: never use lane-pool scheduling (default: automatic on the
                             tensor-core path for launches with more tiles than SMs);
                             bit 3: debug, CTA 0 prints its phase clocks (device printf);
                             bits 4-5: epilogue column groups of the tensor-core kernels, 1..3
                             (0 = default 2; fixes the output-layer summation order, i.e. results);
                             bit 6: never use the two-tile ping-pong kernel; bit 7: use it (experimental,
                             never chosen automatically);
                             bit 8: tensor-core forward with the bf16x3 operand split (six MMAs per
                             fp32 product, any activation range) instead of the default fp16x2
                             split (three MMAs; hidden activations must stay below 65504 / 16 in
                             magnitude -- beyond that trial steps are rejected and, in the physical domain,
                             the lane ends with IKR_TC_RANGE);
                             bit 9: ikr_backward never overlaps the weight-gradient GEMM with the next
                             adjoint round (default: on a second stream when the batch leaves SMs idle);
                             bit 10: the tensor-core adjoint kernel issues all six bf16x3 products per
                             fp32 product (default: three, a1 b1 + a2 b1 + a1 b2, the ~2^-16 per
                             product that the weight-gradient GEMM carries anyway) */
} ikr_desc;

/* One JOB = one protocol table + one batch of trajectories, i.e. the shape of one reference
 * `odeint` call (train-s1.py:326-327).  A forward call takes an array of jobs that share the MLP
 * weights; their trajectory tiles are scheduled through one queue so the whole GPU stays busy.
 * All pointers are DEVICE pointers.  Layouts follow torchdiffeq: y_out is (T, B, 2).            */
typedef struct ikr_io {
  int64_t B;              /* trajectories                                                        */
  int64_t T;              /* output times                                                        */
  int64_t G;              /* rk4: grid points (== T and grid == t_out unless step_size given)    */
  const void* weights;    /* packed MLP parameters (job 0's pointer is used for the launch)      */
  const double* table_t;  /* [table_len] protocol time (ms)                                      */
  const double* table_v;  /* [table_len] protocol voltage (mV)                                   */
  int32_t table_len;      /* protocol table samples                                              */
  int32_t table_uniform;  /* 1: table_t[i] ~= table_t0 + i / table_inv_dt (index hint only)      */
  double table_t0;
  double table_inv_dt;
  double cost_hint;       /* relative cost of one trajectory (e.g. protocol duration); 0 => T    */
  const void* y0;         /* [B,2] state dtype: (a, r)                                           */
  const double* t_out;    /* [T] strictly increasing output times (fp64 copy of `t`)             */
  const double* grid;     /* [G] rk4 step grid (fp64)                                            */
  const double* v_out;    /* [T] V(t_out) for the current / loss epilogue (nullable)             */
  const void* g;          /* [B] state dtype conductance (nullable => 1)                         */
  const void* e_rev;      /* [B] state dtype reversal potential (nullable => use e_scalar)       */
  double e_scalar;
  const void* data;       /* [T, data_B] state dtype measured current (nullable)                 */
  int64_t data_B;         /* 1 (shared trace) or B                                               */
  void* y_out;            /* [T,B,2] state dtype (nullable)                                      */
  void* i_out;            /* [T,B]  state dtype current (nullable; needs v_out)                  */
  double* loss_out;       /* [B,2]  per-trajectory (sum sq. error, sum abs error) (nullable)     */
  int32_t* stats_out;     /* [B,4]  n_accept, n_reject, nfe, status                              */
  /* step checkpoints for ikr_backward (all nullable when no backward is wanted) */
  int64_t ckpt_cap;       /* capacity in accepted steps per trajectory                           */
  double* ckpt_t;         /* [ckpt_cap, B, 2] (t0, dt)                                           */
  void* ckpt_y;           /* [ckpt_cap, B, 16] state dtype (a0, r0, k_a[0..6], k_r[0..6])        */
} ikr_io;

/* Inputs/outputs of the backward sweep (discrete adjoint of the accepted-step sequence: step
 * sizes, stage times and dense-output abscissae are constants, exactly what PyTorch autograd sees
 * through torchdiffeq's non-adjoint odeint).  The forward job passed alongside must have been run
 * with step checkpoints (ckpt_t / ckpt_y) and, for a fused loss, with y_out, v_out and data.     */
typedef struct ikr_bwd_io {
  const void* grad_y;     /* [T,B,2] state dtype dL/dy_out (used when fused_loss == 0)           */
  int32_t fused_loss;     /* 0: use grad_y; 1: L = sum (I - data)^2 ; 2: L = sum |I - data|      */
  int32_t reserved;
  int64_t max_accepted_steps; /* max over trajectories of stats[:,0]; <= 0: assume ckpt_cap      */
  double* grad_weights;   /* [ikr_param_count] fp64, state_dict order (net.0.weight, net.0.bias,
                             net.2.weight, ...): OVERWRITTEN with dL/dtheta summed over B         */
  void* grad_y0;          /* [B,2] state dtype (nullable)                                        */
  void* grad_g;           /* [B]   state dtype dL/dg (nullable; fused loss only)                 */
} ikr_bwd_io;

int ikr_abi_version(void);
const char* ikr_error_string(int code);

/* number of elements (of the MLP dtype) of the packed parameter buffer, and the documented
 * layout: [w0[:,0] | w0[:,1] | b0] (npad each) | L x Wt[k][npad] (forward, K-major)
 * | L x b[npad] | w_last[npad] | b_last (padded to 8) | L x W[o][npad] (backward, original rows).
 * npad = n rounded up to 8.  Source tensors: state_dict keys net.{2i}.weight/.bias
 * (train-s1.py:186-205, 263).                                                                    */
int64_t ikr_packed_weight_elems(const ikr_desc* d);
int64_t ikr_param_count(const ikr_desc* d);

/* out[8] = npad, off_w0, off_wt, off_bh, off_wl, off_wn, total elems, chunk rows kc */
int ikr_packed_layout(const ikr_desc* d, int64_t out[8]);

/* trajectories per CTA the library will use for jobs of the given batch sizes */
int32_t ikr_tile_m(const ikr_desc* d, int32_t n_jobs, const int64_t* B);
/* What ikr_forward will launch for jobs of the given batch sizes (the library's own decision, so
 * that callers report it instead of guessing):
 * out[16] = tile_m, threads/CTA, grid, dynamic smem bytes, n_tiles, kc, chunks/layer (k-steps/layer
 * on the tensor-core path), SM count, scheduling (0: tile queue, longest job first; 1: lane pool,
 * slots refill from one trajectory queue; 2: two-tile ping-pong lane pool), kernel launches of
 * one ikr_forward call, tensor-core path (0/1), epilogue column groups, 16-bit MMAs issued per fp32
 * product on the tensor-core path (3: fp16x2 split, 6: bf16x3 split), rest reserved (0).            */
int ikr_launch_geometry(const ikr_desc* d, int32_t n_jobs, const int64_t* B, int64_t out[16]);

/* 1 when ikr_forward will run this configuration on the tcgen05 tensor-core kernel (fp32 MLP with
 * n_nodes <= 200: hidden layers as split-operand MMAs with fp32 accumulation), 0 for the FFMA2 / DFMA kernel.
 * desc->reserved bit 1 opts out. */
int32_t ikr_uses_tensor_cores(const ikr_desc* d);

/* device workspace (caller-allocated) needed by ikr_forward / ikr_backward for n_jobs jobs.
 * with_backward: 0 forward only; 1 backward with the default stash of ~4 GiB; n > 1: a stash of ~n GiB
 * (at most 128, and never more than 64 reversed steps per round need): fewer, longer rounds.      */
size_t ikr_workspace_bytes(const ikr_desc* d, int32_t n_jobs, int64_t B_total,
                           int32_t with_backward);

/* Forward integration: for every job, B independent trajectories (== B separate B=1 reference
 * calls), each with its own adaptive step size.  `jobs` is a HOST array of n_jobs descriptors. */
int ikr_forward(const ikr_desc* d, const ikr_io* jobs, int32_t n_jobs, void* workspace,
                size_t workspace_bytes, void* cuda_stream);

/* Backward sweep: gradients of a scalar loss w.r.t. the MLP parameters (and optionally y0, g)
 * for ONE forward job (dopri5).  Runs in rounds sized by the workspace: ikr_workspace_bytes(...,
 * with_backward=1) returns a size that holds the fixed buffers plus a stash for a few reversed
 * steps per round; any larger workspace is used for longer rounds.                              */
int ikr_backward(const ikr_desc* d, const ikr_io* io, const ikr_bwd_io* bio, void* workspace,
                 size_t workspace_bytes, void* cuda_stream);

/* Hodgkin-Huxley candidate model without a network (train-d0.py:321-376 `ODEFunc`; the forward
 * model of the PINTS CMA-ES fit, `Model.simulate` train-d0.py:415-439), batched over a population:
 * trajectory b uses hh_params[b][0..7] = p1..p8 (nullable: desc->p for every trajectory).  One
 * job; `desc` supplies method / dtypes / tolerances (n_layers, n_nodes, mlp_dtype are ignored);
 * io->weights and the checkpoint fields are ignored.  No workspace.                              */
int ikr_forward_hh(const ikr_desc* d, const ikr_io* io, const double* hh_params, void* cuda_stream);

/* MLP regression stage of the training scripts (train-s1.py:891-909, train-r1.py:917-925): one
 * full-batch evaluation of  p = net(x) / netscale;  loss = sum (p - y)^2  and d loss / d theta.
 * x [N,2] fp32 = (V / vrange, a), y [N] fp32 = target da/dt; *loss_out (device) receives the fp64
 * loss, grad_weights [ikr_param_count] the fp64 gradient in state_dict order.  Tensor-core path
 * only (fp32 MLP, n_nodes <= 200 with n_nodes % 16 in 1..8): otherwise IKR_ERR_UNSUPPORTED and the
 * caller keeps its own autograd.  The optimiser step stays with the caller (torch.optim.Adam).   */
size_t ikr_regression_workspace_bytes(const ikr_desc* d, int64_t N);
int ikr_regression_loss_grad(const ikr_desc* d, const void* weights, const void* x, const void* y,
                             int64_t N, double* loss_out, double* grad_weights, void* workspace,
                             size_t workspace_bytes, void* cuda_stream);

/* 6-state Markov ground-truth model of the synthetic-data studies (train-d1.py:134-187 `Lambda`:
 * states [c1, c2, i, ic1, ic2, o], rate parameters p1..p12) and its data production step
 * (train-d1.py:539-569): `odeint(true_model, true_y0, t)` then
 * `true_y[:, 0, -1] * (V(t) + 86) + N(0, noise_sigma^2)`.  Batched: trajectory b has its own
 * initial state, parameters (nullable: `p` for everyone), conductance and noise stream (Philox,
 * subsequence b of `seed`).  `desc` supplies method / state dtype / tolerances.  No workspace.     */
typedef struct ikr_markov_io {
  int64_t B, T, G;
  const double* table_t;  /* protocol table, as in ikr_io                                         */
  const double* table_v;
  int32_t table_len, table_uniform;
  double table_t0, table_inv_dt;
  const void* y0;         /* [B,6] state dtype                                                     */
  const double* t_out;    /* [T]                                                                   */
  const double* grid;     /* [G] rk4                                                               */
  const double* v_out;    /* [T] V(t_out); required with i_out                                     */
  const double* params;   /* [B,12] per-trajectory p1..p12 (nullable)                              */
  double p[12];           /* shared p1..p12 when params is null                                    */
  const void* g;          /* [B] state dtype conductance (nullable => 1)                           */
  double e_rev;           /* -86 in the reference                                                  */
  double noise_sigma;     /* 0: noise-free current                                                 */
  uint64_t seed;
  void* y_out;            /* [T,B,6] state dtype (nullable)                                        */
  double* i_out;          /* [T,B] fp64 current (+ noise) (nullable)                               */
  int32_t* stats_out;     /* [B,4] n_accept, n_reject, nfe, status                                 */
} ikr_markov_io;
int ikr_forward_markov(const ikr_desc* d, const ikr_markov_io* io, void* cuda_stream);

/* V(t) of the protocol table at T query times (scipy interp1d linear semantics; out-of-table
 * => -80 like the callers' ValueError branch, train-s1.py:234-237).                             */
int ikr_interp_protocol(const ikr_io* table, const double* t_query, int64_t T, double* v_out,
                        void* cuda_stream);

/* FMA-pipe micro-benchmark used by bench.py for the roofline denominator: runs `iters`
 * dependent-free FFMA (dtype F32) or DFMA (F64) per thread on every SM and returns the elapsed
 * milliseconds through *ms_out (this one entry point synchronises the stream).                  */
/* dtype: IKR_F32 (scalar FFMA), IKR_F64 (DFMA) or 2 (packed FFMA2, fma.rn.f32x2).            */
int ikr_fma_peak(int32_t dtype, int64_t iters, double* tflops_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* IKR_H_ */
