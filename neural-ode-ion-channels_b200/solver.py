"""Drop-in ``odeint`` for the reference's call sites, backed by the fused sm_100a kernels.

    from neural_ode_ion_channels_b200 import odeint          # was: from torchdiffeq import odeint
    func = ODEFunc(); func.load_state_dict(torch.load('d1/model-state-dict.pt')); func.eval()
    func.set_fixed_form_voltage_protocol(t_np, v_np)         # train-s1.py:218-222
    with torch.no_grad():                                    # as every reference call site does
        y = odeint(func, y0, t, method='dopri5')             # (len(t), *y0.shape)

Like ``torchdiffeq.odeint`` the call is differentiable when an MLP parameter requires grad and
grad mode is on: the forward then records step checkpoints (80 B per accepted step and
trajectory in fp32; the capacity is sized from free memory and grown on demand) for the discrete
adjoint in ``adjoint.py``.  Prediction loops should run under ``torch.no_grad()``.

Signature and semantics follow ``torchdiffeq.odeint`` 0.2.x as the reference uses it
(``train-s1.py:322,327``, ``table-1.py:404,413``, ``train-d0.py:436``): ``rtol=1e-7, atol=1e-9``,
``method`` in {None/'dopri5', 'rk4'}, unknown ``options`` only warn.  Differences, all additive:

* ``y0`` may be ``(B, 2)``: B independent trajectories, each with its *own* adaptive step size
  and error control (== B separate B=1 reference calls, not torchdiffeq's shared-dt batching).
* the module is *introspected* (``describe``) -- ``func.forward`` is never called per stage.
* ``integrate`` exposes the fused observation ``I = g a r (V - E)`` and loss reductions.

There is no CPU path: tensors are moved to the current CUDA device and the C-ABI library
``csrc/libikr_b200.so`` does the work; if it is missing this module raises.
"""
import ctypes
import hashlib
import warnings
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi
from .protocols import compact_table

_ADAPTIVE_OPTS = {'first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps'}
_FIXED_OPTS = {'step_size', 'perturb'}
_EXT_OPTS = {'tile_m', 'check_status', 'compact_table', 'ckpt_cap', 'lane_pool', 'tensor_cores',
             'tc_groups', 'tc_timing', 'ping_pong', 'tc_split', 'bwd_overlap', 'stash_gib',
             'adjoint_products'}


# =============================================================================================
# module introspection (SURVEY 8b "module introspection contract")
# =============================================================================================
@dataclass
class ModelSpec:
    n_layers: int
    n_nodes: int
    nn_d: bool
    p: tuple
    vrange: float
    netscale: float
    negative_slope: float
    linears: list            # the nn.Linear modules in order
    mlp_dtype: torch.dtype


def _scalar(x, name):
    if isinstance(x, torch.Tensor):
        if x.numel() != 1:
            raise TypeError('ODE func attribute %s must be a scalar' % name)
        return float(x.reshape(-1)[0].item())
    return float(x)


def describe(func) -> ModelSpec:
    """Flatten a reference-style ODE func (``train-s1.py:181-216`` / ``train-d2.py:191-232``)
    into a descriptor; anything that is not such a module raises ``TypeError``."""
    net = getattr(func, 'net', None)
    if not isinstance(net, nn.Sequential):
        raise TypeError('odeint: func.net must be an nn.Sequential of Linear/LeakyReLU layers '
                        '(got %r); this integrator has no generic-RHS / CPU fallback' % type(net))
    mods = list(net)
    if len(mods) < 5 or len(mods) % 2 == 0:
        raise TypeError('odeint: func.net must be Linear(2,n), LeakyReLU, [Linear(n,n), '
                        'LeakyReLU]*L, Linear(n,1) with L >= 1')
    linears, slope = [], None
    for i, m in enumerate(mods):
        if i % 2 == 0:
            if not isinstance(m, nn.Linear) or m.bias is None:
                raise TypeError('odeint: func.net[%d] must be nn.Linear with bias' % i)
            linears.append(m)
        else:
            if not isinstance(m, nn.LeakyReLU):
                raise TypeError('odeint: func.net[%d] must be nn.LeakyReLU' % i)
            if slope is not None and m.negative_slope != slope:
                raise TypeError('odeint: all LeakyReLU slopes must agree')
            slope = m.negative_slope
    n = linears[0].out_features
    if linears[0].in_features != 2 or linears[-1].out_features != 1:
        raise TypeError('odeint: func.net must map 2 inputs to 1 output')
    for m in linears[1:-1]:
        if m.in_features != n or m.out_features != n:
            raise TypeError('odeint: hidden layers must all be Linear(%d, %d)' % (n, n))
    if linears[-1].in_features != n:
        raise TypeError('odeint: last layer must be Linear(%d, 1)' % n)
    dtypes = {p.dtype for p in net.parameters()}
    if len(dtypes) != 1 or next(iter(dtypes)) not in (torch.float32, torch.float64):
        raise TypeError('odeint: MLP parameters must be all float32 or all float64')
    for k in ('p5', 'p6', 'p7', 'p8', 'vrange', 'netscale'):
        if not hasattr(func, k):
            raise TypeError('odeint: func lacks attribute %r of the reference ODEFunc' % k)
    nn_d = all(hasattr(func, k) for k in ('p1', 'p2', 'p3', 'p4')) and hasattr(func, '_dadt')
    p = tuple(_scalar(getattr(func, 'p%d' % i), 'p%d' % i) if (i >= 5 or nn_d) else 0.0
              for i in range(1, 9))
    return ModelSpec(n_layers=len(linears) - 2, n_nodes=n, nn_d=nn_d, p=p,
                     vrange=_scalar(func.vrange, 'vrange'),
                     netscale=_scalar(func.netscale, 'netscale'),
                     negative_slope=float(slope), linears=linears,
                     mlp_dtype=next(iter(dtypes)))


def _protocol_arrays(func):
    t = getattr(func, '_t_regular', None)
    v = getattr(func, '_v_regular', None)
    if t is None or v is None:
        raise TypeError('odeint: call func.set_fixed_form_voltage_protocol(t, v) first')
    t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(-1))
    v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    if len(t) != len(v) or len(t) < 2:
        raise TypeError('odeint: protocol table needs >= 2 samples of equal length')
    if not np.all(np.diff(t) > 0):
        raise TypeError('odeint: protocol table times must be strictly increasing')
    return t, v


def _uniform_hint(t):
    if len(t) < 3:
        return 0, 0.0, 0.0
    h = (t[-1] - t[0]) / (len(t) - 1)
    if np.all(np.abs(t - (t[0] + h * np.arange(len(t)))) < 0.25 * h):
        return 1, float(t[0]), float(1.0 / h)
    return 0, 0.0, 0.0


# =============================================================================================
# device-side model: packed weights + protocol table
# =============================================================================================
def _make_desc(spec: ModelSpec, state_dtype, method, rtol, atol, opts, time_f32=False):
    d = _cabi.IkrDesc()
    d.n_layers, d.n_nodes, d.nn_d = spec.n_layers, spec.n_nodes, int(spec.nn_d)
    d.method = _cabi.DOPRI5 if method == 'dopri5' else _cabi.RK4
    d.state_dtype = _cabi.F32 if state_dtype == torch.float32 else _cabi.F64
    d.mlp_dtype = _cabi.F32 if spec.mlp_dtype == torch.float32 else _cabi.F64
    d.time_f32 = int(time_f32)
    d.rk4_perturb = int(bool(opts.get('perturb', False)))
    for i in range(8):
        d.p[i] = spec.p[i]
    d.vrange, d.netscale, d.negative_slope = spec.vrange, spec.netscale, spec.negative_slope
    d.rtol, d.atol = float(rtol), float(atol)
    fs = opts.get('first_step', None)
    d.first_step = float(fs) if fs is not None else 0.0
    d.safety = float(opts.get('safety', 0.9))
    d.ifactor = float(opts.get('ifactor', 10.0))
    d.dfactor = float(opts.get('dfactor', 0.2))
    d.max_num_steps = int(opts.get('max_num_steps', 2 ** 31 - 1))
    d.tile_m = int(opts.get('tile_m', 0))
    # bit 0: force the lane-pool kernel; bit 2: forbid it (lane_pool=None: library decides);
    # bit 1: keep the fp32 MLP on the FFMA2 kernel (no tcgen05 path)
    # bit 3: phase-clock printf of CTA 0; bits 4-5: epilogue column groups (tuning; changes the
    # output-layer summation order); bit 6 / 7: forbid / force the two-tile ping-pong kernel
    lp = opts.get('lane_pool', None)
    pp = opts.get('ping_pong', None)
    if opts.get('tc_split', None) not in (None, 'fp16x2', 'bf16x3'):
        raise ValueError("odeint: tc_split must be 'fp16x2' or 'bf16x3'")
    groups = int(opts.get('tc_groups', 0) or 0)
    if groups not in (0, 1, 2, 3):
        raise ValueError('odeint: tc_groups must be 1, 2 or 3')
    # bit 10: bf16 products per fp32 product in the adjoint kernel's MMAs: 'three' (default) / 'six'
    adj = {None: 0, 'three': 0, 'six': 1}.get(opts.get('adjoint_products', None), -1)
    if adj < 0:
        raise ValueError("odeint: adjoint_products must be 'three' or 'six'")
    d.reserved = ((adj << 10) | (1 if lp else 0) | (4 if lp is False else 0) |
                  (0 if opts.get('tensor_cores', True) else 2) |
                  (8 if opts.get('tc_timing', False) else 0) | (groups << 4) |
                  (64 if pp is False else 0) | (128 if pp else 0) |
                  (256 if opts.get('tc_split', None) == 'bf16x3' else 0) |
                  (0 if opts.get('bwd_overlap', True) else 512))
    return d


def pack_weights(spec: ModelSpec, desc, device, out=None):
    """Pack the state-dict tensors into the kernel layout documented in ``include/ikr.h``
    (``[w0[:,0] | w0[:,1] | b0] | L x W^T[k][npad] | L x b[npad] | w_last | b_last | L x W``).
    ``out``: a buffer of a previous call with the same layout (its padding is already zero)."""
    lay = _cabi.packed_layout(desc)
    npad, n, L = lay['npad'], spec.n_nodes, spec.n_layers
    dt = spec.mlp_dtype
    # a launch that is still reading the previous contents is ordered before these copies by the
    # stream; callers on several streams must not share one func object
    buf = out if out is not None else torch.zeros(lay['total'], dtype=dt, device=device)
    lin = spec.linears
    w0 = lin[0].weight.detach().to(device=device, dtype=dt)
    o = lay['off_w0']
    buf[o:o + n] = w0[:, 0]
    buf[o + npad:o + npad + n] = w0[:, 1]
    buf[o + 2 * npad:o + 2 * npad + n] = lin[0].bias.detach().to(device=device, dtype=dt)
    wt = buf[lay['off_wt']:lay['off_wt'] + L * n * npad].view(L, n, npad)
    wn = buf[lay['off_wn']:lay['off_wn'] + L * n * npad].view(L, n, npad)
    bh = buf[lay['off_bh']:lay['off_bh'] + L * npad].view(L, npad)
    for l in range(L):
        w = lin[1 + l].weight.detach().to(device=device, dtype=dt)      # (out, in)
        wt[l, :, :n] = w.t()
        wn[l, :, :n] = w
        bh[l, :n] = lin[1 + l].bias.detach().to(device=device, dtype=dt)
    o = lay['off_wl']
    buf[o:o + n] = lin[-1].weight.detach().to(device=device, dtype=dt).reshape(-1)
    buf[o + npad] = lin[-1].bias.detach().to(device=device, dtype=dt).reshape(-1)[0]
    return buf


def unpack_grads(spec: ModelSpec, flat):
    """Split the flat gradient (state-dict order: w0, b0, (W_l, b_l)*, w_last, b_last) into
    per-parameter tensors shaped like ``spec.linears``' parameters."""
    out, o = [], 0
    for m in spec.linears:
        for p in (m.weight, m.bias):
            k = p.numel()
            out.append(flat[o:o + k].view_as(p))
            o += k
    return out


class _DeviceTable:
    """One protocol table resident on the GPU (compacted when that is bit-exact)."""

    def __init__(self, t, v, device, use_compaction):
        if use_compaction:
            t, v = compact_table(t, v)
        self.uniform = _uniform_hint(t)
        self.len = len(t)
        self.t = torch.from_numpy(t).to(device)
        self.v = torch.from_numpy(v).to(device)
        self.duration = float(t[-1] - t[0])

    def fill(self, io):
        io.table_t, io.table_v, io.table_len = self.t.data_ptr(), self.v.data_ptr(), self.len
        io.table_uniform, io.table_t0, io.table_inv_dt = self.uniform


def _table_key(t, v, use_compaction):
    """Content key of a protocol table (never the host addresses: temporaries of the reference's
    protocol loops are freed and reallocated at the same place with other contents)."""
    h = hashlib.blake2b(digest_size=16)
    h.update(t.tobytes())
    h.update(v.tobytes())
    return (len(t), h.digest(), bool(use_compaction))


class _DeviceModel:
    """Packed weights + protocol tables resident on one GPU (cached on the func object)."""

    def __init__(self, func, device):
        self.spec = describe(func)
        self.device = device
        self.weights = None
        self.weights_key = None
        self.tables = {}

    def table(self, t, v, use_compaction=True):
        t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(-1))
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
        if len(t) != len(v) or len(t) < 2:
            raise TypeError('odeint: protocol table needs >= 2 samples of equal length')
        key = _table_key(t, v, use_compaction)
        tab = self.tables.get(key)
        if tab is None:
            if not np.all(np.diff(t) > 0):
                raise TypeError('odeint: protocol table times must be strictly increasing')
            if len(self.tables) > 256:
                self.tables.clear()
            tab = _DeviceTable(t, v, self.device, use_compaction)
            self.tables[key] = tab
        return tab

    def weights_for(self, desc, private=False):
        """Packed parameter buffer for this call (``private``: a buffer of its own, for results
        whose backward pass reads the weights again later).  Re-packed on EVERY call (a handful of small
        device copies): writes through ``p.data`` or ``optimizer.step()`` leave neither the data
        pointer nor a reliable version counter behind, and integrating with stale weights would be
        silent.  The buffer itself is reused when the layout is unchanged."""
        if private:
            return pack_weights(self.spec, desc, self.device)
        lay = _cabi.packed_layout(desc)
        key = (lay['total'], self.spec.mlp_dtype)
        if self.weights is None or self.weights_key != key:
            self.weights, self.weights_key = None, key
        self.weights = pack_weights(self.spec, desc, self.device, out=self.weights)
        return self.weights


def _device_model(func, device):
    cache = func.__dict__.setdefault('_ikr_b200_cache', {})
    dm = cache.get(device)
    if dm is None:
        dm = _DeviceModel(func, device)
        cache[device] = dm
    else:
        dm.spec = describe(func)
    return dm


# =============================================================================================
# integrate
# =============================================================================================
@dataclass
class IkrResult:
    y: Optional[torch.Tensor]            # (T, B, 2) state dtype
    current: Optional[torch.Tensor]      # (T, B)
    sse: Optional[torch.Tensor]          # (B,) sum_i (I - data)^2   (fp64)
    sae: Optional[torch.Tensor]          # (B,) sum_i |I - data|     (fp64)
    stats: torch.Tensor                  # (B, 4) int32: n_accept, n_reject, nfe, status
    ckpt: Optional[tuple] = None
    geometry: Optional[dict] = None

    @property
    def nfe(self):
        return int(self.stats[:, 2].sum().item())

    def mae(self, T):
        return self.sae / T


def _resolve_device(y0, device):
    if device is not None:
        return torch.device(device)
    if y0.is_cuda:
        return y0.device
    if not torch.cuda.is_available():
        raise RuntimeError('neural_ode_ion_channels_b200.odeint needs a CUDA device (B200, '
                           'sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _rk4_grid(t_cpu, step_size):
    """torchdiffeq's fixed-grid constructor for ``options={'step_size': h}`` (in t's dtype)."""
    if step_size is None:
        return t_cpu.to(torch.float64)
    start, end = t_cpu[0], t_cpu[-1]
    niters = torch.ceil((end - start) / step_size + 1).item()
    grid = torch.arange(0, niters, dtype=t_cpu.dtype) * step_size + start
    grid[-1] = t_cpu[-1]
    return grid.to(torch.float64)


def _split_options(method, options):
    options = dict(options or {})
    known = (_ADAPTIVE_OPTS if method == 'dopri5' else _FIXED_OPTS) | _EXT_OPTS
    unknown = sorted(set(options) - known)
    if unknown:
        # torchdiffeq 0.2.x warns on unknown options (the reference passes the legacy
        # grid_points/eps, train-d0.py:436) -- same here
        warnings.warn('{}: Unexpected arguments {}'.format(
            'Dopri5Solver' if method == 'dopri5' else 'RK4', {k: options[k] for k in unknown}))
        for k in unknown:
            options.pop(k)
    return options


def _raise_on_status(stats):
    bad = stats[:, 3] != 0
    if bool(bad.any().item()):
        idx = torch.nonzero(bad).reshape(-1)
        code = int(stats[idx[0], 3].item())
        raise AssertionError('%s (trajectory %d%s)' % (
            _cabi.STATUS_TEXT.get(code, 'solver failure %d' % code), int(idx[0].item()),
            '' if idx.numel() == 1 else ' and %d more' % (idx.numel() - 1)))


def _prepare_job(dm, desc, method, opts, dev, lib, sptr, job, state_dtype):
    """Upload one job's inputs, allocate its outputs and fill its ``ikr_io``."""
    y0, t = job['y0'], job['t_cpu']
    B, T = y0.shape[0], t.numel()
    proto = job.get('protocol')
    if proto is None:
        proto = _protocol_arrays(job['func'])
    tab = dm.table(proto[0], proto[1], opts.get('compact_table', True))
    io = _cabi.IkrIO()
    io.B, io.T = B, T
    tab.fill(io)
    io.cost_hint = float(t[-1] - t[0]) if T > 1 else 1.0
    y0_d = y0.detach().to(dev, non_blocking=True).contiguous()
    t_d = t.to(torch.float64).to(dev, non_blocking=True)
    io.y0, io.t_out = y0_d.data_ptr(), t_d.data_ptr()
    keep = [y0_d, t_d, tab]
    if method == 'rk4':
        grid = _rk4_grid(t, opts.get('step_size')).to(dev)
        io.grid, io.G = grid.data_ptr(), grid.numel()
        keep.append(grid)
    g, E, data = job.get('g'), job.get('E', -86.0), job.get('data')
    want_current, want_y = job.get('want_current', False), job.get('want_y', True)
    observe = want_current or data is not None
    y_out = cur = loss = None
    if observe:
        v_out = torch.empty(T, dtype=torch.float64, device=dev)
        _cabi.check(lib.ikr_interp_protocol(ctypes.byref(io), t_d.data_ptr(), T,
                                            v_out.data_ptr(), sptr), 'ikr_interp_protocol')
        io.v_out = v_out.data_ptr()
        keep.append(v_out)
        if g is not None:
            g_d = torch.as_tensor(g).to(device=dev, dtype=state_dtype,
                                        non_blocking=True).reshape(-1).contiguous()
            if g_d.numel() == 1:
                g_d = g_d.expand(B).contiguous()
            if g_d.numel() != B:
                raise ValueError('odeint: g must have B entries')
            io.g = g_d.data_ptr()
            keep.append(g_d)
        if isinstance(E, torch.Tensor) and E.numel() > 1:
            e_d = E.to(device=dev, dtype=state_dtype).reshape(-1).contiguous()
            if e_d.numel() != B:
                raise ValueError('odeint: E must be a scalar or have B entries')
            io.e_rev = e_d.data_ptr()
            keep.append(e_d)
        else:
            io.e_scalar = float(E)
        if want_current:
            cur = torch.empty((T, B), dtype=state_dtype, device=dev)
            io.i_out = cur.data_ptr()
        if data is not None:
            d_d = torch.as_tensor(data).to(device=dev, dtype=state_dtype).contiguous()
            if d_d.dim() == 1:
                d_d = d_d.reshape(T, 1)
            if d_d.shape[0] != T or d_d.shape[1] not in (1, B):
                raise ValueError('odeint: data must be (T,) or (T, B)')
            io.data, io.data_B = d_d.data_ptr(), d_d.shape[1]
            loss = torch.empty((B, 2), dtype=torch.float64, device=dev)
            io.loss_out = loss.data_ptr()
            keep.append(d_d)
    if want_y:
        y_out = torch.empty((T, B, 2), dtype=state_dtype, device=dev)
        io.y_out = y_out.data_ptr()
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
    io.stats_out = stats.data_ptr()
    ckpt = None
    if job.get('want_ckpt', False):
        cap = int(opts.get('ckpt_cap', 0)) or (io.G if method == 'rk4' else 4096)
        ck_t = torch.empty((cap, B, 2), dtype=torch.float64, device=dev)
        ck_y = torch.empty((cap, B, 16), dtype=state_dtype, device=dev)
        io.ckpt_cap, io.ckpt_t, io.ckpt_y = cap, ck_t.data_ptr(), ck_y.data_ptr()
        ckpt = (ck_t, ck_y)
    res = IkrResult(y=y_out, current=cur, sse=None if loss is None else loss[:, 0],
                    sae=None if loss is None else loss[:, 1], stats=stats, ckpt=ckpt)
    res._keep = keep
    return io, res


def _check_job_inputs(y0, t, n_state=2):
    if not isinstance(y0, torch.Tensor) or y0.dim() != 2 or y0.shape[1] != n_state:
        raise TypeError('odeint: y0 must be a (B, 2) tensor of (a, r) states' if n_state == 2 else
                        'odeint: y0 must be a (B, %d) state tensor' % n_state)
    if y0.dtype not in (torch.float32, torch.float64):
        raise TypeError('odeint: y0 must be float32 or float64')
    t = torch.as_tensor(t)
    if t.dim() != 1 or t.numel() < 1 or not t.is_floating_point():
        raise TypeError('odeint: t must be a 1-D floating point tensor')
    t_cpu = t.detach().cpu()
    if t_cpu.numel() > 1 and not bool((t_cpu[1:] > t_cpu[:-1]).all()):
        raise ValueError('odeint: t must be strictly increasing')
    return t_cpu


def integrate_many(func, jobs, *, rtol=1e-7, atol=1e-9, method=None, options=None, device=None):
    """Integrate several (protocol, batch) jobs that share ``func``'s MLP in ONE kernel launch.

    ``jobs``: list of dicts with keys ``y0`` (B,2), ``t`` (T,), optional ``protocol=(t_ms, v_mV)``
    (default: ``func``'s current protocol), ``g``, ``E``, ``data``, ``want_y``, ``want_current``,
    ``want_ckpt``.  This is the batched form of the reference's protocol loops
    (``train-s1.py:316-543``, ``table-1.py:401-599``): the tiles of all jobs are scheduled through
    one longest-first queue so every SM stays busy.  Returns one ``IkrResult`` per job."""
    method = method or 'dopri5'
    if method not in ('dopri5', 'rk4'):
        raise ValueError('odeint: method must be dopri5 or rk4 (got %r); the B200 path '
                         'implements the two solvers of the hot path only' % (method,))
    opts = _split_options(method, options)
    if not jobs:
        return []
    prepared = []
    for job in jobs:
        job = dict(job)
        job.setdefault('func', func)
        job['t_cpu'] = _check_job_inputs(job['y0'], job['t'])
        prepared.append(job)
    state_dtype = prepared[0]['y0'].dtype
    t_dtype = prepared[0]['t_cpu'].dtype
    for job in prepared:
        if job['y0'].dtype != state_dtype or job['t_cpu'].dtype != t_dtype:
            raise TypeError('odeint: all jobs of one launch must share the state and time dtypes')

    dev = _resolve_device(prepared[0]['y0'], device)
    with torch.cuda.device(dev):
        dm = _device_model(func, dev)
        spec = dm.spec
        if state_dtype == torch.float32 and spec.mlp_dtype == torch.float64:
            raise TypeError('odeint: float32 state with a float64 MLP is not supported')
        desc = _make_desc(spec, state_dtype, method, rtol, atol, opts,
                          time_f32=(t_dtype == torch.float32))
        weights = dm.weights_for(desc, private=any(j.get('want_ckpt') for j in prepared))
        lib = _cabi.lib()
        stream = torch.cuda.current_stream(dev)
        sptr = ctypes.c_void_p(stream.cuda_stream)
        ios = (_cabi.IkrIO * len(prepared))()
        results = []
        for k, job in enumerate(prepared):
            io, res = _prepare_job(dm, desc, method, opts, dev, lib, sptr, job, state_dtype)
            io.weights = weights.data_ptr()
            ios[k] = io
            results.append(res)
        Bs = [int(j['y0'].shape[0]) for j in prepared]
        ws_bytes = lib.ikr_workspace_bytes(ctypes.byref(desc), len(prepared), sum(Bs), 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _cabi.check(lib.ikr_forward(ctypes.byref(desc), ios, len(prepared), ws.data_ptr(),
                                    ws_bytes, sptr), 'ikr_forward')
        geo = _cabi.launch_geometry(desc, Bs)
        for k, res in enumerate(results):
            res.geometry = geo
            res._keep += [weights, ws]
            res._desc, res._io = desc, ios[k]
        if opts.get('check_status', True):
            for res in results:
                _raise_on_status(res.stats)
        return results


def integrate(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, g=None, E=-86.0,
              data=None, want_y=True, want_current=False, want_ckpt=False, device=None):
    """Forward integration with the fused observation/loss epilogue.

    ``g`` (B,) conductances, ``E`` scalar or (B,) reversal potential, ``data`` (T,) or (T, B)
    measured current: returns ``IkrResult`` with ``current = g a r (V(t) - E)`` and the
    per-trajectory ``sse`` / ``sae`` reductions against ``data`` (``train-s1.py:328-329``,
    ``train-d0.py:509``)."""
    job = dict(y0=y0, t=t, g=g, E=E, data=data, want_y=want_y, want_current=want_current,
               want_ckpt=want_ckpt)
    return integrate_many(func, [job], rtol=rtol, atol=atol, method=method, options=options,
                          device=device)[0]


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """``torchdiffeq.odeint`` replacement: returns ``(len(t), *y0.shape)`` in ``y0.dtype`` on
    ``y0``'s device."""
    if event_fn is not None:
        raise NotImplementedError('odeint: event handling is not part of the reference hot path')
    squeeze = False
    if isinstance(y0, torch.Tensor) and y0.dim() == 1:
        y0, squeeze = y0.reshape(1, -1), True
    from .hh import hh_params_of, integrate_hh, is_hh_func
    if is_hh_func(func):
        # network-free HH candidate (train-d0.py:321-376): same call, dedicated kernel
        y = integrate_hh(hh_params_of(func), y0, t, _protocol_arrays(func), rtol=rtol, atol=atol,
                         method=method, options=options).y
        if not y0.is_cuda:
            y = y.to(y0.device)
        return y[:, 0, :] if squeeze else y
    from .markov import integrate_markov, is_markov_func, markov_params_of
    if is_markov_func(func):
        # 6-state Markov ground truth (train-d1.py:134-187): same call, dedicated kernel
        y = integrate_markov(markov_params_of(func), y0, t, _protocol_arrays(func), rtol=rtol,
                             atol=atol, method=method, options=options).y
        if not y0.is_cuda:
            y = y.to(y0.device)
        return y[:, 0, :] if squeeze else y
    needs_grad = torch.is_grad_enabled() and any(
        p.requires_grad for p in getattr(func, 'net', nn.Sequential()).parameters())
    if needs_grad:
        from .adjoint import odeint_with_grad
        y = odeint_with_grad(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
    else:
        y = integrate(func, y0, t, rtol=rtol, atol=atol, method=method, options=options).y
    if not y0.is_cuda:
        y = y.to(y0.device)
    return y[:, 0, :] if squeeze else y
