"""6-state Markov ground-truth model and the synthetic-data production step.

Reference: ``train-d1.py:134-187`` (``Lambda``: states ``[c1, c2, i, ic1, ic2, o]``, rates from
``p1..p12``) and ``train-d1.py:539-569`` (one ``odeint(true_model, true_y0, t)`` per protocol sweep,
then ``true_y[:, 0, -1] * (V(t) + 86) + N(0, noise_sigma^2)``).  Here every sweep / parameter vector
/ noise realisation is one trajectory of ONE kernel call (``ikr_forward_markov``): trajectory ``b``
has its own initial state, ``p1..p12``, conductance and Philox noise stream.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .solver import (_check_job_inputs, _DeviceTable, _raise_on_status, _resolve_device, _rk4_grid,
                     _split_options)

_MK_KEYS = tuple('p%d' % i for i in range(1, 13))

# train-d1.py:139-150: best of 10 fits for herg25oc1 cell B06, in 1/ms and 1/mV
MARKOV_B06 = tuple(x * 1e-3 for x in (
    5.94625498751561316e-02, 1.21417701632850410e+02, 4.76436985414236425e+00,
    3.49383233960778904e-03, 9.62243079990877703e+01, 2.26404683824047979e+01,
    8.00924780462999131e+00, 2.43749808069009823e+01, 2.06822607368134157e+02,
    3.30791433507312362e+01, 1.26069071928587784e+00, 2.24844970727316245e+01))


class MarkovResult:
    """y (T, B, 6) state dtype or None; current (T, B) fp64 or None; stats (B, 4) int32."""

    def __init__(self, y, current, stats):
        self.y, self.current, self.stats = y, current, stats
        self._keep = []


def is_markov_func(func):
    """A reference-style Markov ODE func: scalars p1..p12, a protocol, and no ``net``."""
    return getattr(func, 'net', None) is None and all(hasattr(func, k) for k in _MK_KEYS)


def markov_params_of(func):
    return [float(getattr(func, k)) for k in _MK_KEYS]


def integrate_markov(params, y0, t, protocol, *, rtol=1e-7, atol=1e-9, method=None, options=None,
                     g=None, E=-86.0, noise_sigma=0.0, seed=0, want_y=True, want_current=False,
                     device=None):
    """Integrate B Markov trajectories.  ``params``: (12,) shared or (B, 12); ``y0`` (B, 6);
    ``protocol = (t_ms, v_mV)``.  ``want_current`` adds ``I = g o (V(t) - E)`` (+ Gaussian noise of
    standard deviation ``noise_sigma``, Philox stream ``(seed, b)``) as an fp64 (T, B) tensor."""
    method = method or 'dopri5'
    if method not in ('dopri5', 'rk4'):
        raise ValueError('method must be dopri5 or rk4')
    opts = _split_options(method, dict(options or {}))
    t_cpu = _check_job_inputs(y0, t, n_state=6)
    B, T = y0.shape[0], t_cpu.numel()
    P = np.asarray(params, dtype=np.float64)
    if P.shape not in ((12,), (B, 12)):
        raise ValueError('params must be (12,) or (B, 12)')
    dev = _resolve_device(y0, device)
    state_dtype = y0.dtype
    with torch.cuda.device(dev):
        d = _cabi.IkrDesc()
        d.n_layers, d.n_nodes, d.nn_d = 1, 1, 0
        d.method = _cabi.DOPRI5 if method == 'dopri5' else _cabi.RK4
        d.state_dtype = _cabi.F32 if state_dtype == torch.float32 else _cabi.F64
        d.mlp_dtype = _cabi.F64
        d.time_f32 = int(t_cpu.dtype == torch.float32)
        d.rk4_perturb = int(bool(opts.get('perturb', False)))
        d.vrange, d.netscale, d.negative_slope = 100.0, 1000.0, 0.01
        d.rtol, d.atol = float(rtol), float(atol)
        fs = opts.get('first_step', None)
        d.first_step = float(fs) if fs is not None else 0.0
        d.safety, d.ifactor, d.dfactor = (float(opts.get('safety', 0.9)),
                                          float(opts.get('ifactor', 10.0)),
                                          float(opts.get('dfactor', 0.2)))
        d.max_num_steps = int(opts.get('max_num_steps', 2 ** 31 - 1))
        lib = _cabi.lib()
        stream = torch.cuda.current_stream(dev)
        sptr = ctypes.c_void_p(stream.cuda_stream)
        tt = np.ascontiguousarray(np.asarray(protocol[0], dtype=np.float64).reshape(-1))
        vv = np.ascontiguousarray(np.asarray(protocol[1], dtype=np.float64).reshape(-1))
        tab = _DeviceTable(tt, vv, dev, opts.get('compact_table', True))
        io = _cabi.IkrMarkovIO()
        io.B, io.T = B, T
        io.table_t, io.table_v, io.table_len = tab.t.data_ptr(), tab.v.data_ptr(), tab.len
        io.table_uniform, io.table_t0, io.table_inv_dt = tab.uniform
        y0_d = y0.detach().to(dev).contiguous()
        t_d = t_cpu.to(torch.float64).to(dev)
        io.y0, io.t_out = y0_d.data_ptr(), t_d.data_ptr()
        keep = [y0_d, t_d, tab]
        if P.ndim == 2:
            p_d = torch.from_numpy(np.array(P, dtype=np.float64, order='C', copy=True)).to(dev)
            io.params = p_d.data_ptr()
            keep.append(p_d)
        else:
            for i in range(12):
                io.p[i] = float(P[i])
        if method == 'rk4':
            grid = _rk4_grid(t_cpu, opts.get('step_size')).to(dev)
            io.grid, io.G = grid.data_ptr(), grid.numel()
            keep.append(grid)
        y_out = cur = None
        if want_current:
            tio = _cabi.IkrIO()
            tab.fill(tio)
            v_out = torch.empty(T, dtype=torch.float64, device=dev)
            _cabi.check(lib.ikr_interp_protocol(ctypes.byref(tio), t_d.data_ptr(), T,
                                                v_out.data_ptr(), sptr), 'ikr_interp_protocol')
            io.v_out = v_out.data_ptr()
            keep.append(v_out)
            if g is not None:
                g_d = torch.as_tensor(g).to(device=dev, dtype=state_dtype).reshape(-1)
                g_d = (g_d.expand(B) if g_d.numel() == 1 else g_d).contiguous()
                io.g = g_d.data_ptr()
                keep.append(g_d)
            cur = torch.empty((T, B), dtype=torch.float64, device=dev)
            io.i_out = cur.data_ptr()
        io.e_rev = float(E)
        io.noise_sigma, io.seed = float(noise_sigma), int(seed)
        if want_y:
            y_out = torch.empty((T, B, 6), dtype=state_dtype, device=dev)
            io.y_out = y_out.data_ptr()
        stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
        io.stats_out = stats.data_ptr()
        _cabi.check(lib.ikr_forward_markov(ctypes.byref(d), ctypes.byref(io), sptr),
                    'ikr_forward_markov')
        res = MarkovResult(y_out, cur, stats)
        res._keep = keep
        if opts.get('check_status', True):
            _raise_on_status(stats)
        return res


class MarkovGroundTruth:
    """The reference's ``Lambda`` ground-truth model (``train-d1.py:134-187``) as a data generator:
    same ``p1..p12`` attributes and ``set_fixed_form_voltage_protocol``; ``simulate_data`` is the
    batched form of ``train-d1.py:539-569`` (all sweeps and noise realisations in one launch per
    protocol)."""

    def __init__(self, params=MARKOV_B06, y0=(0., 1., 0., 0., 0., 0.), E=-86.0, device=None):
        for k, v in zip(_MK_KEYS, params):
            setattr(self, k, float(v))
        self._y0 = np.asarray(y0, dtype=np.float32).reshape(1, 6)   # gt_true_y0s[1], train-d1.py:117
        self._E = float(E)
        self._device = device
        self._t_regular = self._v_regular = None

    def set_fixed_form_voltage_protocol(self, t, v):
        self._t_regular = np.asarray(t, dtype=np.float64)
        self._v_regular = np.asarray(v, dtype=np.float64)

    def simulate_data(self, protocols, t, n_realisations=1, noise_sigma=0.1, seed=0):
        """``protocols``: list of (n, 2) arrays [time, voltage] (``protocol_batches`` of the
        reference).  Returns (open probability (T, n_protocols) fp32, current (T, n_protocols,
        n_realisations) fp64): every realisation re-draws the noise, the ODE is solved once per
        realisation lane (identical lanes give identical solutions)."""
        t_t = torch.as_tensor(np.array(t, copy=True))
        opens, curs = [], []
        for k, pr in enumerate(protocols):
            pr = np.asarray(pr)
            y0 = torch.from_numpy(np.repeat(self._y0, n_realisations, 0))
            res = integrate_markov(markov_params_of(self), y0, t_t, (pr[:, 0], pr[:, 1]),
                                   want_y=True, want_current=True, E=self._E,
                                   noise_sigma=noise_sigma, seed=seed + 1000003 * k,
                                   device=self._device)
            opens.append(res.y[:, 0, 5])
            curs.append(res.current)
        return torch.stack(opens, 1), torch.stack(curs, 1)
