"""Hodgkin-Huxley candidate model without a network, batched over a POPULATION of parameter
vectors: the forward model the reference fits with PINTS CMA-ES.

Reference: ``train-d0.py:321-376`` (``ODEFunc`` with ``p1..p8``, ``set_parameters``) and
``train-d0.py:377-439`` (``Model(pints.ForwardModel)``: ``simulate(x, t)`` -> ``odeint`` ->
``g a r (V + 86)``); the optimiser evaluates one parameter vector per ``simulate`` call, in forked
worker processes (``set_parallel(True)``, ``train-d0.py:538``).  Here a whole CMA-ES generation is
ONE kernel call: trajectory ``b`` integrates with its own ``p1..p8`` (``ikr_forward_hh``).
"""
import ctypes
import warnings

import numpy as np
import torch

from . import _cabi
from .solver import (IkrResult, _check_job_inputs, _DeviceTable, _raise_on_status, _resolve_device,
                     _rk4_grid, _split_options)

_HH_KEYS = tuple('p%d' % i for i in range(1, 9))


def is_hh_func(func):
    """A reference-style HH ODE func: scalars p1..p8, a protocol, and no ``net``."""
    return (getattr(func, 'net', None) is None and all(hasattr(func, k) for k in _HH_KEYS)
            and not hasattr(func, 'p9'))     # p1..p12 => the 6-state Markov model (markov.py)


def hh_params_of(func):
    return [float(getattr(func, k)) for k in _HH_KEYS]


def integrate_hh(params, y0, t, protocol, *, rtol=1e-7, atol=1e-9, method=None, options=None,
                 g=None, E=-86.0, data=None, want_y=True, want_current=False, device=None):
    """Integrate B HH candidates.  ``params``: (8,) shared or (B, 8) per-trajectory ``p1..p8``
    (a (B, 4) array sets ``p1..p4`` and needs ``options={'inactivation': (p5, p6, p7, p8)}``).
    ``protocol = (t_ms, v_mV)``.  Same outputs as ``integrate``."""
    method = method or 'dopri5'
    if method not in ('dopri5', 'rk4'):
        raise ValueError('method must be dopri5 or rk4')
    options = dict(options or {})
    inact = options.pop('inactivation', None)
    opts = _split_options(method, options)
    t_cpu = _check_job_inputs(y0, t)
    B, T = y0.shape[0], t_cpu.numel()
    P = np.asarray(params, dtype=np.float64)
    if P.ndim == 1:
        P = np.broadcast_to(P, (B, P.shape[0]))
    if P.shape[1] == 4:
        if inact is None:
            raise ValueError("4-parameter candidates need options={'inactivation': (p5..p8)}")
        P = np.concatenate([P, np.broadcast_to(np.asarray(inact, dtype=np.float64), (P.shape[0], 4))], 1)
    if P.shape != (B, 8):
        raise ValueError('params must be (8,), (B, 8) or (B, 4)')
    dev = _resolve_device(y0, device)
    state_dtype = y0.dtype
    with torch.cuda.device(dev):
        d = _cabi.IkrDesc()
        d.n_layers, d.n_nodes, d.nn_d = 1, 1, 1
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
        io = _cabi.IkrIO()
        io.B, io.T = B, T
        tab.fill(io)
        y0_d = y0.detach().to(dev).contiguous()
        t_d = t_cpu.to(torch.float64).to(dev)
        p_d = torch.from_numpy(np.array(P, dtype=np.float64, order="C", copy=True)).to(dev)
        io.y0, io.t_out = y0_d.data_ptr(), t_d.data_ptr()
        keep = [y0_d, t_d, p_d, tab]
        if method == 'rk4':
            grid = _rk4_grid(t_cpu, opts.get('step_size')).to(dev)
            io.grid, io.G = grid.data_ptr(), grid.numel()
            keep.append(grid)
        y_out = cur = loss = None
        if want_current or data is not None:
            v_out = torch.empty(T, dtype=torch.float64, device=dev)
            _cabi.check(lib.ikr_interp_protocol(ctypes.byref(io), t_d.data_ptr(), T,
                                                v_out.data_ptr(), sptr), 'ikr_interp_protocol')
            io.v_out = v_out.data_ptr()
            keep.append(v_out)
            if g is not None:
                g_d = torch.as_tensor(g).to(device=dev, dtype=state_dtype).reshape(-1)
                g_d = (g_d.expand(B) if g_d.numel() == 1 else g_d).contiguous()
                io.g = g_d.data_ptr()
                keep.append(g_d)
            io.e_scalar = float(E)
            if want_current:
                cur = torch.empty((T, B), dtype=state_dtype, device=dev)
                io.i_out = cur.data_ptr()
            if data is not None:
                d_d = torch.as_tensor(data).to(device=dev, dtype=state_dtype).contiguous()
                if d_d.dim() == 1:
                    d_d = d_d.reshape(T, 1)
                io.data, io.data_B = d_d.data_ptr(), d_d.shape[1]
                loss = torch.empty((B, 2), dtype=torch.float64, device=dev)
                io.loss_out = loss.data_ptr()
                keep.append(d_d)
        if want_y:
            y_out = torch.empty((T, B, 2), dtype=state_dtype, device=dev)
            io.y_out = y_out.data_ptr()
        stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
        io.stats_out = stats.data_ptr()
        _cabi.check(lib.ikr_forward_hh(ctypes.byref(d), ctypes.byref(io), p_d.data_ptr(), sptr),
                    'ikr_forward_hh')
        res = IkrResult(y=y_out, current=cur, sse=None if loss is None else loss[:, 0],
                        sae=None if loss is None else loss[:, 1], stats=stats)
        res._keep = keep
        if opts.get('check_status', True):
            _raise_on_status(stats)
        return res


class HHPopulationModel:
    """Drop-in for the reference's ``Model(pints.ForwardModel)`` (``train-d0.py:377-439``) with a
    batched ``simulate_population``: same ``n_parameters / n_outputs / set_* / simulate``
    interface (PINTS only duck-types those), so ``pints.SumOfSquaresError`` etc. work unchanged,
    while an optimiser that evaluates a whole generation calls ``simulate_population(X, t)``."""

    def __init__(self, inactivation, y0=(0., 1.), E=-86.0, g=1.0, device=None):
        self._inact = tuple(float(x) for x in inactivation)
        self._ps = None
        self._protocol = None
        self._E, self._g = float(E), float(g)
        self._device = device
        self.set_y0(np.asarray(y0, dtype=np.float64))

    def n_parameters(self):
        return 4

    def n_outputs(self):
        return len(self._ps) if self._ps is not None else 1

    def set_fixed_form_voltage_protocol(self, t, v):
        self._protocol = (np.asarray(t, dtype=np.float64), np.asarray(v, dtype=np.float64))

    def set_voltage_protocol_batches(self, ps=None):
        "ps: list of voltage time series [times, voltages]"
        self._ps = ps

    def set_y0(self, y0=np.asarray([0, 1])):
        self._y0 = np.asarray(y0, dtype=np.float64).reshape(1, -1)

    def set_discontinous(self, discontn=None):   # noqa: D401 - reference spelling
        self.discontn = discontn                 # legacy grid_points: a no-op under torchdiffeq 0.2.1

    def simulate_population(self, X, t):
        """X: (B, 4) candidates -> currents (T, B) [one protocol] or (T, n_protocols, B)."""
        X = np.asarray(X, dtype=np.float64).reshape(-1, 4)
        B = X.shape[0]
        t_t = torch.from_numpy(np.array(t, copy=True))
        y0 = torch.from_numpy(np.repeat(self._y0, B, 0)).float()     # reference: fp32 y0 (:409)
        protos = self._ps if self._ps is not None else [np.stack(self._protocol, 1)]
        outs = []
        for pr in protos:
            pr = np.asarray(pr)
            res = integrate_hh(X, y0, t_t, (pr[:, 0], pr[:, 1]), method='dopri5', g=self._g,
                               E=self._E, want_y=False, want_current=True, device=self._device,
                               options={'inactivation': self._inact, 'check_status': False})
            cur = res.current.double()
            bad = res.stats[:, 3] != 0                                # reference: time-out -> inf
            cur[:, bad] = float('inf')
            outs.append(cur.cpu().numpy())
        return outs[0] if self._ps is None else np.stack(outs, 1)

    def simulate(self, x, t):
        "Pints's forward simulation, x parameters, t time series"
        out = self.simulate_population(np.asarray(x).reshape(1, -1), t)
        return out[..., 0] if self._ps is None else out[..., 0]
