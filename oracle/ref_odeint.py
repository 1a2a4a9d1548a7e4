"""CPU ORACLE (test infrastructure, NOT product code) -- ``odeint`` restated.

The reference calls ``torchdiffeq.odeint`` (``requirements.txt:1`` pins ``torchdiffeq==0.2.1``;
call sites e.g. ``train-s1.py:322,327``, ``table-1.py:404,413``).  That package is a third-party
dependency which is neither vendored under ``/root/reference`` nor installable here (no network,
no wheel), so this file restates its *published algorithm* (Dormand-Prince 5(4) with Shampine's
dense output as implemented by torchdiffeq 0.2.x, and the fixed-grid 3/8-rule ``rk4``), following
SURVEY.md Appendix A.  If the real package is importable, ``odeint`` below defers to it.

PARITY PINNING: the restatement is pinned by the reference's own logged results -- the
``{s1,s2,d1,d2}/log2`` losses that are reproducible with in-repo inputs (``tests/test_oracle_kat.py``
and ``tests/golden/kat_log2.json``).  No tensor-level golden vector exists at the ``odeint``
boundary in the reference, so tensor-level parity rests on those known answers.

dtype semantics reproduced (torchdiffeq 0.2.x): times / step sizes / tolerances live in fp64;
state, stage derivatives ``k``, the tableau and the interpolation coefficients live in
``y0.dtype``; inside one RK step ``t0, dt, t1`` are cast to ``y0.dtype`` first; the RHS wrapper
casts ``t`` to ``y.dtype`` and, for the alpha==1 stages of dopri5, replaces it by
``nextafter(t, t-1)``; step-size control runs under ``no_grad``.

Everything is written with torch ops so that ``loss.backward()`` through this function *is* the
gradient oracle (torchdiffeq non-adjoint autograd semantics).
"""
import math
import warnings

import torch

try:  # pragma: no cover - not available in this image
    import torchdiffeq as _real_tde
except Exception:  # noqa: BLE001
    _real_tde = None

_PREV, _NONE, _NEXT = -1, 0, 1

# Dormand-Prince tableau (Shampine form)
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [
    35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.,
]
_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _rms(x):
    return x.pow(2).mean().sqrt()


class _TimeCast:
    """RHS wrapper: cast t to the state dtype and apply the PREV/NEXT perturbation."""

    def __init__(self, func):
        self.func = func
        self.nfe = 0

    def __call__(self, t, y, perturb=_NONE):
        self.nfe += 1
        t = t.to(y.dtype)
        if perturb == _NEXT:
            t = torch.nextafter(t, t + 1)
        elif perturb == _PREV:
            t = torch.nextafter(t, t - 1)
        return self.func(t, y)


def _initial_step(f, t0, y0, order, rtol, atol, f0):
    """Hairer/Norsett/Wanner starting step size as used by torchdiffeq (one extra RHS eval)."""
    dtype = y0.dtype
    t_dtype = t0.dtype
    t0 = t0.to(dtype)
    scale = atol + torch.abs(y0) * rtol
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype)
    else:
        h0 = 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = f(t0 + h0, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1. / float(order + 1))
    return torch.min(100 * h0, h1).to(t_dtype)


@torch.no_grad()
def _next_step_size(dt, ratio, safety, ifactor, dfactor, order):
    if ratio == 0:
        return dt * ifactor
    if ratio < 1:
        dfactor = torch.ones((), dtype=dt.dtype)
    ratio = ratio.type_as(dt)
    exponent = torch.tensor(order, dtype=dt.dtype).reciprocal()
    factor = torch.min(ifactor, torch.max(safety / ratio ** exponent, dfactor))
    return dt * factor


def _dense_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _dense_eval(coef, t0, t1, t):
    assert (t0 <= t) & (t <= t1), 'invalid interpolation, fails `t0 <= t <= t1`'
    x = ((t - t0) / (t1 - t0)).to(coef[0].dtype)
    total = coef[0] + x * coef[1]
    xp = x
    for c in coef[2:]:
        xp = xp * x
        total = total + xp * c
    return total


class Dopri5:
    order = 5

    def __init__(self, func, y0, rtol, atol, first_step=None, safety=0.9, ifactor=10.0,
                 dfactor=0.2, max_num_steps=2 ** 31 - 1, detach_first_step=False, record=None,
                 replay=None):
        self.f = _TimeCast(func)
        self.y0 = y0
        tdt = torch.promote_types(torch.float64, y0.dtype)
        self.tdtype = tdt
        self.rtol = torch.as_tensor(rtol, dtype=tdt)
        self.atol = torch.as_tensor(atol, dtype=tdt)
        self.first_step = None if first_step is None else torch.as_tensor(first_step, dtype=tdt)
        self.safety = torch.as_tensor(safety, dtype=tdt)
        self.ifactor = torch.as_tensor(ifactor, dtype=tdt)
        self.dfactor = torch.as_tensor(dfactor, dtype=tdt)
        self.max_num_steps = max_num_steps
        self.detach_first_step = detach_first_step
        sd = y0.dtype
        self.alpha = torch.tensor(_ALPHA, dtype=torch.float64).to(sd)
        self.beta = [torch.tensor(b, dtype=torch.float64).to(sd) for b in _BETA]
        self.c_err = torch.tensor(_C_ERR, dtype=torch.float64).to(sd)
        self.c_mid = torch.tensor(_C_MID, dtype=torch.float64).to(sd)
        self.n_accept = 0
        self.n_reject = 0
        self.record = record          # optional list receiving (t0, dt, accepted) per attempt
        # TEST HOOK (not torchdiffeq): `replay` = [(t0, dt), ...] of the accepted steps of another
        # run.  The controller is bypassed and exactly these steps are taken, so that gradients
        # can be compared step-for-step (the step-size sequence is ill-conditioned w.r.t. rounding:
        # two correct implementations drift apart by ~1e-7 relative in dt after a stiff phase).
        self.replay = None if replay is None else list(replay)

    # one Runge-Kutta attempt ---------------------------------------------------------------
    def _rk_step(self, y0, f0, t0, dt, t1):
        sd = y0.dtype
        t0 = t0.to(sd)
        dt = dt.to(sd)
        t1 = t1.to(sd)
        k = [f0.to(sd).reshape(y0.shape)]
        yi = y0
        for i in range(6):
            if _ALPHA[i] == 1.:
                ti, pert = t1, _PREV
            else:
                ti, pert = t0 + self.alpha[i] * dt, _NONE
            acc = k[0] * (self.beta[i][0] * dt)
            for j in range(1, i + 1):
                acc = acc + k[j] * (self.beta[i][j] * dt)
            yi = y0 + acc
            k.append(self.f(ti, yi, perturb=pert).to(sd).reshape(y0.shape))
        y1 = yi
        f1 = k[-1]
        err = k[0] * (dt * self.c_err[0])
        for j in range(1, 7):
            err = err + k[j] * (dt * self.c_err[j])
        return y1, f1, err, k

    def _mid_fit(self, y0, y1, k, dt):
        dt = dt.type_as(y0)
        acc = k[0] * (dt * self.c_mid[0])
        for j in range(1, 7):
            acc = acc + k[j] * (dt * self.c_mid[j])
        y_mid = y0 + acc
        return _dense_fit(y0, y1, y_mid, k[0], k[-1], dt)

    def integrate(self, t):
        y0 = self.y0
        sol = [y0]
        t = t.to(self.tdtype)
        f0 = self.f(t[0], y0)
        if self.first_step is None:
            dt = _initial_step(self.f, t[0], y0, self.order - 1, self.rtol, self.atol, f0)
            if self.detach_first_step:
                dt = dt.detach()
        else:
            dt = self.first_step
        # state: (y, f, t_lo, t_hi, dt, coef)
        y, f, t_lo, t_hi, coef = y0, f0, t[0], t[0], [y0] * 5
        for i in range(1, len(t)):
            n_steps = 0
            while t[i] > t_hi:
                assert n_steps < self.max_num_steps, 'max_num_steps exceeded'
                ts = t_hi
                t1 = ts + dt
                assert ts + dt > ts, 'underflow in dt {}'.format(dt.item())
                assert torch.isfinite(y).all(), 'non-finite values in state `y`: {}'.format(y)
                if self.replay is not None:
                    r_t0, r_dt = self.replay[self.n_accept]
                    ts = torch.as_tensor(r_t0, dtype=self.tdtype)
                    dt = torch.as_tensor(r_dt, dtype=self.tdtype)
                    t1 = ts + dt
                y1, f1, err, k = self._rk_step(y, f, ts, dt, t1)
                tol = self.atol + self.rtol * torch.max(y.abs(), y1.abs())
                ratio = _rms(err / tol)
                accept = bool(ratio <= 1) or self.replay is not None
                if self.record is not None:
                    self.record.append((float(ts), float(dt), accept))
                if accept:
                    coef = self._mid_fit(y, y1, k, dt)
                    y, f, t_lo, t_hi = y1, f1, ts, t1
                    self.n_accept += 1
                else:
                    t_lo, t_hi = ts, ts
                    self.n_reject += 1
                dt = _next_step_size(dt, ratio, self.safety, self.ifactor, self.dfactor,
                                     self.order)
                n_steps += 1
            sol.append(_dense_eval(coef, t_lo, t_hi, t[i]))
        return torch.stack([s.reshape(y0.shape).to(y0.dtype) for s in sol])


class RK4:
    """Fixed-grid 3/8-rule Runge-Kutta ("rk4" in torchdiffeq 0.2.x); grid = t unless
    ``step_size`` is given; outputs linearly interpolated when the grid differs from t."""
    order = 4

    def __init__(self, func, y0, step_size=None, perturb=False, **unused):
        self.f = _TimeCast(func)
        self.y0 = y0
        self.step_size = step_size
        self.perturb = perturb
        self.n_accept = 0
        self.n_reject = 0

    def _grid(self, t):
        if self.step_size is None:
            return t
        h = self.step_size
        start, end = t[0], t[-1]
        niters = torch.ceil((end - start) / h + 1).item()
        grid = torch.arange(0, niters, dtype=t.dtype) * h + start
        grid[-1] = t[-1]
        return grid

    def _step(self, t0, dt, t1, y0):
        f = self.f
        k1 = f(t0, y0, perturb=_NEXT if self.perturb else _NONE).reshape(y0.shape)
        k2 = f(t0 + dt * (1 / 3), y0 + dt * k1 * (1 / 3)).reshape(y0.shape)
        k3 = f(t0 + dt * (2 / 3), y0 + dt * (k2 - k1 * (1 / 3))).reshape(y0.shape)
        k4 = f(t1, y0 + dt * (k1 - k2 + k3),
               perturb=_PREV if self.perturb else _NONE).reshape(y0.shape)
        return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125

    def integrate(self, t):
        grid = self._grid(t)
        assert grid[0] == t[0] and grid[-1] == t[-1]
        y0 = self.y0
        sol = [y0]
        j = 1
        for t0, t1 in zip(grid[:-1], grid[1:]):
            dt = t1 - t0
            y1 = y0 + self._step(t0, dt, t1, y0)
            self.n_accept += 1
            while j < len(t) and t1 >= t[j]:
                if t[j] == t0:
                    sol.append(y0)
                elif t[j] == t1:
                    sol.append(y1)
                else:
                    slope = (t[j] - t0) / (t1 - t0)
                    sol.append(y0 + slope * (y1 - y0))
                j += 1
            y0 = y1
        return torch.stack([s.reshape(self.y0.shape).to(self.y0.dtype) for s in sol])


_KNOWN_ADAPTIVE = {'first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps',
                   'detach_first_step', 'record', 'replay'}
_KNOWN_FIXED = {'step_size', 'perturb'}


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, stats=None,
           prefer_real=False):
    """Restated ``torchdiffeq.odeint`` (B=1 semantics exactly like the reference uses it).

    ``stats`` (dict, optional) receives ``n_accept, n_reject, nfe``.  ``prefer_real=True`` routes
    to the real package when it is importable (it is not in this image)."""
    if prefer_real and _real_tde is not None:  # pragma: no cover
        return _real_tde.odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
    method = method or 'dopri5'
    options = dict(options or {})
    if method == 'dopri5':
        unknown = set(options) - _KNOWN_ADAPTIVE
        for key in unknown:            # torchdiffeq 0.2.x: "unexpected arguments" warning only
            options.pop(key)
        if unknown:
            warnings.warn('Dopri5: Unexpected arguments {}'.format(sorted(unknown)))
        solver = Dopri5(func, y0, rtol, atol, **options)
    elif method == 'rk4':
        unknown = set(options) - _KNOWN_FIXED
        for key in unknown:
            options.pop(key)
        if unknown:
            warnings.warn('RK4: Unexpected arguments {}'.format(sorted(unknown)))
        solver = RK4(func, y0, **options)
    else:
        raise ValueError('oracle restates only dopri5 and rk4, got {!r}'.format(method))
    out = solver.integrate(t)
    if stats is not None:
        stats['n_accept'] = solver.n_accept
        stats['n_reject'] = solver.n_reject
        stats['nfe'] = solver.f.nfe
    return out
