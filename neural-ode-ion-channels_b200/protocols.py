"""Voltage-clamp protocol tables (time in ms, voltage in mV) for the IKr hot path.

The step protocols are the ones the reference *defines in code*; the CSV-backed protocols whose
files are absent from the reference checkout (``.MISSING_LARGE_BLOBS``) are replaced by labelled
synthetic stand-ins.  Citations are relative to ``/root/reference``.

Every builder returns ``(t_ms, v_mV)`` float64 numpy arrays, exactly what
``func.set_fixed_form_voltage_protocol(t, v)`` takes (``train-s1.py:218-222``).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

PR3_STEPS = (-60, -40, -20, 0, 20, 40, 60)                    # train-s1.py:76, :442
PR5_STEPS = (-120, -110, -100, -90, -80, -70, -60, -50, -40)  # train-s1.py:91, :482
PR2_DURATIONS = (3, 10, 30, 100, 300, 1000)                   # train-s1.py:517 (ms)
PR4_STEPS = tuple(range(-100, 60, 10))                        # stand-in: 16 sweeps


def _grid(t_end_ms, per_ms):
    n = int(round(t_end_ms * per_ms)) + 1
    return np.linspace(0., float(t_end_ms), n), per_ms


def pr3_activation(v_step, per_ms=1):
    """Steady-state activation sweep: -80 | v_step (5 s) | -40 (1 s) | -120 (0.5 s) | -80.

    ``per_ms=1``: the 8,001-sample prediction table of ``train-s1.py:431-444``;
    ``per_ms=10``: the 80,001-sample 0.1 ms training table of ``train-s1.py:69-80``."""
    t, s = _grid(8000, per_ms)
    v = np.zeros(t.shape)
    v[:1000 * s] = -80
    v[1000 * s:6000 * s] = v_step
    v[6000 * s:7000 * s] = -40
    v[7000 * s:7500 * s] = -120
    v[7500 * s:] = -80
    return t, v


def pr5_deactivation(v_step, per_ms=1):
    """Deactivation sweep: -80 | +50 (2 s) | v_step (6 s) | -120 (0.5 s) | -80
    (``train-s1.py:471-484`` at 1 ms, ``:84-95`` at 0.1 ms)."""
    t, s = _grid(10000, per_ms)
    v = np.zeros(t.shape)
    v[:1000 * s] = -80
    v[1000 * s:3000 * s] = 50
    v[3000 * s:9000 * s] = v_step
    v[9000 * s:9500 * s] = -120
    v[9500 * s:] = -80
    return t, v


def pr2_time_constant(t_step_ms, per_ms=1):
    """Activation time constant at +40 mV: -80 | +40 (t_step) | -120 (2.5 s) | -80
    (``train-s1.py:511-521``)."""
    t, s = _grid(5000, per_ms)
    n = int(t_step_ms * s)
    v = np.zeros(t.shape)
    v[:1000 * s] = -80
    v[1000 * s:1000 * s + n] = 40
    v[1000 * s + n:3500 * s + n] = -120
    v[3500 * s + n:] = -80
    return t, v


def ap2hz():
    """``test-protocols/ap2hz.csv`` (35,000 samples, 0.1 ms grid, 0-3499.9 ms), seconds -> ms as
    in ``train-s1.py:44-45``.  Shipped as a binary fixture (values bit-identical to the CSV)."""
    blob = np.load(os.path.join(_DATA, 'ap2hz.npz'))
    return blob['t_ms'].copy(), blob['v_mV'].copy()


def load_protocol_csv(path):
    """Reference CSV layout: header line, columns time[s], voltage[mV] (``train-s1.py:44-45``)."""
    arr = np.loadtxt(path, skiprows=1, delimiter=',')
    return arr[:, 0] * 1e3, arr[:, 1].copy()


# ---------------------------------------------------------------------------------------------
# Labelled stand-ins for protocols whose CSVs are missing from the reference checkout
# ---------------------------------------------------------------------------------------------
def staircase_standin(per_ms=10):
    """STAND-IN for ``test-protocols/staircase.csv`` (15 s): hold -80, a -120 -> -80 ramp, +-20 mV
    500 ms stairs -40 ... +40 ... -60, a closing ramp.  150,001 samples at 0.1 ms."""
    t, s = _grid(15000, per_ms)
    v = np.full(t.shape, -80.0)
    v[250 * s:300 * s] = -120
    ramp = slice(300 * s, 700 * s)
    v[ramp] = np.linspace(-120, -80, 400 * s, endpoint=False)
    v[700 * s:900 * s] = -80
    v[900 * s:1900 * s] = 40
    v[1900 * s:2400 * s] = -120
    v[2400 * s:3400 * s] = -80
    levels = [-40, -60, -20, -40, 0, -20, 20, 0, 40, 20, 40, 0, 20, -20, 0, -40, -20, -60, -40]
    for i, lv in enumerate(levels):
        v[(3400 + 500 * i) * s:(3900 + 500 * i) * s] = lv
    end = 3400 + 500 * len(levels)                       # 12,900 ms
    v[end * s:(end + 500) * s] = -80
    v[(end + 500) * s:(end + 1000) * s] = 40
    v[(end + 1000) * s:(end + 1100) * s] = -70
    r2 = slice((end + 1100) * s, (end + 1200) * s)
    v[r2] = np.linspace(-70, -110, 100 * s, endpoint=False)
    v[(end + 1200) * s:(end + 1600) * s] = -120
    v[(end + 1600) * s:] = -80
    return t, v


def pr4_inactivation_standin(v_step, per_ms=10):
    """STAND-IN for Pr4 (inactivation): -80 | +50 (600 ms) | v_step (150 ms) | -80; 1.5 s sweep."""
    t, s = _grid(1500, per_ms)
    v = np.full(t.shape, -80.0)
    v[100 * s:700 * s] = 50
    v[700 * s:850 * s] = v_step
    return t, v


def sinewave_standin(per_ms=10):
    """STAND-IN for ``test-protocols/sinewave.csv`` (Beattie et al. 2018 form, 8 s): steps, then
    -30 + 54 sin(0.007 t') + 26 sin(0.037 t') + 10 sin(0.19 t') on 3000.1-6500.1 ms (window per
    ``train-r1.py:107-108``), then closing steps."""
    t, s = _grid(8000, per_ms)
    v = np.full(t.shape, -80.0)
    v[250 * s:300 * s] = -120
    v[500 * s:1500 * s] = 40
    v[1500 * s:2000 * s] = -120
    lo, hi = 3000 * s + 1, 6500 * s + 1
    tp = t[lo:hi] - 2500.0
    v[lo:hi] = -30 + 54 * np.sin(0.007 * tp) + 26 * np.sin(0.037 * tp) + 10 * np.sin(0.19 * tp)
    v[hi:7000 * s] = -120
    return t, v


PROTOCOL_FAMILIES = ('pr3', 'pr4', 'pr5', 'sinewave', 'aps')


def protocol_set(family, per_ms=1):
    """All sweeps of one family as a list of ``(name, t_ms, v_mV, t_out_ms)``.  ``t_out`` follows
    the reference prediction grids (``train-s1.py:65,268-276,431,471``)."""
    out = []
    if family == 'pr3':
        for vs in PR3_STEPS:
            t, v = pr3_activation(vs, per_ms)
            out.append(('pr3[%+d mV]' % vs, t, v, np.linspace(0., 8000., 8001)))
    elif family == 'pr5':
        for vs in PR5_STEPS:
            t, v = pr5_deactivation(vs, per_ms)
            out.append(('pr5[%+d mV]' % vs, t, v, np.linspace(0., 10000., 10001)))
    elif family == 'pr4':
        for vs in PR4_STEPS:
            t, v = pr4_inactivation_standin(vs)
            out.append(('pr4-standin[%+d mV]' % vs, t, v, np.linspace(0., 1500., 1501)))
    elif family == 'sinewave':
        t, v = sinewave_standin()
        out.append(('sinewave-standin', t, v, np.linspace(0., 8000., 4001)))
    elif family == 'aps':
        t, v = ap2hz()
        out.append(('ap2hz', t, v, np.linspace(0., 3000., 1501)))
    elif family == 'staircase':
        t, v = staircase_standin()
        out.append(('staircase-standin', t, v, np.linspace(0., 15000., 7501)))
    else:
        raise KeyError(family)
    return out


def concatenate_sweeps(sweeps, gap_ms=None):
    """Sweeps of one protocol laid end to end on ONE strictly increasing time axis, the way the
    real-data files hold them (``train-r1.py:80-94``: e.g. the 7 sweeps of Pr3 are one CSV, integrated
    by a single ``odeint`` call whose state carries across sweeps and sliced into ``l = len / 7``
    pieces afterwards, ``train-r1.py:313-329``).  ``sweeps``: list of ``(name, t, v, t_out)`` as
    returned by ``protocol_set``; sweep k is shifted by k x (duration + one sample period).
    Returns ``(t_table, v_table, t_out, n_sweeps)``."""
    ts, vs, outs, off = [], [], [], 0.0
    for _, t, v, t_out in sweeps:
        t = np.asarray(t, dtype=np.float64)
        dt = float(t[1] - t[0]) if gap_ms is None else float(gap_ms)
        ts.append(t - t[0] + off)
        vs.append(np.asarray(v, dtype=np.float64))
        outs.append(np.asarray(t_out, dtype=np.float64) - t[0] + off)
        off += float(t[-1] - t[0]) + dt
    return np.concatenate(ts), np.concatenate(vs), np.concatenate(outs), len(sweeps)


def compact_table(t, v):
    """Drop interior samples of runs where V is *exactly* constant.  scipy's linear ``interp1d``
    evaluates ``slope * (x - x_lo) + y_lo`` with ``slope == 0`` on such a run, so the compacted
    table interpolates bit-identically while step protocols shrink from 8,001-100,001 samples to a
    few dozen breakpoints (small enough to be staged in shared memory)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    n = len(t)
    if n <= 2:
        return t.copy(), v.copy()
    keep = np.ones(n, dtype=bool)
    same_prev = v[1:-1] == v[:-2]
    same_next = v[1:-1] == v[2:]
    keep[1:-1] = ~(same_prev & same_next)
    return t[keep].copy(), v[keep].copy()
