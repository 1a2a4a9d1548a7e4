"""ctypes driver of tests/host_harness.cpp (TEST INFRASTRUCTURE: host build of the lane state
machine in csrc/ikr_math.h with a scalar MLP)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, 'neural-ode-ion-channels_b200', 'csrc')
_SO = os.path.join(HERE, '_host_harness.so')


class Args(ctypes.Structure):
    _fields_ = [
        ('L', ctypes.c_int), ('n', ctypes.c_int), ('nn_d', ctypes.c_int), ('method', ctypes.c_int),
        ('time_f32', ctypes.c_int), ('rk4_perturb', ctypes.c_int),
        ('params', ctypes.c_void_p),
        ('tab_t', ctypes.c_void_p), ('tab_v', ctypes.c_void_p),
        ('tab_len', ctypes.c_int), ('tab_uniform', ctypes.c_int),
        ('tab_t0', ctypes.c_double), ('tab_inv_dt', ctypes.c_double),
        ('p', ctypes.c_void_p),
        ('rtol', ctypes.c_double), ('atol', ctypes.c_double), ('first_step', ctypes.c_double),
        ('t_out', ctypes.c_void_p), ('T', ctypes.c_int),
        ('grid', ctypes.c_void_p), ('G', ctypes.c_int),
        ('y0a', ctypes.c_double), ('y0r', ctypes.c_double),
        ('y_out', ctypes.c_void_p), ('stats', ctypes.c_void_p),
        ('steps', ctypes.c_void_p), ('steps_cap', ctypes.c_int),
    ]


def build():
    src = os.path.join(HERE, 'host_harness.cpp')
    hdr = os.path.join(CSRC, 'ikr_math.h')
    if (not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src),
                                                               os.path.getmtime(hdr))):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-mfma',
                               '-ffp-contract=off', '-I', CSRC, src, '-o', _SO])
    return ctypes.CDLL(_SO)


def flat_params(net, dtype):
    import torch
    parts = []
    for m in net:
        if isinstance(m, torch.nn.Linear):
            parts.append(m.weight.detach().to(dtype).reshape(-1))
            parts.append(m.bias.detach().to(dtype).reshape(-1))
    return torch.cat(parts).contiguous().numpy()


def uniform_hint(t):
    t = np.asarray(t, dtype=np.float64)
    if len(t) < 3:
        return 0, 0.0, 0.0
    h = (t[-1] - t[0]) / (len(t) - 1)
    ok = np.all(np.abs(t - (t[0] + h * np.arange(len(t)))) < 0.25 * h)
    return (1, float(t[0]), float(1.0 / h)) if ok else (0, 0.0, 0.0)


def integrate(net, L, n, nn_d, p8, tab_t, tab_v, y0, t_out, state_f64, mlp_f64, method='dopri5',
              rtol=1e-7, atol=1e-9, first_step=0.0, grid=None, time_f32=False, perturb=False,
              steps_cap=0):
    import torch
    lib = build()
    params = flat_params(net, torch.float64 if mlp_f64 else torch.float32)
    tab_t = np.ascontiguousarray(tab_t, dtype=np.float64)
    tab_v = np.ascontiguousarray(tab_v, dtype=np.float64)
    t_out = np.ascontiguousarray(t_out, dtype=np.float64)
    grid = t_out if grid is None else np.ascontiguousarray(grid, dtype=np.float64)
    p = np.ascontiguousarray(p8, dtype=np.float64)
    y_out = np.zeros((len(t_out), 2))
    stats = np.zeros(4, dtype=np.int32)
    steps = np.zeros((max(steps_cap, 1), 2))
    uni, t0, inv = uniform_hint(tab_t)
    a = Args(L, n, int(nn_d), 0 if method == 'dopri5' else 1, int(time_f32), int(perturb),
             params.ctypes.data, tab_t.ctypes.data, tab_v.ctypes.data, len(tab_t), uni, t0, inv,
             p.ctypes.data, rtol, atol, first_step, t_out.ctypes.data, len(t_out),
             grid.ctypes.data, len(grid), float(y0[0]), float(y0[1]), y_out.ctypes.data,
             stats.ctypes.data, steps.ctypes.data if steps_cap else None, steps_cap)
    rc = lib.harness_integrate(int(state_f64), int(mlp_f64), ctypes.byref(a))
    assert rc == 0
    return y_out, stats, steps[:stats[0]] if steps_cap else None


def table_voltage(tab_t, tab_v, x):
    lib = build()
    lib.harness_table_voltage.restype = ctypes.c_double
    tab_t = np.ascontiguousarray(tab_t, dtype=np.float64)
    tab_v = np.ascontiguousarray(tab_v, dtype=np.float64)
    uni, t0, inv = uniform_hint(tab_t)
    flag = ctypes.c_int(0)
    v = lib.harness_table_voltage(ctypes.c_void_p(tab_t.ctypes.data),
                                  ctypes.c_void_p(tab_v.ctypes.data), len(tab_t), uni,
                                  ctypes.c_double(t0), ctypes.c_double(inv), ctypes.c_double(x),
                                  ctypes.byref(flag))
    return v, bool(flag.value)


class GradArgs(ctypes.Structure):
    _fields_ = [('grad_y', ctypes.c_void_p), ('grad_params', ctypes.c_void_p),
                ('grad_y0', ctypes.c_void_p), ('steps_cap', ctypes.c_int)]


def gradient(net, L, n, nn_d, p8, tab_t, tab_v, y0, t_out, grad_y, state_f64, mlp_f64, rtol=1e-7,
             atol=1e-9, first_step=0.0):
    """Host build of the lane adjoint (ikr_math.h) + scalar MLP backward: dL/dparams (flat,
    state_dict order) and dL/dy0 for L = sum(grad_y * y_out)."""
    import torch
    lib = build()
    params = flat_params(net, torch.float64 if mlp_f64 else torch.float32)
    tab_t = np.ascontiguousarray(tab_t, dtype=np.float64)
    tab_v = np.ascontiguousarray(tab_v, dtype=np.float64)
    t_out = np.ascontiguousarray(t_out, dtype=np.float64)
    p = np.ascontiguousarray(p8, dtype=np.float64)
    y_out = np.zeros((len(t_out), 2))
    stats = np.zeros(4, dtype=np.int32)
    uni, t0, inv = uniform_hint(tab_t)
    a = Args(L, n, int(nn_d), 0, 0, 0, params.ctypes.data, tab_t.ctypes.data, tab_v.ctypes.data,
             len(tab_t), uni, t0, inv, p.ctypes.data, rtol, atol, first_step, t_out.ctypes.data,
             len(t_out), t_out.ctypes.data, len(t_out), float(y0[0]), float(y0[1]),
             y_out.ctypes.data, stats.ctypes.data, None, 0)
    gy = np.ascontiguousarray(grad_y, dtype=np.float64)
    gp = np.zeros(len(params))
    g0 = np.zeros(2)
    g = GradArgs(gy.ctypes.data, gp.ctypes.data, g0.ctypes.data, 0)
    rc = lib.harness_grad(int(state_f64), int(mlp_f64), ctypes.byref(a), ctypes.byref(g))
    assert rc == 0
    return gp, g0, y_out, stats
