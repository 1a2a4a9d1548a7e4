"""Accuracy of the MLP arithmetic of each kernel family, measured through the public API.

One rk4 step of length dt = 1e-3 ms with an fp64 state resolves the RHS to full precision:
(y1 - y0) / dt = mean of the four stage derivatives.  The same step is taken with
  truth : fp64 state + fp64 MLP (DFMA kernel),
  ffma  : fp64 state + fp32 MLP on the FFMA2 kernel (plain fp32 FMAs),
  tc    : fp64 state + fp32 MLP on the tensor-core kernel (split operands, fp32 accumulation in TMEM),
for 4,096 random activation states at five clamp voltages.  Printed per model: the net-output scale,
and for ffma / tc the RMS and the MEAN SIGNED error of da/dt relative to that scale (a non-zero mean
is a bias: tensor-core fp32 accumulation truncates instead of rounding)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
modes = sys.argv[1:] or ['fp16x2', 'bf16x3']
dt = 1e-3
t = torch.tensor([0.0, dt], dtype=torch.float64)
rng = np.random.RandomState(0)
B = 4096
y0 = torch.tensor(np.stack([rng.uniform(0, 1, B), np.ones(B)], 1), dtype=torch.float64).cuda()
for study in ('d1', 's1', 'd2'):
    cls = ikr.ODEFuncNNd if study == 'd2' else ikr.ODEFuncNNf
    w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', '%s-model-state-dict.pt' % study)
    f32 = ikr.load_weights(cls(params='s' if study == 's1' else 'd'), w)
    f64 = ikr.load_weights(cls(params='s' if study == 's1' else 'd'), w).double()
    out = {}
    for v in (-100.0, -60.0, -20.0, 20.0, 40.0):
        tab = (np.array([-1.0, 1.0]), np.array([v, v]))
        for f in (f32, f64):
            f.set_fixed_form_voltage_protocol(*tab)
        with torch.no_grad():
            truth = (ikr.odeint(f64, y0, t, method='rk4')[1, :, 0] - y0[:, 0]) / dt
            runs = {'ffma': {'tensor_cores': False}}
            for m in modes:
                runs[m] = {'tensor_cores': True, 'tc_split': m}
            for name, opts in runs.items():
                opts = {k: v_ for k, v_ in opts.items() if v_ is not None}
                got = (ikr.odeint(f32, y0, t, method='rk4', options=opts)[1, :, 0] - y0[:, 0]) / dt
                out.setdefault(name, []).append((got - truth).cpu().numpy())
        out.setdefault('scale', []).append(truth.abs().cpu().numpy())
    scale = np.sqrt(np.mean(np.concatenate(out['scale']) ** 2))
    line = '%s: rms |da/dt| %.3e;' % (study, scale)
    for name in [k for k in out if k != 'scale']:
        e = np.concatenate(out[name])
        line += '  %s: rms err %.2e, mean signed err %+.2e, max %.2e (relative to scale)' % (
            name, np.sqrt(np.mean(e ** 2)) / scale, e.mean() / scale, np.abs(e).max() / scale)
    print(line, flush=True)
