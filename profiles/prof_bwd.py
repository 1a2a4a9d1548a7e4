"""Single-call driver for timing / ncu captures of the backward path (fused loss + gradient):
   python profiles/prof_bwd.py B family dtype n_out study"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fam = sys.argv[2] if len(sys.argv) > 2 else 'staircase'
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == 'f64') else torch.float32
n_out = int(sys.argv[4]) if len(sys.argv) > 4 else 0
study = sys.argv[5] if len(sys.argv) > 5 else 'd2'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', '%s-model-state-dict.pt' % study)
cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
f = ikr.load_weights(cls(params='d'), w)
if dtype == torch.float64:
    f = f.double()
name, t_tab, v_tab, t_out = protocols.protocol_set(fam)[10 if fam == 'pr4' else 0]
f.set_fixed_form_voltage_protocol(t_tab, v_tab)
rng = np.random.RandomState(0)
y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1), dtype=dtype).cuda()
n_out = n_out or len(t_out)
t = torch.tensor(t_out[:n_out], dtype=dtype)
data = torch.from_numpy(rng.randn(n_out, 1).astype(np.float32) * 0.1).to(dtype)
f.cuda()
cap = int(os.environ.get('CKPT_CAP', '0')) or None
for _ in range(int(os.environ.get('REPS', '2'))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    opts = {'check_status': False, 'tensor_cores': not os.environ.get('NO_TC'),
                                   'tc_timing': bool(os.environ.get('TC_TIMING')),
                                   'tc_split': os.environ.get('TC_SPLIT') or None,
                                   'tc_groups': int(os.environ.get('TC_GROUPS', '0')),
                                   'ping_pong': {'1': True, '0': False}.get(os.environ.get('PP', ''), None),
            'lane_pool': {'1': True, '0': False}.get(os.environ.get('POOL', ''), None),
            'adjoint_products': os.environ.get('ADJ') or None}
    if cap:
        opts['ckpt_cap'] = cap
    res = ikr.integrate(f, y0, t, data=data, want_y=True, want_ckpt=True, options=opts)
    ev[1].record()
    from neural_ode_ion_channels_b200.adjoint import _run_backward
    flat, _, _ = _run_backward(f, res, fused_loss=1, want_y0=False)
    ev[2].record()
    torch.cuda.synchronize()
    st = res.stats.cpu().numpy()
    assert (st[:, 3] == 0).all(), np.unique(st[:, 3])
    nfe_f = int(st[:, 2].sum())
    nfe_b = int((6 * st[:, 0] + 1).sum())
    ms_f, ms_b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    flops = nfe_f * 401200 + nfe_b * 3 * 401200
    print('%s %s B=%d %s: fwd %.1f ms (%.2f M evals/s, %.1f TF/s) | bwd %.1f ms (%.2f M adjoint '
          'evals/s, %.1f TF/s algorithmic) | fwd+bwd %.2f M evals/s, %.1f TF/s; acc steps mean %.1f '
          'max %d; |grad|max %.3e'
          % (study, name, B, dtype, ms_f, nfe_f / ms_f / 1e3, nfe_f * 401200 / ms_f / 1e9, ms_b,
             nfe_b / ms_b / 1e3, nfe_b * 3 * 401200 / ms_b / 1e9, (nfe_f + nfe_b) / (ms_f + ms_b) / 1e3,
             flops / (ms_f + ms_b) / 1e9, st[:, 0].mean(), st[:, 0].max(), float(flat.abs().max())))
