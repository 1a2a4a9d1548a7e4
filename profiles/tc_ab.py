"""A/B of the tcgen05 forward kernel against the FFMA2 forward kernel on the same inputs:
per-sample agreement of the current traces, step counts, and timing (one GPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fam = sys.argv[2] if len(sys.argv) > 2 else 'aps'
n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 200
method = sys.argv[4] if len(sys.argv) > 4 else 'dopri5'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', 'd1-model-state-dict.pt')
f = ikr.load_weights(ikr.ODEFunc(params='d'), w)
name, t_tab, v_tab, t_out = protocols.protocol_set(fam)[10 if fam == 'pr4' else 0]
f.set_fixed_form_voltage_protocol(t_tab, v_tab)
rng = np.random.RandomState(0)
y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                  dtype=torch.float32).cuda()
t = torch.tensor(t_out[:n_out], dtype=torch.float32)
out = {}
with torch.no_grad():
    for tcore in (False, True):
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = ikr.integrate(f, y0, t, want_y=True, want_current=True, method=method,
                              options={'tensor_cores': tcore, 'check_status': False})
            e1.record()
            torch.cuda.synchronize()
        st = r.stats.cpu().numpy()
        nfe = int(st[:, 2].sum())
        ms = e0.elapsed_time(e1)
        print('tensor_cores=%s %s B=%d T=%d %s: %.2f ms NFE %d %.2f M evals/s acc %.1f rej %.1f status %s geo %s'
              % (tcore, name, B, n_out, method, ms, nfe, nfe / ms / 1e3, st[:, 0].mean(),
                 st[:, 1].mean(), np.unique(st[:, 3]), r.geometry), flush=True)
        out[tcore] = (r.y.cpu().double().numpy(), r.current.cpu().double().numpy())
dy = np.abs(out[True][0] - out[False][0])
di = np.abs(out[True][1] - out[False][1])
print('max |y_tc - y_ffma| = %.3e (a %.3e, r %.3e); max |I_tc - I_ffma| = %.3e; max |I| = %.3f'
      % (dy.max(), dy[..., 0].max(), dy[..., 1].max(), di.max(), np.abs(out[False][1]).max()))
print('finite:', np.isfinite(out[True][0]).all())
