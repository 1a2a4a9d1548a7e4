"""Small single-launch driver for ncu captures of the forward kernel (one GPU, one protocol)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 96
fam = sys.argv[2] if len(sys.argv) > 2 else 'pr4'
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == 'f64') else torch.float32
w = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                 'neural-ode-ion-channels_b200', 'data', 'weights', 'd1-model-state-dict.pt')
f = ikr.load_weights(ikr.ODEFunc(params='d'), w)
if dtype == torch.float64:
    f = f.double()
name, t_tab, v_tab, t_out = protocols.protocol_set(fam)[10 if fam == 'pr4' else 0]
f.set_fixed_form_voltage_protocol(t_tab, v_tab)
rng = np.random.RandomState(0)
y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1), dtype=dtype).cuda()
n_out = int(sys.argv[4]) if len(sys.argv) > 4 else len(t_out)
t = torch.tensor(t_out[:n_out], dtype=dtype)
with torch.no_grad():
    for _ in range(int(os.environ.get('REPS', '2'))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ikr.integrate(f, y0, t, want_y=False, want_current=False, data=torch.zeros(len(t)),
                          options={'lane_pool': {'1': True, '0': False}.get(os.environ.get('POOL', ''), None),
                                   'tensor_cores': not os.environ.get('NO_TC'),
                                   'tc_timing': bool(os.environ.get('TC_TIMING')),
                                   'tc_split': os.environ.get('TC_SPLIT') or None,
                                   'tc_groups': int(os.environ.get('TC_GROUPS', '0')),
                                   'ping_pong': {'1': True, '0': False}.get(os.environ.get('PP', ''), None)})
        e1.record()
        torch.cuda.synchronize()
        nfe = int(r.stats[:, 2].sum())
        ms = e0.elapsed_time(e1)
        st = r.stats.cpu().numpy()
        if os.environ.get('REPS'):
            print('  rep: %.2f ms' % ms, flush=True)
        print('%s B=%d %s: %.1f ms, NFE %d, %.2f M evals/s, %.2f TFLOP/s; steps/lane mean %.1f '
              'max %d min %d; geometry %s' % (name, B, dtype, ms, nfe, nfe / ms / 1e3,
                                              nfe * 401200 / ms / 1e9,
                                              (st[:, 0] + st[:, 1]).mean(), (st[:, 0] + st[:, 1]).max(),
                                              (st[:, 0] + st[:, 1]).min(), r.geometry))
