"""BASELINE configs[3]: forward evals/s of every architectures/sNN.py network (s00-s11), fp32 as
shipped and fp64, same inputs (B trajectories of the pr4 stand-in, first 400 outputs), one GPU.
Which kernel family runs each case is printed (tensor cores: fp32 MLP with n <= 200 and n >= 16)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
n_out = int(sys.argv[2]) if len(sys.argv) > 2 else 400
name, t_tab, v_tab, t_out = protocols.protocol_set('pr4')[10]
rng = np.random.RandomState(0)
y0np = np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1)
print('| arch | (L, n) | MACs | fp32: kernel, M evals/s, TFLOP/s | fp64: M evals/s, TFLOP/s |')
print('|---|---|---|---|---|')
for arch, (L, n) in ikr.ARCHITECTURES.items():
    torch.manual_seed(0)
    macs = 2 * n + L * n * n + n
    row = '| %s | (%d, %d) | %d |' % (arch, L, n, macs)
    for dtype in (torch.float32, torch.float64):
        f = ikr.ODEFunc(arch=arch)
        if dtype == torch.float64:
            f = f.double()
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
        y0 = torch.tensor(y0np, dtype=dtype).cuda()
        t = torch.tensor(t_out[:n_out], dtype=dtype)
        with torch.no_grad():
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = ikr.integrate(f, y0, t, want_y=False, data=torch.zeros(len(t)),
                                  options={'check_status': False})
                e1.record()
                torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nfe = int(r.stats[:, 2].sum())
        ok = int((r.stats[:, 3] != 0).sum()) == 0
        kern = 'tcgen05' if r.geometry['tensor_cores'] else ('FFMA2' if dtype == torch.float32 else 'DFMA')
        row += ' %s%s, %.1f, %.1f |' % (kern, '' if ok else ' (status!)', nfe / ms / 1e3,
                                      nfe * 2 * macs / ms / 1e9) if dtype == torch.float32 else \
               ' %.1f, %.1f |' % (nfe / ms / 1e3, nfe * 2 * macs / ms / 1e9)
    print(row, flush=True)
