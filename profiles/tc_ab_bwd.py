"""A/B of the tensor-core backward (adjoint + weight-gradient GEMM on tcgen05) against the FFMA2
backward on the same inputs: loss, flat gradient, grad_y0 and timing (one GPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402
from neural_ode_ion_channels_b200.adjoint import loss_and_grad  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fam = sys.argv[2] if len(sys.argv) > 2 else 'aps'
n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 60
study = sys.argv[4] if len(sys.argv) > 4 else 'd2'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', '%s-model-state-dict.pt' % study)
cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
f = ikr.load_weights(cls(params='d'), w).cuda()
name, t_tab, v_tab, t_out = protocols.protocol_set(fam)[10 if fam == 'pr4' else 0]
f.set_fixed_form_voltage_protocol(t_tab, v_tab)
rng = np.random.RandomState(0)
y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1), dtype=torch.float32).cuda()
t = torch.tensor(t_out[:n_out], dtype=torch.float32)
data = torch.from_numpy(rng.randn(n_out).astype(np.float32) * 0.1)
out = {}
for tcore in (False, True):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        total, per, grads, res = loss_and_grad(f, y0, t, data, want_y0=True,
                                               options={'tensor_cores': tcore, 'check_status': False})
        e1.record()
        torch.cuda.synchronize()
    flat = torch.cat([g.reshape(-1).double() for g in grads]).cpu().numpy()
    print('tensor_cores=%s %s B=%d T=%d: %.2f ms, loss %.9g, |grad|max %.6e, finite %s, acc steps %.1f'
          % (tcore, name, B, n_out, e0.elapsed_time(e1), float(total), np.abs(flat).max(),
             np.isfinite(flat).all(), float(res.stats[:, 0].float().mean())), flush=True)
    out[tcore] = (flat, res.grad_y0.cpu().double().numpy(), float(total))
a, b = out[True], out[False]
n = 200
segs = [('w0', 0, 2 * n), ('b0', 2 * n, 3 * n)]
o = 3 * n
for l in range(5):
    segs += [('W%d' % (l + 1), o, o + n * n), ('b%d' % (l + 1), o + n * n, o + n * n + n)]
    o += n * n + n
segs += [('w_last', o, o + n), ('b_last', o + n, o + n + 1)]
for nm, lo, hi in segs:
    d = np.abs(a[0][lo:hi] - b[0][lo:hi]).max()
    print('  %-7s max|tc - ffma| = %.3e   max|ffma| = %.3e   rel %.2e' % (nm, d, np.abs(b[0][lo:hi]).max(),
                                                                        d / max(np.abs(b[0][lo:hi]).max(), 1e-300)))
print('grad_y0: max|tc - ffma| = %.3e (max %.3e); loss rel diff %.2e'
      % (np.abs(a[1] - b[1]).max(), np.abs(b[1]).max(), abs(a[2] - b[2]) / abs(b[2])))
