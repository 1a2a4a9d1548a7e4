"""Tiny invocations of every tensor-core / neighbouring-stage kernel (forward tile + pool + rk4,
backward, regression, Markov): a one-minute smoke of all paths on one GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import protocols  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', 'd2-model-state-dict.pt')
f = ikr.load_weights(ikr.ODEFuncNNd(params='d'), w).cuda()
t_tab, v_tab = protocols.ap2hz()
f.set_fixed_form_voltage_protocol(t_tab, v_tab)
rng = np.random.RandomState(0)
B = 130
y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1), dtype=torch.float32).cuda()
t = torch.linspace(0., 12., 7)
data = torch.zeros(7)
with torch.no_grad():
    r = ikr.integrate(f, y0, t, data=data, want_current=True)
    print('forward tile', r.geometry['tensor_cores'], int(r.stats[:, 2].sum()))
    r = ikr.integrate(f, y0, t, data=data, options={'lane_pool': True})
    print('forward pool', int(r.stats[:, 2].sum()))
    r = ikr.integrate(f, y0, t, method='rk4')
    print('forward rk4', int(r.stats[:, 2].sum()))
    r = ikr.integrate(f, y0, t, data=data, options={'tc_split': 'bf16x3'})
    print('forward bf16x3 split', r.geometry['mma_split'], int(r.stats[:, 2].sum()))
    r = ikr.integrate(f, torch.cat([y0, y0, y0]), t, data=data, options={'ping_pong': True})
    print('forward two-tile ping-pong', r.geometry['scheduling'], int(r.stats[:, 2].sum()))
    r = ikr.integrate(f, y0, t, data=data, options={'tensor_cores': False})
    print('forward FFMA2', int(r.stats[:, 2].sum()))
total, per, grads, res = ikr.loss_and_grad(f, y0, t, data, want_y0=True)
print('backward', float(total), float(max(g.abs().max() for g in grads)))
total, per, grads, res = ikr.loss_and_grad(f, y0, t, data, method='rk4', want_y0=True)
print('backward rk4', float(total), float(max(g.abs().max() for g in grads)))
total, per, grads, res = ikr.loss_and_grad(f, y0[:40], t, data, options={'tensor_cores': False})
print('backward FFMA2', float(total), float(max(g.abs().max() for g in grads)))
x = torch.rand(300, 2).cuda()
yy = torch.rand(300).cuda() * 1e-3
loss, g = ikr.mse_loss_and_grad(f, x, yy)
print('regression', float(loss))
ym = torch.tensor([[0., 1., 0., 0., 0., 0.]] * 5).cuda()
with torch.no_grad():
    rm = ikr.integrate_markov(ikr.MARKOV_B06, ym, t, (t_tab, v_tab), want_current=True, noise_sigma=0.1)
print('markov', float(rm.current.abs().max()))
torch.cuda.synchronize()
print('done')
