"""Diagnostic: one pr3 lane that ends with 'underflow in dt' under the fp16x2 split."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr
import bench
wl = bench.workload(['pr3'])
fam, name, t_tab, v_tab, t_out = wl[0]
func = ikr.load_weights(ikr.ODEFunc(params='d'), bench.WEIGHTS)
func.set_fixed_form_voltage_protocol(t_tab, v_tab)
y0 = torch.tensor([[0.00575035, 0.950099]], dtype=torch.float32).cuda()
t = torch.tensor(t_out, dtype=torch.float32)
res = {}
for sp in ('fp16x2', 'bf16x3', 'ffma'):
    opts = {'check_status': False, 'ckpt_cap': 4096}
    if sp == 'ffma':
        opts['tensor_cores'] = False
    else:
        opts['tc_split'] = sp
    with torch.no_grad():
        r = ikr.integrate(func, y0, t, want_ckpt=True, options=opts)
    st = r.stats.cpu().numpy()[0]
    y = r.y.cpu().numpy()[:, 0]
    ck_t = r.ckpt[0].cpu().numpy()[:, 0]
    ck_y = r.ckpt[1].cpu().numpy()[:, 0]
    n = st[0]
    print(sp, 'stats', st, 'last accepted steps (t0, dt):', ck_t[max(0, n - 4):n].tolist())
    print('   state at last steps', ck_y[max(0, n - 3):n, :2].tolist(), 'k_a', ck_y[max(0, n - 2):n, 2:9].tolist())
    res[sp] = (ck_t[:n], ck_y[:n])
a, b = res['fp16x2'], res['bf16x3']
m = min(len(a[0]), len(b[0]))
d = np.abs(a[0][:m, 0] - b[0][:m, 0])
first = int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else -1
print('step sequences agree up to step', first, 'of', m)
print('fp16x2 dt sequence tail', a[0][-12:, 1].tolist())
