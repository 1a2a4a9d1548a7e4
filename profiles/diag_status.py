"""Diagnostic: status histogram of the bench mix per protocol family for a given operand split."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr
import bench
split = sys.argv[1] if len(sys.argv) > 1 else 'fp16x2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
wl = bench.workload(['pr3', 'pr4', 'pr5', 'sinewave', 'aps'])
func = ikr.load_weights(ikr.ODEFunc(params='d'), bench.WEIGHTS)
rng = np.random.RandomState(1000)
sizes = [B // 5 + (1 if i < B % 5 else 0) for i in range(5)]
jobs = []
for (fam, name, t_tab, v_tab, t_out), nb in zip(wl, sizes):
    y0 = np.stack([rng.uniform(0, 0.05, nb), rng.uniform(0.95, 1.0, nb)], 1).astype(np.float32)
    g = rng.lognormal(0.0, 0.2, nb).astype(np.float32)
    jobs.append(dict(protocol=(t_tab, v_tab), t=torch.tensor(t_out, dtype=torch.float32), E=-86.0,
                     data=torch.zeros(len(t_out)), want_y=False, y0=torch.from_numpy(y0).cuda(),
                     g=torch.from_numpy(g).cuda()))
with torch.no_grad():
    outs = ikr.integrate_many(func, jobs, options={'check_status': False, 'tc_split': split,
                                                   'tc_timing': bool(os.environ.get('TC_TIMING'))})
torch.cuda.synchronize()
for (fam, *_), r, job in zip(wl, outs, jobs):
    st = r.stats.cpu().numpy()
    codes, counts = np.unique(st[:, 3], return_counts=True)
    print(fam, dict(zip(codes.tolist(), counts.tolist())), 'attempts mean %.1f' % (st[:, 0] + st[:, 1]).mean(), flush=True)
    bad = np.nonzero(st[:, 3] != 0)[0]
    if len(bad):
        b = int(bad[0])
        print('  first bad lane', b, 'stats', st[b], 'y0', job['y0'][b].cpu().numpy(), 'g', float(job['g'][b]))
        for sp in ('fp16x2', 'bf16x3'):
            one = ikr.integrate_many(func, [dict(job, y0=job['y0'][b:b + 1], g=job['g'][b:b + 1])],
                                     options={'check_status': False, 'tc_split': sp})[0]
            print('   alone with', sp, one.stats.cpu().numpy()[0])
print(outs[0].geometry)
