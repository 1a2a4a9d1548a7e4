"""GPU parity tests of the forward path (through the C ABI) against the CPU oracle.

Tolerances (stated per north_star):
* rk4, fp64 state + fp64 MLP: <= 1e-10 relative to the oracle.
* dopri5, fp64/fp64: every accepted step reproduced from its checkpoint to <= 1e-10 relative
  (step-wise parity); whole traces within the solver's own step-sequence sensitivity envelope
  (the adaptive controller on this RHS is chaotic at the 1e-5 level -- see DESIGN.md), and within
  10 x atol on a smooth protocol where the accept/reject sequence is reproduced.
* dopri5, fp32 state "as shipped": reference-logged losses within 5e-5 absolute.
"""
import os
import warnings

import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_models as rm
from oracle import ref_odeint as ro
from tests import kat

pytestmark = pytest.mark.gpu

CUDA = torch.cuda.is_available()


@pytest.fixture(autouse=True)
def _no_grad():
    # the reference runs every odeint call under torch.no_grad() (train-s1.py:311)
    with torch.no_grad():
        yield


def _nn(study, double=False):
    """(product module, oracle module) with identical weights."""
    cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
    pset = 's' if study in ('s1', 's2') else 'd'
    func = ikr.load_weights(cls(params=pset), kat.weights_path(study))
    ofunc = kat.make_nn(study, mlp_follows_state=double)
    if double:
        func = func.double()
        ofunc = ofunc.double()
        ofunc.vrange = ofunc.vrange.double()
        ofunc.netscale = ofunc.netscale.double()
    return func, ofunc


def _oracle_batch(ofunc, y0, t, **kw):
    outs = []
    with torch.no_grad():
        for b in range(y0.shape[0]):
            outs.append(ro.odeint(ofunc, y0[b:b + 1], t, **kw))
    return torch.cat(outs, dim=1)


def _rel(got, want, floor=1e-300):
    return (np.abs(got - want) / np.maximum(np.abs(want), floor)).max()


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_rk4_fp64_matches_oracle_1e10(study):
    torch.set_num_threads(1)
    func, ofunc = _nn(study, double=True)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 120., 241, dtype=torch.float64)
    y0 = torch.tensor([[0., 1.], [0.3, 0.6], [0.9, 0.05]], dtype=torch.float64)
    want = _oracle_batch(ofunc, y0, t, method='rk4').numpy()
    got = ikr.odeint(func, y0.cuda(), t, method='rk4').cpu().numpy()
    assert got.shape == want.shape == (241, 3, 2)
    assert _rel(got, want) <= 1e-10
    # off-grid outputs (options step_size): linear interpolation between grid points
    t2 = torch.linspace(0., 120., 13, dtype=torch.float64)
    want = _oracle_batch(ofunc, y0[:1], t2, method='rk4', options={'step_size': 0.7}).numpy()
    got = ikr.odeint(func, y0[:1].cuda(), t2, method='rk4', options={'step_size': 0.7})
    assert _rel(got.cpu().numpy(), want) <= 1e-10


def test_rk4_fp32_as_shipped():
    torch.set_num_threads(1)
    func, ofunc = _nn('s1')
    t_tab, v_tab = protocols.pr3_activation(40)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(900., 1100., 401)
    y0 = torch.tensor([[0., 1.]])
    want = _oracle_batch(ofunc, y0, t, method='rk4').numpy()
    got = ikr.odeint(func, y0.cuda(), t, method='rk4').cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)   # fp32 state: ~400 steps of eps


@pytest.mark.parametrize('arch', ['s03', 's10', 's06', 's01'])
def test_rk4_fp64_architectures(arch):
    """architectures/sNN.py sweep: tiny (n=10), n=100, wide (n=500), shallow nets."""
    torch.set_num_threads(1)
    torch.manual_seed(0)
    func = ikr.ODEFuncNNf(arch=arch, params='r').double()
    L, n = ikr.ARCHITECTURES[arch]
    ofunc = rm.NNfRhs(net=rm.build_mlp(L, n), inact=tuple(func.__dict__['p%d' % i] for i in (5, 6, 7, 8)),
                      mlp_follows_state=True).double()
    ofunc.net.load_state_dict(func.net.state_dict())
    ofunc.vrange = ofunc.vrange.double()
    ofunc.netscale = ofunc.netscale.double()
    t_tab, v_tab = protocols.pr5_deactivation(-60)
    for f in (func, ofunc):
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(990., 1010., 41, dtype=torch.float64)
    y0 = torch.tensor([[0.02, 0.97], [0.5, 0.5]], dtype=torch.float64)
    want = _oracle_batch(ofunc, y0, t, method='rk4').numpy()
    got = ikr.odeint(func, y0.cuda(), t, method='rk4').cpu().numpy()
    assert _rel(got, want) <= 1e-10


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_dopri5_fp64_stepwise_parity(study):
    """Each accepted step of the CUDA run, restarted on the oracle from the step checkpoint
    (t0, dt, y0, k_0..k_6), must land on the next checkpoint: bit-level parity of the RK arithmetic
    and of the dense output, independent of the chaotic step-size controller."""
    torch.set_num_threads(1)
    func, ofunc = _nn(study, double=True)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 400., 201, dtype=torch.float64)
    y0 = torch.tensor([[0., 1.], [0.2, 0.7]], dtype=torch.float64)
    res = ikr.integrate(func, y0.cuda(), t, want_ckpt=True)
    stats = res.stats.cpu().numpy()
    ck_t, ck_y = res.ckpt[0].cpu(), res.ckpt[1].cpu()
    y_gpu = res.y.cpu()
    solver = ro.Dopri5(ofunc, y0[:1], 1e-7, 1e-9)
    worst = 0.0
    for b in range(2):
        n_acc = int(stats[b, 0])
        assert stats[b, 3] == 0 and n_acc > 20
        j_out = 1
        for j in range(0, n_acc - 1, max(1, n_acc // 40)):   # ~40 steps sampled per trajectory
            ts, dt = ck_t[j, b, 0], ck_t[j, b, 1]
            yy = ck_y[j, b, :2].reshape(1, 2)
            ff = ck_y[j, b, [2, 9]].reshape(1, 2)      # k_0 = f0 of (a, r)
            with torch.no_grad():
                y1, f1, err, k = solver._rk_step(yy, ff, ts, dt, ts + dt)
                coef = solver._mid_fit(yy, y1, k, dt)
            want = torch.cat([y1.reshape(-1), f1.reshape(-1)]).numpy()
            got = ck_y[j + 1, b, [0, 1, 2, 9]].numpy()
            worst = max(worst, _rel(got, want, 1e-30))
            # dense output samples inside this step
            inside = torch.nonzero((t > ts) & (t <= ts + dt)).reshape(-1)
            for i in inside.tolist():
                with torch.no_grad():
                    yi = ro._dense_eval(coef, ts, ts + dt, t[i]).reshape(-1).numpy()
                worst = max(worst, _rel(y_gpu[i, b].numpy(), yi, 1e-30))
    assert worst <= 1e-10, worst


def test_dopri5_fp64_smooth_protocol_within_10_atol():
    """Constant-voltage protocol: the accept/reject sequence is reproduced and the traces agree
    within 10 x atol (1e-8), the loss to 1e-8 relative."""
    torch.set_num_threads(1)
    func, ofunc = _nn('d2', double=True)
    tt = np.linspace(0, 500, 5001)
    vv = np.full_like(tt, 20.0)
    func.set_fixed_form_voltage_protocol(tt, vv)
    ofunc.set_fixed_form_voltage_protocol(tt, vv)
    t = torch.linspace(0., 500., 251, dtype=torch.float64)
    y0 = torch.tensor([[0.1, 0.9]], dtype=torch.float64)
    st = {}
    with torch.no_grad():
        want = ro.odeint(ofunc, y0, t, stats=st)
    data = torch.zeros(251, dtype=torch.float64)
    res = ikr.integrate(func, y0.cuda(), t, data=data, want_current=True, E=-86.0)
    got = res.y.cpu()
    stats = res.stats.cpu().numpy()[0]
    assert (stats[0], stats[1]) == (st['n_accept'], st['n_reject'])
    assert (got - want).abs().max().item() <= 1e-8
    i_want = want[:, 0, 0] * want[:, 0, 1] * (20.0 + 86.0)
    loss_want = i_want.abs().mean().item()
    loss_got = res.sae.cpu()[0].item() / 251
    assert abs(loss_got - loss_want) <= 1e-8 * abs(loss_want)


def test_dopri5_fp64_envelope_on_ap_protocol():
    """Whole-trace agreement on the AP protocol is limited by the controller's sensitivity: the
    oracle differs from *itself* by ~2e-5 when its first step is perturbed by 1e-6.  The CUDA
    trace must sit inside that envelope around a tight-tolerance solution."""
    torch.set_num_threads(1)
    func, ofunc = _nn('s1', double=True)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 600., 301, dtype=torch.float64)
    y0 = torch.tensor([[0., 1.]], dtype=torch.float64)
    with torch.no_grad():
        base = ro.odeint(ofunc, y0, t)
        tight = ro.odeint(ofunc, y0, t, rtol=1e-10, atol=1e-12)
    oracle_err = (base - tight).abs().max().item()
    got = ikr.odeint(func, y0.cuda(), t).cpu()
    assert (got - tight).abs().max().item() <= 3 * oracle_err + 1e-9
    # and a tight-tolerance CUDA run converges to the tight oracle
    got_tight = ikr.odeint(func, y0.cuda(), t, rtol=1e-10, atol=1e-12).cpu()
    assert (got_tight - tight).abs().max().item() <= 5e-8


@pytest.mark.parametrize('study', kat.STUDIES)
def test_logged_ap2hz_loss_fp32_as_shipped(study):
    """Reference-logged AP-2Hz loss ({study}/log2:4) with the CUDA path in the NN leg."""
    torch.set_num_threads(1)
    row = kat.KAT[study][0]
    t_tab, v_tab, t_out = kat.row_protocol(row)
    i_gt = kat.gt_current(study, t_tab, v_tab, t_out).reshape(-1)
    func, _ = _nn(study)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    y = ikr.odeint(func, torch.tensor([[0., 1.]]).cuda(), t_out).cpu()
    i_nn = (y[:, 0, 0] * y[:, 0, 1] * (func._v(t_out).reshape(-1) + 86))
    loss = torch.mean(torch.abs(i_nn - i_gt)).item()
    assert abs(loss - row['loss']) < 5e-5, (loss, row['loss'])


def test_logged_step_protocol_losses_fp32_fused_loss():
    """pr3 +40 mV and pr5 -120 mV rows of s1/log2 through the fused current/loss epilogue, with
    both sweeps' trajectories in one batch each."""
    torch.set_num_threads(1)
    func, _ = _nn('s1')
    for idx in (9, 11):
        row = kat.KAT['s1'][idx]
        t_tab, v_tab, t_out = kat.row_protocol(row)
        i_gt = kat.gt_current('s1', t_tab, v_tab, t_out).reshape(-1)
        func.set_fixed_form_voltage_protocol(t_tab, v_tab)
        y0 = torch.tensor([[0., 1.]] * 5)
        res = ikr.integrate(func, y0.cuda(), t_out, data=i_gt.float(), E=-86.0, want_y=False)
        mae = (res.sae / len(t_out)).cpu().numpy()
        assert np.all(np.abs(mae - row['loss']) < 5e-5), (mae, row['loss'])
        assert np.all(mae == mae[0])          # identical lanes give identical results


def test_batch_is_independent_trajectories_and_tile_invariant():
    func, _ = _nn('d1')
    t_tab, v_tab = protocols.pr4_inactivation_standin(-20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 300., 151)
    rng = np.random.RandomState(3)
    B = 333                                     # ragged: not a multiple of any tile
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    # FFMA2 kernel: bit-identical whatever the tile size and the batch the trajectory sits in
    full = ikr.integrate(func, y0, t, options={'tensor_cores': False})
    for tile in (8, 64):
        part = ikr.integrate(func, y0[:77], t, options={'tile_m': tile})
        assert torch.equal(part.y, full.y[:, :77])
        assert torch.equal(part.stats, full.stats[:77])
    assert full.geometry['n_tiles'] * full.geometry['tile_m'] >= B
    # tcgen05 kernel (default): bit-identical whatever the lane / tile the trajectory sits in
    tc_full = ikr.integrate(func, y0, t)
    assert tc_full.geometry['tensor_cores']
    tc_part = ikr.integrate(func, y0[200:277], t)
    assert torch.equal(tc_part.y, tc_full.y[:, 200:277])
    assert torch.equal(tc_part.stats, tc_full.stats[200:277])


def test_fused_current_and_loss_epilogue():
    func, _ = _nn('d1')
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 150., 76)
    B = 40
    rng = np.random.RandomState(5)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    g = torch.tensor(rng.lognormal(0, 0.2, B), dtype=torch.float32).cuda()
    data = torch.tensor(rng.normal(0, 0.1, (76, B)), dtype=torch.float32).cuda()
    res = ikr.integrate(func, y0, t, g=g, E=-86.0, data=data, want_current=True)
    v = func._v(t).reshape(-1).cuda()
    cur = (g[None, :] * res.y[:, :, 0] * res.y[:, :, 1]).double() * (v[:, None] + 86.0)
    np.testing.assert_allclose(res.current.double().cpu().numpy(), cur.cpu().numpy(), rtol=2e-7,
                               atol=1e-9)
    diff = cur - data.double()
    np.testing.assert_allclose(res.sse.cpu().numpy(), (diff ** 2).sum(0).cpu().numpy(), rtol=1e-10)
    np.testing.assert_allclose(res.sae.cpu().numpy(), diff.abs().sum(0).cpu().numpy(), rtol=1e-10)
    # shared (T,) data trace
    res1 = ikr.integrate(func, y0, t, g=g, data=data[:, 0], want_y=False)
    diff1 = cur - data[:, :1].double()
    np.testing.assert_allclose(res1.sse.cpu().numpy(), (diff1 ** 2).sum(0).cpu().numpy(),
                               rtol=1e-10)


def test_out_of_table_time_uses_minus_80():
    """Times beyond the table end take the V = -80 branch (train-s1.py:234-237): dopri5 overshoots
    the last output time, and an rk4 grid may run past the table."""
    torch.set_num_threads(1)
    func, ofunc = _nn('s1', double=True)
    tt = np.linspace(0, 100, 1001)
    vv = np.where(tt < 50, -80.0, 20.0)
    func.set_fixed_form_voltage_protocol(tt, vv)
    ofunc.set_fixed_form_voltage_protocol(tt, vv)
    y0 = torch.tensor([[0., 1.]], dtype=torch.float64)
    # rk4 grid crossing the table end: stages beyond t = 100 see V = -80.  In that branch the
    # reference forms the HH rates in fp32 (`p6 * tensor([-80])` is a float32 tensor), so parity
    # there is bounded by fp32 exp rounding (torch CPU vs CUDA expf), not by 1e-10.
    t2 = torch.linspace(95., 105., 21, dtype=torch.float64)
    want2 = _oracle_batch(ofunc, y0, t2, method='rk4').numpy()
    got2 = ikr.odeint(func, y0.cuda(), t2, method='rk4').cpu().numpy()
    assert _rel(got2[:11], want2[:11]) <= 1e-10          # inside the table
    assert _rel(got2, want2) <= 1e-6                     # beyond it
    # dopri5 stepping over the end of the table (discontinuous protocol: envelope tolerance)
    t = torch.tensor([0., 30., 60., 100.], dtype=torch.float64)
    want = ro.odeint(ofunc, y0, t)
    res = ikr.integrate(func, y0.cuda(), t)
    assert (res.y.cpu() - want).abs().max().item() < 2e-5


def test_edge_cases_and_status_codes():
    func, _ = _nn('s1')
    t_tab, v_tab = protocols.pr2_time_constant(30)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    y0 = torch.tensor([[0., 1.], [0.1, 0.8]]).cuda()
    # single output time: solution is y0
    y = ikr.odeint(func, y0, torch.tensor([0.]))
    assert torch.equal(y[0], y0)
    # 1-D y0 like torchdiffeq accepts
    y1 = ikr.odeint(func, y0[0], torch.linspace(0., 10., 3))
    assert y1.shape == (3, 2)
    # CPU tensors in, CPU tensors out
    ycpu = ikr.odeint(func, y0.cpu(), torch.linspace(0., 10., 3))
    assert not ycpu.is_cuda
    # max_num_steps -> torchdiffeq's assertion text
    with pytest.raises(AssertionError, match='max_num_steps exceeded'):
        ikr.odeint(func, y0, torch.linspace(0., 2000., 3), options={'max_num_steps': 5})
    # non-finite state
    bad = torch.tensor([[float('nan'), 1.]]).cuda()
    with pytest.raises(AssertionError, match='underflow in dt|non-finite'):
        ikr.odeint(func, bad, torch.linspace(0., 10., 3))
    # legacy option names only warn (train-d0.py:436)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        ikr.odeint(func, y0, torch.linspace(0., 10., 3),
                   options={'grid_points': np.array([1.0]), 'eps': 1e-6})
    assert any('Unexpected arguments' in str(w.message) for w in rec)
    with pytest.raises(ValueError):
        ikr.odeint(func, y0, torch.tensor([0., 2., 1.]))
    with pytest.raises(TypeError):
        ikr.odeint(torch.nn.Linear(2, 2), y0, torch.linspace(0., 10., 3))


def test_reference_module_is_accepted_unchanged():
    """A module that is *not* ours but looks like the reference's ODEFunc (attributes only) is
    introspected and integrated; weights load from the shipped state-dict file."""
    import torch.nn as nn

    class ODEFunc(nn.Module):                  # shaped like train-s1.py:181-216
        def __init__(self):
            super().__init__()
            self.net = ikr.build_net(5, 200)
            self.vrange = torch.tensor([100.])
            self.netscale = torch.tensor([1000.])
            self.p5, self.p6, self.p7, self.p8 = rm.HH_B06[4:]

        def set_fixed_form_voltage_protocol(self, t, v):
            self._t_regular = t
            self._v_regular = v

    f = ODEFunc()
    f.load_state_dict(torch.load(kat.weights_path('s1')))
    f.eval()
    mine, _ = _nn('s1')
    t_tab, v_tab = protocols.pr3_activation(0)
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    mine.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 1500., 301)
    y0 = torch.tensor([[0., 1.]]).cuda()
    with torch.no_grad():
        assert torch.equal(ikr.odeint(f, y0, t), ikr.odeint(mine, y0, t))


def test_interp_kernel_matches_scipy():
    import ctypes
    from scipy.interpolate import interp1d
    from neural_ode_ion_channels_b200 import _cabi, solver
    t_tab, v_tab = protocols.ap2hz()
    rng = np.random.RandomState(0)
    tq = np.concatenate([rng.uniform(0, 3499.9, 5000), t_tab[::700], [3499.9, 3500.0, -1.0]])
    for compact in (False, True):
        tab = solver._DeviceTable(t_tab, v_tab, 'cuda', compact)
        io = _cabi.IkrIO()
        tab.fill(io)
        q = torch.from_numpy(tq).cuda()
        out = torch.empty_like(q)
        rc = _cabi.lib().ikr_interp_protocol(ctypes.byref(io), q.data_ptr(), q.numel(),
                                             out.data_ptr(), None)
        assert rc == 0
        torch.cuda.synchronize()
        want = np.full(len(tq), -80.0)
        inside = (tq >= t_tab[0]) & (tq <= t_tab[-1])
        want[inside] = interp1d(t_tab, v_tab)(tq[inside])
        assert np.array_equal(out.cpu().numpy(), want)      # bit-exact, compacted or not


@pytest.mark.parametrize('tensor_cores', [True, False])
def test_integrate_many_equals_separate_calls(tensor_cores):
    """One launch over several (protocol, batch) jobs == the reference's protocol loop of separate
    odeint calls (train-s1.py:316-543), bit for bit, on the tcgen05 and on the FFMA2 kernel."""
    func, _ = _nn('d1')
    single_opts = {'tensor_cores': True} if tensor_cores else {'tensor_cores': False, 'tile_m': 32}
    rng = np.random.RandomState(11)
    jobs, singles = [], []
    for fam, nb in (('pr3', 37), ('pr4', 200), ('aps', 5)):
        name, t_tab, v_tab, t_out = protocols.protocol_set(fam)[1 if fam != 'aps' else 0]
        t = torch.tensor(t_out[:121], dtype=torch.float32)
        y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, nb), rng.uniform(0.95, 1, nb)], 1),
                          dtype=torch.float32).cuda()
        g = torch.tensor(rng.lognormal(0, 0.2, nb), dtype=torch.float32).cuda()
        data = torch.tensor(rng.normal(0, 0.1, 121), dtype=torch.float32).cuda()
        jobs.append(dict(protocol=(t_tab, v_tab), y0=y0, t=t, g=g, data=data, want_current=True))
        func.set_fixed_form_voltage_protocol(t_tab, v_tab)
        singles.append(ikr.integrate(func, y0, t, g=g, data=data, want_current=True,
                                     options=single_opts))
    many = ikr.integrate_many(func, jobs, options={'tensor_cores': tensor_cores})
    assert len(many) == 3
    for a, b in zip(many, singles):
        assert torch.equal(a.y, b.y) and torch.equal(a.current, b.current)
        assert torch.equal(a.stats, b.stats) and torch.equal(a.sae, b.sae)


def test_lane_pool_kernel_equals_tile_scheduled_kernel():
    """The opt-in lane-pool kernel (slots refill from one trajectory queue, jobs mixed inside a
    CTA) must reproduce the tile-scheduled kernel bit for bit: a lane's arithmetic does not depend
    on its slot or its neighbours."""
    func, _ = _nn('d1')
    rng = np.random.RandomState(11)
    jobs = []
    for k, (tt, vv, T) in enumerate([(*protocols.ap2hz(), 101), (*protocols.pr3_activation(20), 81),
                                     (*protocols.pr5_deactivation(-60), 41)]):
        B = (37, 200, 5)[k]
        y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                          dtype=torch.float32).cuda()
        t = torch.linspace(0., 2. * (T - 1), T)
        g = torch.tensor(rng.lognormal(0, 0.2, B), dtype=torch.float32)
        jobs.append(dict(protocol=(tt, vv), y0=y0, t=t, g=g, data=torch.zeros(T),
                         want_current=True, want_ckpt=True))
    a = ikr.integrate_many(func, jobs, options={'tile_m': 32})
    b = ikr.integrate_many(func, jobs, options={'tile_m': 32, 'lane_pool': True})
    for ra, rb in zip(a, b):
        assert torch.equal(ra.stats, rb.stats)
        assert torch.equal(ra.y, rb.y) and torch.equal(ra.current, rb.current)
        assert torch.equal(ra.sse, rb.sse) and torch.equal(ra.sae, rb.sae)
        n = int(ra.stats[:, 0].max())
        assert torch.equal(ra.ckpt[0][:1], rb.ckpt[0][:1])
        for bb in range(ra.stats.shape[0]):
            k = int(ra.stats[bb, 0])
            assert torch.equal(ra.ckpt[1][:k, bb], rb.ckpt[1][:k, bb])
        assert n > 0


@pytest.mark.parametrize('B,tensor_cores', [(65536, True), (65536, False), (1048576, True)])
def test_full_size_ensembles_duplicate_invariance(B, tensor_cores):
    """BASELINE configs[1] / configs[4] sizes (65,536 and 1M trajectories): 256 distinct (y0, g)
    instances, each repeated B/256 times in shuffled positions.  Size-independent properties:
    every copy of an instance gives bit-identical results wherever it sits (tile / CTA / slot
    invariance), all statuses are ok, and the distinct instances match a separate B=256 launch."""
    func, _ = _nn('d1')
    t_tab, v_tab = protocols.pr3_activation(20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    rng = np.random.RandomState(12)
    base_y0 = np.stack([rng.uniform(0, 0.05, 256), rng.uniform(0.95, 1, 256)], 1).astype(np.float32)
    base_g = rng.lognormal(0, 0.2, 256).astype(np.float32)
    idx = rng.permutation(np.arange(B) % 256)
    t = torch.linspace(0., 120., 13)
    data = torch.zeros(13)
    big = ikr.integrate(func, torch.from_numpy(base_y0[idx]).cuda(), t, g=torch.from_numpy(base_g[idx]),
                        data=data, want_y=False, want_current=False,
                        options={'tensor_cores': tensor_cores})
    small = ikr.integrate(func, torch.from_numpy(base_y0).cuda(), t, g=torch.from_numpy(base_g),
                          data=data, want_y=False, want_current=False,
                          options={'tensor_cores': True} if tensor_cores else
                          {'tensor_cores': False, 'tile_m': 32})
    assert int((big.stats[:, 3] != 0).sum()) == 0
    idx_d = torch.from_numpy(idx).cuda()
    assert torch.equal(big.stats, small.stats[idx_d])
    assert torch.equal(big.sse, small.sse[idx_d]) and torch.equal(big.sae, small.sae[idx_d])


def test_architecture_sweep_s00_to_s11_fp32_and_fp64():
    """BASELINE configs[3]: every architectures/sNN.py net integrates (dopri5) in fp32 and in fp64
    with finite results and ok status; fp64 results agree with fp32 at the fp32 noise level."""
    t_tab, v_tab = protocols.pr5_deactivation(-60)
    y0 = torch.tensor([[0.02, 0.97], [0.4, 0.6], [0.0, 1.0]])
    t = torch.linspace(0., 60., 16)
    for arch in sorted(ikr.ARCHITECTURES):
        torch.manual_seed(0)
        f32 = ikr.ODEFuncNNf(arch=arch, params='r', std=0.05)
        f32.set_fixed_form_voltage_protocol(t_tab, v_tab)
        r32 = ikr.integrate(f32, y0.cuda(), t)
        f64 = ikr.ODEFuncNNf(arch=arch, params='r', std=0.05).double()
        f64.net.load_state_dict({k: v.double() for k, v in f32.net.state_dict().items()})
        f64.set_fixed_form_voltage_protocol(t_tab, v_tab)
        r64 = ikr.integrate(f64, y0.double().cuda(), t.double())
        for r in (r32, r64):
            assert int((r.stats[:, 3] != 0).sum()) == 0, arch
            assert bool(torch.isfinite(r.y).all()), arch
        assert (r32.y.double() - r64.y).abs().max().item() < 5e-4, arch
