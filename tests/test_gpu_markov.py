"""6-state Markov ground-truth generator (train-d1.py:134-187, 539-569) against the
CPU oracle (restated torchdiffeq + the restated ``Lambda`` class, itself pinned to vectors of the
reference class in tests/golden/rhs_vectors.npz)."""
import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_models as rm, ref_odeint as ro

pytestmark = pytest.mark.gpu

Y0 = [[0., 1., 0., 0., 0., 0.], [0., 0., 1., 0., 0., 0.]]     # gt_true_y0s, train-d1.py:117-118


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _oracle(t_tab, v_tab, y0, t, **kw):
    f = rm.MarkovRhs()
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    return ro.odeint(f, y0, t, **kw)


def test_markov_rk4_fp64_matches_oracle_1e10():
    t_tab, v_tab = protocols.ap2hz()
    t = torch.linspace(0., 200., 401, dtype=torch.float64)
    y0 = torch.tensor(Y0, dtype=torch.float64)
    res = ikr.integrate_markov(ikr.MARKOV_B06, y0.cuda(), t, (t_tab, v_tab), method='rk4')
    for b in range(2):
        want = _oracle(t_tab, v_tab, y0[b:b + 1], t, method='rk4')[:, 0].numpy()
        got = res.y[:, b].cpu().numpy()
        assert np.abs(got - want).max() <= 1e-10 * max(np.abs(want).max(), 1.0)
    assert np.abs(res.y.sum(-1).cpu().numpy() - 1.0).max() < 1e-12      # probabilities conserved


def test_markov_dopri5_as_shipped_and_odeint_dispatch():
    """fp32 state like ``gt_true_y0s`` (train-d1.py:117); also through ``odeint`` with a
    reference-style object (p1..p12, protocol, no net), inside and beyond the protocol table."""
    t_tab, v_tab = protocols.ap2hz()
    t = torch.linspace(0., 600., 301)
    y0 = torch.tensor(Y0[:1])

    class Lambda:                                   # attribute layout of train-d1.py:134-160
        pass

    gt = Lambda()
    for k, v in zip(['p%d' % i for i in range(1, 13)], ikr.MARKOV_B06):
        setattr(gt, k, v)
    gt._t_regular, gt._v_regular = t_tab, v_tab
    y = ikr.odeint(gt, y0.cuda(), t, method='dopri5')
    assert y.shape == (301, 1, 6) and y.dtype == torch.float32
    want = _oracle(t_tab, v_tab, y0, t)
    assert (y.cpu() - want).abs().max().item() < 2e-5             # fp32 noise of the adaptive solver
    # beyond the table end (3.5 s): V = -80 fallback, fp32 rates (train-d1.py:163-166)
    t2 = torch.linspace(3400., 3700., 61)
    y2 = ikr.odeint(gt, y0.cuda(), t2)
    want2 = _oracle(t_tab, v_tab, y0, t2)
    assert (y2.cpu() - want2).abs().max().item() < 2e-5
    # fp64 state on the AP protocol: the accept/reject sequence is sensitive to 1-ulp differences
    # (DESIGN.md section 6), so two correct solvers agree at the solver's global-error level
    y3 = ikr.integrate_markov(ikr.MARKOV_B06, y0.double().cuda(), t.double(), (t_tab, v_tab)).y
    want3 = _oracle(t_tab, v_tab, y0.double(), t.double())
    assert (y3.cpu() - want3).abs().max().item() < 2e-5
    # fp64 state on a step protocol (piecewise-constant V): agreement at 10 x atol
    s_tab, sv_tab = protocols.pr3_activation(20)
    ts = torch.linspace(0., 400., 201, dtype=torch.float64)
    y4 = ikr.integrate_markov(ikr.MARKOV_B06, y0.double().cuda(), ts, (s_tab, sv_tab)).y
    want4 = _oracle(s_tab, sv_tab, y0.double(), ts)
    assert (y4.cpu() - want4).abs().max().item() < 1e-8


def test_markov_batch_parameters_and_noise():
    t_tab, v_tab = protocols.pr3_activation(20)
    t = torch.linspace(0., 400., 201)
    rng = np.random.RandomState(4)
    B = 300
    P = np.array(ikr.MARKOV_B06)[None, :] * rng.uniform(0.8, 1.25, (B, 12))
    P[0] = ikr.MARKOV_B06
    y0 = torch.tensor(Y0[:1]).repeat(B, 1).cuda()
    res = ikr.integrate_markov(P, y0, t, (t_tab, v_tab), want_current=True)
    one = ikr.integrate_markov(ikr.MARKOV_B06, y0[:1], t, (t_tab, v_tab), want_current=True)
    assert torch.equal(res.y[:, :1], one.y) and torch.equal(res.current[:, :1], one.current)
    assert int((res.stats[:, 3] != 0).sum()) == 0
    f = rm.MarkovRhs(params=tuple(P[7]))
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    want = ro.odeint(f, y0[:1].cpu(), t)
    assert (res.y[:, 7:8].cpu() - want).abs().max().item() < 2e-5
    # observation (train-d1.py:545): o (V + 86), fp64
    v = np.interp(t.double().numpy(), t_tab, v_tab)
    cur = res.y[:, 7, 5].double().cpu().numpy() * (v + 86.0)
    assert np.abs(res.current[:, 7].cpu().numpy() - cur).max() < 1e-12
    # noise: N(0, sigma^2) per sample, independent streams per trajectory, reproducible by seed
    same = y0[:1].repeat(4096, 1)
    a = ikr.integrate_markov(ikr.MARKOV_B06, same, t, (t_tab, v_tab), want_y=False, want_current=True,
                             noise_sigma=0.1, seed=11)
    b = ikr.integrate_markov(ikr.MARKOV_B06, same, t, (t_tab, v_tab), want_y=False, want_current=True,
                             noise_sigma=0.1, seed=11)
    c = ikr.integrate_markov(ikr.MARKOV_B06, same, t, (t_tab, v_tab), want_y=False, want_current=True,
                             noise_sigma=0.1, seed=12)
    assert torch.equal(a.current, b.current) and not torch.equal(a.current, c.current)
    noise = (a.current - one.current).cpu().numpy()               # (T, 4096)
    assert abs(noise.mean()) < 1e-3 and abs(noise.std() - 0.1) < 1e-3
    assert abs(np.corrcoef(noise[:, 0], noise[:, 1])[0, 1]) < 0.25
    assert abs(np.corrcoef(noise[3], noise[4])[0, 1]) < 0.06


def test_markov_ground_truth_generator():
    """``MarkovGroundTruth.simulate_data``: the protocol loop of train-d1.py:539-557 in batched
    form; the noise-free current equals the oracle's ground-truth trace used by the KAT rows."""
    from tests import kat
    gt = ikr.MarkovGroundTruth()
    row = kat.KAT['d1'][9]
    t_tab, v_tab, t_out = kat.row_protocol(row)
    i_gt = kat.gt_current('d1', t_tab, v_tab, t_out).reshape(-1).numpy()
    o, cur = gt.simulate_data([np.stack([t_tab, v_tab], 1)], t_out.numpy(), n_realisations=3,
                              noise_sigma=0.0)
    assert cur.shape == (len(t_out), 1, 3)
    assert np.abs(cur[:, 0, 0].cpu().numpy() - i_gt).max() < 5e-4
    assert torch.equal(cur[:, 0, 0], cur[:, 0, 2])
