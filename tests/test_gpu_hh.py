"""GPU parity of the network-free HH candidate path (PINTS forward model,
train-d0.py:321-439) against the CPU oracle."""
import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_models as rm
from oracle import ref_odeint as ro

pytestmark = pytest.mark.gpu

X = np.array([[1.13e-4, 7.45e-2, 3.60e-5, 4.49e-2],
              [2.26e-4, 6.99e-2, 3.45e-5, 5.46e-2],
              [5.0e-5, 9.0e-2, 1.0e-4, 3.0e-2]])


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _oracle(x, y0, t, tab, **kw):
    f = rm.HHFitRhs()
    f.set_parameters(list(x))
    f.set_fixed_form_voltage_protocol(*tab)
    return ro.odeint(f, y0, t, **kw)


def test_hh_rk4_fp64_population_matches_oracle_1e10():
    torch.set_num_threads(1)
    tab = protocols.pr3_activation(20)
    t = torch.linspace(0., 400., 401, dtype=torch.float64)
    y0 = torch.tensor([[0., 1.]], dtype=torch.float64)
    res = ikr.integrate_hh(X, y0.repeat(3, 1).cuda(), t, tab, method='rk4',
                           options={'inactivation': rm.INACT_D})
    got = res.y.cpu().numpy()
    for k, x in enumerate(X):
        want = _oracle(x, y0, t, tab, method='rk4').numpy()[:, 0, :]
        rel = np.abs(got[:, k, :] - want) / np.maximum(np.abs(want), 1e-300)
        assert rel.max() <= 1e-10


def test_hh_dopri5_fp32_as_reference_simulate():
    """`Model.simulate` semantics (train-d0.py:415-439): fp32 y0 = [0, 1], dopri5 defaults,
    current = a r (V + 86): population call == per-candidate oracle within the fp32 envelope, and
    the one-candidate `simulate` equals column 0 of `simulate_population` bit for bit."""
    torch.set_num_threads(1)
    t_tab, v_tab = protocols.ap2hz()
    t = np.linspace(0., 600., 301)
    model = ikr.HHPopulationModel(inactivation=rm.INACT_D)
    model.set_fixed_form_voltage_protocol(t_tab, v_tab)
    cur = model.simulate_population(X, t)
    assert cur.shape == (301, 3)
    v = np.interp(t, t_tab, v_tab)
    for k, x in enumerate(X):
        y = _oracle(x, torch.tensor([[0., 1.]]), torch.from_numpy(t), (t_tab, v_tab)).numpy()
        want = y[:, 0, 0] * y[:, 0, 1] * (v + 86)
        assert np.abs(cur[:, k] - want).max() < 2e-4
    one = model.simulate(X[1], t)
    assert np.array_equal(one, cur[:, 1])
    # several protocols at once (set_voltage_protocol_batches)
    p2 = np.stack(protocols.pr3_activation(20), 1)
    model.set_voltage_protocol_batches([np.stack([t_tab, v_tab], 1), p2])
    both = model.simulate_population(X, t)
    assert both.shape == (301, 2, 3) and np.array_equal(both[:, 0, :], cur)


def test_odeint_accepts_reference_style_hh_func():
    class ODEFunc(torch.nn.Module):              # attribute layout of train-d0.py:321-345
        def __init__(self):
            super().__init__()
            self.p1, self.p2, self.p3, self.p4 = X[0]
            self.p5, self.p6, self.p7, self.p8 = rm.INACT_D

        def set_fixed_form_voltage_protocol(self, t, v):
            self._t_regular, self._v_regular = t, v

    torch.set_num_threads(1)
    f = ODEFunc()
    tab = protocols.pr5_deactivation(-60)
    f.set_fixed_form_voltage_protocol(*tab)
    t = torch.linspace(0., 300., 151, dtype=torch.float64)
    y0 = torch.tensor([[0.1, 0.8]], dtype=torch.float64)
    got = ikr.odeint(f, y0, t, method='rk4')
    want = _oracle(X[0], y0, t, tab, method='rk4')
    assert got.shape == want.shape and got.device == y0.device
    assert ((got - want).abs() / want.abs().clamp_min(1e-300)).max().item() <= 1e-10


def test_hh_population_members_are_independent():
    tab = protocols.ap2hz()
    t = torch.linspace(0., 300., 61)
    rng = np.random.RandomState(3)
    P = np.abs(np.concatenate([X, X * rng.uniform(0.5, 2.0, X.shape)]))
    P = np.concatenate([P] * 50)                                    # 300 candidates
    y0 = torch.tensor([[0., 1.]]).repeat(len(P), 1).cuda()
    a = ikr.integrate_hh(P, y0, t, tab, options={'inactivation': rm.INACT_D})
    perm = rng.permutation(len(P))
    b = ikr.integrate_hh(P[perm], y0, t, tab, options={'inactivation': rm.INACT_D})
    assert torch.equal(a.y[:, perm, :], b.y) and torch.equal(a.stats[perm], b.stats)
    assert torch.equal(a.y[:, :6], a.y[:, 6:12])
