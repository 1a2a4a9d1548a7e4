"""The per-lane solver state machine (csrc/ikr_math.h -- the code the kernels run between MLP
evaluations) built for the host and checked against the CPU oracle.  TEST INFRASTRUCTURE: the
harness is never part of the product path."""
import numpy as np
import pytest
import torch

from neural_ode_ion_channels_b200 import protocols
from oracle import ref_odeint as ro
from tests import harness, kat


def _p8(nn):
    return [getattr(nn, 'p%d' % i, 0.0) for i in range(1, 9)]


def _double(nn):
    nn = nn.double()
    nn.vrange = nn.vrange.double()
    nn.netscale = nn.netscale.double()
    return nn


def test_table_voltage_matches_scipy_bitwise():
    from scipy.interpolate import interp1d
    t_tab, v_tab = protocols.ap2hz()
    f = interp1d(t_tab, v_tab)
    rng = np.random.RandomState(0)
    for x in np.concatenate([rng.uniform(0, 3499.9, 500), t_tab[:20], t_tab[-20:]]):
        v, ok = harness.table_voltage(t_tab, v_tab, x)
        assert ok and v == float(f([x])[0])
    assert harness.table_voltage(t_tab, v_tab, 3500.0) == (-80.0, False)
    assert harness.table_voltage(t_tab, v_tab, -1e-9) == (-80.0, False)
    tc, vc = protocols.compact_table(*protocols.pr5_deactivation(-40, per_ms=10))
    t_full, v_full = protocols.pr5_deactivation(-40, per_ms=10)
    g = interp1d(t_full, v_full)
    for x in rng.uniform(0, 10000, 300):
        assert harness.table_voltage(tc, vc, x)[0] == float(g([x])[0])


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_rk4_fp64_host_logic(study):
    torch.set_num_threads(1)
    nn = _double(kat.make_nn(study, mlp_follows_state=True))
    t_tab, v_tab = protocols.ap2hz()
    nn.set_fixed_form_voltage_protocol(t_tab, v_tab)
    y0 = torch.tensor([[0., 1.]], dtype=torch.float64)
    t = torch.linspace(0., 100., 201, dtype=torch.float64)
    with torch.no_grad():
        want = ro.odeint(nn, y0, t, method='rk4').numpy()[:, 0, :]
    got, stats, _ = harness.integrate(nn.net, 5, 200, study == 'd2', _p8(nn), t_tab, v_tab,
                                      [0., 1.], t.numpy(), True, True, method='rk4')
    assert (np.abs(got - want) / np.abs(want).clip(1e-30)).max() <= 1e-10
    assert tuple(stats) == (200, 0, 800, 0)


def test_dopri5_fp64_host_logic_smooth():
    torch.set_num_threads(1)
    nn = _double(kat.make_nn('d2', mlp_follows_state=True))
    tt = np.linspace(0, 500, 5001)
    vv = np.full_like(tt, 20.0)
    nn.set_fixed_form_voltage_protocol(tt, vv)
    y0 = torch.tensor([[0.1, 0.9]], dtype=torch.float64)
    t = torch.linspace(0., 500., 251, dtype=torch.float64)
    st = {}
    with torch.no_grad():
        want = ro.odeint(nn, y0, t, stats=st).numpy()[:, 0, :]
    got, stats, _ = harness.integrate(nn.net, 5, 200, True, _p8(nn), tt, vv, [0.1, 0.9],
                                      t.numpy(), True, True)
    assert (stats[0], stats[1], stats[2]) == (st['n_accept'], st['n_reject'], st['nfe'])
    assert np.abs(got - want).max() <= 1e-8


def test_dopri5_fp32_host_logic_ap2hz():
    torch.set_num_threads(1)
    nn = kat.make_nn('s1')
    t_tab, v_tab = protocols.ap2hz()
    nn.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 1000., 501)
    st = {}
    with torch.no_grad():
        want = ro.odeint(nn, torch.tensor([[0., 1.]]), t, stats=st).numpy()[:, 0, :]
    got, stats, _ = harness.integrate(nn.net, 5, 200, False, _p8(nn), t_tab, v_tab, [0., 1.],
                                      t.double().numpy(), False, False)
    assert stats[3] == 0 and abs(int(stats[0]) - st['n_accept']) < 0.05 * st['n_accept']
    assert np.abs(got - want).max() < 1e-4          # fp32-state noise envelope
