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


# ---------------------------------------------------------------------------------------------
# discrete adjoint (the lane functions the backward kernel runs) vs autograd through the oracle
# ---------------------------------------------------------------------------------------------
def _oracle_grad(nn, y0, t, grad_y, first_step, replay=None):
    """d/dtheta, d/dy0 of sum(grad_y * odeint(...)) by PyTorch autograd through the restated
    torchdiffeq dopri5 (step sizes are constants of the backward pass: `first_step` is given so
    that the initial-step heuristic -- the only place torchdiffeq lets dt carry a graph -- is
    not in play)."""
    for p in nn.net.parameters():
        p.requires_grad_(True)
        p.grad = None
    y0 = y0.clone().requires_grad_(True)
    st = {}
    opts = {'first_step': first_step}
    if replay is not None:
        opts['replay'] = replay
    y = ro.odeint(nn, y0, t, options=opts, stats=st)
    (y[:, 0, :] * grad_y).sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in nn.net.parameters()]).double().numpy()
    for p in nn.net.parameters():
        p.requires_grad_(False)
    return g, y0.grad.reshape(-1).double().numpy(), y.detach().numpy()[:, 0, :], st


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_adjoint_fp64_host_logic_matches_oracle_autograd(study):
    """31 outputs over the first AP upstroke (37 accepted / 17 rejected steps).  With the
    oracle's own controller the two step-size sequences drift apart by ~1e-7 relative after the
    stiff phase (ill-conditioned error estimate), so that comparison is loose; with the accepted
    steps replayed into the oracle the gradients agree to rounding."""
    torch.set_num_threads(1)
    nn = _double(kat.make_nn(study, mlp_follows_state=True))
    t_tab, v_tab = protocols.ap2hz()
    nn.set_fixed_form_voltage_protocol(t_tab, v_tab)
    y0 = torch.tensor([[0.02, 0.97]], dtype=torch.float64)
    t = torch.linspace(0., 60., 31, dtype=torch.float64)
    rng = np.random.RandomState(1)
    gy = rng.randn(len(t), 2)
    args = (nn.net, 5, 200, study == 'd2', _p8(nn), t_tab, v_tab, [0.02, 0.97], t.numpy())
    got_g, got_y0, got_y, stats = harness.gradient(*args, gy, True, True, first_step=0.05)
    _, _, steps = harness.integrate(*args, True, True, first_step=0.05, steps_cap=4096)

    want_g, want_y0, want_y, st = _oracle_grad(nn, y0, t, torch.from_numpy(gy), 0.05)
    if (stats[0], stats[1]) == (st['n_accept'], st['n_reject']):   # same accept/reject decisions
        assert np.abs(got_y - want_y).max() <= 1e-10
        assert np.abs(got_g - want_g).max() <= 1e-5 * np.abs(want_g).max()
        assert np.abs(got_y0 - want_y0).max() <= 1e-5 * np.abs(want_y0).max()
    else:                                                          # controller took another path
        assert study != 's1'
        assert np.abs(got_y - want_y).max() <= 5e-6   # global-error envelope at rtol 1e-7

    want_g, want_y0, want_y, st = _oracle_grad(nn, y0, t, torch.from_numpy(gy), 0.05,
                                               replay=[tuple(r) for r in steps])
    assert np.abs(got_y - want_y).max() <= 1e-13
    assert np.abs(got_g - want_g).max() <= 1e-11 * np.abs(want_g).max()
    assert np.abs(got_y0 - want_y0).max() <= 1e-11 * np.abs(want_y0).max()


def test_adjoint_fp32_as_shipped_host_logic():
    """fp32 state + fp32 MLP (the reference's shipped precision): gradient vs the oracle's
    autograd on the replayed step sequence, fp32 noise envelope."""
    torch.set_num_threads(1)
    nn = kat.make_nn('d2')
    t_tab, v_tab = protocols.ap2hz()
    nn.set_fixed_form_voltage_protocol(t_tab, v_tab)
    y0 = torch.tensor([[0.02, 0.97]])
    t = torch.linspace(0., 40., 21)
    gy = np.random.RandomState(3).randn(len(t), 2).astype(np.float32)
    args = (nn.net, 5, 200, True, _p8(nn), t_tab, v_tab, [0.02, 0.97], t.double().numpy())
    got_g, got_y0, _, stats = harness.gradient(*args, gy, False, False, first_step=0.05)
    _, _, steps = harness.integrate(*args, False, False, first_step=0.05, steps_cap=4096)
    want_g, want_y0, _, _ = _oracle_grad(nn, y0, t, torch.from_numpy(gy), 0.05,
                                         replay=[tuple(r) for r in steps])
    assert np.abs(got_g - want_g).max() <= 2e-4 * np.abs(want_g).max()
    assert np.abs(got_y0 - want_y0).max() <= 2e-4 * np.abs(want_y0).max()


def test_table_cache_key_is_content_not_address():
    """a freed protocol array reallocated at the same address with the same length, span
    and sum (a time-shifted pulse) must not hit the stale device table."""
    from neural_ode_ion_channels_b200 import solver
    t = np.linspace(0.0, 100.0, 1001)
    v1 = np.full_like(t, -80.0)
    v1[100:200] = 20.0
    v2 = np.full_like(t, -80.0)
    v2[300:400] = 20.0                       # same length, span and sum
    assert v1.sum() == v2.sum()
    k1, k2 = solver._table_key(t, v1, True), solver._table_key(t, v2, True)
    assert k1 != k2
    buf = v1.copy()
    ka = solver._table_key(t, buf, True)
    buf[:] = v2                              # in-place edit: same address, new content
    assert solver._table_key(t, buf, True) == k2 != ka
    assert solver._table_key(t, v1, True) != solver._table_key(t, v1, False)
