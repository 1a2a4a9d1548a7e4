"""Parity of the tcgen05 forward kernel (bf16x3 split MMAs, activations and accumulators in TMEM):
against the CPU oracle on fixed-step rk4 (no accept/reject decisions => a clean arithmetic check),
against the FFMA2 kernel on the same inputs, and against the reference's logged losses."""
import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_odeint as ro
from tests import kat

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _pair(study):
    """(product module, oracle module) with identical weights."""
    cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
    pset = 's' if study in ('s1', 's2') else 'd'
    func = ikr.load_weights(cls(params=pset), kat.weights_path(study))
    return func, kat.make_nn(study)


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_tc_rk4_matches_oracle_fp32(study):
    """fp32 state + fp32 MLP, fixed grid: every RHS evaluation of the tensor-core path must agree
    with the oracle's fp32 torch MLP to fp32 rounding (tolerance 2e-6 abs on a, r in [0, 1])."""
    func, ofunc = _pair(study)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 120., 241)
    y0 = torch.tensor([[0.01, 0.97], [0.0, 1.0], [0.04, 0.95]], dtype=torch.float32)
    res = ikr.integrate(func, y0.cuda(), t, method='rk4')
    assert res.geometry['tile_m'] == 128 and res.geometry['threads'] in (192, 320, 448)
    with torch.no_grad():
        for b in range(3):
            want = ro.odeint(ofunc, y0[b:b + 1], t, method='rk4')
            err = (res.y[:, b:b + 1].cpu() - want).abs().max().item()
            assert err < 2e-6, (study, b, err)


def test_tc_matches_ffma_kernel_and_fp64_state_mode():
    func, _ = _pair('d1')
    t_tab, v_tab = protocols.pr4_inactivation_standin(-20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 300., 151)
    rng = np.random.RandomState(7)
    B = 300
    y0 = np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1)
    # 150 rk4 steps: fp32 rounding of the two MLP arithmetics accumulates to a few 1e-6
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-5)):
        y = torch.tensor(y0, dtype=dtype).cuda()
        tt = t.to(dtype)
        a = ikr.integrate(func, y, tt, method='rk4')
        b = ikr.integrate(func, y, tt, method='rk4', options={'tensor_cores': False})
        assert a.geometry['tile_m'] == 128 and b.geometry['threads'] not in (192, 320, 448)
        assert (a.y - b.y).abs().max().item() < tol
        # adaptive: same solver, decisions may flip at knife-edge ratios => solver-level agreement
        c = ikr.integrate(func, y, tt)
        d = ikr.integrate(func, y, tt, options={'tensor_cores': False})
        assert (c.y - d.y).abs().max().item() < 5e-4
        assert abs(int(c.stats[:, 2].sum()) - int(d.stats[:, 2].sum())) < 0.02 * int(d.stats[:, 2].sum())


def test_tc_logged_loss_and_narrow_architectures():
    # the reference's logged AP-2Hz MAE (s1/log2:4) through the tensor-core path
    func, _ = _pair('s1')
    row = kat.KAT['s1'][0]
    t_tab, v_tab, t_out = kat.row_protocol(row)
    i_gt = kat.gt_current('s1', t_tab, v_tab, t_out).reshape(-1)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    res = ikr.integrate(func, torch.tensor([[0., 1.]]).cuda(), t_out, data=i_gt.float(), E=-86.0,
                        want_y=False)
    assert res.geometry['tile_m'] == 128
    assert abs(float(res.sae[0]) / len(t_out) - row['loss']) < 5e-5
    # widths with other k-step / tail geometries: n = 100 (tail of 4), 72 (tail of 8), 64 (no tail),
    # 48 (odd number of k-steps), 90 (zero-padded k-step)
    t_tab, v_tab = protocols.ap2hz()
    t = torch.linspace(0., 60., 61)
    y0 = torch.tensor([[0.01, 0.97]] * 3, dtype=torch.float32).cuda()
    for n, L in ((100, 5), (72, 2), (64, 1), (48, 3), (90, 2), (200, 1)):
        torch.manual_seed(n)
        f = ikr.ODEFuncNNf(arch=(L, n))
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
        a = ikr.integrate(f, y0, t, method='rk4')
        b = ikr.integrate(f, y0, t, method='rk4', options={'tensor_cores': False})
        assert a.geometry['tile_m'] == 128, n
        assert (a.y - b.y).abs().max().item() < 5e-6, (n, L)
