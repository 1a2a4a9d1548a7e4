"""The oracle's RHS classes against vectors produced by the reference's OWN classes
(``tests/golden/make_golden.py`` exec-s them from /root/reference at generation time)."""
import os

import numpy as np
import pytest
import torch

from neural_ode_ion_channels_b200 import protocols
from tests import kat

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rhs_vectors.npz'))


@pytest.mark.parametrize('study', kat.STUDIES)
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_rhs_matches_reference_classes(study, tag):
    torch.set_num_threads(1)
    dt = torch.float32 if tag == 'f32' else torch.float64
    t_tab, v_tab = protocols.ap2hz()
    f = kat.make_nn(study)
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    gt, _ = kat.make_gt(study)
    gt.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ygt = GOLD[study + '_ygt']
    with torch.no_grad():
        for i in range(len(GOLD['t'])):
            t = torch.tensor(GOLD['t'][i]).to(dt)
            y = torch.tensor([[GOLD['a'][i], GOLD['r'][i]]]).to(dt)
            got = f(t, y).double().numpy().reshape(-1)
            want = GOLD['%s_nn_%s' % (study, tag)][i]
            # same torch ops on the same host library: identical up to MKL thread-count noise
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)
            got = gt(t, torch.tensor(ygt[i:i + 1]).to(dt)).double().numpy()
            np.testing.assert_allclose(got, GOLD['%s_gt_%s' % (study, tag)][i], rtol=1e-12,
                                       atol=1e-15)


def test_fallback_voltage_is_minus_80():
    # train-s1.py:234-237: t beyond the table -> V = -80 (int tensor => fp32 HH arithmetic)
    f = kat.make_nn('s1')
    t_tab, v_tab = protocols.ap2hz()
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    with torch.no_grad():
        out = f(torch.tensor(1e4), torch.tensor([[0.3, 0.6]]))
    assert out.dtype == torch.float32          # no fp64 V in the fallback branch
    k3 = f.p5 * np.exp(f.p6 * -80.0)
    k4 = f.p7 * np.exp(-f.p8 * -80.0)
    assert abs(out[0, 1].item() - (-k3 * 0.6 + k4 * 0.4)) < 1e-7


@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_hh_fit_rhs_matches_reference_class(tag):
    """HH candidate ODEFunc of the PINTS fit (train-d0.py:321-376): restated class vs vectors
    produced by the reference's own class (tests/golden/make_golden.py::make_hh_fit_vectors)."""
    import os
    from neural_ode_ion_channels_b200 import protocols
    from oracle import ref_models as rm
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'hh_fit_vectors.npz'))
    f = rm.HHFitRhs()
    f.set_fixed_form_voltage_protocol(*protocols.ap2hz())
    dt = torch.float32 if tag == 'f32' else torch.float64
    with torch.no_grad():
        for k, x in enumerate(g['X']):
            f.set_parameters(list(x))
            for i in range(len(g['t'])):
                out = f(torch.tensor(g['t'][i]).to(dt),
                        torch.tensor([[g['a'][i], g['r'][i]]]).to(dt)).double().numpy().reshape(-1)
                assert np.array_equal(out, g['out_' + tag][k, i])
