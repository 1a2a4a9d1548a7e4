"""``reporting.predict_current`` (table-1.py:409-416 ``predict``) through the B200 path: reference
tensor layout, the reference's logged AP-2Hz loss, and a cache round trip."""
import os

import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import reporting as rp
from tests import kat

pytestmark = pytest.mark.gpu


def test_predict_current_layout_loss_and_cache(tmp_path):
    func = ikr.load_weights(ikr.ODEFuncNNf(params='d'), kat.weights_path('d1'))
    row = kat.KAT['d1'][0]                                    # d1/log2:4  AP 2Hz, 0.116660
    t_tab, v_tab, t_out = kat.row_protocol(row)
    i_gt = kat.gt_current('d1', t_tab, v_tab, t_out).reshape(-1).numpy()
    lines = []
    pred = rp.predict_current(func, t_tab, v_tab, t_out, 1.0, [0., 1.], -86.0, name='AP 2Hz',
                              data=i_gt, log=lines.append)
    assert pred.shape == (1, len(t_out)) and pred.dtype == torch.float64
    assert abs(rp.mean_abs_loss(pred, i_gt) - row['loss']) < 5e-5
    assert lines and lines[0].startswith('AP 2Hz prediction | Total Loss 0.1166')
    rp.save_prediction_cache(str(tmp_path), 'aps', data=i_gt, **{'1': pred})
    back = rp.load_prediction_cache(str(tmp_path), 'aps', keys=('1',))
    assert torch.equal(back['1'], pred) and np.array_equal(back['c'].numpy(), i_gt)
    assert os.path.exists(os.path.join(tmp_path, 'yc-aps.pt'))
