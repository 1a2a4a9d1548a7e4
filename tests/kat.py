"""Known-answer cases of the reference (``{s1,s2,d1,d2}/log2``) expressed against the oracle.

A case = (study, protocol row).  ``oracle_case`` reproduces exactly what the reference's
``--pred`` branch does for that row (``train-s1.py:311-329, 431-546``): ground-truth model and
trained NN model integrated with dopri5 from fp32 y0 on an fp32 ``linspace`` grid, currents formed
with E = -86 and the loss is mean |I_nn - I_gt|."""
import json
import os

import numpy as np
import torch

from neural_ode_ion_channels_b200 import protocols
from oracle import ref_models as rm
from oracle import ref_odeint as ro

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS = os.path.join(os.path.dirname(HERE), 'neural-ode-ion-channels_b200', 'data', 'weights')

with open(os.path.join(HERE, 'golden', 'kat_log2.json')) as _fh:
    KAT = json.load(_fh)

STUDIES = ('s1', 's2', 'd1', 'd2')


def weights_path(study):
    return os.path.join(WEIGHTS, '%s-model-state-dict.pt' % study)


def make_nn(study, mlp_follows_state=False):
    if study == 's1':
        f = rm.NNfRhs(inact=rm.HH_B06[4:], mlp_follows_state=mlp_follows_state)
    elif study == 's2':
        f = rm.NNdRhs(act=rm.HH_B06[:4], inact=rm.HH_B06[4:], mlp_follows_state=mlp_follows_state)
    elif study == 'd1':
        f = rm.NNfRhs(inact=rm.INACT_D, mlp_follows_state=mlp_follows_state)
    elif study == 'd2':
        f = rm.NNdRhs(act=rm.HH_B06[:4], inact=rm.INACT_D, mlp_follows_state=mlp_follows_state)
    else:
        raise KeyError(study)
    return rm.load_state_dict_file(f, weights_path(study))


def make_gt(study):
    if study in ('s1', 's2'):
        return rm.HHRhs(), torch.tensor([[0., 1.]])
    return rm.MarkovRhs(), torch.tensor([[0., 1., 0., 0., 0., 0.]])


def row_protocol(row):
    """(t_table, v_table, t_out fp32 tensor) for one log2 row."""
    sec = row['section']
    if sec == 'top' and row['name'] == 'AP 2Hz':
        t, v = protocols.ap2hz()
        return t, v, torch.linspace(0., 3000, 1501)
    if sec == 'Activation':
        t, v = protocols.pr3_activation(row['value'])
        return t, v, torch.linspace(0., 8000., 8001)
    if sec == 'Deactivation':
        t, v = protocols.pr5_deactivation(row['value'])
        return t, v, torch.linspace(0., 10000., 10001)
    if sec.startswith('Activation time constant'):
        t, v = protocols.pr2_time_constant(row['value'])
        return t, v, torch.linspace(0., 5000., 5001)
    raise KeyError(row)


def gt_current(study, t_tab, v_tab, t_out):
    gt, y0 = make_gt(study)
    gt.set_fixed_form_voltage_protocol(t_tab, v_tab)
    with torch.no_grad():
        y = ro.odeint(gt, y0, t_out, method='dopri5')
        if study in ('s1', 's2'):
            return y[:, 0, 0] * y[:, 0, 1] * (gt._v(t_out) + 86)
        return y[:, 0, -1] * (gt._v(t_out) + 86)


def oracle_case(study, row, stats=None):
    t_tab, v_tab, t_out = row_protocol(row)
    i_gt = gt_current(study, t_tab, v_tab, t_out)
    f = make_nn(study)
    f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    with torch.no_grad():
        y = ro.odeint(f, torch.tensor([[0., 1.]]), t_out, stats=stats)
        i_nn = y[:, 0, 0] * y[:, 0, 1] * (f._v(t_out) + 86)
    return torch.mean(torch.abs(i_nn - i_gt)).item(), i_nn.reshape(-1).numpy(), \
        i_gt.reshape(-1).numpy(), y.numpy()
