"""Prediction caches and table rows in the formats the reference's reporting scripts consume
(SURVEY.md 8f-4).

Reference: ``table-1.py:401-417`` (``sim_data`` / ``predict``), ``:468-523`` (the ``--cached`` files
``table-1/y{c,o,1,2}-<protocol>.pt``), ``:525-584`` (mean-absolute-error table in LaTeX),
``table-s1.py:241-287`` (``<arch>-y1-<protocol>.pt``).  ``predict_current`` goes through the B200
``odeint`` path; the file layouts are the reference's: a prediction is a ``(1, T)`` float64 tensor
(fp32 state product times the fp64 ``(1, T)`` voltage of ``ODEFunc._v``), the data trace ``yc`` a
``(T,)`` float64 numpy array.
"""
import os

import numpy as np
import torch

from .solver import integrate

PROTOCOL_KEYS = ('pr3', 'pr5', 'pr4', 'sinewave', 'aps')      # table column order (table-1.py:560)


def predict_current(func, time, voltage, time_torch, g, y0, e, name=None, data=None, log=print):
    """``predict`` of ``table-1.py:409-416``: set the protocol, integrate, observe; returns the
    ``(1, T)`` float64 prediction and (when ``data`` is given) prints the reference's loss line."""
    func.set_fixed_form_voltage_protocol(time, voltage)
    y0 = torch.as_tensor(y0)
    with torch.no_grad():
        res = integrate(func, y0.reshape(1, -1), time_torch, g=torch.tensor([float(g)]), E=float(e),
                        want_y=True)
    pred_y = res.y                                           # (T, 1, 2) in y0.dtype
    v = torch.from_numpy(np.interp(torch.as_tensor(time_torch).double().numpy(),
                                   np.asarray(time, dtype=np.float64),
                                   np.asarray(voltage, dtype=np.float64))).reshape(1, -1)
    pred_yo = (float(g) * pred_y[:, 0, 0] * pred_y[:, 0, 1]).cpu() * (v - float(e))   # (1, T) fp64
    if data is not None and name is not None and log is not None:
        loss = torch.mean(torch.abs(pred_yo - torch.from_numpy(np.asarray(data))))
        log('{:s} prediction | Total Loss {:.6f}'.format(name, loss.item()))
    return pred_yo


def mean_abs_loss(x, y):
    """``loss`` of ``table-1.py:525-527`` (broadcasts a (1, T) prediction against a (T,) trace)."""
    return torch.mean(torch.abs(torch.as_tensor(x) - torch.as_tensor(y))).item()


def pr4_window(n_samples):
    """``table-1.py:535-538``: Pr4 is scored on sweeps 1..3 of its 16 concatenated steps."""
    seg = int(n_samples / 16)
    return slice(seg * 1, seg * (3 + 1))


def save_prediction_cache(directory, protocol, data=None, prefix='', **predictions):
    """Write ``<prefix>y<k>-<protocol>.pt`` for every ``k=tensor`` keyword (``o``, ``1``, ``2``) and
    ``yc-<protocol>.pt`` for the data trace, exactly the files ``--cached`` reloads."""
    os.makedirs(directory, exist_ok=True)
    if data is not None:
        torch.save(np.asarray(data, dtype=np.float64).reshape(-1),
                   os.path.join(directory, '%syc-%s.pt' % (prefix, protocol)))
    for k, pred in predictions.items():
        pred = torch.as_tensor(pred).detach().cpu().to(torch.float64).reshape(1, -1)
        torch.save(pred, os.path.join(directory, '%sy%s-%s.pt' % (prefix, k, protocol)))


def load_prediction_cache(directory, protocol, keys=('o', '1', '2'), prefix=''):
    """The ``--cached`` branch of ``table-1.py:420-440``."""
    out = {}
    path = os.path.join(directory, '%syc-%s.pt' % (prefix, protocol))
    if os.path.exists(path):
        out['c'] = torch.from_numpy(torch.load(path, weights_only=False))
    for k in keys:
        out[k] = torch.load(os.path.join(directory, '%sy%s-%s.pt' % (prefix, k, protocol)),
                            weights_only=False)
    return out


def table_losses(caches, keys=('o', '1', '2')):
    """``caches``: {protocol: {'c': data, 'o': ..., '1': ..., '2': ...}} -> {key: [5 losses]} in the
    column order Pr3, Pr5, Pr4 (windowed), Sinusoidal, APs (``table-1.py:529-548``)."""
    rows = {}
    for k in keys:
        row = []
        for proto in PROTOCOL_KEYS:
            c = caches[proto]
            if proto == 'pr4':
                w = pr4_window(c['c'].reshape(-1).shape[0])
                row.append(mean_abs_loss(c[k].reshape(-1)[w], c['c'].reshape(-1)[w]))
            else:
                row.append(mean_abs_loss(c[k], c['c']))
        rows[k] = row
    return rows


def table1_latex(rows, labels=(('o', 'Original'), ('1', 'NN-f'), ('2', 'NN-d'))):
    """The LaTeX body written to ``table-1/table-1.txt`` (``table-1.py:550-584``), same rounding."""
    out = "{@{}XXXcXXX@{}}\n"
    out += "\\toprule\n"
    out += "         & \\multicolumn{2}{c}{Training} & \\phantom{a} & \\multicolumn{3}{c}{Prediction} \\\\\n"
    out += "           \\cmidrule{2-3}                               \\cmidrule{5-7}\n"
    out += "         & Pr3 & Pr5 & & Pr4 & Sinusoidal & APs \\\\\n"
    out += "\\midrule\n"
    for key, label in labels:
        r = [round(x, 3) for x in rows[key]]
        out += "%-8s & %s & %s & & %s & %s & %s \\\\\n" % (label, r[0], r[1], r[2], r[3], r[4])
    out += "\\bottomrule"
    return out
