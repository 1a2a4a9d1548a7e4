"""Reporting formats against the reference's own cached predictions and table:
``tests/golden/table1_cache.json`` holds the shapes / dtypes / losses derived from
``/root/reference/table-1/*.pt`` and the text of ``table-1/table-1.txt`` (make_golden.py)."""
import json
import os

import numpy as np
import torch

from neural_ode_ion_channels_b200 import reporting as rp

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'table1_cache.json')) as _fh:
    GOLD = json.load(_fh)


def test_table1_latex_is_reproduced_from_the_reference_losses():
    # Pr3 / Pr5 / Pr4 columns: the table's own numbers (their caches are not in the checkout);
    # Sinusoidal / APs: recomputed from the reference's cached tensors with mean_abs_loss
    txt = GOLD['table_txt']
    printed = {}
    for line in txt.splitlines():
        for label, key in (('Original', 'o'), ('NN-f', '1'), ('NN-d', '2')):
            if line.startswith(label):
                printed[key] = [float(x) for x in line.replace('\\\\', '').split('&')[1:] if x.strip()]
    rows = {}
    for key in ('o', '1', '2'):
        rows[key] = printed[key][:3] + [GOLD['losses']['sinewave'][key], GOLD['losses']['aps'][key]]
        assert round(rows[key][3], 3) == printed[key][3] and round(rows[key][4], 3) == printed[key][4]
    assert rp.table1_latex(rows) == txt


def test_cache_files_have_the_reference_layout(tmp_path):
    T = 1000
    rng = np.random.RandomState(0)
    data = rng.randn(T)
    preds = {k: torch.from_numpy(rng.randn(1, T)) for k in ('o', '1', '2')}
    for proto in rp.PROTOCOL_KEYS:
        rp.save_prediction_cache(str(tmp_path), proto, data=data, **preds)
    # what `table-1.py --cached` does (table-1.py:421-440)
    yc = torch.from_numpy(torch.load(os.path.join(tmp_path, 'yc-aps.pt'), weights_only=False))
    y1 = torch.load(os.path.join(tmp_path, 'y1-aps.pt'), weights_only=False)
    ref = GOLD['shapes']['aps']
    assert str(yc.dtype) == ref['c'][1] and yc.dim() == len(ref['c'][0])
    assert str(y1.dtype) == ref['1'][1] and y1.dim() == len(ref['1'][0]) and y1.shape[0] == 1
    caches = {p: rp.load_prediction_cache(str(tmp_path), p) for p in rp.PROTOCOL_KEYS}
    rows = rp.table_losses(caches)
    want = float(np.mean(np.abs(preds['1'].numpy().reshape(-1) - data)))
    assert abs(rows['1'][4] - want) < 1e-15 and abs(rows['1'][0] - want) < 1e-15
    w = rp.pr4_window(T)
    assert (w.start, w.stop) == (62, 248)
    want4 = float(np.mean(np.abs(preds['2'].numpy().reshape(-1)[w] - data[w])))
    assert abs(rows['2'][2] - want4) < 1e-15
    # table-s1.py:241-287 naming
    rp.save_prediction_cache(str(tmp_path), 'pr4', prefix='s03-', **{'1': preds['1']})
    assert os.path.exists(os.path.join(tmp_path, 's03-y1-pr4.pt'))
