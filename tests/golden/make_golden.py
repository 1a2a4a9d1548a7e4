#!/usr/bin/env python
"""Generate the golden fixtures under ``tests/golden/`` from the reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py [--traces]

Produces
* ``kat_log2.json``    -- the losses the reference itself logged (``{s1,s2,d1,d2}/log2``), parsed
                          verbatim; rows whose input CSVs are missing are kept but flagged.
* ``rhs_vectors.npz``  -- outputs of the reference's OWN ``ODEFunc`` / ``Lambda`` classes (their
                          source is ``exec``-ed straight out of ``/root/reference/train-*.py`` at
                          generation time, nothing is copied into this repo) on seeded (t, y)
                          probes, including out-of-table times (the V = -80 fallback).
* ``traces_*.npz``     -- (``--traces``) oracle trajectories/currents used by the GPU parity tests.
"""
import argparse
import ast
import json
import os
import re
import sys

import numpy as np
import torch
import torch.nn as nn
from scipy.interpolate import interp1d

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)


def reference_classes(script, names=('Lambda', 'ODEFunc')):
    """exec the named top-level classes of a reference script in an isolated namespace."""
    path = os.path.join(REF, script)
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src)
    ns = {'torch': torch, 'nn': nn, 'np': np, 'interp1d': interp1d,
          'device': torch.device('cpu')}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.get_source_segment(src, node), path, 'exec'), ns)
    return {n: ns[n] for n in names if n in ns}


def parse_log2(model):
    rows = []
    section = None
    with open(os.path.join(REF, model, 'log2')) as fh:
        for line in fh:
            line = line.rstrip('\n')
            m = re.match(r'^(.*) prediction \| Total Loss ([0-9.]+)$', line)
            if m:
                rows.append({'section': 'top', 'name': m.group(1), 'loss': float(m.group(2))})
                continue
            m = re.match(r'^(.*) prediction:$', line)
            if m:
                section = m.group(1)
                continue
            m = re.match(r'^\s+(-?[0-9.]+)(mV|ms) \| Total Loss ([0-9.]+)$', line)
            if m:
                rows.append({'section': section, 'name': m.group(1) + m.group(2),
                             'value': float(m.group(1)), 'loss': float(m.group(3))})
    for r in rows:
        r['inputs_present'] = not (r['section'] == 'top' and r['name'] != 'AP 2Hz')
    return rows


SCRIPT = {'s1': 'train-s1.py', 's2': 'train-s2.py', 'd1': 'train-d1.py', 'd2': 'train-d2.py'}


def make_kat():
    out = {m: parse_log2(m) for m in SCRIPT}
    with open(os.path.join(HERE, 'kat_log2.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    n = sum(sum(r['inputs_present'] for r in rows) for rows in out.values())
    print('kat_log2.json: %d reproducible rows' % n)


def make_rhs_vectors():
    from neural_ode_ion_channels_b200 import protocols
    t_tab, v_tab = protocols.ap2hz()
    rng = np.random.RandomState(1234)
    n = 48
    tq = np.concatenate([rng.uniform(0, 3499.9, n - 8), [0.0, 3499.9, 3499.95, 3600.0, 1e4],
                         rng.uniform(3500, 5000, 3)])
    aq = rng.uniform(0, 1, n)
    rq = rng.uniform(0, 1, n)
    blob = {'t': tq, 'a': aq, 'r': rq}
    for model, script in SCRIPT.items():
        cls = reference_classes(script)
        func = cls['ODEFunc']()
        func.load_state_dict(torch.load(os.path.join(REF, model, 'model-state-dict.pt')))
        func.eval()
        func.set_fixed_form_voltage_protocol(t_tab, v_tab)
        gt = cls['Lambda']()
        gt.set_fixed_form_voltage_protocol(t_tab, v_tab)
        n_state = 2 if model in ('s1', 's2') else 6
        ygt = rng.uniform(0, 1, (n, n_state))
        blob[model + '_ygt'] = ygt
        for st_dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            o_nn, o_gt = [], []
            with torch.no_grad():
                for i in range(n):
                    t = torch.tensor(tq[i]).to(st_dtype)
                    y = torch.tensor([[aq[i], rq[i]]]).to(st_dtype)
                    o_nn.append(func(t, y).double().numpy().reshape(-1))
                    o_gt.append(gt(t, torch.tensor(ygt[i:i + 1]).to(st_dtype)).double().numpy())
            blob['%s_nn_%s' % (model, tag)] = np.array(o_nn)
            blob['%s_gt_%s' % (model, tag)] = np.array(o_gt)
    np.savez_compressed(os.path.join(HERE, 'rhs_vectors.npz'), **blob)
    print('rhs_vectors.npz: %d probes x 4 models' % n)


def nested_class(script, name):
    """exec a class defined anywhere in a reference script (train-d0.py nests its ODEFunc under
    the ``--myokit`` else-branch)."""
    path = os.path.join(REF, script)
    with open(path) as fh:
        src = fh.read()
    ns = {'torch': torch, 'nn': nn, 'np': np, 'interp1d': interp1d,
          'device': torch.device('cpu')}
    import textwrap
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == name:
            lines = src.splitlines()[node.lineno - 1:node.end_lineno]
            exec(compile(textwrap.dedent('\n'.join(lines)), path, 'exec'), ns)
            return ns[name]
    raise KeyError(name)


def make_hh_fit_vectors():
    """Outputs of the reference's own HH candidate ``ODEFunc`` (train-d0.py:321-376) for a few
    CMA-ES-like parameter vectors on seeded (t, y) probes -> ``hh_fit_vectors.npz``."""
    from neural_ode_ion_channels_b200 import protocols
    t_tab, v_tab = protocols.ap2hz()
    rng = np.random.RandomState(4321)
    n = 32
    tq = np.concatenate([rng.uniform(0, 3499.9, n - 4), [0.0, 3499.9, 3600.0, 1e4]])
    aq, rq = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    X = np.array([[1.13e-4, 7.45e-2, 3.60e-5, 4.49e-2],
                  [2.26e-4, 6.99e-2, 3.45e-5, 5.46e-2],
                  [5.0e-5, 9.0e-2, 1.0e-4, 3.0e-2]])
    func = nested_class('train-d0.py', 'ODEFunc')()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    blob = {'t': tq, 'a': aq, 'r': rq, 'X': X}
    for st_dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        out = np.zeros((len(X), n, 2))
        with torch.no_grad():
            for k, x in enumerate(X):
                func.set_parameters(list(x))
                for i in range(n):
                    t = torch.tensor(tq[i]).to(st_dtype)
                    y = torch.tensor([[aq[i], rq[i]]]).to(st_dtype)
                    out[k, i] = func(t, y).double().numpy().reshape(-1)
        blob['out_' + tag] = out
    np.savez_compressed(os.path.join(HERE, 'hh_fit_vectors.npz'), **blob)
    print('hh_fit_vectors.npz: %d probes x %d parameter vectors' % (n, len(X)))


def make_table1_cache_fixture():
    """Formats and losses of the reference's own cached predictions (`table-1/*.pt`, written by
    table-1.py:468-523) and its LaTeX table -> ``table1_cache.json`` (the tensors themselves stay in
    the reference checkout: only shapes, dtypes and the derived losses are committed)."""
    from neural_ode_ion_channels_b200 import reporting as rp
    out = {'source': 'derived from /root/reference/table-1/*.pt and table-1.txt by '
                     'tests/golden/make_golden.py'}
    losses = {}
    for proto in ('pr4', 'sinewave', 'aps'):
        c = {}
        p = os.path.join(REF, 'table-1', 'yc-%s.pt' % proto)
        if os.path.exists(p):
            c['c'] = torch.from_numpy(torch.load(p, weights_only=False))
        for k in ('o', '1', '2'):
            c[k] = torch.load(os.path.join(REF, 'table-1', 'y%s-%s.pt' % (k, proto)), weights_only=False)
        out.setdefault('shapes', {})[proto] = {k: [list(c[k].shape), str(c[k].dtype)] for k in c}
        if 'c' in c:
            losses[proto] = {k: rp.mean_abs_loss(c[k], c['c']) for k in ('o', '1', '2')}
    out['losses'] = losses
    with open(os.path.join(REF, 'table-1', 'table-1.txt')) as fh:
        out['table_txt'] = fh.read()
    with open(os.path.join(HERE, 'table1_cache.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print('table1_cache.json: %d protocols with data traces' % len(losses))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--traces', action='store_true')
    args = ap.parse_args()
    torch.set_num_threads(1)
    make_kat()
    make_rhs_vectors()
    make_hh_fit_vectors()
    make_table1_cache_fixture()
    if args.traces:
        from tests.golden import make_traces
        make_traces.main()


if __name__ == '__main__':
    main()
