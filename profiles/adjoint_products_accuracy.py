"""Accuracy of the adjoint kernel's product modes (desc.reserved bits 10-11) on IDENTICAL step
checkpoints: one tensor-core forward, then the backward with six / three bf16 products per fp32
product (desc.reserved bit 10), each compared with the FFMA2 backward kernels (plain fp32 FMAs) on
the same checkpoints.  Per parameter block: max and rms deviation relative to the block maximum."""
import copy
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_ode_ion_channels_b200 as ikr  # noqa: E402
from neural_ode_ion_channels_b200 import _cabi  # noqa: E402
from neural_ode_ion_channels_b200.adjoint import _run_backward  # noqa: E402
from tests.test_gpu_parity_r2 import _nn, _segments, _set, _staircase_window  # noqa: E402

for study, B in (('d2', 300), ('d1', 64), ('d2', 2048)):
    func, _ = _nn(study)
    t_tab, v_tab, t, y0, data = _staircase_window(B, 22, n_out=21)
    _set((func,), t_tab, v_tab)
    func.cuda()
    res = ikr.integrate(func, y0.cuda(), t, data=data, E=-86.0, want_y=True, want_ckpt=True,
                        options={'first_step': 0.05})
    assert res.geometry['tensor_cores']

    def backward(bits):
        r = copy.copy(res)
        r._desc = copy.copy(res._desc)
        r._desc.reserved = (res._desc.reserved & ~(1 << 10)) | bits
        flat, gy0, gg = _run_backward(func, r, fused_loss=1, want_y0=True, want_g=True)
        return flat.cpu().numpy(), gy0.double().cpu().numpy(), gg.double().cpu().numpy()
    ref = backward(2)                     # bit 1: FFMA2 backward kernels
    for name, bits in (('six', 1 << 10), ('three', 0)):
        got = backward(bits)
        out = {}
        for seg, lo, hi in _segments():
            d = got[0][lo:hi] - ref[0][lo:hi]
            m = np.abs(ref[0][lo:hi]).max()
            out[seg] = '%.1e/%.1e' % (np.abs(d).max() / m, np.sqrt((d ** 2).mean()) / m)
        out['grad_y0'] = '%.1e' % (np.abs(got[1] - ref[1]).max() / np.abs(ref[1]).max())
        out['grad_g'] = '%.1e' % (np.abs(got[2] - ref[2]).max() / np.abs(ref[2]).max())
        print(study, B, name, json.dumps(out), flush=True)
