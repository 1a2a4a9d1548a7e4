"""Builds tests/golden/d2_regression.npz from the reference checkout (run in the build container;
/root/reference does not exist on the GPU box).

The reference's d2 study fits the NN-d discrepancy network by derivative regression
(train-d2.py:880-915) on the 69,361 (V, a, da/dt) points it stores as d2/{v,a,dadt}.pt; d2/log:4-24
holds the loss curve of that fit (0.06535 -> 0.014476 over 8,000 Adam iterations, target 0.015568).
Stored as float32 (the reference casts x_av and y_dadt to float32 before the loop)."""
import os

import numpy as np
import torch

REF = os.environ.get('IKR_REFERENCE', '/root/reference')
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'd2_regression.npz')
arrs = {k: torch.load(os.path.join(REF, 'd2', k + '.pt')).double().numpy() for k in ('v', 'a', 'dadt')}
np.savez_compressed(out, v=arrs['v'].astype(np.float32), a=arrs['a'].astype(np.float32),
                    dadt=arrs['dadt'].astype(np.float32),
                    logged_first=np.float64(0.06535385549068451),      # d2/log:5  "Iter 0"
                    logged_last=np.float64(0.014476394280791283),      # d2/log:24 "Iter 7600"
                    logged_target=np.float64(0.015568403527140617))    # d2/log:4  "Target Loss"
print(out, os.path.getsize(out))
