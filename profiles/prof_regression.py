"""Timing of one regression iteration (forward + MSE(sum) + backward) on N (V, a) points:
tensor-core kernels of this repo vs PyTorch autograd (cuBLAS fp32) on the same GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_ode_ion_channels_b200 as ikr  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 214000
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = os.path.join(root, 'neural-ode-ion-channels_b200', 'data', 'weights', 's1-model-state-dict.pt')
f = ikr.load_weights(ikr.ODEFuncNNf(params='s'), w).cuda()
rng = np.random.RandomState(0)
x = torch.tensor(np.stack([rng.uniform(-1.3, 0.7, N), rng.uniform(0, 1, N)], 1), dtype=torch.float32).cuda()
y = torch.tensor(1e-3 * rng.randn(N), dtype=torch.float32).cuda()
torch.backends.cuda.matmul.allow_tf32 = False


def run(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_step():
    for p in f.net.parameters():
        p.grad = None
    loss = torch.nn.MSELoss(reduction='sum')((f.net(x) / 1000.0).reshape(-1), y)
    loss.backward()


ms_b200 = run(lambda: ikr.mse_loss_and_grad(f, x, y))
ms_torch = run(torch_step)
flop = 3 * 2 * 200600 * N
print('regression iteration, N = %d: tcgen05 kernels %.3f ms (%.1f TFLOP/s algorithmic) | PyTorch '
      'autograd fp32 (cuBLAS, TF32 off) %.3f ms (%.1f TFLOP/s)' % (N, ms_b200, flop / ms_b200 / 1e9,
                                                                 ms_torch, flop / ms_torch / 1e9))
