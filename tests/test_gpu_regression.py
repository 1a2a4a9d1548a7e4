"""MLP regression stage on the tensor cores (train-s1.py:891-909) against PyTorch
autograd of the same loss (fp32 MLP, MSELoss(reduction='sum'))."""
import os
import tempfile

import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from tests import kat

pytestmark = pytest.mark.gpu


def _data(N, seed=0):
    rng = np.random.RandomState(seed)
    v = rng.uniform(-130, 70, N)
    a = rng.uniform(-0.02, 1.02, N)                 # a few points outside (0, 1) for the filter
    dadt = 1e-3 * np.sin(v / 40.0) * (1 - a) + 2e-4 * rng.randn(N)
    return torch.tensor(v), torch.tensor(a), torch.tensor(dadt)


def _torch_loss_grad(func, x, y):
    for p in func.net.parameters():
        p.grad = None
    pred = func.net(x) / float(func.netscale.reshape(-1)[0])
    loss = torch.nn.MSELoss(reduction='sum')(pred.reshape(-1), y)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1).double() for p in func.net.parameters()]).cpu().numpy()
    for p in func.net.parameters():
        p.grad = None
    return float(loss), flat


@pytest.mark.parametrize('N', [1000, 70001])
def test_regression_loss_and_gradient_match_autograd(N):
    func = ikr.load_weights(ikr.ODEFuncNNf(params='s'), kat.weights_path('s1')).cuda()
    v, a, dadt = _data(N)
    x = torch.stack([v / 100.0, a]).T.float().cuda().contiguous()
    y = dadt.float().cuda()
    want_l, want_g = _torch_loss_grad(func, x, y)
    loss, grads = ikr.mse_loss_and_grad(func, x, y)
    got_g = torch.cat([g.reshape(-1).double() for g in grads]).cpu().numpy()
    assert abs(float(loss) - want_l) <= 2e-5 * abs(want_l)
    # fp32 forward on both sides + bf16x2 products in the weight-gradient GEMM (2^-16, unbiased)
    assert np.abs(got_g - want_g).max() <= 2e-4 * np.abs(want_g).max()
    o = 0
    for p in func.net.parameters():                 # every parameter block individually
        k = p.numel()
        ref = np.abs(want_g[o:o + k]).max()
        assert np.abs(got_g[o:o + k] - want_g[o:o + k]).max() <= 1e-3 * ref + 1e-12, tuple(p.shape)
        o += k


def test_fit_regression_follows_the_reference_loop():
    torch.manual_seed(0)
    f_ref = ikr.ODEFuncNNf(params='s').cuda()
    f_b200 = ikr.ODEFuncNNf(params='s').cuda()
    f_b200.load_state_dict(f_ref.state_dict())
    v, a, dadt = _data(20000, seed=1)
    hist = ikr.fit_regression(f_b200, v, a, dadt, n_iter=60, log_every=20)
    # the same loop in plain PyTorch (train-s1.py:885-909)
    x = torch.stack([v / 100.0, a]).T.cuda()
    keep = (x[:, 1] > 0) & (x[:, 1] < 1)
    x, y = x[keep].float(), dadt.cuda()[keep].float()
    opt = torch.optim.Adam(f_ref.net.parameters(), lr=0.001)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=0.9)
    ref_hist = []
    for itr in range(60):
        loss = torch.nn.MSELoss(reduction='sum')((f_ref.net(x) / 1000.0).reshape(-1), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if itr % 20 == 0:
            ref_hist.append(float(loss))
    assert len(hist) == 3 and hist[-1] < hist[0]
    assert np.allclose(hist, ref_hist, rtol=2e-2)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'checkpoint.pt')
        ikr.save_checkpoint(f_b200, f_b200._ikr_regression_optimizer, 60, [0.1, 0.2], path)
        ck = torch.load(path, weights_only=False)
        assert set(ck) == {'epoch', 'state_dict', 'optimizer', 'loss'} and ck['epoch'] == 60
        assert set(ck['state_dict']) == set(f_ref.state_dict())


def test_d2_regression_fit_reaches_the_reference_logged_loss():
    """The reference's own d2 fit (train-d2.py:880-915, data d2/{v,a,dadt}.pt, curve d2/log:4-24):
    NN-d network initialised N(0, 1e-3^2), Adam(1e-3) + StepLR(400, 0.9), full batch of the 69,361
    stored points, loss = sum (net(x) / netscale + HH(a, V) - da/dt)^2.  Every iteration's loss and
    gradient come from the tensor-core kernels (adjoint MMAs bf16x3, weight-gradient GEMM with bf16x2
    products).  The first loss must equal the logged 0.065354 (the network starts at ~0, so this pins
    data + HH term + reduction), and the fit must get below the reference's target loss 0.015568 --
    the logged run sat on the 0.0652 plateau for 4,000 iterations and ended at 0.014476 after 8,000;
    when the plateau is left depends on the random initialisation, so the loop may run to 16,000."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'd2_regression.npz'))
    torch.manual_seed(0)
    func = ikr.ODEFuncNNd(params='d').cuda()
    v = torch.from_numpy(gold['v']).double().cuda()
    a = torch.from_numpy(gold['a']).double().cuda()
    dadt = torch.from_numpy(gold['dadt']).double().cuda()
    with torch.no_grad():
        hh = func._dadt(a, v)                  # k1 (1 - a) - k2 a, train-d2.py:247-250
    target = float(gold['logged_target'])
    hist = ikr.fit_regression(func, v, a, dadt - hh, n_iter=16001, lr=1e-3, step_size=400, gamma=0.9,
                              log_every=400, stop_below=target)
    print('d2 regression fit, loss every 400 iterations:', ['%.6f' % h for h in hist])
    assert abs(hist[0] - float(gold['logged_first'])) < 2e-5, hist[0]
    assert hist[-1] < target, hist
