"""MLP regression stage of the reference's training scripts on the B200.

Reference: ``train-s1.py:891-909`` / ``train-d1.py`` / ``train-r1.py:917-959``: 4,000-16,000 full-batch
Adam iterations of ``loss = MSELoss(reduction='sum')(net(x_av) / netscale, y_dadt)`` over the
69k-214k ``(V / vrange, a)`` points estimated from the data.  ``mse_loss_and_grad`` is one
iteration's forward + backward on the tensor-core kernels (``ikr_regression_loss_grad``); the
optimiser, the learning-rate schedule and the checkpoint dictionary stay PyTorch objects, so
``fit_regression`` reads like the reference loop and ``save_checkpoint`` writes the reference's
``{'epoch', 'state_dict', 'optimizer', 'loss'}`` layout (``train-r1.py:947-952``).
"""
import ctypes

import torch

from . import _cabi
from .solver import _device_model, _make_desc, _resolve_device, describe, unpack_grads


def mse_loss_and_grad(func, x_av, y_dadt, *, device=None, accumulate=False):
    """``x_av`` (N, 2) = (V / vrange, a), ``y_dadt`` (N,).  Returns (loss fp64 0-d tensor, grads)
    with ``grads`` shaped like ``func.net.parameters()``; ``accumulate=True`` also adds them to
    ``.grad`` (what ``loss.backward()`` does in ``train-s1.py:906``)."""
    dev = _resolve_device(x_av, device)
    with torch.cuda.device(dev):
        dm = _device_model(func, dev)
        spec = dm.spec
        if spec.mlp_dtype != torch.float32:
            raise TypeError('mse_loss_and_grad: the tensor-core regression path needs an fp32 MLP')
        desc = _make_desc(spec, torch.float32, 'dopri5', 1e-7, 1e-9, {})
        lib = _cabi.lib()
        x = x_av.detach().to(device=dev, dtype=torch.float32).contiguous()
        y = y_dadt.detach().to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
        if x.dim() != 2 or x.shape[1] != 2 or y.numel() != x.shape[0]:
            raise ValueError('mse_loss_and_grad: x_av must be (N, 2) and y_dadt (N,)')
        N = x.shape[0]
        ws_bytes = lib.ikr_regression_workspace_bytes(ctypes.byref(desc), N)
        if ws_bytes == 0:
            raise TypeError('mse_loss_and_grad: architecture (%d layers x %d nodes) is outside the '
                            'tensor-core regression path' % (spec.n_layers, spec.n_nodes))
        weights = dm.weights_for(desc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float64, device=dev)
        flat = torch.empty(lib.ikr_param_count(ctypes.byref(desc)), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev)
        _cabi.check(lib.ikr_regression_loss_grad(
            ctypes.byref(desc), weights.data_ptr(), x.data_ptr(), y.data_ptr(), N, loss.data_ptr(),
            flat.data_ptr(), ws.data_ptr(), ws_bytes, ctypes.c_void_p(stream.cuda_stream)),
            'ikr_regression_loss_grad')
        ws.record_stream(stream)
    plist = [q for m in spec.linears for q in (m.weight, m.bias)]
    grads = [g.to(device=q.device, dtype=q.dtype) for g, q in zip(unpack_grads(spec, flat), plist)]
    if accumulate:
        for q, g in zip(plist, grads):
            q.grad = g.clone() if q.grad is None else q.grad + g
    return loss, grads


def fit_regression(func, v, a, dadt, *, n_iter=4000, lr=0.001, step_size=100, gamma=0.9,
                   log_every=400, log=None, stop_below=None):
    """The reference's regression loop (``train-s1.py:885-909``): keep 0 < a < 1, Adam(lr) with
    StepLR(step_size, gamma), ``n_iter`` full-batch iterations (fewer when a logged loss falls below
    ``stop_below``).  Returns the list of logged losses."""
    describe(func)
    dev = next(func.net.parameters()).device
    vr = float(func.vrange) if not isinstance(func.vrange, torch.Tensor) else float(func.vrange.reshape(-1)[0])
    x_av = torch.stack([torch.as_tensor(v).reshape(-1).to(dev) / vr,
                        torch.as_tensor(a).reshape(-1).to(dev)]).T
    y = torch.as_tensor(dadt).reshape(-1).to(dev)
    keep = (x_av[:, 1] > 0) & (x_av[:, 1] < 1)
    x_av, y = x_av[keep].float().contiguous(), y[keep].float().contiguous()
    opt = torch.optim.Adam(func.net.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=gamma)
    history = []
    for itr in range(n_iter):
        opt.zero_grad()
        loss, _ = mse_loss_and_grad(func, x_av, y, accumulate=True)
        opt.step()
        sched.step()
        if log_every and itr % log_every == 0:
            history.append(float(loss))
            if log is not None:
                log('Iter %d LR %g Loss %g' % (itr, opt.param_groups[0]['lr'], history[-1]))
            if stop_below is not None and history[-1] < stop_below:
                break
    func._ikr_regression_optimizer = opt
    return history


def save_checkpoint(func, optimizer, epoch, losses, path):
    """``train-r1.py:947-952``: {'epoch', 'state_dict', 'optimizer', 'loss'}."""
    torch.save({'epoch': int(epoch), 'state_dict': func.state_dict(),
                'optimizer': optimizer.state_dict(), 'loss': list(losses)}, path)
