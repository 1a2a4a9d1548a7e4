"""Gradient through ``odeint``: discrete adjoint of the accepted-step sequence (the semantics of
torchdiffeq's non-adjoint autograd -- step sizes are constants of the backward pass).

The reference never differentiates through ``odeint`` (SURVEY.md finding 3); this is the new
capability the north star asks for, offered behind the same ``odeint`` call: when any MLP
parameter requires grad (and grad mode is on) the forward records per-step checkpoints and
``loss.backward()`` runs the fused backward kernel (``ikr_backward``)."""
import ctypes

import torch

from . import _cabi
from .solver import _resolve_device, integrate, unpack_grads


class _OdeintFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, y0, t, kwargs, *params):
        res = integrate(func, y0, t, want_ckpt=True, **kwargs)
        ctx.res = res
        ctx.func = func
        ctx.n_params = len(params)
        ctx.y0_requires_grad = y0.requires_grad
        return res.y

    @staticmethod
    def backward(ctx, grad_y):
        from .solver import _device_model
        res = ctx.res
        desc, io = res._desc, res._io
        dev = res.y.device
        with torch.cuda.device(dev):
            dm = _device_model(ctx.func, dev)
            spec = dm.spec
            lib = _cabi.lib()
            n_par = lib.ikr_param_count(ctypes.byref(desc))
            acc_dtype = torch.float64 if res.y.dtype == torch.float64 else torch.float32
            grad_flat = torch.zeros(n_par, dtype=acc_dtype, device=dev)
            gy = grad_y.contiguous().to(res.y.dtype)
            B = res.y.shape[1]
            grad_y0 = torch.zeros((B, 2), dtype=res.y.dtype, device=dev)
            bio = _cabi.IkrBwdIO()
            bio.grad_y = gy.data_ptr()
            bio.fused_loss = 0
            bio.weights_bwd = io.weights
            bio.grad_weights = grad_flat.data_ptr()
            bio.grad_y0 = grad_y0.data_ptr()
            ws_bytes = lib.ikr_workspace_bytes(ctypes.byref(desc), 1, B, 1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev)
            _cabi.check(lib.ikr_backward(ctypes.byref(desc), ctypes.byref(io), ctypes.byref(bio),
                                         ws.data_ptr(), ws_bytes,
                                         ctypes.c_void_p(stream.cuda_stream)), 'ikr_backward')
            grads = [g.to(p.dtype) for g, p in
                     zip(unpack_grads(spec, grad_flat),
                         [q for m in spec.linears for q in (m.weight, m.bias)])]
        return (None, grad_y0 if ctx.y0_requires_grad else None, None, None) + tuple(grads)


def odeint_with_grad(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None):
    dev = _resolve_device(y0, None)
    kwargs = dict(rtol=rtol, atol=atol, method=method, options=options, device=dev)
    from .solver import describe
    spec = describe(func)
    params = [q for m in spec.linears for q in (m.weight, m.bias)]
    return _OdeintFn.apply(func, y0.to(dev), t, kwargs, *params)
