"""Host-side mirrors of the reference's ODE-func ``nn.Module`` classes.

The reference re-declares the same classes in every script; these are the two that sit on the
hot path, with the same attribute names so that ``state_dict`` files load unchanged and so that
``odeint`` (which *introspects* the module instead of calling ``forward`` per stage) accepts
either these classes or the reference's own:

* ``ODEFunc``   / ``ODEFuncNNf`` -- NN-f, ``train-s1.py:181-247`` (d-study constants
  ``train-d1.py:220-223``, cell-5 ``train-r1.py:171-174``)
* ``ODEFuncNNd``                -- NN-d, ``train-d2.py:191-272`` (cell-5 ``train-r2.py:167-174``)
* ``build_net`` / ``ARCHITECTURES`` -- ``train-r1-tune.py:155-163`` + ``architectures/s00..s11.py``

``forward(t, y)`` keeps the reference's signature for direct calls (plots, rate surfaces) and is
batched over ``y`` of shape (B, 2); it is *not* used by ``odeint`` -- the integration runs in the
fused CUDA kernels only.
"""
import numpy as np
import torch
import torch.nn as nn

#: (n_layers, n_nodes) per architectures/sNN.py
ARCHITECTURES = {
    's00': (5, 200), 's01': (1, 200), 's02': (10, 200), 's03': (5, 10), 's04': (1, 10),
    's05': (10, 10), 's06': (5, 500), 's07': (1, 500), 's08': (10, 500), 's09': (5, 100),
    's10': (1, 100), 's11': (10, 100),
}

_B06 = dict(p1=1.12592345582957387e-01 * 1e-3, p2=8.26751134920666146e+01 * 1e-3,
            p3=3.38768033864048357e-02 * 1e-3, p4=4.67106147665183542e+01 * 1e-3,
            p5=8.47769667061995875e+01 * 1e-3, p6=2.04001345352499328e+01 * 1e-3,
            p7=1.02860743916105211e+01 * 1e-3, p8=2.78201179336874098e+01 * 1e-3)
_D = dict(p5=9.62243079990877703e+01 * 1e-3, p6=2.26404683824047979e+01 * 1e-3,
          p7=8.00924780462999131e+00 * 1e-3, p8=2.43749808069009823e+01 * 1e-3)
_CELL5_F = dict(p5=8.7324e-2, p6=7.3338e-3, p7=6.1655e-3, p8=3.1574e-2)
_CELL5_D = dict(p1=2.1055e-4, p2=6.5799e-2, p3=3.3172e-6, p4=7.4310e-2)

#: named constant sets: 's' synthetic study (B06), 'd' discrepancy study, 'r' real cell-5 data
PARAMETER_SETS = {
    's': dict(_B06),
    'd': dict(_B06, **_D),
    'r': dict(_B06, **_CELL5_F, **_CELL5_D),
}


def build_net(n_layers=5, n_nodes=200, std=0.1):
    """``Linear(2,n) LeakyReLU [Linear(n,n) LeakyReLU]*n_layers Linear(n,1)`` with the
    reference's initialisation (weights N(0, std^2), biases 0)."""
    layers = [nn.Linear(2, n_nodes), nn.LeakyReLU()]
    for _ in range(n_layers):
        layers += [nn.Linear(n_nodes, n_nodes), nn.LeakyReLU()]
    layers += [nn.Linear(n_nodes, 1)]
    net = nn.Sequential(*layers)
    for m in net.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0, std=std)
            nn.init.constant_(m.bias, val=0)
    return net


class _IKrBase(nn.Module):
    def __init__(self, arch, std, params):
        super().__init__()
        if isinstance(arch, str):
            arch = ARCHITECTURES[arch]
        self.net = build_net(arch[0], arch[1], std)
        self.vrange = torch.tensor([100.])
        self.netscale = torch.tensor([1000.])
        self.unity = torch.tensor([1])
        for k, v in params.items():
            setattr(self, k, float(v))
        self._t_regular = None
        self._v_regular = None

    # -- protocol ---------------------------------------------------------------------------
    def set_fixed_form_voltage_protocol(self, t, v):
        """Same contract as the reference (``train-s1.py:218-222``): regular time-series arrays
        (ms, mV); linear interpolation in between, V = -80 outside."""
        self._t_regular = np.asarray(t, dtype=np.float64)
        self._v_regular = np.asarray(v, dtype=np.float64)

    def _v(self, t):
        """V at times ``t`` (tensor) -> fp64 tensor shaped like the reference's (leading 1).
        Raises ``ValueError`` outside the table like scipy's ``interp1d`` does."""
        tt = np.atleast_1d(t.detach().cpu().numpy().astype(np.float64))
        if np.any(tt < self._t_regular[0]) or np.any(tt > self._t_regular[-1]):
            raise ValueError('A value in t is outside the protocol table range.')
        idx = np.clip(np.searchsorted(self._t_regular, tt), 1, len(self._t_regular) - 1)
        x_lo, x_hi = self._t_regular[idx - 1], self._t_regular[idx]
        y_lo, y_hi = self._v_regular[idx - 1], self._v_regular[idx]
        out = (y_hi - y_lo) / (x_hi - x_lo) * (tt - x_lo) + y_lo
        return torch.from_numpy(out.reshape((1,) + tt.shape))

    def voltage(self, t):
        return self._v(t).numpy()

    def _v_or_holding(self, t):
        try:
            return self._v(t).reshape(-1)
        except ValueError:
            return torch.tensor([-80.], dtype=torch.float64)

    def _inactivation(self, r, v):
        k3 = self.p5 * torch.exp(self.p6 * v)
        k4 = self.p7 * torch.exp(-self.p8 * v)
        return -k3 * r + k4 * (1 - r)

    def _net_term(self, v, a):
        w = self.net[0].weight
        x = torch.stack([(v / 100.).expand_as(a), a], dim=-1).to(w)
        return (self.net(x) / self.netscale.to(w)).reshape(a.shape)


class ODEFuncNNf(_IKrBase):
    """NN-f: ``da/dt = net([V/100, a]) / 1000``; ``dr/dt = -k3 r + k4 (1 - r)``."""

    def __init__(self, arch=(5, 200), params='s', std=0.1):
        p = PARAMETER_SETS[params] if isinstance(params, str) else params
        super().__init__(arch, std, {k: p[k] for k in ('p5', 'p6', 'p7', 'p8')})

    def forward(self, t, y):
        a, r = torch.unbind(y, dim=1)
        v = self._v_or_holding(t).to(y.device)
        dadt = self._net_term(v, a).to(torch.float64)
        drdt = self._inactivation(r, v)
        return torch.stack([dadt, drdt], dim=1)


class ODEFuncNNd(_IKrBase):
    """NN-d: ``da/dt = k1 (1 - a) - k2 a + net([V/100, a]) / 1000``; same ``dr/dt``."""

    def __init__(self, arch=(5, 200), params='d', std=1e-3):
        p = PARAMETER_SETS[params] if isinstance(params, str) else params
        super().__init__(arch, std, {k: p[k] for k in ('p1', 'p2', 'p3', 'p4', 'p5', 'p6', 'p7',
                                                        'p8')})

    def _dadt(self, a, v):
        k1 = self.p1 * torch.exp(self.p2 * v)
        k2 = self.p3 * torch.exp(-self.p4 * v)
        return k1 * (1 - a) - k2 * a

    def _drdt(self, r, v):
        return self._inactivation(r, v)

    def forward(self, t, y):
        a, r = torch.unbind(y, dim=1)
        v = self._v_or_holding(t).to(y.device)
        dadt = self._dadt(a, v) + self._net_term(v, a).to(torch.float64)
        drdt = self._inactivation(r, v)
        return torch.stack([dadt, drdt], dim=1)


#: the name every reference script uses for the NN-f class
ODEFunc = ODEFuncNNf


def load_weights(func, path):
    """``func.load_state_dict(torch.load(path))`` accepting both ``model-state-dict.pt`` files and
    ``{epoch, state_dict, optimizer, loss}`` checkpoints (``train-r1.py:61-66``)."""
    blob = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(blob, dict) and 'state_dict' in blob:
        blob = blob['state_dict']
    func.load_state_dict(blob)
    func.eval()
    return func
