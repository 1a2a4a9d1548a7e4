"""CPU ORACLE (test infrastructure, NOT product code) -- RHS models of the hERG/IKr hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  The product package never does.

What is restated here (citations relative to ``/root/reference``):

* ``NNfRhs``   -- ``ODEFunc`` NN-f,  ``train-s1.py:181-247`` (= ``train-d1.py:191-256``,
                  ``train-r1.py:141-209``, ``table-1.py:210-278``)
* ``NNdRhs``   -- ``ODEFunc`` NN-d,  ``train-d2.py:191-272`` (= ``train-s2.py:180-259``,
                  ``train-r2.py:137-219``)
* ``HHRhs``    -- 2-state Hodgkin-Huxley ``Lambda``, ``train-s1.py:134-177``
* ``HHFitRhs`` -- the network-free HH candidate ``ODEFunc`` fitted by PINTS, ``train-d0.py:321-376``
* ``MarkovRhs``-- 6-state Markov ``Lambda`` (d-study ground truth), ``train-d1.py:134-187``
* ``build_mlp``-- architecture builder, ``train-r1-tune.py:155-163`` with
                  ``architectures/sNN.py:1-2`` hyper-parameters

The arithmetic is evaluated with the *same torch dtypes and promotions* the reference classes
produce (finding 5 of SURVEY.md): V(t) comes back fp64 from scipy, the HH part is fp64 (fp32 when
the out-of-table fallback ``tensor([-80])`` fires), the MLP is evaluated in fp32
(``.float()``) and the stacked result is fp64.  ``mlp_follows_state=True`` selects the
dtype-following variant needed for the fp64 1e-10 parity bar (the shipped class cannot run a
``.double()`` net because of its ``.float()`` cast).

The classes are pinned against the reference's own classes by ``tests/golden/make_golden.py``
(which ``exec``s the class source straight out of ``/root/reference`` at generation time) and
against the reference's logged losses (``s1/log2`` ...), see ``tests/test_oracle_kat.py``.
"""
import numpy as np
import torch
import torch.nn as nn
from scipy.interpolate import interp1d

# ---------------------------------------------------------------------------------------------
# Parameter sets (ms^-1 / mV^-1 units: the reference multiplies the published values by 1e-3)
# ---------------------------------------------------------------------------------------------
_B06 = [1.12592345582957387e-01, 8.26751134920666146e+01, 3.38768033864048357e-02,
        4.67106147665183542e+01, 8.47769667061995875e+01, 2.04001345352499328e+01,
        1.02860743916105211e+01, 2.78201179336874098e+01]

#: HH parameters p1..p8 of the synthetic "s" study, train-s1.py:139-146
HH_B06 = tuple(x * 1e-3 for x in _B06)

#: p5..p8 used by the "d" study NN-f / NN-d classes, train-d1.py:220-223 / train-d2.py:226-229
INACT_D = tuple(x * 1e-3 for x in (9.62243079990877703e+01, 2.26404683824047979e+01,
                                   8.00924780462999131e+00, 2.43749808069009823e+01))

#: 12 parameters of the 6-state Markov ground truth, train-d1.py:139-150
MARKOV_B06 = tuple(x * 1e-3 for x in (
    5.94625498751561316e-02, 1.21417701632850410e+02, 4.76436985414236425e+00,
    3.49383233960778904e-03, 9.62243079990877703e+01, 2.26404683824047979e+01,
    8.00924780462999131e+00, 2.43749808069009823e+01, 2.06822607368134157e+02,
    3.30791433507312362e+01, 1.26069071928587784e+00, 2.24844970727316245e+01))

#: (n_layers, n_nodes) of architectures/s00.py ... s11.py
ARCHITECTURES = {
    's00': (5, 200), 's01': (1, 200), 's02': (10, 200), 's03': (5, 10), 's04': (1, 10),
    's05': (10, 10), 's06': (5, 500), 's07': (1, 500), 's08': (10, 500), 's09': (5, 100),
    's10': (1, 100), 's11': (10, 100),
}


def build_mlp(n_layers=5, n_nodes=200, std=0.1, seed=None):
    """Linear(2,n) LeakyReLU [Linear(n,n) LeakyReLU]*n_layers Linear(n,1); N(0,std^2) weights,
    zero bias (train-r1-tune.py:155-163, init :165-168)."""
    if seed is not None:
        torch.manual_seed(seed)
    mods = [nn.Linear(2, n_nodes), nn.LeakyReLU()]
    for _ in range(n_layers):
        mods += [nn.Linear(n_nodes, n_nodes), nn.LeakyReLU()]
    mods.append(nn.Linear(n_nodes, 1))
    net = nn.Sequential(*mods)
    for m in net.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0, std=std)
            nn.init.constant_(m.bias, val=0)
    return net


class _ProtocolMixin:
    """Protocol table + linear interpolation, train-s1.py:218-229 (scipy ``interp1d`` with its
    default ``bounds_error`` -> ``ValueError`` outside the table)."""

    def set_fixed_form_voltage_protocol(self, t, v):
        self._t_regular = t
        self._v_regular = v
        self._interp = interp1d(t, v)

    def _v(self, t):
        return torch.from_numpy(self._interp([t.cpu().detach().numpy()]))

    def voltage(self, t):
        return self._v(t).numpy()

    def _v_or_holding(self, t):
        # train-s1.py:234-237: out-of-table time -> int64 tensor([-80])
        try:
            return self._v(t)
        except ValueError:
            return torch.tensor([-80])


class HHRhs(nn.Module, _ProtocolMixin):
    """2-state HH model [a, r]; returns shape (2,) like the reference Lambda."""

    def __init__(self, params=HH_B06):
        super().__init__()
        self.p = tuple(params)
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        a, r = torch.unbind(y[0])
        v = self._v_or_holding(t)
        p1, p2, p3, p4, p5, p6, p7, p8 = self.p
        k1 = p1 * torch.exp(p2 * v)
        k2 = p3 * torch.exp(-p4 * v)
        k3 = p5 * torch.exp(p6 * v)
        k4 = p7 * torch.exp(-p8 * v)
        dadt = k1 * (1. - a) - k2 * a
        drdt = -k3 * r + k4 * (1. - r)
        return torch.stack([dadt[0], drdt[0]])


class HHFitRhs(nn.Module, _ProtocolMixin):
    """HH candidate model of the CMA-ES fit (train-d0.py:321-376): ``set_parameters(x)`` sets
    p1..p4, p5..p8 are fixed; returns shape (1, 2); ``unity`` is an int64 tensor like the
    reference's, so ``unity - a`` follows the state dtype."""

    def __init__(self, inact=INACT_D):
        super().__init__()
        self.p1, self.p2, self.p3, self.p4 = 1.13e-4, 7.45e-2, 3.60e-5, 4.49e-2
        self.p5, self.p6, self.p7, self.p8 = inact
        self.unity = torch.tensor([1])
        self.nfe = 0

    def set_parameters(self, x):
        self.p1, self.p2, self.p3, self.p4 = x

    def forward(self, t, y):
        self.nfe += 1
        a, r = torch.unbind(y, dim=1)
        v = self._v_or_holding(t)
        k1 = self.p1 * torch.exp(self.p2 * v)
        k2 = self.p3 * torch.exp(-self.p4 * v)
        k3 = self.p5 * torch.exp(self.p6 * v)
        k4 = self.p7 * torch.exp(-self.p8 * v)
        dadt = k1 * (self.unity - a) - k2 * a
        drdt = -k3 * r + k4 * (self.unity - r)
        return torch.stack([dadt[0], drdt[0]]).reshape(1, -1)


class MarkovRhs(nn.Module, _ProtocolMixin):
    """6-state Markov model [c1, c2, i, ic1, ic2, o]; open probability is the last state."""

    def __init__(self, params=MARKOV_B06):
        super().__init__()
        self.p = tuple(params)
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        c1, c2, i, ic1, ic2, o = torch.unbind(y[0])
        v = self._v_or_holding(t)
        p = self.p
        a1 = p[0] * torch.exp(p[1] * v)
        b1 = p[2] * torch.exp(-p[3] * v)
        bh = p[4] * torch.exp(p[5] * v)
        ah = p[6] * torch.exp(-p[7] * v)
        a2 = p[8] * torch.exp(p[9] * v)
        b2 = p[10] * torch.exp(-p[11] * v)
        dc1 = a1 * c2 + ah * ic1 + b2 * o - (b1 + bh + a2) * c1
        dc2 = b1 * c1 + ah * ic2 - (a1 + bh) * c2
        di = a2 * ic1 + bh * o - (b2 + ah) * i
        dic1 = a1 * ic2 + bh * c1 + b2 * i - (b1 + ah + a2) * ic1
        dic2 = b1 * ic1 + bh * c2 - (ah + a1) * ic2
        do = a2 * c1 + ah * i - (b2 + bh) * o
        return torch.stack([dc1[0], dc2[0], di[0], dic1[0], dic2[0], do[0]])


class _NNRhsBase(nn.Module, _ProtocolMixin):
    def __init__(self, net, inact, mlp_follows_state):
        super().__init__()
        self.net = net
        self.vrange = torch.tensor([100.])
        self.netscale = torch.tensor([1000.])
        self.p5, self.p6, self.p7, self.p8 = inact
        self.unity = torch.tensor([1])
        self.mlp_follows_state = mlp_follows_state
        self.nfe = 0

    def _mlp_dtype(self):
        return next(self.net.parameters()).dtype

    def _net_term(self, nv, a):
        x = torch.stack([nv[0], a[0]])
        if self.mlp_follows_state:
            # dtype-following variant: evaluate in the net's own dtype, scale in that dtype
            x = x.to(self._mlp_dtype())
            return self.net(x) / self.netscale.to(x.dtype)
        return self.net(x.float()) / self.netscale          # train-s1.py:245

    def _drdt(self, r, v):
        k3 = self.p5 * torch.exp(self.p6 * v)
        k4 = self.p7 * torch.exp(-self.p8 * v)
        return -k3 * r + k4 * (self.unity - r)


class NNfRhs(_NNRhsBase):
    """NN-f: da/dt = net([V/100, a]) / 1000, dr/dt HH (train-s1.py:231-247)."""

    def __init__(self, net=None, inact=HH_B06[4:], mlp_follows_state=False):
        super().__init__(net if net is not None else build_mlp(), inact, mlp_follows_state)

    def forward(self, t, y):
        self.nfe += 1
        a, r = torch.unbind(y, dim=1)
        v = self._v_or_holding(t)
        nv = v / self.vrange
        drdt = self._drdt(r, v)
        dadt = self._net_term(nv, a)
        return torch.stack([dadt[0], drdt[0]]).reshape(1, -1)


class NNdRhs(_NNRhsBase):
    """NN-d: da/dt = k1 (1-a) - k2 a + net([V/100, a]) / 1000 (train-d2.py:247-272)."""

    def __init__(self, net=None, act=HH_B06[:4], inact=INACT_D, mlp_follows_state=False):
        super().__init__(net if net is not None else build_mlp(std=1e-3), inact,
                         mlp_follows_state)
        self.p1, self.p2, self.p3, self.p4 = act

    def _dadt(self, a, v):
        k1 = self.p1 * torch.exp(self.p2 * v)
        k2 = self.p3 * torch.exp(-self.p4 * v)
        return k1 * (self.unity - a) - k2 * a

    def forward(self, t, y):
        self.nfe += 1
        a, r = torch.unbind(y, dim=1)
        v = self._v_or_holding(t)
        nv = v / self.vrange
        drdt = self._drdt(r, v)
        dadt = self._dadt(a, v).reshape(-1)
        ddadt = self._net_term(nv, a)
        dadt = dadt + ddadt.reshape(-1)       # reference: in-place "+=" (fp64 += fp32)
        return torch.stack([dadt[0], drdt[0]]).reshape(1, -1)


def load_state_dict_file(module, path):
    """Load ``model-state-dict.pt`` or a ``{epoch,state_dict,optimizer,loss}`` checkpoint
    (train-s1.py:263, table-2.py:314-319)."""
    blob = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(blob, dict) and 'state_dict' in blob:
        blob = blob['state_dict']
    module.load_state_dict(blob)
    module.eval()
    return module


def observe_current(func, y, t, g=1.0, e=-86.0):
    """I = g a r (V - E), train-s1.py:328 / table-1.py:414."""
    return g * y[:, 0, 0] * y[:, 0, 1] * (func._v(t) - e)


def observe_open_current(func, y, t, e=-86.0):
    """Markov ground truth: I = O (V - E), train-d1.py:299."""
    return y[:, 0, -1] * (func._v(t) - e)
