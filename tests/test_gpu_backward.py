"""GPU parity tests of the backward path (adjoint sweep + weight-gradient GEMM, through the C ABI)
against PyTorch autograd through the CPU oracle (restated torchdiffeq dopri5 + reference RHS).

The oracle replays the accepted steps the CUDA run took (``options={'replay': ...}``, see
oracle/ref_odeint.py): the step-size controller is ill-conditioned with respect to rounding, so two
correct implementations drift apart in dt by ~1e-7 relative after a stiff phase; with the steps
pinned, the gradients must agree to rounding.

Tolerances: fp64 state + fp64 MLP: 1e-8 relative to the largest gradient entry (north star: equal
loss / gradients to 1e-8 relative); fp32 "as shipped": 2e-3 (fp32 accumulation noise).
"""
import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_odeint as ro
from tests import kat
from tests.test_gpu_forward import _nn

pytestmark = pytest.mark.gpu


def _flat(grads):
    return torch.cat([g.reshape(-1).double().cpu() for g in grads]).numpy()


def _oracle_grads(ofunc, y0, t, steps_per_lane, first_step, loss_fn):
    """sum over lanes of d loss_fn(b, y_b)/d theta by autograd through the oracle, replaying the
    accepted steps of each lane; returns (flat param grad, grad_y0 (B,2), losses)."""
    params = list(ofunc.net.parameters())
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    gy0, losses = [], []
    for b in range(y0.shape[0]):
        yb = y0[b:b + 1].clone().requires_grad_(True)
        y = ro.odeint(ofunc, yb, t, options={'first_step': first_step,
                                             'replay': steps_per_lane[b]})
        lb = loss_fn(b, y[:, 0, :])
        lb.backward()
        gy0.append(yb.grad.reshape(-1).double())
        losses.append(float(lb.detach()))
    g = torch.cat([p.grad.reshape(-1).double() for p in params]).numpy()
    for p in params:
        p.requires_grad_(False)
        p.grad = None
    return g, torch.stack(gy0).numpy(), np.array(losses)


def _steps(res):
    ck_t = res.ckpt[0].cpu().numpy()
    st = res.stats.cpu().numpy()
    return [[(float(ck_t[j, b, 0]), float(ck_t[j, b, 1])) for j in range(st[b, 0])]
            for b in range(st.shape[0])]


def _setup(study, double, t_end=60., n_out=31):
    torch.set_num_threads(1)
    func, ofunc = _nn(study, double=double)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    dt = torch.float64 if double else torch.float32
    t = torch.linspace(0., t_end, n_out, dtype=dt)
    return func, ofunc, t


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_autograd_through_odeint_fp64_matches_oracle(study):
    func, ofunc, t = _setup(study, True)
    y0 = torch.tensor([[0.02, 0.97], [0.0, 1.0], [0.3, 0.6]], dtype=torch.float64)
    rng = np.random.RandomState(4)
    w = torch.from_numpy(rng.randn(len(t), 3, 2))
    func.cuda()
    for p in func.net.parameters():
        p.requires_grad_(True)
    y0g = y0.cuda().requires_grad_(True)
    y = ikr.odeint(func, y0g, t, options={'first_step': 0.05})
    (y * w.cuda()).sum().backward()
    got = _flat([p.grad for p in func.net.parameters()])
    got_y0 = y0g.grad.cpu().numpy()
    # replay the accepted steps on the oracle
    with torch.no_grad():
        res = ikr.integrate(func, y0.cuda(), t, options={'first_step': 0.05}, want_ckpt=True)
    want, want_y0, _ = _oracle_grads(ofunc, y0, t, _steps(res), 0.05,
                                     lambda b, yb: (yb * w[:, b, :]).sum())
    assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()
    assert np.abs(got_y0 - want_y0).max() <= 1e-8 * np.abs(want_y0).max()


def test_fused_sse_loss_and_grad_fp64_matches_oracle():
    func, ofunc, t = _setup('d2', True)
    B = 5
    rng = np.random.RandomState(5)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.9, 1, B)], 1))
    g = torch.tensor(rng.lognormal(0, 0.2, B))
    data = torch.from_numpy(rng.randn(len(t), B) * 0.1)
    func.cuda()
    total, per, grads, res = ikr.loss_and_grad(func, y0.cuda(), t, data, g=g, E=-86.0,
                                               options={'first_step': 0.05}, want_y0=True,
                                               want_g=True)
    v = torch.from_numpy(np.interp(t.numpy(), *protocols.ap2hz()))

    def loss_fn(b, yb):
        cur = g[b] * yb[:, 0] * yb[:, 1] * (v + 86.0)
        return ((cur - data[:, b]) ** 2).sum()

    want, want_y0, want_l = _oracle_grads(ofunc, y0, t, _steps(res), 0.05, loss_fn)
    assert np.abs(per.cpu().numpy() - want_l).max() <= 1e-8 * np.abs(want_l).max()
    assert abs(float(total) - want_l.sum()) <= 1e-8 * want_l.sum()
    assert np.abs(_flat(grads) - want).max() <= 1e-8 * np.abs(want).max()
    assert np.abs(res.grad_y0.cpu().numpy() - want_y0).max() <= 1e-8 * np.abs(want_y0).max()
    # dL/dg by finite differences of the fused loss itself (g enters the loss linearly per sample)
    eps = 1e-6
    with torch.no_grad():
        lp = ikr.integrate(func, y0.cuda(), t, g=g + eps, data=data, want_y=False,
                           options={'first_step': 0.05}).sse.cpu().numpy()
        lm = ikr.integrate(func, y0.cuda(), t, g=g - eps, data=data, want_y=False,
                           options={'first_step': 0.05}).sse.cpu().numpy()
    fd = (lp - lm) / (2 * eps)
    assert np.abs(res.grad_g.cpu().numpy() - fd).max() <= 1e-5 * np.abs(fd).max()


def test_fp32_as_shipped_gradient_envelope():
    func, ofunc, t = _setup('d1', False, t_end=40., n_out=21)
    y0 = torch.tensor([[0.02, 0.97], [0.0, 1.0]])
    rng = np.random.RandomState(6)
    data = torch.from_numpy((rng.randn(len(t), 1) * 0.1).astype(np.float32))
    func.cuda()
    total, per, grads, res = ikr.loss_and_grad(func, y0.cuda(), t, data, loss='sae',
                                               options={'first_step': 0.05}, want_y0=True)
    v = torch.from_numpy(np.interp(t.double().numpy(), *protocols.ap2hz()))

    def loss_fn(b, yb):
        cur = (yb[:, 0] * yb[:, 1]).double() * (v + 86.0)
        return (cur - data[:, 0].double()).abs().sum()

    want, want_y0, want_l = _oracle_grads(ofunc, y0, t, _steps(res), 0.05, loss_fn)
    assert np.abs(per.cpu().numpy() - want_l).max() <= 1e-4 * np.abs(want_l).max()
    assert np.abs(_flat(grads) - want).max() <= 2e-3 * np.abs(want).max()
    assert np.abs(res.grad_y0.cpu().numpy() - want_y0).max() <= 2e-3 * np.abs(want_y0).max()


def test_gradient_is_round_and_tile_invariant():
    """Many trajectories (several tiles, ragged last tile); one reversed step per round versus
    the default round size; two tile sizes: same gradient up to summation order."""
    func, _, t = _setup('s1', True, t_end=30., n_out=16)
    B = 77
    rng = np.random.RandomState(7)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.9, 1, B)], 1))
    data = torch.from_numpy(rng.randn(len(t), B) * 0.1)
    func.cuda()
    ref = None
    for opts, ws in (({}, None), ({'tile_m': 16}, None), ({'tile_m': 32}, 'small')):
        o = dict(opts, first_step=0.05)
        kw = {}
        if ws == 'small':
            # room for exactly one reversed step per round
            import ctypes
            from neural_ode_ion_channels_b200 import _cabi
            res0 = ikr.integrate(func, y0.cuda(), t, options=o, data=data, want_ckpt=True)
            full = _cabi.lib().ikr_workspace_bytes(ctypes.byref(res0._desc), 1, B, 1)
            kw['workspace_bytes'] = full // 3
        total, per, grads, res = ikr.loss_and_grad(func, y0.cuda(), t, data, options=o, **kw)
        flat = _flat(grads)
        if ref is None:
            ref = (flat, per.cpu().numpy())
            assert np.isfinite(flat).all() and np.abs(flat).max() > 0
        else:
            assert np.abs(flat - ref[0]).max() <= 1e-10 * np.abs(ref[0]).max()
            assert np.abs(per.cpu().numpy() - ref[1]).max() <= 1e-12 * np.abs(ref[1]).max()


@pytest.mark.parametrize('arch', ['s03', 's10', 's01'])
def test_gradient_fp64_architectures(arch):
    torch.set_num_threads(1)
    torch.manual_seed(0)
    func = ikr.ODEFuncNNf(arch=arch, params='r').double()
    from oracle import ref_models as rm
    L, n = ikr.ARCHITECTURES[arch]
    ofunc = rm.NNfRhs(net=rm.build_mlp(L, n),
                      inact=tuple(func.__dict__['p%d' % i] for i in (5, 6, 7, 8)),
                      mlp_follows_state=True).double()
    ofunc.net.load_state_dict(func.net.state_dict())
    ofunc.vrange = ofunc.vrange.double()
    ofunc.netscale = ofunc.netscale.double()
    t_tab, v_tab = protocols.pr3_activation(20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 1200., 25, dtype=torch.float64)
    y0 = torch.tensor([[0.01, 0.9], [0.1, 0.5]], dtype=torch.float64)
    w = torch.from_numpy(np.random.RandomState(8).randn(len(t), 2, 2))
    func.cuda()
    for p in func.net.parameters():
        p.requires_grad_(True)
    y = ikr.odeint(func, y0.cuda(), t, options={'first_step': 0.05})
    (y * w.cuda()).sum().backward()
    got = _flat([p.grad for p in func.net.parameters()])
    with torch.no_grad():
        res = ikr.integrate(func, y0.cuda(), t, options={'first_step': 0.05}, want_ckpt=True)
    want, _, _ = _oracle_grads(ofunc, y0, t, _steps(res), 0.05,
                               lambda b, yb: (yb * w[:, b, :]).sum())
    assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()


def test_gradient_is_additive_over_batch_shards():
    """Multi-GPU contract checked on one device: the gradient of the whole batch equals
    the sum of the gradients of its shards (what the flat all-reduce adds up), and so does the
    loss -- independent of how `parallel.shard_bounds` cuts the batch."""
    from neural_ode_ion_channels_b200 import parallel
    func, _, t = _setup('d2', True, t_end=30., n_out=16)
    B = 13
    rng = np.random.RandomState(9)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.9, 1, B)], 1))
    g = torch.tensor(rng.lognormal(0, 0.2, B))
    data = torch.from_numpy(rng.randn(len(t), B) * 0.1)
    func.cuda()
    opts = {'first_step': 0.05}
    total, _, grads, _ = ikr.loss_and_grad(func, y0.cuda(), t, data, g=g, options=opts)
    whole = _flat(grads)
    for world in (2, 4):
        acc, loss = np.zeros_like(whole), 0.0
        for rank in range(world):
            lo, hi = parallel.shard_bounds(B, world, rank)
            tl, _, gr, _ = ikr.loss_and_grad(func, y0[lo:hi].cuda(), t, data[:, lo:hi], g=g[lo:hi],
                                             options=opts)
            acc += _flat(gr)
            loss += float(tl)
        assert np.abs(acc - whole).max() <= 1e-10 * np.abs(whole).max()
        assert abs(loss - float(total)) <= 1e-12 * abs(float(total))


def test_full_size_training_batch_4096_properties():
    """BASELINE configs[2] size (4,096 datasets) on a short window: statuses ok, gradient finite,
    and equal to 16 x the gradient of the 256 distinct datasets it is made of (fp32 as shipped,
    so equality up to fp32 / summation-order noise)."""
    func, _, t = _setup('d2', False, t_end=40., n_out=21)
    rng = np.random.RandomState(10)
    y0s = np.stack([rng.uniform(0, 0.05, 256), rng.uniform(0.9, 1, 256)], 1).astype(np.float32)
    ds = (rng.randn(len(t), 256) * 0.1).astype(np.float32)
    idx = rng.permutation(np.arange(4096) % 256)
    func.cuda()
    tot_b, per_b, g_b, res_b = ikr.loss_and_grad(func, torch.from_numpy(y0s[idx]).cuda(), t,
                                                 torch.from_numpy(ds[:, idx]))
    tot_s, per_s, g_s, res_s = ikr.loss_and_grad(func, torch.from_numpy(y0s).cuda(), t,
                                                 torch.from_numpy(ds))
    assert int((res_b.stats[:, 3] != 0).sum()) == 0
    assert torch.equal(per_b, per_s[torch.from_numpy(idx).cuda()])
    fb, fs = _flat(g_b), _flat(g_s)
    assert np.isfinite(fb).all()
    assert np.abs(fb - 16 * fs).max() <= 1e-4 * np.abs(fb).max()


# ---------------------------------------------------------------------------------------------
# rk4 (3/8 rule, fixed grid) backward -- VERDICT r1 missing #4
# ---------------------------------------------------------------------------------------------
def _oracle_grads_rk4(ofunc, y0, t, loss_fn, options=None):
    params = list(ofunc.net.parameters())
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    gy0, losses = [], []
    for b in range(y0.shape[0]):
        yb = y0[b:b + 1].clone().requires_grad_(True)
        y = ro.odeint(ofunc, yb, t, method='rk4', options=options)
        lb = loss_fn(b, y[:, 0, :])
        lb.backward()
        gy0.append(yb.grad.reshape(-1).double())
        losses.append(float(lb.detach()))
    g = torch.cat([p.grad.reshape(-1).double() for p in params]).numpy()
    for p in params:
        p.requires_grad_(False)
        p.grad = None
    return g, torch.stack(gy0).numpy(), np.array(losses)


@pytest.mark.parametrize('study,step_size', [('s1', None), ('d2', None), ('d2', 0.7)])
def test_rk4_autograd_fp64_matches_oracle(study, step_size):
    """`odeint(..., method='rk4')` with parameters that require grad: gradient w.r.t. the MLP
    parameters and y0 against PyTorch autograd through the oracle's fixed-grid solver, fp64 state +
    fp64 MLP, <= 1e-8 of the largest entry; with `step_size` the outputs are interpolated between
    grid points (their adjoints split between the two ends of the step)."""
    func, ofunc, _ = _setup(study, True)
    t = torch.linspace(0., 60., 31, dtype=torch.float64) if step_size is None else \
        torch.linspace(0., 21., 8, dtype=torch.float64)
    opts = None if step_size is None else {'step_size': step_size}
    y0 = torch.tensor([[0.02, 0.97], [0.0, 1.0], [0.3, 0.6]], dtype=torch.float64)
    rng = np.random.RandomState(14)
    w = torch.from_numpy(rng.randn(len(t), 3, 2))
    func.cuda()
    for p in func.net.parameters():
        p.requires_grad_(True)
    y0g = y0.cuda().requires_grad_(True)
    y = ikr.odeint(func, y0g, t, method='rk4', options=opts)
    (y * w.cuda()).sum().backward()
    got = _flat([p.grad for p in func.net.parameters()])
    got_y0 = y0g.grad.cpu().numpy()
    want, want_y0, _ = _oracle_grads_rk4(ofunc, y0, t, lambda b, yb: (yb * w[:, b, :]).sum(), opts)
    assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()
    assert np.abs(got_y0 - want_y0).max() <= 1e-8 * np.abs(want_y0).max()


def test_rk4_fused_loss_fp32_tensor_core_backward():
    """rk4 training step on the tensor cores (fp32 as shipped, fused SSE loss, B = 40): loss and
    gradient against oracle autograd, 1e-4 of each parameter block's largest entry."""
    func, ofunc, _ = _setup('d2', False)
    t = torch.linspace(0., 40., 41)
    B = 40
    rng = np.random.RandomState(15)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.9, 1, B)], 1), dtype=torch.float32)
    data = torch.from_numpy((rng.randn(len(t), B) * 0.1).astype(np.float32))
    func.cuda()
    total, per, grads, res = ikr.loss_and_grad(func, y0.cuda(), t, data, method='rk4', want_y0=True)
    assert res.geometry['tensor_cores']
    v = torch.from_numpy(np.interp(t.double().numpy(), *protocols.ap2hz()))

    def loss_fn(b, yb):
        cur = (yb[:, 0] * yb[:, 1]).double() * (v + 86.0)
        return ((cur - data[:, b].double()) ** 2).sum()

    want, want_y0, want_l = _oracle_grads_rk4(ofunc, y0, t, loss_fn)
    got = _flat(grads)
    assert np.abs(per.cpu().numpy() - want_l).max() <= 1e-5 * np.abs(want_l).max()
    n, o = 200, 0
    for size in [2 * n, n] + [n * n, n] * 5 + [n, 1]:
        ref = np.abs(want[o:o + size]).max()
        assert np.abs(got[o:o + size] - want[o:o + size]).max() <= 1e-4 * ref
        o += size
    assert np.abs(res.grad_y0.cpu().numpy() - want_y0).max() <= 1e-4 * np.abs(want_y0).max()


def test_concurrent_fits_on_separate_streams_equal_serial_fits():
    """bench.py's `sweep` leg runs the fits of a rank concurrently, one host thread and one CUDA
    stream each (a 256-dataset fit is a latency-bound chain on 2 of 148 SMs).  The library keeps no
    global state, so every fit must return what it returns alone: same loss and accepted steps,
    gradients equal to the summation-order noise of the weight-gradient partials.  Covers the
    three kernel families (tcgen05 fp32, FFMA2 n = 500, DFMA fp64)."""
    import threading
    name, t_tab, v_tab, t_out = protocols.protocol_set('pr4')[10]
    t_out = t_out[:60]
    rng = np.random.RandomState(5)
    B = 200
    y0np = np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1.0, B)], 1)
    noise = rng.normal(0, 0.05, (len(t_out), B))
    cases = [('s00', False), ('s06', False), ('s03', True), ('s10', True), ('s02', False), ('s00', True)]

    def build(arch, f64):
        torch.manual_seed(0)
        func = ikr.ODEFuncNNf(arch=arch, params='r')
        return (func.double() if f64 else func).cuda()

    def fit(func, f64, out, k):
        dtype = torch.float64 if f64 else torch.float32
        func.set_fixed_form_voltage_protocol(t_tab, v_tab)
        t = torch.tensor(t_out, dtype=dtype)
        y0 = torch.tensor(y0np, dtype=dtype).cuda()
        data = torch.tensor(0.01 * noise, dtype=dtype).cuda()
        total, per, grads, res = ikr.loss_and_grad(func, y0, t, data, g=0.16, E=-93.4,
                                                   options={'ckpt_cap': 512, 'stash_gib': 2})
        out[k] = (float(total), res.stats[:, :2].cpu().numpy().copy(), _flat(grads))

    serial = [None] * len(cases)
    for k, (arch, f64) in enumerate(cases):
        fit(build(arch, f64), f64, serial, k)
    torch.cuda.synchronize()

    funcs = [build(arch, f64) for arch, f64 in cases]
    conc = [None] * len(cases)
    errors = []
    gate = threading.Barrier(len(cases), timeout=300)

    def worker(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):
                gate.wait()
                for _ in range(2):            # two rounds each, so that the launches interleave
                    fit(funcs[k], cases[k][1], conc, k)
                torch.cuda.current_stream().synchronize()
        except Exception as exc:              # noqa: BLE001 -- reported below, never swallowed
            errors.append((cases[k], repr(exc)))
            gate.abort()
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(cases))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for (arch, f64), a, b in zip(cases, serial, conc):
        tol = 1e-10 if f64 else 2e-5
        assert (a[1] == b[1]).all(), (arch, f64)                       # same accepted / rejected steps
        assert abs(a[0] - b[0]) <= tol * abs(a[0]), (arch, f64, a[0], b[0])
        assert np.abs(a[2] - b[2]).max() <= tol * np.abs(a[2]).max(), (arch, f64)
