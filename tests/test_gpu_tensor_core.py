"""Parity of the tcgen05 forward kernel (bf16x3 split MMAs, activations and accumulators in TMEM):
against the CPU oracle on fixed-step rk4 (no accept/reject decisions => a clean arithmetic check),
against the FFMA2 kernel on the same inputs, and against the reference's logged losses."""
import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import protocols
from oracle import ref_odeint as ro
from tests import kat

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _pair(study):
    """(product module, oracle module) with identical weights."""
    cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
    pset = 's' if study in ('s1', 's2') else 'd'
    func = ikr.load_weights(cls(params=pset), kat.weights_path(study))
    return func, kat.make_nn(study)


@pytest.mark.parametrize('study', ['s1', 'd2'])
def test_tc_rk4_matches_oracle_fp32(study):
    """fp32 state + fp32 MLP, fixed grid: every RHS evaluation of the tensor-core path must agree
    with the oracle's fp32 torch MLP to fp32 rounding (tolerance 2e-6 abs on a, r in [0, 1])."""
    func, ofunc = _pair(study)
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    ofunc.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 120., 241)
    y0 = torch.tensor([[0.01, 0.97], [0.0, 1.0], [0.04, 0.95]], dtype=torch.float32)
    res = ikr.integrate(func, y0.cuda(), t, method='rk4')
    assert res.geometry['tensor_cores'] and res.geometry['threads'] in (192, 320, 448)
    with torch.no_grad():
        for b in range(3):
            want = ro.odeint(ofunc, y0[b:b + 1], t, method='rk4')
            err = (res.y[:, b:b + 1].cpu() - want).abs().max().item()
            assert err < 2e-6, (study, b, err)


def test_tc_matches_ffma_kernel_and_fp64_state_mode():
    func, _ = _pair('d1')
    t_tab, v_tab = protocols.pr4_inactivation_standin(-20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 300., 151)
    rng = np.random.RandomState(7)
    B = 300
    y0 = np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1)
    # 150 rk4 steps: fp32 rounding of the two MLP arithmetics accumulates to a few 1e-6
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-5)):
        y = torch.tensor(y0, dtype=dtype).cuda()
        tt = t.to(dtype)
        a = ikr.integrate(func, y, tt, method='rk4')
        b = ikr.integrate(func, y, tt, method='rk4', options={'tensor_cores': False})
        assert a.geometry['tensor_cores'] and not b.geometry['tensor_cores']
        assert (a.y - b.y).abs().max().item() < tol
        # adaptive: same solver, decisions may flip at knife-edge ratios => solver-level agreement
        c = ikr.integrate(func, y, tt)
        d = ikr.integrate(func, y, tt, options={'tensor_cores': False})
        assert (c.y - d.y).abs().max().item() < 5e-4
        assert abs(int(c.stats[:, 2].sum()) - int(d.stats[:, 2].sum())) < 0.02 * int(d.stats[:, 2].sum())


def test_tc_logged_loss_and_narrow_architectures():
    # the reference's logged AP-2Hz MAE (s1/log2:4) through the tensor-core path
    func, _ = _pair('s1')
    row = kat.KAT['s1'][0]
    t_tab, v_tab, t_out = kat.row_protocol(row)
    i_gt = kat.gt_current('s1', t_tab, v_tab, t_out).reshape(-1)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    res = ikr.integrate(func, torch.tensor([[0., 1.]]).cuda(), t_out, data=i_gt.float(), E=-86.0,
                        want_y=False)
    assert res.geometry['tensor_cores']
    assert abs(float(res.sae[0]) / len(t_out) - row['loss']) < 5e-5
    # widths with other k-step / tail geometries: n = 100 (tail of 4), 72 (tail of 8), 64 (no tail),
    # 48 (odd number of k-steps), 90 (zero-padded k-step)
    t_tab, v_tab = protocols.ap2hz()
    t = torch.linspace(0., 60., 61)
    y0 = torch.tensor([[0.01, 0.97]] * 3, dtype=torch.float32).cuda()
    # 32 / 16: a single unit per layer pass
    for n, L in ((100, 5), (72, 2), (64, 1), (48, 3), (90, 2), (200, 1), (32, 2), (16, 1), (24, 3)):
        torch.manual_seed(n)
        f = ikr.ODEFuncNNf(arch=(L, n))
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
        a = ikr.integrate(f, y0, t, method='rk4')
        b = ikr.integrate(f, y0, t, method='rk4', options={'tensor_cores': False})
        assert a.geometry['tensor_cores'], n
        assert (a.y - b.y).abs().max().item() < 5e-6, (n, L)


# ---------------------------------------------------------------------------------------------
# tensor-core backward (adjoint MMAs + weight-gradient GEMM over the bf16x2 stash)
# ---------------------------------------------------------------------------------------------
def _flat(grads):
    return torch.cat([g.reshape(-1).double() for g in grads]).cpu().numpy()


def _segments(n=200, L=5):
    segs, o = [('w0', 0, 2 * n), ('b0', 2 * n, 3 * n)], 3 * n
    for l in range(L):
        segs += [('W%d' % (l + 1), o, o + n * n), ('b%d' % (l + 1), o + n * n, o + n * n + n)]
        o += n * n + n
    return segs + [('w_last', o, o + n), ('b_last', o + n, o + n + 1)]


@pytest.mark.parametrize('study,B', [('d2', 6), ('d1', 300)])
def test_tc_backward_matches_ffma_backward(study, B):
    """Same inputs through the tensor-core and the FFMA2 training step (fp32 as shipped): every
    parameter block of the gradient, grad_y0 and the loss agree at the fp32 noise level of the
    adaptive forward (the two forwards accept slightly different step sequences, so 1e-2 of the
    block maximum; test_gpu_backward.py holds the tensor-core path to 2e-3 against the oracle with the
    accepted steps replayed).  B = 300 spans several tiles."""
    func, _ = _pair(study)
    func.cuda()
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 80., 41)
    rng = np.random.RandomState(8)
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    data = torch.from_numpy((rng.randn(len(t), B) * 0.1).astype(np.float32))
    out = {}
    with torch.enable_grad():
        for tcore in (True, False):
            total, per, grads, res = ikr.loss_and_grad(
                func, y0, t, data, want_y0=True, options={'tensor_cores': tcore, 'first_step': 0.05})
            assert res.geometry['tensor_cores'] == tcore
            out[tcore] = (_flat(grads), res.grad_y0.cpu().double().numpy(), float(total))
    a, b = out[True], out[False]
    assert np.isfinite(a[0]).all() and np.abs(a[0]).max() > 0
    for name, lo, hi in _segments():
        ref = np.abs(b[0][lo:hi]).max()
        assert np.abs(a[0][lo:hi] - b[0][lo:hi]).max() <= 1e-2 * ref, (name, ref)
    assert np.abs(a[1] - b[1]).max() <= 1e-2 * np.abs(b[1]).max()
    assert abs(a[2] - b[2]) <= 1e-4 * abs(b[2])


@pytest.mark.parametrize('arch', [(2, 72), (3, 100), (1, 200), (7, 200), (1, 64)])
def test_tc_backward_other_architectures(arch):
    """Widths / depths other than s00 through the tensor-core training step (n % 16 in 1..8: the
    constant-1 bias column rides in the tail group; (1, 64) has no tail and must take the FFMA2
    backward on its own): gradient blocks agree with the FFMA2 step at the same 1e-2 of the block
    maximum as above."""
    L, n = arch
    torch.manual_seed(100 * L + n)
    func = ikr.ODEFuncNNf(arch=(L, n)).cuda()
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 60., 31)
    rng = np.random.RandomState(L + n)
    B = 40
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    data = torch.from_numpy((rng.randn(len(t), B) * 0.1).astype(np.float32))
    out = {}
    with torch.enable_grad():
        for tcore in (True, False):
            total, per, grads, res = ikr.loss_and_grad(
                func, y0, t, data, want_y0=True, options={'tensor_cores': tcore, 'first_step': 0.05})
            out[tcore] = (_flat(grads), res.grad_y0.cpu().double().numpy(), float(total))
    a, b = out[True], out[False]
    assert np.isfinite(a[0]).all() and np.abs(a[0]).max() > 0
    for name, lo, hi in _segments(n, L):
        ref = np.abs(b[0][lo:hi]).max()
        assert np.abs(a[0][lo:hi] - b[0][lo:hi]).max() <= 1e-2 * ref, (arch, name, ref)
    assert np.abs(a[1] - b[1]).max() <= 1e-2 * np.abs(b[1]).max()
    assert abs(a[2] - b[2]) <= 1e-4 * abs(b[2])


def test_tc_backward_round_and_shard_invariance():
    """One reversed step per round (small workspace) versus the default round size, and the sum of
    two batch shards versus the whole batch: same gradient up to fp32 summation order."""
    import ctypes
    from neural_ode_ion_channels_b200 import _cabi
    func, _ = _pair('d2')
    func.cuda()
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 40., 21)
    rng = np.random.RandomState(9)
    B = 150
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    data = torch.from_numpy((rng.randn(len(t), B) * 0.1).astype(np.float32))
    opts = {'first_step': 0.05}
    with torch.enable_grad():
        total, per, grads, res = ikr.loss_and_grad(func, y0, t, data, options=opts)
        ref = _flat(grads)
        full = _cabi.lib().ikr_workspace_bytes(ctypes.byref(res._desc), 1, B, 1)
        # a quarter of the default stash: about four times as many (shorter) rounds
        from neural_ode_ion_channels_b200.adjoint import loss_and_grad
        total2, per2, grads2, _ = loss_and_grad(func, y0, t, data, options=opts,
                                                workspace_bytes=full // 4)
        assert torch.equal(per, per2)
        assert np.abs(_flat(grads2) - ref).max() <= 2e-5 * np.abs(ref).max()
        h = B // 2
        ta, pa, ga, _ = ikr.loss_and_grad(func, y0[:h], t, data[:, :h], options=opts)
        tb, pb, gb, _ = ikr.loss_and_grad(func, y0[h:], t, data[:, h:], options=opts)
        assert torch.equal(torch.cat([pa, pb]), per)
        assert np.abs(_flat(ga) + _flat(gb) - ref).max() <= 2e-5 * np.abs(ref).max()


def test_tc_lane_pool_kernel_equals_tile_scheduled_kernel():
    """The tensor-core lane-pool kernel (slots refill from one trajectory queue, jobs mixed inside a
    CTA) reproduces the tile-scheduled tensor-core kernel bit for bit, checkpoints included."""
    func, _ = _pair('d1')
    rng = np.random.RandomState(11)
    jobs = []
    for k, (tt, vv, T) in enumerate([(*protocols.ap2hz(), 101), (*protocols.pr3_activation(20), 81),
                                     (*protocols.pr5_deactivation(-60), 41)]):
        B = (37, 300, 5)[k]
        y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                          dtype=torch.float32).cuda()
        t = torch.linspace(0., 2. * (T - 1), T)
        g = torch.tensor(rng.lognormal(0, 0.2, B), dtype=torch.float32)
        jobs.append(dict(protocol=(tt, vv), y0=y0, t=t, g=g, data=torch.zeros(T),
                         want_current=True, want_ckpt=True))
    a = ikr.integrate_many(func, jobs)
    b = ikr.integrate_many(func, jobs, options={'lane_pool': True})
    assert a[0].geometry['tensor_cores'] and b[0].geometry['tensor_cores']
    for ra, rb in zip(a, b):
        assert torch.equal(ra.stats, rb.stats)
        assert torch.equal(ra.y, rb.y) and torch.equal(ra.current, rb.current)
        assert torch.equal(ra.sse, rb.sse) and torch.equal(ra.sae, rb.sae)
        for bb in range(ra.stats.shape[0]):
            k = int(ra.stats[bb, 0])
            assert torch.equal(ra.ckpt[1][:k, bb], rb.ckpt[1][:k, bb])
            assert torch.equal(ra.ckpt[0][:k, bb], rb.ckpt[0][:k, bb])


def test_fp16x2_split_survives_wild_trial_steps():
    """The default operand split of the tensor-core forward (fp16x2, three MMAs per fp32 product)
    survives activations beyond the fp16 range: a first step of 2,000 ms throws the stage states far outside
    [0, 1] (hidden activations beyond the fp16 range); the reference computes finite garbage there
    and rejects the step, and so must this path -- every lane ends with status ok and the traces
    agree with the bf16x3 split and the FFMA2 kernel at the solver's fp32 noise level."""
    func, _ = _pair('d1')
    t_tab, v_tab = protocols.pr3_activation(20)
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 8000., 801)
    rng = np.random.RandomState(12)
    B = 256
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    out = {}
    for name, o in (('fp16x2', {'tc_split': 'fp16x2'}), ('bf16x3', {'tc_split': 'bf16x3'}),
                    ('ffma', {'tensor_cores': False})):
        r = ikr.integrate(func, y0, t, options=dict(o, first_step=2000.0, check_status=False))
        assert int((r.stats[:, 3] != 0).sum()) == 0, name
        assert int(r.stats[:, 1].min()) >= 1          # the wild first step was rejected
        out[name] = r.y
    assert out['fp16x2'].isfinite().all()
    assert (out['fp16x2'] - out['ffma']).abs().max().item() < 5e-4
    assert (out['bf16x3'] - out['ffma']).abs().max().item() < 5e-4
    g = ikr.integrate(func, y0[:4], t).geometry
    assert g['mma_split'] == 'fp16x2 split' and g['mma_products'] == 3


@pytest.mark.parametrize('sizes', [(37, 300, 5), (130,), (1, 1), (700, 513)])
def test_ping_pong_kernel_equals_two_group_tile_kernel(sizes):
    """The two-tile ping-pong lane pool (two tiles per CTA whose evaluations alternate on the tensor
    pipe; every evaluation produced by two column groups) reproduces the tile-scheduled kernel run
    with two column groups bit for bit -- states, currents, losses, statistics and step checkpoints --
    for mixed jobs, a tile that runs dry long before the other, and CTAs with a single trajectory."""
    func, _ = _pair('d1')
    rng = np.random.RandomState(21)
    protos = [(*protocols.ap2hz(), 101), (*protocols.pr3_activation(20), 81),
              (*protocols.pr5_deactivation(-60), 41)]
    jobs = []
    for k, B in enumerate(sizes):
        tt, vv, T = protos[k % 3]
        y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                          dtype=torch.float32).cuda()
        t = torch.linspace(0., 2. * (T - 1), T)
        g = torch.tensor(rng.lognormal(0, 0.2, B), dtype=torch.float32)
        jobs.append(dict(protocol=(tt, vv), y0=y0, t=t, g=g, data=torch.zeros(T),
                         want_current=True, want_ckpt=True))
    a = ikr.integrate_many(func, jobs, options={'tc_groups': 2, 'lane_pool': False})
    b = ikr.integrate_many(func, jobs, options={'ping_pong': True})
    assert a[0].geometry['scheduling'].startswith('tile queue') and a[0].geometry['column_groups'] == 2
    assert b[0].geometry['scheduling'].startswith('two-tile ping-pong')
    for ra, rb in zip(a, b):
        assert torch.equal(ra.stats, rb.stats)
        assert torch.equal(ra.y, rb.y) and torch.equal(ra.current, rb.current)
        assert torch.equal(ra.sse, rb.sse) and torch.equal(ra.sae, rb.sae)
        for bb in range(0, ra.stats.shape[0], 7):
            k = int(ra.stats[bb, 0])
            assert torch.equal(ra.ckpt[1][:k, bb], rb.ckpt[1][:k, bb])
            assert torch.equal(ra.ckpt[0][:k, bb], rb.ckpt[0][:k, bb])


def test_tc_backward_overlap_of_wgrad_and_next_adjoint_round():
    """With fewer tiles than half the SMs `ikr_backward` runs the weight-gradient GEMM of round r on a
    second stream under the adjoint kernel of round r + 1 (double-buffered stash and counters).  Same
    gradient as the serial schedule up to the summation order of the (shorter) rounds, per-trajectory
    losses and grad_y0 bit-identical."""
    func, _ = _pair('d2')
    func.cuda()
    t_tab, v_tab = protocols.ap2hz()
    func.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 120., 61)
    rng = np.random.RandomState(31)
    B = 300
    y0 = torch.tensor(np.stack([rng.uniform(0, 0.05, B), rng.uniform(0.95, 1, B)], 1),
                      dtype=torch.float32).cuda()
    data = torch.from_numpy((rng.randn(len(t), B) * 0.1).astype(np.float32))
    out = {}
    with torch.enable_grad():
        for ov in (True, False):
            total, per, grads, res = ikr.loss_and_grad(func, y0, t, data, want_y0=True,
                                                       options={'first_step': 0.05, 'bwd_overlap': ov})
            out[ov] = (_flat(grads), per.clone(), res.grad_y0.clone())
    assert torch.equal(out[True][1], out[False][1]) and torch.equal(out[True][2], out[False][2])
    ref = out[False][0]
    assert np.abs(out[True][0] - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.abs(ref).max() > 0


def test_fp16x2_range_violation_on_an_accepted_step_is_reported():
    """A network whose hidden activations leave the fp16x2 range IN the physical domain (layer 4 scaled
    by 1,000 and the next layer by 1 / 1,000: the same function, LeakyReLU is positively homogeneous)
    must not integrate silently wrong: the lane ends with status IKR_TC_RANGE and `odeint` raises with
    the remedy in the message; the bf16x3 split (no range restriction) integrates it like the FFMA2
    kernel integrates the unscaled network."""
    func, _ = _pair('d1')
    big, _ = _pair('d1')
    with torch.no_grad():
        big.net[8].weight *= 1000.0
        big.net[8].bias *= 1000.0
        big.net[10].weight /= 1000.0
    t_tab, v_tab = protocols.pr3_activation(20)
    for f in (func, big):
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
    t = torch.linspace(0., 2000., 201)
    y0 = torch.tensor([[0.01, 0.97], [0.03, 0.96]], dtype=torch.float32).cuda()
    with pytest.raises(AssertionError, match='fp16x2'):
        ikr.integrate(big, y0, t)
    r = ikr.integrate(big, y0, t, options={'check_status': False})
    assert set(r.stats[:, 3].tolist()) == {5}
    ok = ikr.integrate(big, y0, t, options={'tc_split': 'bf16x3'})
    ref = ikr.integrate(func, y0, t, options={'tensor_cores': False})
    assert int((ok.stats[:, 3] != 0).sum()) == 0
    assert (ok.y - ref.y).abs().max().item() < 5e-4
