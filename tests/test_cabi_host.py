"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ikr.h declares, validates arguments without touching a GPU, and the Python host layer
(introspection, packing, options, protocol tables) behaves like the reference interface."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
import torch

import neural_ode_ion_channels_b200 as ikr
from neural_ode_ion_channels_b200 import _cabi, protocols, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    return _cabi.lib()


def test_library_exports_every_declared_symbol(lib):
    with open(os.path.join(ROOT, 'include', 'ikr.h')) as fh:
        hdr = fh.read()
    declared = set(re.findall(r'\b(ikr_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for sym in declared:
        assert hasattr(lib, sym), 'libikr_b200.so does not export %s' % sym
    assert set(_cabi.EXPORTS) <= declared
    assert lib.ikr_abi_version() == 2
    assert lib.ikr_error_string(-1) == b'invalid argument'


def test_struct_sizes_match_header(lib):
    # compile a tiny C program against include/ikr.h and compare sizeof with the ctypes mirrors
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "ikr.h"\nint main(){printf("%zu %zu %zu", ' \
          'sizeof(ikr_desc), sizeof(ikr_io), sizeof(ikr_bwd_io));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 's.c')
        with open(c, 'w') as fh:
            fh.write(src)
        exe = os.path.join(td, 's')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_cabi.IkrDesc), ctypes.sizeof(_cabi.IkrIO),
                     ctypes.sizeof(_cabi.IkrBwdIO)]


def _desc(func=None, state=torch.float32, method='dopri5', **opts):
    func = func or ikr.ODEFunc()
    return solver._make_desc(ikr.describe(func), state, method, 1e-7, 1e-9, opts)


def test_argument_validation_without_gpu(lib):
    d = _desc()
    io = _cabi.IkrIO()
    assert lib.ikr_forward(None, ctypes.byref(io), 1, None, 0, None) == -1
    assert lib.ikr_forward(ctypes.byref(d), ctypes.byref(io), 1, None, 0, None) == -1   # B = 0
    assert lib.ikr_forward(ctypes.byref(d), ctypes.byref(io), 0, None, 0, None) == -1   # no jobs
    mio = _cabi.IkrMarkovIO()
    assert lib.ikr_forward_markov(None, ctypes.byref(mio), None) == -1
    assert lib.ikr_forward_markov(ctypes.byref(d), ctypes.byref(mio), None) == -1      # B = 0
    bad = _desc()
    bad.n_layers = 0
    assert lib.ikr_packed_weight_elems(ctypes.byref(bad)) == -1
    bad = _desc()
    bad.state_dtype, bad.mlp_dtype = _cabi.F32, _cabi.F64
    one = (ctypes.c_int64 * 1)(10)
    assert lib.ikr_tile_m(ctypes.byref(bad), 1, one) == -1


def test_packed_layout_and_param_count(lib):
    for arch, (L, n) in ikr.ARCHITECTURES.items():
        f = ikr.ODEFunc(arch=arch)
        d = _desc(f)
        lay = _cabi.packed_layout(d)
        npad = (n + 7) // 8 * 8
        assert lay['npad'] == npad
        assert lay['total'] == 3 * npad + 2 * L * n * npad + L * npad + npad + 8
        assert lib.ikr_param_count(ctypes.byref(d)) == sum(p.numel() for p in f.parameters())
        geo = _cabi.launch_geometry(d, 65536)
        assert geo['smem'] <= 227 * 1024 and geo['threads'] <= 512 and geo['tile_m'] % 8 == 0
        assert geo['n_tiles'] * geo['tile_m'] >= 65536
        multi = _cabi.launch_geometry(d, [13107] * 4 + [13108])
        # tile-scheduled: whole tiles per job; lane-pool (large tensor-core launches): one pool
        assert multi['n_tiles'] * multi['tile_m'] >= 65536
        assert multi['n_tiles'] <= sum(-(-b // multi['tile_m']) for b in [13107] * 4 + [13108])
    # the default network (fp32 MLP, n = 200) runs on the tcgen05 kernel: 128-trajectory tiles (one
    # per TMEM lane), 3 x 128 lane threads + MMA warp + weight-producer warp
    d = _desc()
    geo = _cabi.launch_geometry(d, [13107] * 4 + [13108])
    assert lib.ikr_uses_tensor_cores(ctypes.byref(d)) == 1 and geo['tensor_cores']
    assert (geo['tile_m'], geo['threads']) == (128, 320)      # two column groups + two engine warps
    # opting out (reserved bit 1) gives the FFMA2 kernel its 16-warp, 160-trajectory tile
    d.reserved = 2
    geo = _cabi.launch_geometry(d, [13107] * 4 + [13108])
    assert lib.ikr_uses_tensor_cores(ctypes.byref(d)) == 0 and not geo['tensor_cores']
    assert (geo['tile_m'], geo['threads']) == (160, 512)
    # n = 500 (s06-s08) and fp64 MLPs do not fit the TMEM budget / are not fp32: FFMA2 / DFMA kernel
    big = _desc(ikr.ODEFunc(arch='s06'))
    assert lib.ikr_uses_tensor_cores(ctypes.byref(big)) == 0


def test_pack_weights_roundtrip(lib):
    torch.manual_seed(1)
    f = ikr.ODEFunc(arch='s09')
    spec = ikr.describe(f)
    d = _desc(f)
    buf = solver.pack_weights(spec, d, 'cpu')
    lay = _cabi.packed_layout(d)
    n, L, npad = 100, 5, lay['npad']
    wt = buf[lay['off_wt']:lay['off_wt'] + L * n * npad].view(L, n, npad)
    wn = buf[lay['off_wn']:lay['off_wn'] + L * n * npad].view(L, n, npad)
    for l in range(L):
        w = f.net[2 + 2 * l].weight.detach()
        assert torch.equal(wt[l, :, :n], w.t())
        assert torch.equal(wn[l, :, :n], w)
        assert float(wt[l, :, n:].abs().sum()) == 0.0
    assert torch.equal(buf[lay['off_wl']:lay['off_wl'] + n], f.net[-1].weight.detach().reshape(-1))
    assert buf[lay['off_wl'] + npad] == f.net[-1].bias.detach()[0]


def test_describe_contract():
    f = ikr.ODEFunc()
    s = ikr.describe(f)
    assert (s.n_layers, s.n_nodes, s.nn_d) == (5, 200, False)
    assert s.p[4:] == (f.p5, f.p6, f.p7, f.p8) and s.vrange == 100.0 and s.netscale == 1000.0
    s = ikr.describe(ikr.ODEFuncNNd(arch='s10'))
    assert (s.n_layers, s.n_nodes, s.nn_d) == (1, 100, True) and s.p[0] > 0
    with pytest.raises(TypeError):
        ikr.describe(torch.nn.Linear(2, 2))
    g = ikr.ODEFunc()
    g.net = torch.nn.Sequential(torch.nn.Linear(2, 8), torch.nn.ReLU(), torch.nn.Linear(8, 8),
                                torch.nn.ReLU(), torch.nn.Linear(8, 1))
    with pytest.raises(TypeError):
        ikr.describe(g)


def test_state_dict_files_load_unchanged():
    from tests import kat
    for study in kat.STUDIES:
        cls = ikr.ODEFuncNNd if study in ('s2', 'd2') else ikr.ODEFuncNNf
        f = cls()
        f.load_state_dict(torch.load(kat.weights_path(study)))     # train-s1.py:263 verbatim
        assert sum(p.numel() for p in f.parameters()) == 201801
    ck = {'epoch': 3, 'state_dict': f.state_dict(), 'optimizer': {}, 'loss': [0.1, 0.2]}
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'best-model-checkpoint-2.pt')
        torch.save(ck, path)                                        # train-r1.py:61-66 format
        g = ikr.load_weights(ikr.ODEFuncNNd(), path)
    assert torch.equal(g.net[0].weight, f.net[0].weight)


def test_module_forward_matches_oracle_rhs():
    from tests import kat
    from oracle import ref_models as rm
    t_tab, v_tab = protocols.ap2hz()
    for study in ('s1', 'd2'):
        cls = ikr.ODEFuncNNd if study == 'd2' else ikr.ODEFuncNNf
        f = ikr.load_weights(cls(params='s' if study == 's1' else 'd'), kat.weights_path(study))
        o = kat.make_nn(study)
        f.set_fixed_form_voltage_protocol(t_tab, v_tab)
        o.set_fixed_form_voltage_protocol(t_tab, v_tab)
        with torch.no_grad():
            for tq, a, r in ((12.3, 0.1, 0.9), (1000.05, 0.7, 0.2), (5000.0, 0.3, 0.3)):
                y = torch.tensor([[a, r]])
                np.testing.assert_allclose(f(torch.tensor(tq), y).numpy(),
                                           o(torch.tensor(tq), y).double().numpy(), rtol=2e-6,
                                           atol=1e-12)


def test_options_and_grid_semantics():
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        o = solver._split_options('dopri5', {'grid_points': 1, 'eps': 1e-6, 'first_step': 0.1})
    assert o == {'first_step': 0.1} and len(rec) == 1
    t = torch.linspace(0., 10., 6, dtype=torch.float64)
    from oracle import ref_odeint as ro
    for h in (0.7, 2.0, 3.3):
        mine = solver._rk4_grid(t, h)
        want = ro.RK4(None, torch.zeros(1, 2), step_size=h)._grid(t)
        assert torch.equal(mine, want.double())
    assert torch.equal(solver._rk4_grid(t, None), t)
    with pytest.raises(ValueError):
        ikr.integrate(ikr.ODEFunc(), torch.zeros(1, 2), t, method='adams')


def test_protocol_tables_follow_reference_definitions():
    t, v = protocols.pr3_activation(20)
    assert len(t) == 8001 and v[999] == -80 and v[1000] == 20 and v[5999] == 20
    assert v[6000] == -40 and v[7000] == -120 and v[7500] == -80
    t, v = protocols.pr3_activation(20, per_ms=10)
    assert len(t) == 80001 and v[9999] == -80 and v[10000] == 20 and v[60000] == -40
    t, v = protocols.pr5_deactivation(-90)
    assert len(t) == 10001 and v[1000] == 50 and v[3000] == -90 and v[9000] == -120
    t, v = protocols.pr2_time_constant(30)
    assert len(t) == 5001 and v[1000] == 40 and v[1030] == -120 and v[3530] == -80
    t, v = protocols.ap2hz()
    assert len(t) == 35000 and t[1] == 1.000000000000000048e-04 * 1e3 and v[0] == -80
    for fam in protocols.PROTOCOL_FAMILIES + ('staircase',):
        for name, tt, vv, to in protocols.protocol_set(fam):
            assert len(tt) == len(vv) and np.all(np.diff(tt) > 0) and to[-1] <= tt[-1]


def test_compact_table_is_bit_exact_for_interp1d():
    from scipy.interpolate import interp1d
    rng = np.random.RandomState(0)
    for t, v in (protocols.pr5_deactivation(-70), protocols.staircase_standin(),
                 protocols.ap2hz()):
        tc, vc = protocols.compact_table(t, v)
        assert len(tc) <= len(t)
        q = np.concatenate([rng.uniform(t[0], t[-1], 4000), t[::97], tc])
        assert np.array_equal(interp1d(t, v)(q), interp1d(tc, vc)(q))
    tc, _ = protocols.compact_table(*protocols.pr3_activation(40, per_ms=10))
    assert len(tc) <= 12


def test_no_cpu_fallback():
    f = ikr.ODEFunc()
    f.set_fixed_form_voltage_protocol(*protocols.pr3_activation(0))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            ikr.odeint(f, torch.tensor([[0., 1.]]), torch.linspace(0., 1., 3))


def test_shipped_sass_keeps_thread_index_and_geometry_in_registers(lib):
    """Code-generation guard for the tensor-core forward kernels.  nvcc's `-split-compile` splits
    the NVVM optimiser and was seen to emit two different PTX files for the same source from one
    build to the next; in one of them %tid and the tile geometry are rematerialised at every use
    (114 instead of 10-15 S2R in the tile kernel, 159 instead of 161-168 registers) and the
    18,944-trajectory tile benchmark takes 32.5 instead of 26.3 ms.  csrc/build.py therefore
    parallelises ptxas only; this test reads the shipped library's SASS so that a slow variant
    cannot ship unnoticed."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    path = _cabi.LIB_PATH
    sass = subprocess.run([cuobjdump, '-sass', path], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur and ' S2R ' in line:
            counts[cur] = counts.get(cur, 0) + 1
    hot = {k: v for k, v in counts.items()
           if re.match(r'_ZN3ikr(21ikr_forward_tc_kernel|26ikr_forward_tc_pool_kernel)IfLi2ELi2E', k)}
    assert len(hot) == 2, sorted(counts)[:5]
    for name, n in hot.items():
        assert n < 40, '%s: %d S2R -- the slow code-generation variant (see csrc/build.py)' % (name, n)
