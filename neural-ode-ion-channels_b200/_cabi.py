"""ctypes binding of the C ABI in ``include/ikr.h`` (``csrc/libikr_b200.so``).

Thin by design: plain pointers and sizes cross the boundary, torch only provides device memory
and streams.  There is NO fallback: if the shared library is missing or was not built for this
GPU, importing/using it raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IKR_B200_LIB: another build of the same library (A/B runs of kernel variants on one GPU box)
LIB_PATH = os.environ.get('IKR_B200_LIB') or os.path.join(_HERE, 'csrc', 'libikr_b200.so')

F32, F64 = 0, 1
DOPRI5, RK4 = 0, 1
STATUS_TEXT = {
    1: 'underflow in dt',
    2: 'max_num_steps exceeded',
    3: 'non-finite values in state `y`',
    4: 'step-checkpoint capacity exceeded',
    5: "a hidden activation left the range of the fp16x2 tensor-core split on an accepted step; "
       "use options={'tc_split': 'bf16x3'} or {'tensor_cores': False}",
}

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class IkrDesc(ctypes.Structure):
    _fields_ = [
        ('n_layers', c_i32), ('n_nodes', c_i32), ('nn_d', c_i32), ('method', c_i32),
        ('state_dtype', c_i32), ('mlp_dtype', c_i32), ('time_f32', c_i32), ('rk4_perturb', c_i32),
        ('p', c_f64 * 8),
        ('vrange', c_f64), ('netscale', c_f64), ('negative_slope', c_f64),
        ('rtol', c_f64), ('atol', c_f64), ('first_step', c_f64),
        ('safety', c_f64), ('ifactor', c_f64), ('dfactor', c_f64),
        ('max_num_steps', c_i64),
        ('tile_m', c_i32), ('reserved', c_i32),
    ]


class IkrIO(ctypes.Structure):
    _fields_ = [
        ('B', c_i64), ('T', c_i64), ('G', c_i64),
        ('weights', c_vp), ('table_t', c_vp), ('table_v', c_vp),
        ('table_len', c_i32), ('table_uniform', c_i32), ('table_t0', c_f64),
        ('table_inv_dt', c_f64), ('cost_hint', c_f64),
        ('y0', c_vp), ('t_out', c_vp),
        ('grid', c_vp), ('v_out', c_vp), ('g', c_vp), ('e_rev', c_vp), ('e_scalar', c_f64),
        ('data', c_vp), ('data_B', c_i64),
        ('y_out', c_vp), ('i_out', c_vp), ('loss_out', c_vp), ('stats_out', c_vp),
        ('ckpt_cap', c_i64), ('ckpt_t', c_vp), ('ckpt_y', c_vp),
    ]


class IkrMarkovIO(ctypes.Structure):
    _fields_ = [
        ('B', c_i64), ('T', c_i64), ('G', c_i64),
        ('table_t', c_vp), ('table_v', c_vp), ('table_len', c_i32), ('table_uniform', c_i32),
        ('table_t0', c_f64), ('table_inv_dt', c_f64),
        ('y0', c_vp), ('t_out', c_vp), ('grid', c_vp), ('v_out', c_vp), ('params', c_vp),
        ('p', c_f64 * 12), ('g', c_vp), ('e_rev', c_f64), ('noise_sigma', c_f64),
        ('seed', ctypes.c_uint64), ('y_out', c_vp), ('i_out', c_vp), ('stats_out', c_vp),
    ]


class IkrBwdIO(ctypes.Structure):
    _fields_ = [
        ('grad_y', c_vp), ('fused_loss', c_i32), ('reserved', c_i32),
        ('max_accepted_steps', c_i64), ('grad_weights', c_vp), ('grad_y0', c_vp), ('grad_g', c_vp),
    ]


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is absent -- no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libikr_b200.so is not built (%s). Run `python -c "import __graft_entry__ as g; '
            'g.build()"` or `python neural-ode-ion-channels_b200/csrc/build.py`. There is no CPU '
            'fallback for this path.' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.ikr_abi_version.restype = c_i32
    L.ikr_error_string.restype = ctypes.c_char_p
    L.ikr_error_string.argtypes = [c_i32]
    L.ikr_packed_weight_elems.restype = c_i64
    L.ikr_packed_weight_elems.argtypes = [ctypes.POINTER(IkrDesc)]
    L.ikr_packed_layout.restype = c_i32
    L.ikr_packed_layout.argtypes = [ctypes.POINTER(IkrDesc), ctypes.POINTER(c_i64)]
    L.ikr_param_count.restype = c_i64
    L.ikr_param_count.argtypes = [ctypes.POINTER(IkrDesc)]
    L.ikr_uses_tensor_cores.restype = c_i32
    L.ikr_uses_tensor_cores.argtypes = [ctypes.POINTER(IkrDesc)]
    L.ikr_tile_m.restype = c_i32
    L.ikr_tile_m.argtypes = [ctypes.POINTER(IkrDesc), c_i32, ctypes.POINTER(c_i64)]
    L.ikr_launch_geometry.restype = c_i32
    L.ikr_launch_geometry.argtypes = [ctypes.POINTER(IkrDesc), c_i32, ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_i64)]
    L.ikr_workspace_bytes.restype = ctypes.c_size_t
    L.ikr_workspace_bytes.argtypes = [ctypes.POINTER(IkrDesc), c_i32, c_i64, c_i32]
    L.ikr_forward.restype = c_i32
    L.ikr_forward.argtypes = [ctypes.POINTER(IkrDesc), ctypes.POINTER(IkrIO), c_i32, c_vp,
                              ctypes.c_size_t, c_vp]
    L.ikr_backward.restype = c_i32
    L.ikr_backward.argtypes = [ctypes.POINTER(IkrDesc), ctypes.POINTER(IkrIO),
                               ctypes.POINTER(IkrBwdIO), c_vp, ctypes.c_size_t, c_vp]
    L.ikr_forward_hh.restype = c_i32
    L.ikr_forward_hh.argtypes = [ctypes.POINTER(IkrDesc), ctypes.POINTER(IkrIO), c_vp, c_vp]
    L.ikr_regression_workspace_bytes.restype = ctypes.c_size_t
    L.ikr_regression_workspace_bytes.argtypes = [ctypes.POINTER(IkrDesc), c_i64]
    L.ikr_regression_loss_grad.restype = c_i32
    L.ikr_regression_loss_grad.argtypes = [ctypes.POINTER(IkrDesc), c_vp, c_vp, c_vp, c_i64, c_vp, c_vp,
                                           c_vp, ctypes.c_size_t, c_vp]
    L.ikr_forward_markov.restype = c_i32
    L.ikr_forward_markov.argtypes = [ctypes.POINTER(IkrDesc), ctypes.POINTER(IkrMarkovIO), c_vp]
    L.ikr_interp_protocol.restype = c_i32
    L.ikr_interp_protocol.argtypes = [ctypes.POINTER(IkrIO), c_vp, c_i64, c_vp, c_vp]
    L.ikr_fma_peak.restype = c_i32
    L.ikr_fma_peak.argtypes = [c_i32, c_i64, ctypes.POINTER(c_f64), c_vp]
    if L.ikr_abi_version() != 2:
        raise RuntimeError('libikr_b200.so ABI version mismatch')
    _lib = L
    return L


EXPORTS = ('ikr_abi_version', 'ikr_error_string', 'ikr_packed_weight_elems', 'ikr_packed_layout',
           'ikr_param_count', 'ikr_uses_tensor_cores', 'ikr_tile_m', 'ikr_launch_geometry', 'ikr_workspace_bytes',
           'ikr_forward', 'ikr_backward', 'ikr_forward_hh', 'ikr_forward_markov', 'ikr_regression_workspace_bytes',
           'ikr_regression_loss_grad', 'ikr_interp_protocol',
           'ikr_fma_peak')


def check(code, what):
    if code != 0:
        raise RuntimeError('%s failed: %s (%d)' % (what, lib().ikr_error_string(code).decode(), code))


def packed_layout(desc):
    out = (c_i64 * 8)()
    check(lib().ikr_packed_layout(ctypes.byref(desc), out), 'ikr_packed_layout')
    return {'npad': out[0], 'off_w0': out[1], 'off_wt': out[2], 'off_bh': out[3],
            'off_wl': out[4], 'off_wn': out[5], 'total': out[6], 'kc': out[7]}


def launch_geometry(desc, B):
    out = (c_i64 * 16)()
    Bs = [int(B)] if not isinstance(B, (list, tuple)) else [int(b) for b in B]
    arr = (c_i64 * len(Bs))(*Bs)
    check(lib().ikr_launch_geometry(ctypes.byref(desc), len(Bs), arr, out), 'ikr_launch_geometry')
    return {'tile_m': out[0], 'threads': out[1], 'grid': out[2], 'smem': out[3],
            'n_tiles': out[4], 'kc': out[5], 'cpl': out[6], 'sms': out[7],
            'scheduling': ('tile queue (longest job first)', 'lane pool (slots refill from one '
                           'trajectory queue)', 'two-tile ping-pong lane pool')[out[8]],
            'launches': out[9], 'tensor_cores': bool(out[10]), 'column_groups': out[11],
            'mma_products': out[12],
            'mma_split': {3: 'fp16x2 split', 6: 'bf16x3 split'}.get(out[12], 'none')}
