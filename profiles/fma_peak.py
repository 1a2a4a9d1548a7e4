"""FMA-pipe peaks measured on the device: scalar FFMA, packed FFMA2, DFMA (roofline denominators)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401
from neural_ode_ion_channels_b200 import _cabi  # noqa: E402
torch.cuda.init()
for name, code in (('FFMA f32', 0), ('FFMA2 f32x2', 2), ('DFMA f64', 1)):
    v = ctypes.c_double(0)
    for it in (20000, 200000):
        _cabi.check(_cabi.lib().ikr_fma_peak(code, it, ctypes.byref(v), None), 'peak')
    print('%-12s %.2f TFLOP/s' % (name, v.value))
