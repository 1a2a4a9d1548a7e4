"""The tcgen05 building blocks of csrc/ikr_tc.cuh checked in isolation (tests/tc_probe.cu, built by
__graft_entry__.build()): the TMEM A-operand layout + no-swizzle K-major B descriptor of the forward
/ adjoint MMAs, and the MN-major shared-memory operands of the weight-gradient MMAs."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

PROBE = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_tc_probe')


@pytest.mark.parametrize('args', [['0', '208', '13'], ['0', '112', '7'], ['mn', '0', '208'], ['mn', '0', '16']])
def test_tcgen05_layouts(args):
    if not os.path.exists(PROBE):
        pytest.skip('tests/_tc_probe not built (python -c "import __graft_entry__ as g; g.build()")')
    res = subprocess.run([PROBE] + args, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'bad 0/' in res.stdout
