"""Build libikr_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'libikr_b200.so')
SOURCES = ['ikr_capi.cu']
HEADERS = ['ikr_math.h', 'ikr_device.cuh', 'ikr_forward.cuh', 'ikr_backward.cuh', 'ikr_hh.cuh',
           'ikr_tc.cuh', 'ikr_forward_tc.cuh', 'ikr_backward_tc.cuh', 'ikr_markov.cuh', 'ikr_regress_tc.cuh',
           os.path.join('..', '..', 'include', 'ikr.h')]

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
    '-fmad=false',          # solver arithmetic must not be contracted (torchdiffeq semantics);
                            # the MLP inner loops use explicit fma intrinsics
    '-Xcompiler', '-fPIC', '-shared', '--expt-relaxed-constexpr',
    # link the CUDA runtime dynamically: the process already holds libcudart.so.12 (torch's), and the
    # shipped binary then carries none of the runtime's unused entry-point names
    '-cudart', 'shared', '-Xlinker', '-rpath,/usr/local/cuda/lib64',
    # PTX -> SASS of the kernels in parallel.  Only ptxas: nvcc's own -split-compile also splits the
    # NVVM optimiser, whose output then differs from build to build (two PTX variants of the same
    # source were observed; in one of them %tid and the tile geometry are re-read at every use and
    # the tensor-core forward is 24 % slower) -- tests/test_cabi_host.py checks the shipped SASS.
    '-Xptxas', '-v,-split-compile=0',
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: a variant build next to the shipped library (A/B runs select it with
    IKR_B200_LIB); the default call builds libikr_b200.so."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + [os.path.join(HERE, s) for s in SOURCES] + ['-o', out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log') if out == OUT else out + '.build.log', 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + log)
    if verbose or res.returncode != 0:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libikr_b200.so (see csrc/build.log)')
    return out


if __name__ == '__main__':
    # python build.py [variant.so -DNAME=VALUE ...]
    if len(sys.argv) > 1:
        print(build(force=True, out=os.path.abspath(sys.argv[1]), defines=sys.argv[2:]))
    else:
        print(build(force=True, verbose=True))
