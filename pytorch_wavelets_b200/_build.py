"""Build libb200wave.so (CUDA, sm_100a) in-tree with nvcc.  No GPU is needed to compile.

    python -m pytorch_wavelets_b200._build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libb200wave.so')
SOURCES = ['b200wave.cu']
HEADERS = ['common.h', 'tile_kernels.h', 'launch_params.h', 'fast_kernels.cuh', 'fast_dtcwt.cuh', 'fast_inverse.cuh', os.path.join('..', '..', 'include', 'b200wave.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
    '--expt-relaxed-constexpr',
    '-cudart', 'static',
]


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libb200wave.so')
    return exe


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, out=None, extra_flags=()):
    """Compile the library.  ``out`` / ``extra_flags`` build an experimental variant next to the default one
    (see profiles/r01_notes.md, "A/B discipline"); the package itself only ever loads libb200wave.so unless B200W_LIB points elsewhere."""
    target = out or SO
    if out is None and not force and not needs_build():
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', target] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc wants the system g++
    env.pop('CC', None)
    env.pop('CXX', None)
    r = subprocess.run(cmd + ['-ccbin', '/usr/bin/g++'], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout)
    if verbose:
        print(r.stdout)
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
