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
OBJ = os.path.join(HERE, 'build')   # object files (git-ignored)

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
]
LINK_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-cudart', 'static', '-ldl']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))] + \
        [os.path.join(HERE, '..', 'include', 'b200wave.h'), os.path.abspath(__file__)]


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libb200wave.so')
    return exe


def _newest_header():
    return max(os.path.getmtime(h) for h in headers() if os.path.exists(h))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return _newest_header() > t or any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in sources())


def _run(cmd):
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc wants the system g++
    env.pop('CC', None)
    env.pop('CXX', None)
    r = subprocess.run(cmd + ['-ccbin', '/usr/bin/g++'], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False, out=None, extra_flags=(), only=None):
    """Compile every translation unit of csrc/ (in parallel, one nvcc per .cu, objects cached under build/) and link
    libb200wave.so.  ``out`` / ``extra_flags`` build an experimental variant next to the default one (objects are
    then not cached); ``only`` = the translation units the extra flags apply to (the rest link from the cache).
    The package only ever loads libb200wave.so unless B200W_LIB points elsewhere."""
    from concurrent.futures import ThreadPoolExecutor
    target = out or SO
    if out is None and not force and not needs_build():
        return SO
    variant = out is not None or bool(extra_flags)
    objdir = os.path.join(OBJ, 'variant_%d' % os.getpid()) if variant else OBJ
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for src in sources():
        special = variant and (only is None or src in only)
        o = os.path.join(objdir if special else OBJ, src[:-3] + '.o')
        objs.append(o)
        sp = os.path.join(CSRC, src)
        if force or special or not os.path.exists(o) or os.path.getmtime(o) < max(hdr_t, os.path.getmtime(sp)):
            jobs.append([nvcc()] + NVCC_FLAGS + (list(extra_flags) if special else []) +
                        (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', o, sp])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        logs = list(ex.map(_run, jobs))
    log = _run([nvcc()] + LINK_FLAGS + ['-o', target] + objs)
    if verbose:
        print('\n'.join(logs) + log)
    if variant:
        shutil.rmtree(objdir, ignore_errors=True)
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
