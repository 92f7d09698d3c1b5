"""Build the C-ABI shared library (csrc/*.cu -> lib/libddsp_b200.so) with nvcc for sm_100a.

In-tree on purpose: the .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libddsp_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build libddsp_b200.so')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))) + \
        [os.path.join(INCLUDE, 'ddsp_b200.h')]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources.  Returns the .so path.

    Safe under concurrent callers (e.g. 8 torchrun ranks importing at once): an flock serialises
    the builders, late comers re-check staleness, and the library is moved into place atomically."""
    if not force and not is_stale():
        return LIB
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            tmp = f'{LIB}.tmp.{os.getpid()}'
            cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
                   '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v', '-I', INCLUDE,
                   '-o', tmp, os.path.join(CSRC, 'ddsp_b200.cu')]
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = res.stdout + res.stderr
            with open(os.path.join(LIBDIR, 'build.log'), 'w') as f:
                f.write(' '.join(cmd) + '\n' + log)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed:\n' + log)
            os.replace(tmp, LIB)
            if verbose:
                print(log)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
