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


TORCH_LIB = os.path.join(LIBDIR, 'ddsp_b200_torch.so')
TORCH_SRC = os.path.join(CSRC, 'torch_ext.cpp')


def torch_ext_is_stale():
    if not os.path.exists(TORCH_LIB):
        return True
    t = os.path.getmtime(TORCH_LIB)
    return os.path.getmtime(TORCH_SRC) > t or os.path.getmtime(os.path.join(INCLUDE, 'ddsp_b200.h')) > t


def build_torch_ext(force=False):
    """Compile the PyTorch-extension host (csrc/torch_ext.cpp: TORCH_LIBRARY(ddsp_b200) operators over the C ABI) into
    lib/ddsp_b200_torch.so, linked against lib/libddsp_b200.so (rpath $ORIGIN) and this interpreter's torch."""
    if not force and not torch_ext_is_stale():
        return TORCH_LIB
    import fcntl
    import sys
    import torch
    from torch.utils import cpp_extension as ce
    build()                                                    # the C-ABI library it links
    with open(os.path.join(LIBDIR, '.build_torch.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not torch_ext_is_stale():
                return TORCH_LIB
            tmp = f'{TORCH_LIB}.tmp.{os.getpid()}'
            cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), 'include')
            cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
                   '-DTORCH_API_INCLUDE_EXTENSION_H', TORCH_SRC, '-o', tmp, '-I', cuda_inc]
            for inc in ce.include_paths():
                cmd += ['-isystem', inc]
            for lp in ce.library_paths():
                cmd += ['-L', lp, '-Wl,-rpath,' + lp]
            cmd += ['-L', LIBDIR, '-lddsp_b200', '-Wl,-rpath,$ORIGIN', '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch']
            res = subprocess.run(cmd, capture_output=True, text=True)
            with open(os.path.join(LIBDIR, 'build_torch.log'), 'w') as f:
                f.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('g++ failed on torch_ext.cpp:\n' + (res.stdout + res.stderr)[-3000:])
            os.replace(tmp, TORCH_LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return TORCH_LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
    print(build_torch_ext(force=True))
