#!/usr/bin/env python
"""Build variants of libddsp_b200.so with different -D switches into ddsp-svc-official_b200/lib/variants/
(git-ignored, shipped to the GPU box).  usage: python profiles/build_variants.py name="-DX=1 -DY=2" ...
Variants compile in parallel.  Pair with profiles/run_variants.py on the GPU box."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ddsp-svc-official_b200')
OUT = os.path.join(PKG, 'lib', 'variants')
os.makedirs(OUT, exist_ok=True)


def build(item):
    name, flags = item
    out = os.path.join(OUT, f'{name}.so')
    cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
           '-shared', '-Xptxas', '-v', '-I', os.path.join(ROOT, 'include'), '-o', out] + flags.split() + \
          [os.path.join(PKG, 'csrc', 'ddsp_b200.cu')]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    info = ''
    lines = log.split('\n')
    for i, l in enumerate(lines):
        if 'combsubfast_kernelILb0' in l and 'Compiling' in l:
            info = ' | '.join(x.strip() for x in lines[i + 1:i + 3])
    return name, r.returncode, info if r.returncode == 0 else log[-2000:]


if __name__ == '__main__':
    items = []
    for a in sys.argv[1:]:
        n, _, f = a.partition('=')
        items.append((n, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, rc, info in ex.map(build, items):
            print(f'{name:24s} rc={rc} {info}')
