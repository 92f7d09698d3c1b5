#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference synthesizer modules, copied where they lie.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ddsp_oracle.py).  The reference
(tarepan/DDSP-SVC-official) is a plain Python tree: "building" it means copying the files the hot path
imports -- ddsp/{__init__,core,vocoder,unit2control,pcmer,loss}.py -- from /root/reference into
oracle/_ref/ddsp/.  oracle/_ref/ is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels to the GPU box, where `bench.py --impl reference` and the
`cpu_baseline` leg run it on the host cores through oracle/ref_shim.py.

    python oracle/make_ref.py            # (re)creates oracle/_ref/ ; no-op when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['ddsp/__init__.py', 'ddsp/core.py', 'ddsp/vocoder.py', 'ddsp/unit2control.py', 'ddsp/pcmer.py', 'ddsp/loss.py']


def make(src=None):
    src = src or os.environ.get('DDSP_REFERENCE', '/root/reference')
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            raise FileNotFoundError(s)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'sha256': manifest}, f, indent=1)
    return DST


if __name__ == '__main__':
    out = make(sys.argv[1] if len(sys.argv) > 1 else None)
    print(out or 'reference tree not found: oracle/_ref left as is')
