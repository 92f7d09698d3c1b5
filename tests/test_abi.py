"""CPU-side checks of the C-ABI boundary: the shared library builds/loads without a GPU and
exports every symbol include/ddsp_b200.h declares; the ctypes prototypes cover them all; argument
validation rejects bad calls before any CUDA work (no compute calls here)."""
import ctypes
import os
import re

import pytest

from ddsp_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ddsp_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ddsp_b200_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _cabi.lib()
    assert lib.ddsp_b200_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 15
    handle = ctypes.CDLL(build.LIB)
    for s in syms:
        assert hasattr(handle, s), f'{s} declared in include/ddsp_b200.h but not exported'
    assert sorted(_cabi.exported_symbols()) == syms, 'ctypes prototypes out of sync with the header'


def test_strerror_and_argument_validation_without_gpu():
    lib = _cabi.lib()
    assert lib.ddsp_b200_strerror(0) == b'ok'
    assert b'block_size' in lib.ddsp_b200_strerror(-2)
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.ddsp_b200_phase(0, 0, 0, 1, 4, 512, 44100.0, 0, 1, 0, 0, 0, 0) == -1
    assert lib.ddsp_b200_upsample(0, 0, 0, 0, 1, 1, 1, 512, 0, 0) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert lib.ddsp_b200_phase(p, 4, 1, 1, 4, 256, 44100.0, 0, 1, p, p, 0, 0) == -2      # hop != 512
    assert lib.ddsp_b200_fo_to_rot_workspace_bytes(2, 5000) == 2 * 3 * 8
    assert lib.ddsp_b200_combsub_workspace_bytes(2, 10, 256, 512, 256) == 2 * 2 * 10 * 512 * 4 + 2 * 2 * 10 * 1024 * 8
    with pytest.raises(ValueError):
        _cabi.check(-5)
    with pytest.raises(_cabi.DDSPB200Error):
        _cabi.check(-1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'ddsp-svc-official_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('the CPU oracle in `oracle/`', ''), f
