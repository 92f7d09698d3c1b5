"""Loader of the PyTorch-extension host (lib/ddsp_b200_torch.so: TORCH_LIBRARY(ddsp_b200) operators over the C ABI).

`ops()` returns `torch.ops.ddsp_b200` once the library is loaded (built in-tree on first use when stale and a
compiler is present).  When a differently built C-ABI library is forced through DDSP_B200_LIB (kernel
experiments) the extension is not used, because it links lib/libddsp_b200.so itself; the ctypes path in core.py
then carries the calls -- both paths enqueue the same kernels.
"""
import os

import torch

from . import build as _build

_ops = None
_tried = False


def ops():
    global _ops, _tried
    if _tried:
        return _ops
    _tried = True
    if os.environ.get('DDSP_B200_LIB') or os.environ.get('DDSP_B200_NO_TORCH_EXT') == '1':
        return None
    path = _build.TORCH_LIB
    try:
        if not os.path.exists(path) or (_build.torch_ext_is_stale() and os.environ.get('DDSP_B200_NO_REBUILD') != '1'):
            path = _build.build_torch_ext()
    except Exception:
        if not os.path.exists(_build.TORCH_LIB):
            return None
        path = _build.TORCH_LIB
    from . import _cabi
    _cabi.lib()                                   # the C-ABI library first: the extension resolves its symbols against it
    torch.ops.load_library(path)
    _ops = torch.ops.ddsp_b200
    return _ops
