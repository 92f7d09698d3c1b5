"""Importable name for the package that lives in `ddsp-svc-official_b200/`.

The product directory carries the reference repo's name (with a hyphen, so it cannot be
imported directly); this shim points the package path at it, so `import ddsp_b200` and
`from ddsp_b200.vocoder import CombSubFast` resolve to `ddsp-svc-official_b200/*.py`.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'ddsp-svc-official_b200')
__path__ = [_impl]
with open(_os.path.join(_impl, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, '__init__.py'), 'exec'))
del _f
