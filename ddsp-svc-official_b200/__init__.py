"""ddsp_b200 -- B200-native DDSP-SVC synthesizer forward (Sins / CombSub / CombSubFast).

Product code lives here; the CPU oracle in `oracle/` is test infrastructure and is never
imported from this package.
"""
__version__ = '0.1.0'
