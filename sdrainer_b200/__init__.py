"""sdrainer_b200 -- B200-native (sm_100a) DSP hot path of SDRainer behind a C ABI.

The product is `libsdrgpu.so` (hand-written CUDA, see csrc/ and include/sdrgpu.h).  This Python
package is only the ctypes binding used by the tests and bench.py, plus the synthetic IQ generator.
There is NO CPU fallback: importing `capi` without the built library raises.
"""
from . import _build  # noqa: F401

__all__ = ["capi", "synth", "_build"]
