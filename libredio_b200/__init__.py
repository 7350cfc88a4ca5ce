"""libredio_b200 -- B200 (sm_100a) implementation of LibRedio's sample-stream DSP hot path.

The product is libredio_b200/libredio_cuda.so (hand-written CUDA behind the C ABI of
include/libredio_cuda.h).  This package is the thin host-side mirror used by tests and bench.py.
Importing the package does not need a GPU; creating a Context does.
"""
from . import capi, synth  # noqa: F401
from .capi import LrcError, WINDOW_HANN, WINDOW_NONE  # noqa: F401

__all__ = ["capi", "synth", "blocks", "LrcError", "WINDOW_HANN", "WINDOW_NONE"]


def __getattr__(name):
    if name == "blocks":           # needs torch; keep `import libredio_b200` light
        import importlib
        return importlib.import_module(".blocks", __name__)
    raise AttributeError(name)
