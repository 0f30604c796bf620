"""sketch-for-rna-seq_b200: B200-native (sm_100a) quant hot path of Codfishz/Sketch-for-RNA-seq.

Product = `libsketchquant.so` (hand-written CUDA kernels behind the C ABI of include/sketchquant.h) plus the
C++17 command-line host in host/ (drop-in for `./build/test -o index|quant`).  The Python modules here are
the ctypes binding used by tests and bench.py, the host-side mirror of the reference's functions for the
path, and data helpers (2-bit packing, index file I/O, synthetic data).  Nothing in this package imports
or executes anything under oracle/.
"""
from . import capi, packing, index_io, synth, api  # noqa: F401
from .capi import Engine, SketchQuantError, lib_path, load_library  # noqa: F401
