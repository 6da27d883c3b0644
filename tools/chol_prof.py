"""Timing experiment for k_band_chol (development aid): prints the fp64 microbenchmark and the
per-phase cycle profile of the band Cholesky on a random system of the bench shape."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlesfm_b200 import _lib

L = _lib.lib()
r, lat = C.c_double(), C.c_double()
L.psfm_measure_dfma(C.byref(r), C.byref(lat))
print("dfma/s %.4g (%.2f TFLOP/s fp64), dependent DFMA latency %.1f cycles" % (r.value, 2e-12 * r.value, lat.value))
