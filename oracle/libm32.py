"""Oracle (test infrastructure): float32 sin/cos exactly as numexpr evaluates them.

numexpr's float32 opcodes `sin_ff` / `cos_ff` call the C library's `sinf` / `cosf` (no VML in
the pip/conda-forge builds the reference pins, numexpr 2.8.6).  numpy's own float32 sin/cos are
a different (SIMD, <=1.5 ulp) algorithm and differ from glibc's in ~15 % of the values, so the
restatement calls glibc's functions directly through ctypes.  float64 sin/cos need no such care:
numpy's results are bit-identical to glibc's `sin`/`cos` in this image (checked on 2 M values).
"""
import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library('m') or 'libm.so.6')
for _n in ('sinf', 'cosf'):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]


def _apply(fn, x):
    x = np.asarray(x, dtype=np.float32)
    flat = x.ravel()
    uniq, inv = np.unique(flat, return_inverse=True)     # lat/lon maps repeat values a lot
    vals = np.fromiter((fn(float(v)) for v in uniq), dtype=np.float32, count=uniq.size)
    return vals[inv].reshape(x.shape)


def sinf(x):
    return _apply(_libm.sinf, x)


def cosf(x):
    return _apply(_libm.cosf, x)


def sin(x):
    """numexpr `sin`: float32 in -> libm sinf; otherwise numpy float64 (== libm sin here)."""
    return sinf(x) if isinstance(x, np.ndarray) and x.dtype == np.float32 else np.sin(x)


def cos(x):
    return cosf(x) if isinstance(x, np.ndarray) and x.dtype == np.float32 else np.cos(x)
