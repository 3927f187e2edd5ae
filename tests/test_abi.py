"""CPU-side checks of the C-ABI boundary: libg2pgpu.so loads, exports every symbol that
include/g2p.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pyg2p_b200 import _lib as L

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'g2p.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(g2p_[a-z0-9_]+)\s*\(', text)))


def test_header_and_prototypes_agree():
    syms = _declared_symbols()
    assert len(syms) >= 15
    assert sorted(L.PROTOTYPES) == syms


def test_library_exports_every_declared_symbol():
    handle = C.CDLL(L.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(handle, s), s
    assert L.lib().g2p_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    from pyg2p_b200 import device
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        device.SourceIndex(lon=[0.0, 1.0], lat=[0.0, 1.0], radius=1.0)
    # argument validation does not need a device
    lib = L.lib()
    assert lib.g2p_apply(None, None, 1, 1, None, None, 1, 1, 1, 1, 0.0, 0, None, None, None) == -1
    assert b'null' in lib.g2p_last_error()


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under pyg2p_b200/ may import, call or mention it."""
    pkg = os.path.join(REPO, 'pyg2p_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                # ... nor reach for a CPU implementation of the path (scipy's cKDTree, numexpr)
                assert not re.search(r'^\s*(from|import)\s+(scipy|numexpr)\b', text, flags=re.M), f


@pytest.mark.parametrize('threads', [0, 1, 2, 4])
def test_host_pick_is_pure_host_staging(threads):
    """g2p_host_pick is the one entry point that works on HOST memory and needs no device: the touched values of a decoded
    message into a compact buffer - with their mask bytes, or with the marker NaN folded in - on 1..4 threads."""
    import numpy as np
    lib = L.lib()
    rng = np.random.default_rng(5 + threads)
    n, nc = 300_000, 131_075                      # large enough to be split over the helper threads
    touched = np.sort(rng.choice(n, nc, replace=False)).astype(np.int32)
    x = rng.standard_normal(n)
    mask = (rng.random(n) < 0.03).view(np.uint8)
    out, om = np.empty(nc), np.full(nc, 7, np.uint8)
    assert lib.g2p_host_pick(x.ctypes.data, mask.ctypes.data, touched.ctypes.data, nc, out.ctypes.data, om.ctypes.data, 0,
                             threads) == 0
    np.testing.assert_array_equal(out, x[touched])
    np.testing.assert_array_equal(om, mask[touched])
    assert lib.g2p_host_pick(x.ctypes.data, None, touched.ctypes.data, nc, out.ctypes.data, om.ctypes.data, 0, threads) == 0
    np.testing.assert_array_equal(out, x[touched])
    assert not om.any()
    assert lib.g2p_host_pick(x.ctypes.data, mask.ctypes.data, touched.ctypes.data, nc, out.ctypes.data, None, 1, threads) == 0
    marked = mask[touched] != 0
    assert (out[marked].view(np.uint64) == L.MASK_MARKER).all()
    np.testing.assert_array_equal(out[~marked], x[touched][~marked])
    assert lib.g2p_host_pick(None, None, touched.ctypes.data, nc, out.ctypes.data, None, 0, threads) == -1
