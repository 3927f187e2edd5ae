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
