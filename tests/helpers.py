"""Test doubles for the reference's execution context (reference tests/__init__.py:43-73) and a writer of
PCRaster CSF maps so that the Interpolator can be driven exactly as the reference's tests drive it."""
import struct

import numpy as np


class IntertablesConf:
    """stand-in for IntertablesConfiguration: `.vars`, `.user_vars`, `.check_write()`, `.dump()`."""

    def __init__(self, entries=None):
        self.vars = dict(entries or {})
        self.user_vars = {}
        self.dumps = 0

    def check_write(self):
        return True

    def dump(self):
        self.dumps += 1


class _Configuration:
    def __init__(self, entries=None):
        self.intertables = IntertablesConf(entries)


class MockedExecutionContext:
    """dict-backed ExecutionContext: `.get`, `.configuration`, `.is_with_grib_interpolation`, `.geo_file`."""

    def __init__(self, d, gribinterp=False, entries=None):
        self._vars = d
        self.is_with_grib_interpolation = gribinterp
        self.configuration = _Configuration(entries)

    def get(self, param, default=None):
        return self._vars.get(param, default)

    def geo_file(self, _):
        return None


def write_csf_real4(path, array, x_ul, y_ul, cell):
    """float32 2-D array -> PCRaster CSF v2 REAL4 scalar map; NaN / <= -3e38 cells become MV (all bits set)."""
    a = np.asarray(array, dtype='<f4').copy()
    mv = ~np.isfinite(a) | (a <= -3e38)
    valid = a[~mv]
    raw = a.view('<u4')
    raw[mv] = 0xffffffff
    hdr = bytearray(256)
    hdr[0:27] = b'RUU CROSS SYSTEM MAP FORMAT'
    struct.pack_into('<H', hdr, 32, 2)            # version
    struct.pack_into('<I', hdr, 34, 0)            # gisFileId
    struct.pack_into('<H', hdr, 38, 1)            # projection: y decreases top to bottom
    struct.pack_into('<I', hdr, 40, 0)            # attrTable
    struct.pack_into('<H', hdr, 44, 1)            # mapType
    struct.pack_into('<I', hdr, 46, 1)            # byteOrder
    struct.pack_into('<HH', hdr, 64, 0xeb, 0x5a)  # VS_SCALAR, CR_REAL4
    struct.pack_into('<f', hdr, 68, float(valid.min()) if valid.size else 0.0)
    struct.pack_into('<f', hdr, 76, float(valid.max()) if valid.size else 0.0)
    struct.pack_into('<dd', hdr, 84, x_ul, y_ul)
    struct.pack_into('<II', hdr, 100, a.shape[0], a.shape[1])
    struct.pack_into('<ddd', hdr, 108, cell, cell, 0.0)
    with open(path, 'wb') as f:
        f.write(bytes(hdr))
        f.write(raw.tobytes())
