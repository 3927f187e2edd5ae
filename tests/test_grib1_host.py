"""CPU tests of the minimal GRIB1 reader (pyg2p_b200/grib1.py) and of the reference fixtures it decodes."""
import os
import struct

import numpy as np

from pyg2p_b200 import grib1


def _ibm32_bytes(x):
    """encode a float as IBM System/360 single precision (exact for the values used here)."""
    if x == 0.0:
        return b'\x00\x00\x00\x00'
    sign = 0x80 if x < 0 else 0
    x = abs(x)
    e = 64
    while x >= 1.0:
        x /= 16.0
        e += 1
    while x < 1.0 / 16.0:
        x *= 16.0
        e -= 1
    return bytes([sign | e]) + int(round(x * 16777216.0)).to_bytes(3, 'big')


def _message(values_int, ref, e_scale, d_scale, nbits, bitmap=None, ni=4, nj=3):
    """a GRIB1 message: PDS (28 bytes), regular_ll GDS (32 bytes), optional BMS, BDS with simple packing."""
    def s2(v):
        return ((0x8000 | -v) if v < 0 else v).to_bytes(2, 'big')
    pds = bytearray(28)
    pds[0:3] = (28).to_bytes(3, 'big')
    pds[3], pds[4], pds[7], pds[8], pds[9] = 2, 98, 0x80 | (0x40 if bitmap is not None else 0), 11, 105
    pds[10:12] = (2).to_bytes(2, 'big')
    pds[26:28] = s2(d_scale)
    gds = bytearray(32)
    gds[0:3] = (32).to_bytes(3, 'big')
    gds[5] = 0
    gds[6:8], gds[8:10] = ni.to_bytes(2, 'big'), nj.to_bytes(2, 'big')
    gds[10:13] = (50000).to_bytes(3, 'big')
    gds[13:16] = (0).to_bytes(3, 'big')
    gds[17:20] = (48000).to_bytes(3, 'big')
    gds[20:23] = (3000).to_bytes(3, 'big')
    bms = b''
    if bitmap is not None:
        bits = np.packbits(np.asarray(bitmap, dtype=np.uint8))
        unused = bits.size * 8 - len(bitmap)
        body = bytes([unused]) + (0).to_bytes(2, 'big') + bits.tobytes()
        bms = (len(body) + 3).to_bytes(3, 'big') + body
    stream = ''.join(format(int(v), f'0{nbits}b') for v in values_int)
    unused = (-len(stream)) % 8
    stream += '0' * unused
    data = int(stream, 2).to_bytes(len(stream) // 8, 'big') if stream else b''
    bds = bytes([unused]) + s2(e_scale) + _ibm32_bytes(ref) + bytes([nbits]) + data
    bds = (len(bds) + 3).to_bytes(3, 'big') + bds
    body = bytes(pds) + bytes(gds) + bms + bds + b'7777'
    return b'GRIB' + (len(body) + 8).to_bytes(3, 'big') + b'\x01' + body


def test_simple_packing_roundtrip():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2 ** 11, size=12)
    msg = _message(x, ref=273.0, e_scale=-3, d_scale=1, nbits=11)
    v, meta = grib1.read_values(msg)
    np.testing.assert_array_equal(v, (273.0 + x * 2.0 ** -3) * 10.0 ** -1)
    assert (meta['indicatorOfParameter'], meta['indicatorOfTypeOfLevel'], meta['level'], meta['decimalScaleFactor']) == (11, 105, 2, 1)
    lats, lons, det = grib1.read_grid(msg)
    assert lats.size == 12 and det['gridType'] == 'regular_ll' and det['Nj'] == 3
    # bitmap: 12 grid points, 9 packed values, the others come out as the missing value
    bitmap = np.array([1, 1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1], bool)
    msg = _message(x[:9], ref=-16.5, e_scale=2, d_scale=0, nbits=13, bitmap=bitmap)
    v, _ = grib1.read_values(msg, missing_value=9999.0)
    want = np.full(12, 9999.0)
    want[bitmap] = -16.5 + x[:9] * 4.0
    np.testing.assert_array_equal(v, want)
    assert grib1.split_messages(b'junk' + msg + msg + b'tail') == [msg, msg]
    # constant field: zero bits per value
    v, _ = grib1.read_values(_message([], ref=5.0, e_scale=0, d_scale=0, nbits=0, bitmap=bitmap))
    np.testing.assert_array_equal(v[bitmap], 5.0)


def test_reference_fixtures(golden_dir):
    """the messages and maps of the reference's own tests, as committed by make_golden.py."""
    with open(os.path.join(golden_dir, 'ref_input_2t_step0.grib'), 'rb') as f:
        msg = f.read()
    v, meta = grib1.read_values(msg)
    lats, lons, det = grib1.read_grid(msg)
    assert v.size == lats.size == 511 * 415 and det['grid_id'] == '-15.75$16.125$511$415$212065$rotated_ll'
    assert (meta['centre'], meta['indicatorOfParameter'], meta['indicatorOfTypeOfLevel'], meta['level'], meta['P1']) == (80, 11, 105, 2, 0)
    assert 275.0 < v.min() < v.max() < 313.0                       # 2 m temperature over Europe, K
    with open(os.path.join(golden_dir, 'ref_geopotential.grib'), 'rb') as f:
        z, zmeta = grib1.read_values(f.read())
    assert z.size == 212065 and zmeta['indicatorOfParameter'] == 6 and -900.0 < z.min() and 34000.0 < z.max() < 35000.0
    dem = np.load(os.path.join(golden_dir, 'ref_efas_dem_map.npz'))
    d, mv = dem['dem'], float(dem['mv'])
    valid = d[d != np.float32(mv)]
    # what reference tests/test_readers.py:52-54 asserts for tests/data/dem.map
    assert d.shape == (810, 680) and valid.min() == -9.25 and valid.max() == 3539.3720703125 and mv == -3.4028234663852886e+38
