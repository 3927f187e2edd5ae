"""Minimal GRIB edition 1 reader: the Grid Description Section of `regular_ll` (type 0) and `rotated_ll`
(type 10) messages -> 1-D latitude / longitude arrays in scanning order plus the geodetic keys the
interpolation reads (`radius`, `Nj`, `gridType`, south pole of rotation) - what the reference obtains from
ecCodes through `GribGridDetails` (src/pyg2p/__init__.py:60-144) - and the values of simply packed grid-point
data (`read_values`).  ecCodes stays the decoder in production; this module exists so that intertables can be
(re)built and checked where ecCodes is not installed, e.g. to regenerate the reference's golden table from its own
`tests/data/input.grib` and to run the reference's own test scenarios on its own messages.

Rotated grids are unrotated the way ecCodes' lat/lon iterator does since 1.14.3 (pole rotation on the unit
sphere, result rounded to 6 decimals in single precision).
"""
import numpy as np

EARTH_RADIUS_GRIB1 = 6367470.0          # GRIB1 spherical earth


def _s3(b):
    v = int.from_bytes(b, 'big')
    return -(v & 0x7fffff) if v & 0x800000 else v


def _s2(b):
    v = int.from_bytes(b, 'big')
    return -(v & 0x7fff) if v & 0x8000 else v


def unrotate(lat_r, lon_r, south_pole_lat, south_pole_lon, angle=0.0):
    d2r = np.pi / 180.0
    latr, lonr = np.asarray(lat_r) * d2r, np.asarray(lon_r) * d2r
    xd, yd, zd = np.cos(lonr) * np.cos(latr), np.sin(lonr) * np.cos(latr), np.sin(latr)
    t, o = -(90.0 + south_pole_lat), -south_pole_lon
    st, ct, so, co = np.sin(d2r * t), np.cos(d2r * t), np.sin(d2r * o), np.cos(d2r * o)
    x = ct * co * xd + so * yd + st * co * zd
    y = -ct * so * xd + co * yd - st * so * zd
    z = np.clip(-st * xd + ct * zd, -1.0, 1.0)
    lat = np.degrees(np.arcsin(z))
    lon = np.degrees(np.arctan2(y, x))
    lat = np.round(np.float32(lat * 1e6)).astype(np.float64) / 1e6
    lon = np.round(np.float32(lon * 1e6)).astype(np.float64) / 1e6
    return lat, lon - angle


class GridDetails(dict):
    def get(self, key, default=None):
        return dict.get(self, key, default)


def read_grid(source):
    """`source`: path or bytes of a GRIB1 file (first message).  -> (lats[n], lons[n], GridDetails)."""
    raw = source if isinstance(source, (bytes, bytearray)) else open(source, 'rb').read(4096)
    if raw[:4] != b'GRIB' or raw[7] != 1:
        raise ValueError('not a GRIB edition 1 message')
    pds_len = int.from_bytes(raw[8:11], 'big')
    if not raw[8 + 7] & 0x80:
        raise ValueError('message has no grid description section')
    g = raw[8 + pds_len:]
    grid_type = g[5]
    if grid_type not in (0, 10):
        raise ValueError(f'GRIB1 data representation type {grid_type} is not regular_ll / rotated_ll')
    ni, nj = int.from_bytes(g[6:8], 'big'), int.from_bytes(g[8:10], 'big')
    la1, lo1, la2, lo2 = _s3(g[10:13]) / 1e3, _s3(g[13:16]) / 1e3, _s3(g[17:20]) / 1e3, _s3(g[20:23]) / 1e3
    scan = g[27]
    i_neg, j_pos, j_consecutive = bool(scan & 0x80), bool(scan & 0x40), bool(scan & 0x20)
    if lo2 < lo1 and not i_neg:
        lo2 += 360.0
    lon = lo1 + np.arange(ni) * ((lo2 - lo1) / (ni - 1) if ni > 1 else 0.0)
    lat = la1 + np.arange(nj) * ((la2 - la1) / (nj - 1) if nj > 1 else 0.0)
    if j_consecutive:
        lats, lons = np.meshgrid(lat, lon)          # adjacent points in j
    else:
        lons, lats = np.meshgrid(lon, lat)          # adjacent points in i (the usual case)
    lats, lons = lats.ravel(), lons.ravel()
    details = GridDetails(gridType='rotated_ll' if grid_type == 10 else 'regular_ll', radius=EARTH_RADIUS_GRIB1, Nj=nj,
                          Ni=ni, numberOfValues=ni * nj, jScansPositively=j_pos)
    if grid_type == 10:
        sp_lat, sp_lon = _s3(g[32:35]) / 1e3, _s3(g[35:38]) / 1e3
        details.update(latitudeOfSouthernPoleInDegrees=sp_lat, longitudeOfSouthernPoleInDegrees=sp_lon)
        lats, lons = unrotate(lats, lons, sp_lat, sp_lon)
    # the reference's grid id: first / last longitude as coded, Ni, Nj, number of values, grid type
    # (src/pyg2p/__init__.py:104-111), e.g. '-15.75$16.125$511$415$212065$rotated_ll'
    details['grid_id'] = f"{lo1}${_s3(g[20:23]) / 1e3}${ni}${nj}${ni * nj}${details['gridType']}"
    return lats, lons, details


def _ibm32(b):
    """IBM System/360 single precision (GRIB1 reference value): sign, 7-bit base-16 exponent, 24-bit fraction."""
    v = int.from_bytes(b, 'big')
    sign = -1.0 if v & 0x80000000 else 1.0
    return sign * (v & 0xffffff) / 16777216.0 * 16.0 ** (((v >> 24) & 0x7f) - 64)


def split_messages(raw):
    """byte strings of the GRIB1 messages in `raw` (a whole file)."""
    out, i = [], 0
    while True:
        i = raw.find(b'GRIB', i)
        if i < 0 or i + 8 > len(raw):
            return out
        n = int.from_bytes(raw[i + 4:i + 7], 'big')
        if raw[i + 7] == 1 and raw[i + n - 4:i + n] == b'7777':
            out.append(raw[i:i + n])
            i += n
        else:
            i += 4


def read_values(message, missing_value=9999.0):
    """values of one GRIB1 message with simple packing of grid-point data (BDS flag 0), as ecCodes' `values` key:
    (reference + X * 2**E) * 10**-D, points cleared in the bitmap section get `missing_value`.
    -> (values float64 [n], meta dict with the product-definition keys the reference selects messages by)."""
    raw = bytes(message)
    if raw[:4] != b'GRIB' or raw[7] != 1:
        raise ValueError('not a GRIB edition 1 message')
    pds = raw[8:]
    pds_len = int.from_bytes(pds[0:3], 'big')
    flag = pds[7]
    meta = dict(table2Version=pds[3], centre=pds[4], indicatorOfParameter=pds[8], indicatorOfTypeOfLevel=pds[9],
                level=int.from_bytes(pds[10:12], 'big'), unitOfTimeRange=pds[17], P1=pds[18], P2=pds[19],
                timeRangeIndicator=pds[20], decimalScaleFactor=_s2(pds[26:28]))
    at = 8 + pds_len
    if flag & 0x80:
        at += int.from_bytes(raw[at:at + 3], 'big')             # grid description section
    bitmap = None
    if flag & 0x40:
        bms_len = int.from_bytes(raw[at:at + 3], 'big')
        if int.from_bytes(raw[at + 4:at + 6], 'big') != 0:
            raise ValueError('predefined bitmaps are not supported')
        bits = np.unpackbits(np.frombuffer(raw, dtype=np.uint8, count=bms_len - 6, offset=at + 6))
        bitmap = bits[:bits.size - raw[at + 3]].astype(bool)
        at += bms_len
    bds_len = int.from_bytes(raw[at:at + 3], 'big')
    bflag, unused = raw[at + 3] >> 4, raw[at + 3] & 0x0f
    if bflag & 0b1100:
        raise ValueError('only simple packing of grid-point data is supported')
    e_scale = _s2(raw[at + 4:at + 6])
    ref = _ibm32(raw[at + 6:at + 10])
    nbits = raw[at + 10]
    n_packed = ((bds_len - 11) * 8 - unused) // nbits if nbits else (int(bitmap.sum()) if bitmap is not None else 0)
    if nbits:
        bits = np.unpackbits(np.frombuffer(raw, dtype=np.uint8, count=bds_len - 11, offset=at + 11))[:n_packed * nbits]
        weights = (1 << np.arange(nbits - 1, -1, -1)).astype(np.uint64)
        x = bits.reshape(n_packed, nbits).astype(np.uint64) @ weights
    else:
        x = np.zeros(n_packed, dtype=np.uint64)
    packed = (ref + x.astype(np.float64) * 2.0 ** e_scale) * 10.0 ** (-meta['decimalScaleFactor'])
    if bitmap is None:
        return packed, meta
    values = np.full(bitmap.size, float(missing_value))
    values[bitmap] = packed[:int(bitmap.sum())]
    return values, meta

