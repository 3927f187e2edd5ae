"""Target-grid coordinate maps for the Interpolator: the host-side stand-in for the reference's
`LatLong` (main/interpolation/latlong.py:9-39) and `Dem` (:42-57).

The reference reads PCRaster maps through GDAL and netCDF maps through netCDF4; neither library is part of
this path ("stays on host").  What the interpolation needs from them is small - the 2-D
latitude / longitude arrays WITH THEIR DTYPE (PCRaster REAL4 maps make the whole build run in float32,
SURVEY App. A.3), the missing value, and the identifier string that goes into the intertable id - so this
module reads PCRaster CSF files directly (256-byte header + row-major cells) and otherwise accepts arrays
or `.npy` files.  netCDF target maps have to be handed over as arrays by the caller.
"""
import os
import struct

import numpy as np

from .exceptions import NOT_EXISTING_MAPS, ApplicationException

PCR_MV_F32 = -3.4028234663852886e38       # what GDAL reports (and writes into the array) for REAL4 MV cells
NC_FILL = {'f4': 9.969209968386869e36, 'f8': 9.969209968386869e36}

_CELL_REPR = {0x5a: ('<f4', 4), 0xdb: ('<f8', 8), 0x26: ('<i4', 4), 0x00: ('u1', 1), 0x04: ('i1', 1),
              0x15: ('<i2', 2), 0x11: ('<u2', 2), 0x22: ('<u4', 4)}


class PCRasterMap:
    """Minimal PCRaster CSF v2 reader: `.values`, `.mv`, `.identifier()` as readers/pcr.py:9-42."""

    def __init__(self, path):
        path = os.fspath(path)
        if not os.path.exists(path):
            raise ApplicationException.get_exc(NOT_EXISTING_MAPS, path)
        with open(path, 'rb') as f:
            raw = f.read()
        if not raw.startswith(b'RUU CROSS SYSTEM MAP FORMAT'):
            raise ValueError(f'{path}: not a PCRaster CSF map')
        cell_repr = struct.unpack_from('<H', raw, 66)[0]
        dtype, size = _CELL_REPR[cell_repr]
        self._origX, self._origY = struct.unpack_from('<dd', raw, 84)
        self._rows, self._cols = struct.unpack_from('<II', raw, 100)
        cell_x, cell_y = struct.unpack_from('<dd', raw, 108)
        self._pxlW, self._pxlH = cell_x, -cell_y            # GDAL geotransform convention (north-up)
        self.min = float(np.frombuffer(raw, dtype, 1, 68)[0])
        self.max = float(np.frombuffer(raw, dtype, 1, 76)[0])
        data = np.frombuffer(raw, dtype, self._rows * self._cols, 256).reshape(self._rows, self._cols).copy()
        if dtype == '<f4':
            self.mv = PCR_MV_F32
            data[data.view(np.uint32) == 0xffffffff] = np.float32(PCR_MV_F32)
        elif dtype == '<f8':
            self.mv = -1.7976931348623157e308
            data[data.view(np.uint64) == 0xffffffffffffffff] = self.mv
        else:
            self.mv = {'<i4': -2147483648, 'u1': 255, 'i1': -128, '<i2': -32768, '<u2': 65535,
                       '<u4': 4294967295}[dtype]
        self.values = data

    def identifier(self):
        return (f'{int(self._origX)}_{int(self._origY)}_{int(self._pxlW)}_{int(self._pxlH)}_'
                f'{self.min:.2f}_{self.max:.2f}')

    def close(self):
        pass


def _array_identifier(a):
    a = np.asarray(a)
    fin = a[np.abs(a) < 1e30] if a.dtype.kind == 'f' else a
    return f'{a.shape[0]}x{a.shape[-1]}_{float(fin.min()):.2f}_{float(fin.max()):.2f}'


class LatLong:
    """`.lats`, `.lons` (2-D, dtype preserved), `.mv`, `.identifier` - the attributes Interpolator reads."""

    def __init__(self, lat_map, long_map, mv=None, identifier=None):
        if isinstance(lat_map, (str, os.PathLike)) and os.fspath(lat_map).endswith('.npy'):
            lat_map, long_map = np.load(lat_map), np.load(long_map)
        if isinstance(lat_map, (str, os.PathLike)):
            if os.fspath(lat_map).endswith('.nc'):
                raise ApplicationException.get_exc(
                    NOT_EXISTING_MAPS, 'netCDF target maps: read them with netCDF4 on the host and pass the 2-D '
                                       'lat/lon arrays (pyg2p_b200 does not link netCDF4)')
            r1, r2 = PCRasterMap(lat_map), PCRasterMap(long_map)
            self.lats, self.lons, self.mv = r1.values, r2.values, r1.mv
            self._id = f'{r1.identifier()}_{r2.identifier()}'
        else:
            self.lats, self.lons = np.asarray(lat_map), np.asarray(long_map)
            if self.lats.shape != self.lons.shape:
                raise ValueError('lat and lon maps differ in shape')
            if self.lats.ndim == 1:
                self.lats, self.lons = self.lats.reshape(1, -1), self.lons.reshape(1, -1)
            if mv is None:   # PCRaster REAL4 convention for float32 maps, netCDF default fill otherwise
                mv = PCR_MV_F32 if self.lats.dtype == np.float32 else NC_FILL['f8']
            self.mv = mv
            self._id = identifier or f'{_array_identifier(self.lats)}_{_array_identifier(self.lons)}'

    @property
    def identifier(self):
        return self._id


class Dem:
    """`.values`, `.mv` of the altitude map used by the Corrector (latlong.py:42-57)."""

    def __init__(self, dem_map, mv=None):
        if isinstance(dem_map, (str, os.PathLike)):
            r = PCRasterMap(dem_map)
            self.values, self.mv = r.values, r.mv
        else:
            self.values = np.asarray(dem_map)
            self.mv = PCR_MV_F32 if mv is None else mv
