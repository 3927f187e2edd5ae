"""Synthetic grids and fields of the shapes named in BASELINE.json.

Host-side input generators only (numpy); nothing here is on the compute path.  The
same bytes feed the CUDA path, the oracle and the CPU baseline so results are
comparable bit for bit.

Source grids mimic what ecCodes hands to pyg2p (`GribGridDetails.latlons`,
reference src/pyg2p/__init__.py:127-133): 1-D float64 `latitudes`/`longitudes`
arrays, longitudes in 0..360, row-major north→south, plus the geodetic keys the
interpolation reads (`radius`, `Nj`, `gridType`; scipy_interpolation_lib.py:529,570,685).
"""
import functools

import numpy as np

EARTH_RADIUS = 6371229.0  # GRIB2 shapeOfTheEarth=6
PCR_MV_F32 = np.float32(-3.4028234663852886e38)  # GDAL nodata of PCRaster REAL4 maps
NC_FILL_F64 = 9.969209968386869e36  # netCDF4.default_fillvals['f8']
GRIB_MV = 9999.0  # ecCodes default missingValue


class GridDetails(dict):
    """dict-backed stand-in for `GribGridDetails` (only `.get(key)` is used on the path)."""

    def get(self, key, default=None):  # same call shape as GribGridDetails.get
        return dict.get(self, key, default)


@functools.lru_cache(maxsize=8)
def _gauss_lats(n_o):
    x, _ = np.polynomial.legendre.leggauss(2 * n_o)
    return np.degrees(np.arcsin(x))[::-1].copy()


def octahedral_grid(n_o):
    """O<n_o> octahedral reduced Gaussian grid: ring i (1-based from each pole) has 4*i+16 points."""
    lats_1d = _gauss_lats(n_o)
    nrings = 2 * n_o
    i = np.arange(1, n_o + 1)
    pl_half = 4 * i + 16
    pl = np.concatenate([pl_half, pl_half[::-1]])
    n = int(pl.sum())
    lats = np.repeat(lats_1d, pl)
    starts = np.concatenate([[0], np.cumsum(pl)[:-1]])
    j = np.arange(n) - np.repeat(starts, pl)
    lons = j * (360.0 / np.repeat(pl, pl))
    details = GridDetails(gridType='reduced_gg', radius=EARTH_RADIUS, Nj=nrings, numberOfValues=n)
    assert lats.size == 4 * n_o * (n_o + 9)
    return lats, lons, details


def regular_ll_grid(lat0, lat1, nlat, lon0, lon1, nlon, radius=EARTH_RADIUS):
    """Regional regular lat/lon source grid (flattened row-major N→S), e.g. a limited-area model."""
    lat = np.linspace(lat0, lat1, nlat)
    lon = np.linspace(lon0, lon1, nlon)
    lons, lats = np.meshgrid(lon, lat)
    details = GridDetails(gridType='regular_ll', radius=radius, Nj=nlat, Ni=nlon, numberOfValues=nlat * nlon)
    return lats.ravel().copy(), lons.ravel().copy(), details


def _inverse_laea(x, y, lat0=52.0, lon0=10.0, x0=4321000.0, y0=3210000.0, r=EARTH_RADIUS):
    """Spherical inverse Lambert azimuthal equal-area (ETRS89-LAEA parameters)."""
    xx = x - x0
    yy = y - y0
    rho = np.hypot(xx, yy)
    c = 2.0 * np.arcsin(np.clip(rho / (2.0 * r), -1.0, 1.0))
    phi0 = np.radians(lat0)
    sin_c, cos_c = np.sin(c), np.cos(c)
    with np.errstate(invalid='ignore', divide='ignore'):
        lat = np.arcsin(cos_c * np.sin(phi0) + np.where(rho > 0, yy * sin_c * np.cos(phi0) / rho, 0.0))
        lon = np.radians(lon0) + np.arctan2(xx * sin_c, rho * np.cos(phi0) * cos_c - yy * np.sin(phi0) * sin_c)
    return np.degrees(lat), np.degrees(lon)


def efas_target(rows=950, cols=1000, cell=5000.0, x_ul=2500000.0, y_ul=5500000.0, dtype=np.float64):
    """EFAS 5 km ETRS89-LAEA target: 2-D lat/lon maps of cell centres."""
    xs = x_ul + (np.arange(cols) + 0.5) * cell
    ys = y_ul - (np.arange(rows) + 0.5) * cell
    x, y = np.meshgrid(xs, ys)
    lat, lon = _inverse_laea(x, y)
    return lat.astype(dtype), lon.astype(dtype)


def latlon_target(res, lat_max=90.0, lat_min=-60.0, dtype=np.float64):
    """GloFAS-style global regular target (cell centres): 0.1 → 1500x3600, 0.05 → 3000x7200."""
    nlon = int(round(360.0 / res))
    nlat = int(round((lat_max - lat_min) / res))
    lon = -180.0 + (np.arange(nlon) + 0.5) * res
    lat = lat_max - (np.arange(nlat) + 0.5) * res
    lons, lats = np.meshgrid(lon, lat)
    return lats.astype(dtype), lons.astype(dtype)


def with_mv_region(lat, lon, frac=0.3, mv=PCR_MV_F32):
    """PCRaster-style maps: a contiguous block of missing-value cells (same cells in both maps)."""
    lat = lat.copy()
    lon = lon.copy()
    rows = lat.shape[0]
    r0 = int(rows * (1.0 - frac))
    lat[r0:, :] = mv
    lon[r0:, :] = mv
    return lat, lon


def _relief(lat_deg, lon_deg):
    la = np.radians(lat_deg)
    lo = np.radians(lon_deg)
    h = 0.5 + 0.25 * np.cos(3 * lo) * np.cos(2 * la) + 0.25 * np.sin(5 * lo + 1.0) * np.sin(4 * la)
    return np.clip(h, 0.0, 1.0)


def field_2t(lats, lons, member=0, step=0):
    rng = np.random.default_rng(20260101 + 1000 * member + step)
    return 288.0 - 45.0 * np.sin(np.radians(lats)) ** 2 + 3.0 * rng.standard_normal(lats.shape)


def field_tp_increment(lats, lons, member=0, step=0):
    """Non-negative per-step precipitation increment (metres); cumulate over steps for `tp`."""
    rng = np.random.default_rng(20260101 + 1000 * member + step)
    smooth = 0.5 * (1.0 + np.cos(2 * np.radians(lons) + 0.3 * step) * np.cos(3 * np.radians(lats)))
    return np.maximum(0.0, rng.standard_normal(lats.shape)) * 2e-3 * smooth


def field_geopotential(lats, lons):
    return 9.81 * np.maximum(0.0, 2500.0 * _relief(lats, lons))


def field_dem(lat2d, lon2d, mv, sea_frac=0.3, seed=7):
    rng = np.random.default_rng(seed)
    h = 2500.0 * _relief(np.asarray(lat2d, dtype=np.float64), np.asarray(lon2d, dtype=np.float64))
    h = h + 20.0 * rng.standard_normal(h.shape)
    sea = rng.random(h.shape) < sea_frac
    h = h.astype(np.asarray(lat2d).dtype)
    h[sea] = mv
    return h


def missing_mask(lats, n_band_deg=5.0, frac=0.02, seed=11):
    """GRIB bitmap: 2 % random points + one latitude band; True = missing."""
    rng = np.random.default_rng(seed)
    m = rng.random(lats.shape) < frac
    m |= (lats > 40.0) & (lats < 40.0 + n_band_deg)
    return m
