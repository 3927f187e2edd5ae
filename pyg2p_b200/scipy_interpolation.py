"""Drop-in for the reference's `ScipyInterpolation` (main/interpolation/scipy_interpolation_lib.py:516-727,
:774-812, :1018-1076) for the modes `nearest`, `invdist` and `adw` (Shepard, k = 11): same constructor signature, same
`interpolate(lonefas, latefas) -> (result, weights, indexes)` with the same shapes and dtypes, computed by
the CUDA library (cube-sphere index + k-NN + weights) instead of scipy's cKDTree and numpy.

No CPU fallback: the other modes (cdd, bilinear, triangulation, bilinear_delaunay) are outside this
package's scope and raise INVALID_INTERPOL_METHOD, as the reference does for unknown modes (:655).
"""
import numpy as np
import numpy.ma as ma
import torch

from . import _lib as L
from . import device, parallel
from .exceptions import INVALID_INTERPOL_METHOD, WEIRD_STUFF, ApplicationException

_MODES = {'nearest': (L.MODE_NEAREST, 1), 'invdist': (L.MODE_INVDIST, None), 'adw': (L.MODE_RAW, 11)}


class ScipyInterpolation:
    # ecCodes >= 1.14.3 hands over unrotated coordinates, so sources always take to_3d's default branch (:551)
    rotated_bugfix_gribapi = True

    def __init__(self, longrib, latgrib, grid_details, source_values, nnear, mv_target, mv_source,
                 target_is_rotated=False, parallel=False, mode='nearest', cdd_map='', cdd_mode='', cdd_options=None,
                 use_broadcasting=False, num_of_splits=None, device_index=None):
        if mode not in _MODES or (_MODES[mode][1] is not None and nnear != _MODES[mode][1]):
            raise ApplicationException.get_exc(
                INVALID_INTERPOL_METHOD, f'interpolation method not supported (mode = {mode}, nnear = {nnear})')
        self.geodetic_info = grid_details
        self.source_grid_is_rotated = 'rotated' in (grid_details.get('gridType') or '')
        self.target_grid_is_rotated = target_is_rotated
        self.njobs = 1 if not parallel else -1          # kept for interface parity; the GPU is always "parallel"
        self.longrib, self.latgrib = longrib, latgrib
        self.nnear, self.mode = nnear, mode
        self.num_of_splits = num_of_splits
        self._mv_target, self._mv_source = mv_target, mv_source
        self.z = np.asarray(ma.getdata(source_values), dtype=np.float64).ravel()
        n = np.asarray(longrib).size
        if n != self.z.size:
            # the reference builds this exception and drops it (:553-556); a size mismatch cannot work here
            raise ApplicationException.get_exc(WEIRD_STUFF, f'len(coordinates) {n} != len(values) {self.z.size}')
        dev = None if device_index is None else torch.device('cuda', device_index)
        rot = None if self.rotated_bugfix_gribapi else self._south_pole()
        self.index = device.SourceIndex(np.asarray(longrib, dtype=np.float64), np.asarray(latgrib, dtype=np.float64),
                                        grid_details.get('radius'), rotation=rot, device=dev)
        # self query k=2 -> max -> dmax + dmax*4/Nj (:568-570); "not used in adw (Shepard) algorithm" (:561-562)
        self.min_upper_bound = None if mode == 'adw' else self.index.min_upper_bound(grid_details.get('Nj'))
        self.ref_radius = None
        self.table = None
        self.flags = None

    def _south_pole(self):
        return (self.geodetic_info.get('latitudeOfSouthernPoleInDegrees'),
                self.geodetic_info.get('longitudeOfSouthernPoleInDegrees'))

    def interpolate(self, lonefas, latefas):
        """-> (result[M], weights[M] | [M,k], indexes[M] | [M,k]) as numpy arrays, like the reference (:572-663).
        `num_of_splits` (row chunks, :588-607) needs no chunking on a 180 GB device; with an initialised
        torch.distributed process group the target rows are split over the ranks instead and the shards are
        all-gathered (pyg2p_b200/parallel.py)."""
        lonefas, latefas = np.asarray(lonefas), np.asarray(latefas)
        table = self.build_table(lonefas, latefas)
        indexes, weights = table.indexes_coeffs()
        src = torch.from_numpy(self.z).to(table.device)
        res = table.apply(src, float(self._mv_target)).cpu().numpy()
        result = ma.masked_array(res, fill_value=self._mv_target, copy=False)
        return result, weights, indexes

    def build_table(self, lonefas, latefas):
        """device-resident SoA table (pyg2p_b200.device.DeviceTable) for the given target maps."""
        mode_code, _ = _MODES[self.mode]
        rot = self._south_pole() if self.target_grid_is_rotated else None
        m = lonefas.size
        cols = lonefas.shape[-1] if lonefas.ndim > 1 else m
        rank, ws = parallel.world()
        lo, hi = parallel.shard_bounds(m, rank, ws, align=cols) if ws > 1 else (0, m)
        flat_lon, flat_lat = lonefas.reshape(-1)[lo:hi], latefas.reshape(-1)[lo:hi]
        out = self.index.knn(flat_lon, flat_lat, k=self.nnear, mode=mode_code, mub=self.min_upper_bound or 0.0,
                             radius=self.geodetic_info.get('radius'), rotation=rot, want_flags=True)
        idx, coef = out['idx'], out['coef']
        self.flags = out['flags']
        if self.mode == 'adw':
            # Shepard weights from the raw neighbours (:817-838, :953-1052).  The reference radius is the mean 7th distance
            # over ALL targets - of this build (:339), or of the k=7 query a split build runs first (:576-585): the same
            # numbers, so one global value serves both (the ranks add their partial sums)
            code = out['coef_dtype']
            dev = idx.device
            self.ref_radius = device.adw_reference_radius(coef, code, m_total=m)
            t_lat = device.to_device(flat_lat, torch.float32 if code == L.F32 else torch.float64, dev)
            t_lon = device.to_device(flat_lon, torch.float32 if code == L.F32 else torch.float64, dev)
            s_lat = device.to_device(np.asarray(self.latgrib, dtype=np.float64).ravel(), torch.float64, dev)
            s_lon = device.to_device(np.asarray(self.longrib, dtype=np.float64).ravel(), torch.float64, dev)
            device.weights_adw(idx, coef, t_lat, t_lon, s_lat, s_lon, self.ref_radius, code)
        if ws > 1:
            idx, coef = parallel.all_gather_table(idx, coef, m, align=cols)
        # float32 maps: nearest stores float32 distances (:627-635), adw float32 weights (`w / sums` of float32 arrays,
        # :1019-1021); invdist stores float32-valued weights into z.dtype = float64 arrays (:787)
        f32 = lonefas.dtype == np.float32 and self.mode in ('nearest', 'adw')
        shape = lonefas.shape if lonefas.ndim == 2 else None
        self.table = device.DeviceTable(idx, coef, self.z.size, coef_f32=f32, grid_shape=shape)
        return self.table

    def tie_rows(self):
        """number of target rows (of this rank's shard) whose best k+1 distances contain an exact float64 tie:
        on those rows the neighbour ORDER is a cKDTree artefact and may differ."""
        return int((self.flags & L.FLAG_TIE).ne(0).sum().item()) if self.flags is not None else 0
