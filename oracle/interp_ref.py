"""Oracle (test infrastructure): numpy/scipy restatement of the reference's intertable
build and apply.  Citations are to /root/reference/src/pyg2p/main/interpolation/.

Not imported by the product package.  See oracle/__init__.py.
"""
from math import radians, cos, sin

import numpy as np
import numpy.ma as ma
from scipy.spatial import cKDTree

from . import libm32

LEAFSIZE = 30  # scipy_interpolation_lib.py:559


# --------------------------------------------------------------------------- xyz
def to_3d(lons, lats, radius, rotation=None):
    """scipy_interpolation_lib.py:665-696.

    Default branch: `r*cos(lon)*cos(lat)`, `r*sin(lon)*cos(lat)`, `r*sin(lat)` evaluated
    left to right, `r` a double constant (numexpr promotes f32 operands to double when a
    Python float is involved), result cast back to float32 when the inputs are float32
    (:691-694).  `rotation=(lat_sp, lon_sp)` selects the `to_regular` branch (:673-678),
    which works on the UNIT sphere (no radius factor).
    """
    lons = np.radians(lons)
    lats = np.radians(lats)
    # numexpr: float32 operands go through libm sinf/cosf, float64 through libm sin/cos (== numpy's here)
    cx = libm32.cos(lons)
    sx = libm32.sin(lons)
    cy = libm32.cos(lats)
    sy = libm32.sin(lats)
    if rotation is None:
        r = np.float64(radius)
        x = (r * cx) * cy
        y = (r * sx) * cy
        z = r * sy
    else:
        lat_sp, lon_sp = rotation
        teta = -radians(90 + lat_sp)
        fi = -radians(lon_sp)
        ux = cx * cy
        uy = sx * cy
        uz = sy
        x = (np.float64(cos(teta) * cos(fi)) * ux + np.float64(sin(fi)) * uy) + np.float64(sin(teta) * cos(fi)) * uz
        y = (np.float64(cos(teta) * sin(fi)) * ux + np.float64(cos(fi)) * uy) - np.float64(sin(teta) * sin(fi)) * uz
        z = np.float64(-sin(teta)) * ux + np.float64(cos(teta)) * uz
    if lons.dtype == np.float32:
        x = np.float32(x)
        y = np.float32(y)
        z = np.float32(z)
    return x, y, z


def stack_xyz(x, y, z):
    """`np.vstack((x.ravel(), y.ravel(), z.ravel())).T` (:552, :625)."""
    return np.vstack((x.ravel(), y.ravel(), z.ravel())).T


# --------------------------------------------------------------------------- k-NN
def build_tree(source_xyz):
    return cKDTree(source_xyz, leafsize=LEAFSIZE)


def min_upper_bound(tree, source_xyz, nj, workers=1):
    """Self query k=2, `dmax + dmax*4/Nj` with dmax over BOTH columns (:568-570)."""
    d, _ = tree.query(source_xyz, k=2, workers=workers)
    dmax = np.max(d)
    return dmax + dmax * 4 / nj


def query(tree, target_xyz, k, workers=1):
    """:626-628 - distances are cast to float32 when the target xyz are float32."""
    d, i = tree.query(target_xyz, k=k, workers=workers)
    if target_xyz.dtype == np.float32:
        d = np.float32(d)
    return d, i


def brute_knn(source_xyz, target_xyz, k):
    """Exhaustive k-NN in the cKDTree distance arithmetic:
    d2 = ((dx*dx)+(dy*dy))+(dz*dz) in float64, ties broken by lower source index.
    Independent check of cKDTree on small cases; also returns a per-row flag for exact
    float64 ties among the best k+1 squared distances."""
    s = np.asarray(source_xyz, dtype=np.float64)
    t = np.asarray(target_xyz, dtype=np.float64)
    m = t.shape[0]
    idx = np.empty((m, k), dtype=np.int64)
    dist = np.empty((m, k), dtype=np.float64)
    tie = np.zeros(m, dtype=bool)
    for j in range(m):
        dx = t[j, 0] - s[:, 0]
        dy = t[j, 1] - s[:, 1]
        dz = t[j, 2] - s[:, 2]
        d2 = ((dx * dx) + (dy * dy)) + (dz * dz)
        order = np.lexsort((np.arange(s.shape[0]), d2))[:k + 1]
        idx[j] = order[:k]
        dist[j] = np.sqrt(d2[order[:k]])
        dd = d2[order]
        tie[j] = bool(np.any(dd[1:] == dd[:-1]))
    return dist, idx, tie


# --------------------------------------------------------------------------- weights
def build_nn_loop(distances, indexes, z, mub, mv_target):
    """`_build_nn` as the reference runs it: a per-point Python loop (:698-727).
    Kept literal because it is what the CPU baseline times."""
    n = z.size
    result = ma.masked_array(data=np.empty((len(distances),)), fill_value=mv_target, copy=False)
    idxs = np.full((len(indexes),), n, dtype=int)
    j = 0
    for dist, ix in zip(distances, indexes):
        if dist <= mub:
            wz = z[ix]
            idxs[j] = ix
        else:
            wz = mv_target
        result[j] = wz
        j += 1
    return result, idxs


def build_nn(distances, indexes, z, mub, mv_target):
    """Vectorised equivalent of `build_nn_loop` (checked equal in tests); `coeffs` of a
    nearest table are the raw distances for every row (:634-635)."""
    n = z.size
    inside = distances <= mub
    idxs = np.where(inside, indexes, n).astype(int)
    result = np.where(inside, z[np.minimum(indexes, n - 1)], mv_target)
    return result, idxs


def build_weights_invdist(distances, indexes, z, mub, mv_target, nnear):
    """`_build_weights_invdist(adw_type=None)` (:774-812, :1018-1026, :1036, :1054-1056, :1076).

    A: d0 <= 1e-10 -> w = [1,0,..], idx = cKDTree row.  B: ~A & d0 <= mub -> w = 1/d**2
    over ALL k columns, normalised by the row sum.  Else: w = NaN, idx = N.
    Arithmetic dtype follows `distances` (float32 for float32 target maps); stored into
    `z.dtype` weights."""
    n = z.size
    m = len(distances)
    result = ma.masked_array(data=np.empty((m,), dtype=z.dtype), fill_value=mv_target, copy=False)
    weights = np.full((m, nnear), np.nan, dtype=z.dtype)
    idxs = np.full((m, nnear), n, dtype=int)
    a = distances[:, 0] <= 1e-10
    b = np.logical_and(~a, distances[:, 0] <= mub)
    first = np.zeros(nnear)
    first[0] = 1
    weights[a] = first
    idxs[a] = indexes[a]
    result[a] = z[indexes[a][:, 0]]
    w = np.zeros_like(distances)
    w[b] = 1. / distances[b] ** 2
    sums = np.sum(w[b], axis=1, keepdims=True)
    weights[b] = w[b] / sums
    wz = np.einsum('ij,ij->i', weights, z[indexes])
    idxs[b] = indexes[b]
    result[b] = wz[b]
    return result, weights, idxs


def adw_cutoff_weights(distances, ref_radius=None):
    """`adw_compute_weights_from_cutoff_distances` (:322-381), Shepard 1968 cut-off radii.

    r_ref = mean of the 7th distances (or the caller's global value, :579-585); r' = d(5) where d(4) > r_ref, r_ref
    where d(4) <= r_ref < d(10), else d(11); s = 1/d for d <= r'/3, (27/(4r'))*(d/r' - 1)^2 for r'/3 < d <= r', else 0;
    w = s^2.  The reference runs this under numba with parallel=True: its mean is a parallel reduction whose rounding
    depends on the number of threads (r_ref moves by a few ulp between machines, and weights of stations near the cut-off
    radius - d/r' - 1 cancels - move by up to ~1e-11 relative with it).  At NUMBA_NUM_THREADS=1 the mean is the sequential
    float64 sum restated here; the golden vectors were generated that way.  float32 maps: distances, r', s and w are
    float32 arrays, the arithmetic between them follows numba's typing (see the comments below)."""
    f32 = distances.dtype == np.float32
    # numba at one thread: a sequential float64 sum (float32 distances included); the global value otherwise
    r_ref = np.cumsum(distances[:, 6], dtype=np.float64)[-1] / len(distances) if ref_radius is None else ref_radius
    r = distances[:, 10].copy()
    f1 = distances[:, 3] > r_ref
    r[f1] = distances[f1, 4]
    f2 = np.logical_and(~f1, distances[:, 9] > r_ref)
    r[f2] = r_ref
    rb = r.reshape(-1, 1)
    rb64 = rb.astype(np.float64)            # numba typing: float32 array (op) int64 literal -> float64
    s = np.zeros_like(distances)
    near = distances <= rb64 / 3
    with np.errstate(all='ignore'):
        s[near] = (1. / distances.astype(np.float64))[near]
        mid = np.logical_and(~near, distances <= rb)
        # `distances / r` stays float32 for float32 maps (array / array), `- 1` and `** 2` are float64
        s[mid] = ((27 / (4 * rb64)) * (((distances / rb).astype(np.float64) - 1) ** 2))[mid]
    return s ** 2, s, r_ref


def adw_directional(s, indexes, tlat, tlon, slat, slon):
    """the non-broadcasting loop of `_build_weights_invdist` (:953-973): angular term on the lat/lon plane,
    `sum_j (1 - cos(theta_ij)) s_j / sum_j s_j` over j != i; NaN (no other station in range) -> 1 (:988)."""
    k = indexes.shape[1]
    wd = np.zeros_like(s)
    with np.errstate(all='ignore'):
        for i in range(k):
            xj = tlat[:, np.newaxis] - slat[indexes]
            yj = tlon[:, np.newaxis] - slon[indexes]
            dj = np.sqrt(xj ** 2 + yj ** 2)
            cos_theta = (xj[:, i, np.newaxis] * xj + yj[:, i, np.newaxis] * yj) / (dj[:, i, np.newaxis] * dj)
            others = np.concatenate([np.arange(i), np.arange(i + 1, k)])
            cos_theta = cos_theta[:, others]
            sj = s[:, others]
            wd[:, i] = np.sum((1 - cos_theta) * sj, axis=1) / np.sum(sj, axis=1)
    return np.nan_to_num(wd, nan=1)


def build_weights_adw(distances, indexes, z, tlat, tlon, slat, slon, nnear=11, ref_radius=None):
    """`_build_weights_invdist(adw_type='Shepard')` (:817-838, :953-1001, :1019-1021, :1036-1052): no
    min_upper_bound, no sentinel rows; w = s^2 (1 + directional), normalised; d0 <= 1e-10 -> [1, 0, ...]."""
    # cKDTree hands over C-contiguous arrays; numpy's summation order below depends on the memory layout
    distances, indexes = np.ascontiguousarray(distances), np.ascontiguousarray(indexes)
    w, s, r_ref = adw_cutoff_weights(distances, ref_radius)
    wd = adw_directional(s, indexes, np.asarray(tlat).ravel(), np.asarray(tlon).ravel(), np.asarray(slat).ravel(),
                         np.asarray(slon).ravel())
    w = np.multiply(w, 1 + wd)
    with np.errstate(all='ignore'):
        weights = w / np.sum(w, axis=1, keepdims=True)
        result = np.einsum('ij,ij->i', weights, z[indexes])
    exact = distances[:, 0] <= 1e-10
    first = np.zeros(nnear, dtype=weights.dtype)
    first[0] = 1
    weights[exact] = first
    result[exact] = z[indexes[exact][:, 0]]
    return result, weights, indexes


class BuildOracle:
    """`ScipyInterpolation.__init__` + `interpolate` for modes nearest / invdist / adw
    (:523-570, :572-663) with cKDTree as the neighbour search."""

    def __init__(self, longrib, latgrib, grid_details, source_values, nnear, mv_target, mv_source,
                 parallel=False, mode='nearest', source_rotation=None, target_rotation=None):
        self.workers = -1 if parallel else 1
        self.nnear = nnear
        self.mode = mode
        self.mv_target = mv_target
        self.z = source_values
        self.radius = grid_details.get('radius')
        self.target_rotation = target_rotation
        x, y, zz = to_3d(longrib, latgrib, self.radius, rotation=source_rotation)
        self.source_xyz = stack_xyz(x, y, zz)
        self.tree = build_tree(self.source_xyz)
        self.longrib, self.latgrib = longrib, latgrib
        # "not used in adw (Shepard) algorithm" (:561-562)
        self.min_upper_bound = None if mode == 'adw' else min_upper_bound(self.tree, self.source_xyz,
                                                                          grid_details.get('Nj'), self.workers)

    def interpolate(self, lonefas, latefas, python_loop=False, num_of_splits=None):
        x, y, z = to_3d(lonefas, latefas, self.radius, rotation=self.target_rotation)
        self.target_xyz = stack_xyz(x, y, z)
        distances, indexes = query(self.tree, self.target_xyz, self.nnear, self.workers)
        self.raw_distances, self.raw_indexes = distances, indexes
        if self.mode == 'nearest' and self.nnear == 1:
            fn = build_nn_loop if python_loop else build_nn
            result, idxs = fn(distances, indexes, self.z, self.min_upper_bound, self.mv_target)
            return result, distances, idxs
        if self.mode == 'invdist':
            return build_weights_invdist(distances, indexes, self.z, self.min_upper_bound, self.mv_target, self.nnear)
        if self.mode == 'adw' and self.nnear == 11:
            ref_radius = None
            if num_of_splits is not None:
                # :576-585 - the reference radius of a split build is numpy's mean over a k=7 query of ALL targets; the
                # row blocks (:597-607) are then independent, so they are evaluated in one piece here
                d7, _ = query(self.tree, self.target_xyz, 7, self.workers)
                ref_radius = np.mean(d7[:, 6])
            return build_weights_adw(distances, indexes, self.z, latefas, lonefas, self.latgrib, self.longrib, self.nnear,
                                     ref_radius=ref_radius)
        raise ValueError(f'oracle covers nearest / invdist / adw only, got {self.mode}')


# --------------------------------------------------------------------------- table record
def to_record(indexes, weights):
    """interpolation/__init__.py:251 - array-of-structs {int64 indexes, f8|f4 coeffs}."""
    return np.rec.fromarrays((indexes, weights), names=('indexes', 'coeffs'))


# --------------------------------------------------------------------------- apply
def apply_as_reference(v, weights, indexes, nnear, mv_out):
    """`Interpolator._interpolate_scipy_append_mv` literally (interpolation/__init__.py:142-162).
    Note the quirk: `np.append` drops the mask, so masked inputs
    contribute their payload and the result is a plain ndarray."""
    v = np.append(v, mv_out)
    orig_mask = False if not isinstance(v, ma.core.MaskedArray) else v.mask
    if nnear == 1:
        if isinstance(orig_mask, np.ndarray):
            result = ma.masked_where(orig_mask[indexes], v[indexes], copy=False)
        else:
            result = v[indexes]
    else:
        result = np.einsum('ij,ij->i', weights, v[indexes])
        if isinstance(orig_mask, np.ndarray):
            mask = None
            for i in range(0, indexes.shape[1]):
                mask = orig_mask[indexes.T[i]] if mask is None else mask | orig_mask[indexes.T[i]]
            result = ma.masked_where(mask, result, copy=False)
    return result


def apply_sequential(v, weights, indexes, nnear, mv_out):
    """Same values as `apply_as_reference` but with the evaluation order pinned to
    ((w0*v0 + w1*v1) + w2*v2) + w3*v3, separate multiply and add (this
    is what einsum does on table views loaded from file).  Bit-exact target for the GPU."""
    vv = np.append(np.asarray(v, dtype=np.float64), mv_out)
    if nnear == 1:
        return vv[indexes]
    w = np.asarray(weights, dtype=np.float64)
    acc = w[:, 0] * vv[indexes[:, 0]]
    for j in range(1, indexes.shape[1]):
        acc = acc + w[:, j] * vv[indexes[:, j]]
    return acc


def apply_buffered(v, weights, indexes, nnear, mv_out):
    """The evaluation order of the same einsum (:156) when numpy BUFFERS its operands - float32 coefficients (tables
    built for PCRaster REAL4 maps) or contiguous freshly built arrays: its 2-lane float64 inner loop accumulates the even
    and the odd products separately, i.e. (w0*v0 + w2*v2) + (w1*v1 + w3*v3) for nnear = 4, separate multiply and add
    [probe: bit-equal to np.einsum on 200 000 random rows for both f4 views and contiguous f8].  Other nnear: the einsum
    itself."""
    vv = np.append(np.asarray(v, dtype=np.float64), mv_out)
    if nnear == 1:
        return vv[indexes]
    w = np.asarray(weights, dtype=np.float64)
    if nnear != 4:
        return np.einsum('ij,ij->i', np.ascontiguousarray(w), vv[indexes])
    p = [w[:, j] * vv[indexes[:, j]] for j in range(4)]
    return (p[0] + p[2]) + (p[1] + p[3])


def apply_propagate(v, weights, indexes, nnear, mv_out, src_mask):
    """Intended mask semantics of :148-161 (what the code would do with `ma.append`): the
    output is masked where ANY of the k source points is masked; the sentinel slot N is
    never masked.  Returns (`ma.filled(result, mv_out)`, mask): the data under a mask is
    unspecified in numpy.ma, and every consumer in the reference fills it with the target
    missing value (`result_masked`, util/numeric.py:21-27), so that is the pinned form."""
    # float32 coefficients make the einsum buffer (see apply_buffered); float64 views of a loaded table do not
    order = apply_buffered if np.asarray(weights).dtype == np.float32 else apply_sequential
    data = order(np.asarray(ma.getdata(v)), weights, indexes, nnear, mv_out)
    m = np.append(np.asarray(src_mask, dtype=bool), False)
    if nnear == 1:
        out = m[indexes]
    else:
        out = np.zeros(indexes.shape[0], dtype=bool)
        for j in range(indexes.shape[1]):
            out |= m[indexes[:, j]]
    return np.where(out, mv_out, data), out


# --------------------------------------------------------------------------- grib_* table layouts (apply only)
def _result_masked(values, fill_value):
    """util/numeric.py:21-27."""
    if not isinstance(values, ma.core.MaskedArray):
        return values
    return ma.filled(ma.masked_where(values.mask, values.data, copy=False), fill_value)


def apply_grib_nearest(v, xs, ys, idxs, shape, mv_out):
    """interpolation/__init__.py:286-287, 312: `result.fill(mv); result[xs, ys] = result_masked(v[idxs], mv)`."""
    result = np.empty(shape)
    result.fill(mv_out)
    result[xs, ys] = _result_masked(v[idxs], mv_out)
    return result


def apply_grib_invdist(v, rec, shape, mv_out):
    """:327-328, 348, 369-370: `res = v[i1]*c1 + v[i2]*c2 + v[i3]*c3 + v[i4]*c4; result[xs, ys] = result_masked(res)`."""
    ind, cf = rec['indexes'], rec['coeffs']
    result = np.empty(shape)
    result.fill(mv_out)
    res = v[ind[2]] * cf[0] + v[ind[3]] * cf[1] + v[ind[4]] * cf[2] + v[ind[5]] * cf[3]
    result[ind[0], ind[1]] = _result_masked(res, mv_out)
    return result


def writer_post(values, clone_mask, mv_write, mv_output=None, value_mask=None):
    """What the writers do to an interpolated field before it reaches the file:
    writers/__init__.py:66 `out_v[out_v == mv_output] = nan` (netCDF path only: pass mv_output);
    writers/netcdf.py:149-151 + :170-183 / writers/pcr.py:33 + :44-50: `_mask_values` fills the clone-map mask
    (OR the value mask when the values are a MaskedArray) with the file's missing value, NaNs become it too.
    values: [..., rows, cols]; clone_mask: [rows, cols] bool."""
    v = np.array(values, dtype=np.float64, copy=True)
    bad = np.broadcast_to(np.asarray(clone_mask, dtype=bool), v.shape).copy()
    if value_mask is not None:
        bad |= np.asarray(value_mask, dtype=bool)
    if mv_output is not None:
        bad |= v == mv_output
    bad |= np.isnan(v)
    v[bad] = mv_write
    return v

