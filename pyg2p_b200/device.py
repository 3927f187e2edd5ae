"""Device-side objects of the interpolation path: source index, k-NN/weights build, SoA
intertable and batched apply.  Thin host code over the C ABI in include/g2p.h; torch is used
for device memory, pinned staging buffers and streams only.

There is no CPU fallback: every function here needs libg2pgpu.so and a CUDA device and
raises otherwise.
"""
import concurrent.futures
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

_NP_DTYPE_CODE = {np.dtype('float64'): L.F64, np.dtype('float32'): L.F32}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('pyg2p_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return L.lib()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _bind_device(device):
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    L.check(L.lib().g2p_set_device(dev.index))
    return dev


def to_device(a, dtype=None, device=None):
    """numpy array / tensor -> contiguous device tensor (no copy if already there)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(np.ma.getdata(a))
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=False).contiguous()


class SourceIndex:
    """Cube-sphere spatial index over the source (GRIB) points on one GPU.

    Replaces `KDTree(source_locations, leafsize=30)` of ScipyInterpolation.__init__
    (reference scipy_interpolation_lib.py:551-559).
    """

    def __init__(self, lon=None, lat=None, radius=None, rotation=None, xyz=None, device=None):
        lib = require_cuda()
        self.device = _bind_device(device)
        self._h = C.c_void_p()
        self._lib = lib
        with torch.cuda.device(self.device):
            if xyz is not None:
                x = to_device(xyz, torch.float64, self.device).reshape(-1, 3)
                self.n = x.shape[0]
                L.check(lib.g2p_index_create_xyz(_ptr(x), self.n, C.byref(self._h), _stream()))
            else:
                lo = to_device(np.asarray(lon).ravel() if not isinstance(lon, torch.Tensor) else lon.reshape(-1),
                               torch.float64, self.device)
                la = to_device(np.asarray(lat).ravel() if not isinstance(lat, torch.Tensor) else lat.reshape(-1),
                               torch.float64, self.device)
                if lo.numel() != la.numel():
                    raise ValueError('lon/lat length mismatch')
                self.n = lo.numel()
                L.check(lib.g2p_index_create(_ptr(lo), _ptr(la), self.n, float(radius), L.rotation_arg(rotation),
                                             C.byref(self._h), _stream()))
        self.radius = radius

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.g2p_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        inf = L.IndexInfo()
        L.check(self._lib.g2p_index_get_info(self._h, C.byref(inf)))
        return {f: getattr(inf, f) for f, _ in L.IndexInfo._fields_}

    def xyz(self):
        """(n,3) float64 device tensor: source xyz in original order (to_3d output, :551-552)."""
        _bind_device(self.device)
        with torch.cuda.device(self.device):
            out = torch.empty((self.n, 3), dtype=torch.float64, device=self.device)
            L.check(self._lib.g2p_index_get_xyz(self._h, _ptr(out), _stream()))
        return out

    def max_nn_dist(self):
        """max over sources of the distance to the nearest OTHER source (k=2 self query + np.max,
        reference :568-570).  Under an initialised process group every rank queries its share of the
        sources and the maximum is all-reduced (the result is identical on every rank and to the
        single-process one: max is exact)."""
        from . import parallel
        _bind_device(self.device)
        d = C.c_double()
        rank, ws = parallel.world()
        lo, hi = parallel.shard_bounds(self.n, rank, ws) if ws > 1 else (0, self.n)
        with torch.cuda.device(self.device):
            L.check(self._lib.g2p_index_max_nn_dist_part(self._h, lo, hi, C.byref(d), _stream()))
        return parallel.max_over_ranks(d.value, self.device) if ws > 1 else d.value

    def min_upper_bound(self, nj):
        """`dmax + dmax*4/Nj` in Python float arithmetic, as the reference (:570)."""
        dmax = self.max_nn_dist()
        return dmax + dmax * 4 / nj

    def knn(self, tlon=None, tlat=None, k=1, mode=L.MODE_RAW, mub=0.0, radius=None, rotation=None, txyz=None,
            want_flags=True, want_xyz=False, into=None):
        """k nearest sources of every target; returns dict(idx int32 [k,m], coef f64 [k,m], flags u8 [m], xyz).
        `into` = (idx_full [k, M], coef_full [k, M], offset): write the result as columns [offset, offset + m) of a larger
        table (g2p_knn_block) - the returned idx / coef are views of that block.

        tlon/tlat: numpy arrays or device tensors, float64 or float32 (PCRaster REAL4 maps keep
        float32: the reference then rounds xyz and distances to float32, :627-628, :691-694)."""
        _bind_device(self.device)
        with torch.cuda.device(self.device):
            if txyz is not None:
                tx = to_device(txyz, torch.float64, self.device).reshape(-1, 3)
                m = tx.shape[0]
                lo = la = None
                code = L.F64
            else:
                lo = to_device(tlon, None, self.device).reshape(-1)
                la = to_device(tlat, None, self.device).reshape(-1)
                if lo.dtype != la.dtype or lo.dtype not in (torch.float64, torch.float32):
                    lo = lo.to(torch.float64)
                    la = la.to(torch.float64)
                if lo.numel() != la.numel():
                    raise ValueError('target lon/lat length mismatch')
                m = lo.numel()
                code = L.F32 if lo.dtype == torch.float32 else L.F64
                tx = None
            if into is None:
                idx = torch.empty((k, m), dtype=torch.int32, device=self.device)
                coef = torch.empty((k, m), dtype=torch.float64, device=self.device)
                base_i, base_c, stride, off = idx, coef, m, 0
            else:
                base_i, base_c, off = into
                stride = base_i.shape[1]
                if (base_i.shape != base_c.shape or base_i.shape[0] != k or base_i.dtype != torch.int32
                        or base_c.dtype != torch.float64 or not base_i.is_contiguous() or not base_c.is_contiguous()
                        or off < 0 or off + m > stride):
                    raise ValueError('knn(into=): idx int32 / coef float64 [k, M] contiguous tables and a block inside them')
                idx, coef = base_i[:, off:off + m], base_c[:, off:off + m]
            flags = torch.empty((m,), dtype=torch.uint8, device=self.device) if want_flags else None
            xyz = torch.empty((m, 3), dtype=torch.float64, device=self.device) if want_xyz else None
            r = float(radius) if radius is not None else (float(self.radius) if self.radius else 0.0)
            if m > 0:
                L.check(self._lib.g2p_knn_block(self._h, _ptr(lo), _ptr(la), _ptr(tx), m, code, r, L.rotation_arg(rotation),
                                                int(k), int(mode), float(mub), _ptr(base_i), _ptr(base_c), int(stride),
                                                int(off), _ptr(flags), _ptr(xyz), _stream()))
        return {'idx': idx, 'coef': coef, 'flags': flags, 'xyz': xyz, 'coef_dtype': code}


def weights(mode, idx, coef, n_src, mub, dtype_code=L.F64):
    """Standalone K5 (in place on idx/coef): raw (idx, dist) -> nearest / invdist table."""
    lib = require_cuda()
    _bind_device(idx.device)
    k, m = idx.shape
    with torch.cuda.device(idx.device):
        L.check(lib.g2p_weights(int(mode), int(k), int(m), int(n_src), float(mub), int(dtype_code), _ptr(idx),
                                _ptr(coef), _stream()))
    return idx, coef


def adw_reference_radius(dist, dtype_code=L.F64, m_total=None):
    """`np.mean(distances[:, 6])` (reference scipy_interpolation_lib.py:339, :585) from the raw k = 11 (or k = 7) distances
    [k, m] of this rank's targets: an exact (double-double) column sum on the device; under an initialised process group the
    partial sums of the ranks are added (all-gather of two doubles per rank) before dividing by the global target count."""
    from . import parallel
    lib = require_cuda()
    _bind_device(dist.device)
    k, m = dist.shape
    pair = (C.c_double * 2)()
    with torch.cuda.device(dist.device):
        L.check(lib.g2p_adw_distance_sum(_ptr(dist), int(m), int(k), 6, int(dtype_code), pair, _stream()))
    parts = parallel.gather_pairs((pair[0], pair[1]), dist.device)
    total = parallel.sum_over_ranks(m, dist.device) if m_total is None else m_total
    hi = lo = 0.0
    for a, b in parts:                      # two-sum merge: the result does not depend on the rank partition
        t = hi + a
        bb = t - hi
        lo += (hi - (t - bb)) + (a - bb)
        lo += b
        hi = t
    return (hi + lo) / int(round(total))


def weights_adw(idx, dist, tlat, tlon, slat, slon, ref_radius, dtype_code=L.F64):
    """Standalone K5b (in place on dist -> weights): raw k = 11 (idx, dist) -> Shepard angular-distance weights.
    tlat / tlon: the targets' ORIGINAL coordinates (device, float64 or float32 = dtype_code), slat / slon: float64 device."""
    lib = require_cuda()
    _bind_device(idx.device)
    k, m = idx.shape
    want = torch.float32 if dtype_code == L.F32 else torch.float64
    if tlat.dtype != want or tlon.dtype != want or slat.dtype != torch.float64 or slon.dtype != torch.float64:
        raise ValueError('weights_adw: target coordinates in the map dtype, source coordinates in float64')
    if tlat.numel() != m or tlon.numel() != m or slat.numel() != slon.numel():
        raise ValueError('weights_adw: coordinate arrays do not match the table')
    with torch.cuda.device(idx.device):
        L.check(lib.g2p_weights_adw(int(m), int(k), int(dtype_code), float(ref_radius), _ptr(tlat), _ptr(tlon), _ptr(slat),
                                    _ptr(slon), int(slat.numel()), _ptr(idx), _ptr(dist), _stream()))
    return idx, dist


def latlon_to_xyz(lon, lat, radius, rotation=None, device=None):
    """K1: `to_3d` (reference scipy_interpolation_lib.py:665-696) on the device; (n,3) float64."""
    lib = require_cuda()
    dev = _bind_device(device)
    with torch.cuda.device(dev):
        lo = to_device(lon, None, dev).reshape(-1)
        la = to_device(lat, None, dev).reshape(-1)
        code = L.F32 if lo.dtype == torch.float32 else L.F64
        if code == L.F64:
            lo, la = lo.to(torch.float64), la.to(torch.float64)
        out = torch.empty((lo.numel(), 3), dtype=torch.float64, device=dev)
        L.check(lib.g2p_latlon_to_xyz(_ptr(lo), _ptr(la), lo.numel(), code, float(radius), L.rotation_arg(rotation),
                                      _ptr(out), _stream()))
    return out


REC_F8 = np.dtype([('indexes', '<i8'), ('coeffs', '<f8')])
REC_F4 = np.dtype([('indexes', '<i8'), ('coeffs', '<f4')])


class DeviceTable:
    """Intertable resident in HBM as SoA: idx int32 [k][m], coef float64 [k][m].

    On disk / at the Python boundary the reference layout is kept: structured array
    {int64 indexes, f8|f4 coeffs}, shape (m,) for k=1 and (m,k) otherwise
    (reference interpolation/__init__.py:251)."""

    def __init__(self, idx, coef, n_src, coef_f32=False, shape_1d=None, grid_shape=None):
        self.idx = idx
        self.coef = coef
        self.k, self.m = idx.shape
        self.n_src = int(n_src)
        self.coef_f32 = bool(coef_f32)
        self.shape_1d = (self.k == 1) if shape_1d is None else shape_1d
        self.device = idx.device
        self.grid_shape = tuple(int(x) for x in grid_shape) if grid_shape is not None else None
        if self.grid_shape is not None and self.grid_shape[0] * self.grid_shape[1] != self.m:
            raise ValueError(f'grid_shape {self.grid_shape} does not hold {self.m} targets')
        # k = 4 summation order (G2P_SUM_PAIRWISE): numpy's einsum adds ((p0 + p1) + p2) + p3 on the strided float64
        # views of a loaded table and (p0 + p2) + (p1 + p3) when it buffers - which it does for float32 coefficients
        self.sum_pairwise = self.coef_f32
        self._plan = None            # C.c_void_p of a g2p_plan, False once found not windowable
        self.plan_error = None

    # ---- window plan for the shared-memory staged apply (K7w)
    def plan(self):
        """g2p_plan handle for this table (built on first use), or None when the table's indexes are too
        scattered for a shared-memory window (the generic gather kernel is used then)."""
        if self._plan is None:
            lib = require_cuda()
            _bind_device(self.device)
            h = C.c_void_p()
            rows, cols = self.grid_shape if self.grid_shape is not None else (0, 0)
            with torch.cuda.device(self.device):
                rc = lib.g2p_plan_create(_ptr(self.idx), self.m, self.k, self.n_src, rows, cols, C.byref(h), _stream())
            if rc == L.ERR_UNSUPPORTED:
                self._plan = False
                self.plan_error = lib.g2p_last_error().decode('utf-8', 'replace')
            else:
                L.check(rc)
                self._plan = h
        return self._plan or None

    def plan_info(self):
        h = self.plan()
        if h is None:
            return None
        inf = L.PlanInfo()
        L.check(L.lib().g2p_plan_get_info(h, C.byref(inf)))
        return {f: getattr(inf, f) for f, _ in L.PlanInfo._fields_}

    def __del__(self):
        try:
            if self._plan:
                L.lib().g2p_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ---- record layout
    @classmethod
    def from_record(cls, rec, n_src, device=None, grid_shape=None):
        """structured array as loaded from an intertable file -> device SoA (K6 unpack)."""
        lib = require_cuda()
        dev = _bind_device(device)
        rec = np.asarray(rec)
        if rec.dtype.names != ('indexes', 'coeffs'):
            raise ValueError(f'not an intertable record dtype: {rec.dtype}')
        f32 = rec.dtype['coeffs'] == np.dtype('<f4')
        want = REC_F4 if f32 else REC_F8
        if rec.dtype != want:
            rec = rec.astype(want)
        rec = np.ascontiguousarray(rec)
        one_d = rec.ndim == 1
        m = rec.shape[0]
        k = 1 if one_d else rec.shape[1]
        with torch.cuda.device(dev):
            raw = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
            idx = torch.empty((k, m), dtype=torch.int32, device=dev)
            coef = torch.empty((k, m), dtype=torch.float64, device=dev)
            L.check(lib.g2p_table_unpack_rec(_ptr(raw), m, k, L.F32 if f32 else L.F64, _ptr(idx), _ptr(coef),
                                               _stream()))
            torch.cuda.current_stream().synchronize()
        return cls(idx, coef, n_src, coef_f32=f32, shape_1d=one_d, grid_shape=grid_shape)

    @classmethod
    def from_grib_nearest(cls, intertable, n_src, grid_shape, device=None):
        """`grib_nearest` intertable `[xs, ys, idxs]` (scatter form: target row / column -> source index,
        reference interpolation/__init__.py:296,312) -> k=1 gather table over the whole target grid; targets the
        table does not list get the sentinel (the reference pre-fills the result with mv_out, :286-287)."""
        require_cuda()
        dev = _bind_device(device)
        t = np.asarray(intertable)
        xs, ys, idxs = (torch.from_numpy(np.ascontiguousarray(t[i]).astype(np.int64)).to(dev) for i in range(3))
        rows, cols = grid_shape
        idx = torch.full((1, rows * cols), int(n_src), dtype=torch.int32, device=dev)
        idx[0, xs * cols + ys] = idxs.to(torch.int32)
        coef = torch.zeros((1, rows * cols), dtype=torch.float64, device=dev)
        return cls(idx, coef, n_src, grid_shape=grid_shape)

    @classmethod
    def from_grib_invdist(cls, rec, n_src, grid_shape, device=None):
        """`grib_invdist` intertable: record array of shape (6, M') with indexes = [xs, ys, idx1..idx4] and
        coeffs = [c1..c4, 0, 0] (reference :348-350, :369-370) -> k=4 gather table; v1*c1 + v2*c2 + v3*c3 + v4*c4
        is evaluated left to right, the order of the apply kernels.  Unlisted targets: sentinel with weights
        [1, 0, 0, 0], i.e. exactly mv_out."""
        require_cuda()
        dev = _bind_device(device)
        ind = np.asarray(rec['indexes'])
        cf = np.asarray(rec['coeffs'], dtype=np.float64)
        rows, cols = grid_shape
        m = rows * cols
        j = torch.from_numpy((ind[0].astype(np.int64) * cols + ind[1].astype(np.int64))).to(dev)
        idx = torch.full((4, m), int(n_src), dtype=torch.int32, device=dev)
        coef = torch.zeros((4, m), dtype=torch.float64, device=dev)
        coef[0] = 1.0
        for kk in range(4):
            idx[kk, j] = torch.from_numpy(ind[2 + kk].astype(np.int32)).to(dev)
            coef[kk, j] = torch.from_numpy(np.ascontiguousarray(cf[kk])).to(dev)
        return cls(idx, coef, n_src, grid_shape=grid_shape)

    def to_record(self):
        """device SoA -> numpy structured array in the reference's file layout (K6 pack + D2H)."""
        lib = require_cuda()
        _bind_device(self.device)
        dt = REC_F4 if self.coef_f32 else REC_F8
        with torch.cuda.device(self.device):
            raw = torch.empty((self.m * self.k * dt.itemsize,), dtype=torch.uint8, device=self.device)
            L.check(lib.g2p_table_pack_rec(_ptr(self.idx), _ptr(self.coef), self.m, self.k,
                                           L.F32 if self.coef_f32 else L.F64, _ptr(raw), _stream()))
            host = d2h_numpy(raw)
        rec = host.view(dt)
        return rec.reshape((self.m,) if self.shape_1d else (self.m, self.k))

    def indexes_coeffs(self):
        """(indexes int64, coeffs) numpy arrays shaped like ScipyInterpolation.interpolate's output."""
        rec = self.to_record()
        return np.ascontiguousarray(rec['indexes']), np.ascontiguousarray(rec['coeffs'])

    # ---- apply
    def _window_ok(self, src, src_mask, propagate, bits=False):
        """alignment rules of g2p_apply_planned (16-byte cp.async chunks need 16-byte aligned rows)."""
        path = os.environ.get('G2P_APPLY_PATH', 'auto')
        if self.k not in (1, 4) or path == 'generic' or getattr(self, '_window_refused', False):
            return False
        k1_needs_compact_tiles = self.k == 1 and path != 'window'
        n_pad = -(-self.n_src // 16) * 16
        if src.data_ptr() % 16 or src.stride(0) % 2 or src.stride(0) < n_pad:
            return False
        if propagate and (src_mask.data_ptr() % 16 or src.stride(0) % 16):
            return False
        if bits and src.stride(0) % 8:
            return False
        if self.plan() is None:
            return False
        if k1_needs_compact_tiles:
            # one gather per target.  Where the target rows run along the source rings (global lat/lon targets: the plan
            # picks 1 x 1024 tiles) the generic kernel's gathers are nearly coalesced and staging does not pay
            # (O1280 -> 0.1 deg: generic 67-77 % of HBM peak, windowed 49-80 %); where they cut across the rings (projected
            # regional targets: compact 16 x 64 tiles) the window wins (O1280 -> EFAS nearest, 383 messages:
            # 0.81 / 0.88 ms without / with masks against 1.05 / 1.68 ms), profiles/r01_apply_sweep.md
            if getattr(self, '_k1_window', None) is None:
                self._k1_window = self.plan_info()['tile_rows'] >= 4
            return self._k1_window
        return True

    def compact(self):
        """(table over the compact numbering of the touched source points, touched int32 [nc] sorted).
        Only the source values an intertable references are needed downstream (4 % of an O1280 field for the
        EFAS grid); `g2p_manip` gathers exactly those, and this table reads them back by compact index."""
        if getattr(self, '_compact', None) is None:
            lib = require_cuda()
            _bind_device(self.device)
            with torch.cuda.device(self.device):
                touched = torch.empty((min(self.n_src, self.k * self.m),), dtype=torch.int32, device=self.device)
                cidx = torch.empty_like(self.idx)
                nc = C.c_int64()
                L.check(lib.g2p_table_compact(_ptr(self.idx), self.m, self.k, self.n_src, _ptr(touched), _ptr(cidx),
                                              C.byref(nc), _stream()))
            nc = int(nc.value)
            self._compact = (DeviceTable(cidx, self.coef, nc, coef_f32=self.coef_f32, shape_1d=self.shape_1d,
                                         grid_shape=self.grid_shape), touched[:nc].contiguous())
        return self._compact

    def apply(self, src, mv_out, src_mask=None, mask_policy=L.MASK_AS_REFERENCE, out=None, out_mask=None,
              correction=None, post=None):
        """Batched K7 on device-resident messages.  src: float64 [b, n] (or [n]); returns out [b, m]
        (and out_mask uint8 [b, m] under PROPAGATE; masked outputs hold mv_out under PROPAGATE and FILL).
        Asynchronous on the current stream.
        Uses the windowed kernel (g2p_apply_planned) when the table and the buffers allow it, else the
        generic gather kernel (g2p_apply); both are bit-identical.  `correction` (a Corrector) and `post` (a
        WriterPost) run in the epilogue of the windowed kernel, or as their own passes after the generic one."""
        lib = require_cuda()
        _bind_device(self.device)
        squeeze = src.dim() == 1
        if squeeze:
            src = src.unsqueeze(0)
            src_mask = src_mask.unsqueeze(0) if src_mask is not None else None
        if src.dtype != torch.float64 or src.device != self.device:
            raise ValueError('src must be a float64 tensor on the table device')
        b, n = src.shape
        if n != self.n_src:
            raise ValueError(f'source field has {n} values, intertable was built for {self.n_src}')
        if src.stride(1) != 1:
            src = src.contiguous()
        propagate = mask_policy in (L.MASK_PROPAGATE, L.MASK_FILL)        # byte masks in
        bits = mask_policy == L.MASK_PROPAGATE_BITS                       # bit-packed masks in and out
        marked = mask_policy == L.MASK_MARKED                             # no mask in; bit-packed mask out (optional)
        out_f32 = post is not None and post.out_f32
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((b, self.m), dtype=torch.float32 if out_f32 else torch.float64, device=self.device)
            elif out.dtype != (torch.float32 if out_f32 else torch.float64):
                raise ValueError('out must be float32 with a float32 WriterPost, float64 otherwise')
            if propagate:
                if src_mask is None:
                    raise ValueError('PROPAGATE / FILL need src_mask')
                if src_mask.dtype == torch.bool:
                    src_mask = src_mask.view(torch.uint8)
                if src_mask.stride(1) != 1 or src_mask.stride(0) != src.stride(0):
                    # the C ABI uses one stride for values and masks
                    src = src.contiguous()
                    src_mask = src_mask.contiguous()
                if out_mask is None and mask_policy == L.MASK_PROPAGATE:
                    out_mask = torch.empty((b, self.m), dtype=torch.uint8, device=self.device)
            if bits:
                if src_mask is None or src_mask.dtype != torch.uint8:
                    raise ValueError('PROPAGATE_BITS needs a bit-packed uint8 src_mask (pack_mask_bits)')
                if src.stride(0) % 8 or src_mask.stride(1) != 1 or src_mask.stride(0) != src.stride(0) // 8:
                    raise ValueError('bit-packed src_mask rows must hold src.stride(0) / 8 bytes '
                                     '(pad the source rows to a multiple of 16 values)')
            if bits or (marked and out_mask is not False):
                if out_mask is None:
                    out_mask = torch.empty((b, bits_row_bytes(out.stride(0))), dtype=torch.uint8, device=self.device)
                elif out_mask.stride(0) != bits_row_bytes(out.stride(0)) or out_mask.data_ptr() % 4:
                    raise ValueError('bit-packed out_mask rows must hold bits_row_bytes(out.stride(0)) bytes')
            if marked and out_mask is False:
                out_mask = None
            takes_mask = propagate or bits
            policy_arg = int(mask_policy) | (L.SUM_PAIRWISE if self.sum_pairwise else 0)
            gives_mask = mask_policy == L.MASK_PROPAGATE or bits or (marked and out_mask is not None)
            done = False
            if b > 0 and self.m > 0 and self._window_ok(src, src_mask, propagate, bits):
                if correction is not None or post is not None:
                    rc = lib.g2p_apply_planned_post(
                        self.plan(), _ptr(self.coef) if self.k > 1 else None, _ptr(src),
                        _ptr(src_mask) if takes_mask else None, b, src.stride(0), out.stride(0), float(mv_out),
                        policy_arg, C.byref(correction.c_struct()) if correction is not None else None,
                        C.byref(post.c_struct()) if post is not None else None, _ptr(out),
                        _ptr(out_mask) if gives_mask else None, _stream())
                else:
                    rc = lib.g2p_apply_planned(self.plan(), _ptr(self.coef) if self.k > 1 else None, _ptr(src),
                                               _ptr(src_mask) if takes_mask else None, b, src.stride(0),
                                               out.stride(0), float(mv_out), policy_arg, _ptr(out),
                                               _ptr(out_mask) if gives_mask else None, _stream())
                if rc == L.ERR_UNSUPPORTED:
                    # e.g. the window does not fit in shared memory: remembered, the generic kernel serves the table
                    self._window_refused = True
                    self.plan_error = lib.g2p_last_error().decode('utf-8', 'replace')
                else:
                    L.check(rc)
                    self.last_path = 'window'
                    done = True
            if not done and out_f32:
                raise NotImplementedError('float32 writer output is an epilogue of the windowed apply kernel; this table has '
                                          'no window plan: ' + str(self.plan_error))
            if not done and b > 0 and self.m > 0:
                self.last_path = 'generic'
                # no plan for this table: the Corrector and the writers' post-processing run as their own passes.
                # Both leave masked cells alone / turn them into the file's missing value, so under FILL (masked
                # outputs hold mv_out, no mask array for the caller) a scratch mask array carries them across.
                g_policy, g_mask = mask_policy, (out_mask if gives_mask else None)
                if (correction is not None or post is not None) and (bits or marked):
                    raise NotImplementedError('Corrector / writer post-processing after the generic kernel take byte '
                                              'masks: use PROPAGATE / FILL for tables without a window plan')
                if mask_policy == L.MASK_FILL and (correction is not None or post is not None):
                    g_policy = L.MASK_PROPAGATE                  # one row stride for values and masks in the C ABI
                    g_mask = torch.empty((b, out.stride(0)), dtype=torch.uint8, device=self.device)[:, :self.m]
                L.check(lib.g2p_apply(_ptr(self.idx), _ptr(self.coef) if self.k > 1 else None, self.m, self.k, _ptr(src),
                                      _ptr(src_mask) if takes_mask else None, n, b, src.stride(0), out.stride(0),
                                      float(mv_out), int(g_policy) | (policy_arg & L.SUM_PAIRWISE), _ptr(out),
                                      _ptr(g_mask) if g_mask is not None else None, _stream()))
                if correction is not None:
                    L.check(lib.g2p_correct(_ptr(out), _ptr(g_mask) if g_mask is not None else None,
                                            self.m, b, out.stride(0), C.byref(correction.c_struct()), _stream()))
                if post is not None:
                    L.check(lib.g2p_writer_post_apply(_ptr(out), _ptr(g_mask) if g_mask is not None else None, self.m, b,
                                                      out.stride(0), C.byref(post.c_struct()), _stream()))
        if squeeze:
            out = out[0]
            out_mask = out_mask[0] if out_mask is not None else None
        return (out, out_mask) if gives_mask else out


def bits_row_bytes(stride):
    """G2P_BITS_ROW_BYTES: bytes of one bit-packed output mask row (whole 32-bit words)."""
    return ((int(stride) + 31) // 32) * 4


def pack_mask_bits(mask, row_values=None):
    """bool / uint8 mask [b, n] (numpy or tensor, host or device) -> bit-packed uint8 [b, row_values / 8] in the
    layout of G2P_MASK_PROPAGATE_BITS: bit j of a row = bit (j & 7) of byte (j >> 3), i.e.
    `numpy.packbits(mask, axis=1, bitorder='little')` - the GRIB bitmap section holds the same information, one bit
    per point.  `row_values`: the padded row length of the source buffer (multiple of 16; default n rounded up)."""
    if isinstance(mask, np.ndarray):
        b, n = mask.shape
        rv = -(-n // 16) * 16 if row_values is None else int(row_values)
        out = np.zeros((b, rv // 8), dtype=np.uint8)
        out[:, :(n + 7) // 8] = np.packbits(mask.astype(bool), axis=1, bitorder='little')
        return out
    b, n = mask.shape
    rv = -(-n // 16) * 16 if row_values is None else int(row_values)
    full = torch.zeros((b, rv), dtype=torch.uint8, device=mask.device)
    full[:, :n] = mask.to(torch.uint8) != 0
    wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=mask.device)
    return (full.view(b, rv // 8, 8) * wts).sum(dim=2, dtype=torch.uint8)


def unpack_mask_bits(bits, m):
    """bit-packed rows [b, bytes] -> bool [b, m] (numpy)"""
    a = bits.cpu().numpy() if isinstance(bits, torch.Tensor) else np.asarray(bits)
    return np.unpackbits(a, axis=1, bitorder='little')[:, :m].astype(bool)


def mark_masked(src, src_mask, bits=False):
    """In place on device fields [b, n]: v = mask ? G2P_MASK_MARKER : v (g2p_mark_masked) - the producer-side half of
    the MARKED mask policy for fields that reach the device with a separate mask."""
    lib = require_cuda()
    _bind_device(src.device)
    if src.dim() == 1:
        src, src_mask = src.unsqueeze(0), src_mask.unsqueeze(0)
    if src_mask.dtype == torch.bool:
        src_mask = src_mask.view(torch.uint8)
    b, n = src.shape
    want = src.stride(0) // 8 if bits else src.stride(0)
    if src.stride(1) != 1 or src_mask.stride(1) != 1 or src_mask.stride(0) != want:
        raise ValueError('values and masks must share the row stride (bit-packed: stride / 8 bytes per row)')
    with torch.cuda.device(src.device):
        L.check(lib.g2p_mark_masked(_ptr(src), _ptr(src_mask), 1 if bits else 0, n, b, src.stride(0), _stream()))
    return src


class WriterPost:
    """What the reference's writers do to a field before it reaches the file (writers/__init__.py:66,
    writers/netcdf.py:149-151,170-183, writers/pcr.py:32-34,44-50), as a per-target epilogue of the apply:
    `out = (clone_mask | value mask | isnan(v) | v == match) ? mv_write : v`.
    clone_mask: bool/uint8 array over the target grid (missing cells of the clone map) or None; mv_write: the
    missing value of the file (netCDF: fill value * scale_factor + offset; PCRaster: nodata of the clone map);
    match: the Interpolator's mv_output on the netCDF path (`out_v[out_v == mv_output] = nan`), None on the
    PCRaster path."""

    def __init__(self, clone_mask, mv_write, match=None, device='cuda', out_dtype=None):
        """out_dtype=np.float32: the fields leave the device as float32 (windowed apply only) - what a PCRaster REAL4 band
        or a netCDF `f4` variable stores anyway (GDAL / netCDF4 cast the float64 array they are handed), rounded to
        nearest like `values.astype(np.float32)`; half the bytes over PCIe."""
        require_cuda()
        self.out_f32 = out_dtype is not None and np.dtype(out_dtype) == np.float32
        if out_dtype is not None and not self.out_f32 and np.dtype(out_dtype) != np.float64:
            raise ValueError('WriterPost out_dtype: float64 (default) or float32')
        self.clone = None
        if clone_mask is not None:
            cm = np.ascontiguousarray(np.asarray(clone_mask)).astype(np.uint8).ravel()
            self.clone = torch.from_numpy(cm).to(device)
        self.mv_write = float(mv_write)
        self.match = None if match is None else float(match)
        self._c = L.WriterPost(self.clone.data_ptr() if self.clone is not None else None,
                               self.match if self.match is not None else 0.0, 0 if self.match is None else 1, self.mv_write,
                               1 if self.out_f32 else 0)

    def c_struct(self):
        return self._c

    def apply_(self, values, mask=None):
        """in place on a device tensor [b, m] (mask: the PROPAGATE out_mask or None)."""
        lib = require_cuda()
        v2 = values if values.dim() == 2 else values.unsqueeze(0)
        m2 = mask if mask is None or mask.dim() == 2 else mask.unsqueeze(0)
        if self.clone is not None and self.clone.numel() != v2.shape[1]:
            raise ValueError('clone mask and fields differ in size')
        L.check(lib.g2p_writer_post_apply(_ptr(v2), _ptr(m2) if m2 is not None else None, v2.shape[1], v2.shape[0],
                                          v2.stride(0), C.byref(self._c), _stream()))
        return values


_D2H_CHUNK = 64 << 20


def d2h_numpy(t):
    """device uint8 tensor [nbytes] -> fresh numpy array.  A plain `.cpu()` of a GB-sized tensor runs at ~2 GB/s
    (pageable destination: the driver stages it piecewise and every destination page faults inside the copy);
    here the chunks go through two page-locked slots at PCIe speed while the host copies the previous chunk into
    the destination on worker threads (O1280 -> GloFAS 0.05 deg table, 1.38 GB: 0.62 s -> 0.15 s)."""
    n = t.numel()
    if n <= _D2H_CHUNK:
        return t.cpu().numpy()
    out = np.empty(n, dtype=np.uint8)
    n_slots = 4
    slots = [pinned_empty((_D2H_CHUNK,), torch.uint8) for _ in range(n_slots)]
    evs = [torch.cuda.Event() for _ in range(n_slots)]
    futs = [None] * n_slots
    cur = torch.cuda.current_stream(t.device)
    side = torch.cuda.Stream(device=t.device)
    side.wait_stream(cur)

    def drain(s, lo, hi):                  # runs in a worker thread: numpy releases the GIL inside the copy
        evs[s].synchronize()
        out[lo:hi] = slots[s][:hi - lo].numpy()

    with concurrent.futures.ThreadPoolExecutor(n_slots) as pool, torch.cuda.stream(side):
        for i, lo in enumerate(range(0, n, _D2H_CHUNK)):
            hi, s = min(n, lo + _D2H_CHUNK), i % n_slots
            if futs[s] is not None:
                futs[s].result()           # the slot's previous chunk has reached the destination
            slots[s][:hi - lo].copy_(t[lo:hi], non_blocking=True)
            evs[s].record(side)
            futs[s] = pool.submit(drain, s, lo, hi)
        for f in futs:
            if f is not None:
                f.result()
    cur.wait_stream(side)
    return out


def launch_count():
    return int(L.lib().g2p_launch_count())


def pinned_empty(shape, dtype=torch.float64):
    """Page-locked host buffer (torch tensor; `.numpy()` gives a zero-copy ndarray view).  The GRIB
    decoder writes message values here so that the H2D copy is a plain async DMA."""
    require_cuda()
    return torch.empty(shape, dtype=dtype, pin_memory=True)


_PINNED_OUT_LIMIT = 16 << 30


def _check_implicit_out(b, m, vbytes):
    """a batch whose results do not fit a reasonable page-locked allocation must bring its own `out=` (or come in
    smaller batches): 383 O1280 messages to the 0.05 deg grid are 66 GB"""
    if b * m * vbytes > _PINNED_OUT_LIMIT:
        raise ValueError(f'{b} messages x {m} targets need {b * m * vbytes / 2**30:.0f} GiB of page-locked memory for the '
                         f'results: pass out= (e.g. Interpolator.pinned_output_buffer of a smaller batch) or split the batch')


def _out_buffers(pipe, post):
    """(device result slots, host dtype, bytes per value) of a host pipeline for this call: the float64 slots, or -
    allocated on first use - float32 ones when the writers' epilogue hands over float32 fields"""
    if post is not None and post.out_f32:
        if getattr(pipe, 'd_out32', None) is None:
            pipe.d_out32 = [torch.empty(o.shape, dtype=torch.float32, device=o.device) for o in pipe.d_out]
        return pipe.d_out32, torch.float32, 4
    return pipe.d_out, torch.float64, 8


class HostApplyPipeline:
    """Host-buffer front end of K7: messages in (pinned) host memory -> interpolated fields in pinned
    host memory, chunked and double-buffered so that H2D, the apply kernel and D2H overlap.

    This is the path behind `Interpolator.interpolate_batch`; the bench's `e2e` number times it."""

    def __init__(self, table, chunk_messages=8, with_mask=False):
        self.t = table
        self.c = int(chunk_messages)
        self.with_mask = bool(with_mask)
        dev = table.device
        n, m = table.n_src, table.m
        with torch.cuda.device(dev):
            self.d_src = [torch.empty((self.c, n), dtype=torch.float64, device=dev) for _ in range(2)]
            self.d_out = [torch.empty((self.c, m), dtype=torch.float64, device=dev) for _ in range(2)]
            self.d_msk = [torch.empty((self.c, n), dtype=torch.uint8, device=dev) for _ in range(2)] if with_mask else None
            self.d_omk = [torch.empty((self.c, m), dtype=torch.uint8, device=dev) for _ in range(2)] if with_mask else None
            self.s_in = torch.cuda.Stream(device=dev)
            self.s_out = torch.cuda.Stream(device=dev)
        self._stage = None          # pinned staging for pageable inputs, allocated on first use
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staged(self, host, lo, hi, which):
        """pageable numpy rows -> pinned staging slot (CPU memcpy); pinned tensors pass through."""
        if isinstance(host, torch.Tensor):
            return host[lo:hi]
        if self._stage is None:
            self._stage = {}
        key = (which, host.dtype.str)
        if key not in self._stage:
            tdt = torch.float64 if host.dtype == np.float64 else torch.uint8
            self._stage[key] = [pinned_empty((self.c, host.shape[1]), tdt) for _ in range(2)]
        slot = self._stage[key][(lo // self.c) % 2]
        view = slot[:hi - lo]
        if hasattr(host, 'copy_rows'):
            host.copy_rows(lo, hi, view.numpy())      # a list of separate pageable arrays (Interpolator._RowStack)
        else:
            view.numpy()[...] = host[lo:hi]
        return view

    def run(self, src, mv_out, src_mask=None, mask_policy=L.MASK_AS_REFERENCE, out=None, out_mask=None,
            correction=None, post=None):
        """src: [b, n] float64 pinned tensor (or ndarray); src_mask: [b, n] uint8/bool or None.
        Returns (out [b, m] pinned tensor, out_mask [b, m] pinned uint8 tensor or None).  With `post` (a WriterPost)
        the fields arrive ready for the writer - masked cells already hold its missing value - and no mask array
        is copied back."""
        t = self.t
        propagate = mask_policy in (L.MASK_PROPAGATE, L.MASK_FILL)   # masks go in
        dev_omask = mask_policy == L.MASK_PROPAGATE                  # ... the kernel writes a mask array
        want_omask = dev_omask and post is None                      # ... and it comes back to the host
        if propagate and not self.with_mask:
            raise ValueError('pipeline was created without mask buffers')
        b = src.shape[0]
        if src.shape[1] != t.n_src:
            raise ValueError(f'source field has {src.shape[1]} values, intertable was built for {t.n_src}')
        if isinstance(src_mask, np.ndarray) and src_mask.dtype == np.bool_:
            src_mask = src_mask.view(np.uint8)
        if isinstance(src_mask, torch.Tensor) and src_mask.dtype == torch.bool:
            src_mask = src_mask.view(torch.uint8)
        d_out, out_dtype, vbytes = _out_buffers(self, post)
        if out is None:
            _check_implicit_out(b, t.m, vbytes)
            out = pinned_empty((b, t.m), out_dtype)
        if want_omask and out_mask is None:
            out_mask = pinned_empty((b, t.m), torch.uint8)
        cur = torch.cuda.current_stream(t.device)
        ev_in = [None, None]
        ev_k = [None, None]
        ev_out = [None, None]
        start = torch.cuda.Event()
        start.record(cur)
        self.s_in.wait_event(start)
        self.s_out.wait_event(start)
        for ci, lo in enumerate(range(0, b, self.c)):
            hi = min(b, lo + self.c)
            s = ci % 2
            nb = hi - lo
            if ev_in[s] is not None and not isinstance(src, torch.Tensor):
                ev_in[s].synchronize()                     # staging slot is free again
            hs = self._staged(src, lo, hi, 'v')
            hm = self._staged(src_mask, lo, hi, 'm') if propagate else None
            with torch.cuda.stream(self.s_in):
                if ev_k[s] is not None:
                    self.s_in.wait_event(ev_k[s])          # slot's previous kernel has consumed its input
                self.d_src[s][:nb].copy_(hs, non_blocking=True)
                self.h2d_bytes += hs.numel() * 8
                if propagate:
                    self.d_msk[s][:nb].copy_(hm, non_blocking=True)
                    self.h2d_bytes += hm.numel()
                ev_in[s] = torch.cuda.Event()
                ev_in[s].record(self.s_in)
            cur.wait_event(ev_in[s])
            if ev_out[s] is not None:
                cur.wait_event(ev_out[s])                  # slot's previous output has left the device
            if propagate:
                t.apply(self.d_src[s][:nb], mv_out, src_mask=self.d_msk[s][:nb], mask_policy=mask_policy,
                        out=d_out[s][:nb], out_mask=self.d_omk[s][:nb] if dev_omask else None,
                        correction=correction, post=post)
            else:
                t.apply(self.d_src[s][:nb], mv_out, out=d_out[s][:nb], correction=correction, post=post)
            ev_k[s] = torch.cuda.Event()
            ev_k[s].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_k[s])
                out[lo:hi].copy_(d_out[s][:nb], non_blocking=True)
                self.d2h_bytes += nb * t.m * vbytes
                if want_omask:
                    out_mask[lo:hi].copy_(self.d_omk[s][:nb], non_blocking=True)
                    self.d2h_bytes += nb * t.m
                ev_out[s] = torch.cuda.Event()
                ev_out[s].record(self.s_out)
        for e in ev_out:
            if e is not None:
                cur.wait_event(e)
        return out, (out_mask if want_omask else None)


def compact_gather(src_ptr, mask_ptr, n, b, src_stride, touched, out, out_mask):
    """g2p_manip with COPY items: out[i][c] = src[i][touched[c]] for b messages.  `src_ptr` / `mask_ptr` are raw
    addresses: device memory or PINNED HOST memory (unified addressing lets the kernel read it over PCIe)."""
    lib = require_cuda()
    items = (L.ManipItem * b)()
    for i in range(b):
        items[i].kind, items[i].n_in = L.AGG_COPY, 1
        items[i].inputs[0] = i
        items[i].divisor = 1.0
    desc = L.ManipDesc()
    desc.conv_op1 = desc.conv_op2 = L.OP_NONE
    desc.n_out = b
    desc.items = C.cast(items, C.POINTER(L.ManipItem))
    L.check(lib.g2p_manip(C.byref(desc), C.c_void_p(src_ptr), C.c_void_p(mask_ptr) if mask_ptr else None, n, b,
                          src_stride, _ptr(touched), touched.numel(), out.stride(0), _ptr(out),
                          _ptr(out_mask) if mask_ptr else None, _stream()))


class CompactHostPipeline:
    """Host-buffer front end for tables that touch a small part of the source grid (regional targets: the EFAS
    table reads 4 % of an O1280 field).  Instead of copying whole messages to the device, a gather kernel reads
    exactly the touched values out of the pinned host buffers over PCIe (zero-copy), the windowed apply runs
    on the compact fields through the renumbered table, and the results return by D2H - kernel and D2H of
    neighbouring chunks overlap on two streams.  Same results as HostApplyPipeline, ~17x less H2D traffic."""

    def __init__(self, table, chunk_messages=16, with_mask=False):
        self.t = table
        self.ct, self.touched = table.compact()
        self.c = int(min(chunk_messages, 16))          # <= 16 items travel in the kernel parameters (no sync)
        self.with_mask = bool(with_mask)
        dev = table.device
        nc_pad = -(-self.ct.n_src // 16) * 16
        m = table.m
        with torch.cuda.device(dev):
            self.d_c = [torch.zeros((self.c, nc_pad), dtype=torch.float64, device=dev) for _ in range(2)]
            self.d_cm = [torch.zeros((self.c, nc_pad), dtype=torch.uint8, device=dev) for _ in range(2)] if with_mask else None
            self.d_out = [torch.empty((self.c, m), dtype=torch.float64, device=dev) for _ in range(2)]
            self.d_omk = [torch.empty((self.c, m), dtype=torch.uint8, device=dev) for _ in range(2)] if with_mask else None
            self.s_out = torch.cuda.Stream(device=dev)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, src, mv_out, src_mask=None, mask_policy=L.MASK_AS_REFERENCE, out=None, out_mask=None,
            correction=None, post=None):
        """src: [b, n] float64 PINNED tensor; src_mask: [b, n] uint8 pinned tensor or None -> (out, out_mask) pinned.
        `correction` / `post` as in HostApplyPipeline.run (with `post` no mask array is copied back)."""
        t = self.t
        masks_in = mask_policy in (L.MASK_PROPAGATE, L.MASK_FILL)
        dev_omask = mask_policy == L.MASK_PROPAGATE
        want_omask = dev_omask and post is None
        if masks_in and not self.with_mask:
            raise ValueError('pipeline was created without mask buffers')
        if not (isinstance(src, torch.Tensor) and src.is_pinned() and src.dtype == torch.float64 and src.stride(1) == 1):
            raise ValueError('src must be a pinned float64 tensor [b, n] (device.pinned_empty)')
        if masks_in and not (isinstance(src_mask, torch.Tensor) and src_mask.is_pinned() and src_mask.stride(0) == src.stride(0)):
            raise ValueError('src_mask must be a pinned uint8 tensor with the row stride of src')
        b, n = src.shape
        if n != t.n_src:
            raise ValueError(f'source field has {n} values, intertable was built for {t.n_src}')
        d_out, out_dtype, vbytes = _out_buffers(self, post)
        if out is None:
            _check_implicit_out(b, t.m, vbytes)
            out = pinned_empty((b, t.m), out_dtype)
        if want_omask and out_mask is None:
            out_mask = pinned_empty((b, t.m), torch.uint8)
        _bind_device(t.device)
        nc = self.ct.n_src
        cur = torch.cuda.current_stream(t.device)
        ev_out = [None, None]
        # chunk bounds: a short first chunk, so that the D2H engine - the bottleneck - starts early
        bounds, lo = [], 0
        while lo < b:
            hi = min(b, lo + (max(1, self.c // 4) if lo == 0 and b > self.c else self.c))
            bounds.append((lo, hi))
            lo = hi
        with torch.cuda.device(t.device):
            for ci, (lo, hi) in enumerate(bounds):
                s, nb = ci % 2, hi - lo
                if ev_out[s] is not None:
                    cur.wait_event(ev_out[s])                  # the slot's previous results have left the device
                compact_gather(src.data_ptr() + lo * src.stride(0) * 8,
                               (src_mask.data_ptr() + lo * src.stride(0)) if masks_in else None, n, nb, src.stride(0),
                               self.touched, self.d_c[s][:nb], self.d_cm[s][:nb] if masks_in else None)
                self.h2d_bytes += nb * nc * (9 if masks_in else 8)     # what the gather pulls over PCIe (touched values)
                self.ct.apply(self.d_c[s][:nb, :nc], mv_out, src_mask=self.d_cm[s][:nb, :nc] if masks_in else None,
                              mask_policy=mask_policy, out=d_out[s][:nb],
                              out_mask=self.d_omk[s][:nb] if dev_omask else None, correction=correction, post=post)
                ev_k = torch.cuda.Event()
                ev_k.record(cur)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_k)
                    out[lo:hi].copy_(d_out[s][:nb], non_blocking=True)
                    self.d2h_bytes += nb * t.m * vbytes
                    if want_omask:
                        out_mask[lo:hi].copy_(self.d_omk[s][:nb], non_blocking=True)
                        self.d2h_bytes += nb * t.m
                    ev_out[s] = torch.cuda.Event()
                    ev_out[s].record(self.s_out)
            for e in ev_out:
                if e is not None:
                    cur.wait_event(e)
        return out, (out_mask if want_omask else None)

    def touched_host(self):
        """the touched source indexes as a host int32 array (for `pick` on pageable fields)."""
        if getattr(self, '_touched_host', None) is None:
            self._touched_host = np.ascontiguousarray(self.touched.cpu().numpy().astype(np.int32))
        return self._touched_host

    def pick(self, field, out_values, out_mask=None, threads=0):
        """g2p_host_pick: the touched values of one pageable message (ndarray / MaskedArray, n values) -> out_values
        [>= nc] (a numpy view of a pinned buffer), their mask bytes -> out_mask (zeros for an unmasked message).  Pure
        host staging; releases the GIL."""
        th = self.touched_host()
        data = np.ascontiguousarray(np.ma.getdata(field), dtype=np.float64).reshape(-1)
        if data.size != self.t.n_src:
            raise ValueError(f'source field has {data.size} values, intertable was built for {self.t.n_src}')
        mask = None
        if out_mask is not None and isinstance(field, np.ma.MaskedArray) and field.mask is not np.ma.nomask:
            mask = np.ascontiguousarray(np.broadcast_to(np.ma.getmaskarray(field).reshape(-1), data.shape)).view(np.uint8)
        L.check(L.lib().g2p_host_pick(data.ctypes.data, mask.ctypes.data if mask is not None else None, th.ctypes.data,
                                      th.size, out_values.ctypes.data, out_mask.ctypes.data if out_mask is not None else None,
                                      0, int(threads)))

    def run_compact(self, h_c, mv_out, h_cm=None, mask_policy=L.MASK_AS_REFERENCE, out=None, out_mask=None,
                    correction=None, post=None):
        """as `run`, for messages that lie in PAGEABLE host memory: the caller has picked the touched values of every
        message (`numpy.take(field, self.touched_host(), out=h_c[i, :nc])`) into the pinned tensor h_c [b, >= nc]
        (h_cm: their mask bytes) - 2.2 MB per O1280 message for the EFAS table instead of a 52.8 MB copy into a
        page-locked staging buffer.  H2D of the compact chunk, windowed apply, D2H."""
        t = self.t
        masks_in = mask_policy in (L.MASK_PROPAGATE, L.MASK_FILL)
        dev_omask = mask_policy == L.MASK_PROPAGATE
        want_omask = dev_omask and post is None
        if masks_in and not self.with_mask:
            raise ValueError('pipeline was created without mask buffers')
        b = h_c.shape[0]
        nc = self.ct.n_src
        d_out, out_dtype, vbytes = _out_buffers(self, post)
        if out is None:
            _check_implicit_out(b, t.m, vbytes)
            out = pinned_empty((b, t.m), out_dtype)
        if want_omask and out_mask is None:
            out_mask = pinned_empty((b, t.m), torch.uint8)
        _bind_device(t.device)
        cur = torch.cuda.current_stream(t.device)
        ev_out = [None, None]
        with torch.cuda.device(t.device):
            for ci, lo in enumerate(range(0, b, self.c)):
                hi = min(b, lo + self.c)
                s, nb = ci % 2, hi - lo
                if ev_out[s] is not None:
                    cur.wait_event(ev_out[s])
                self.d_c[s][:nb, :nc].copy_(h_c[lo:hi, :nc], non_blocking=True)
                self.h2d_bytes += nb * nc * 8
                if masks_in:
                    self.d_cm[s][:nb, :nc].copy_(h_cm[lo:hi, :nc], non_blocking=True)
                    self.h2d_bytes += nb * nc
                self.ct.apply(self.d_c[s][:nb, :nc], mv_out, src_mask=self.d_cm[s][:nb, :nc] if masks_in else None,
                              mask_policy=mask_policy, out=d_out[s][:nb],
                              out_mask=self.d_omk[s][:nb] if dev_omask else None, correction=correction, post=post)
                ev_k = torch.cuda.Event()
                ev_k.record(cur)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_k)
                    out[lo:hi].copy_(d_out[s][:nb], non_blocking=True)
                    self.d2h_bytes += nb * t.m * vbytes
                    if want_omask:
                        out_mask[lo:hi].copy_(self.d_omk[s][:nb], non_blocking=True)
                        self.d2h_bytes += nb * t.m
                    ev_out[s] = torch.cuda.Event()
                    ev_out[s].record(self.s_out)
            for e in ev_out:
                if e is not None:
                    cur.wait_event(e)
        return out, (out_mask if want_omask else None)
