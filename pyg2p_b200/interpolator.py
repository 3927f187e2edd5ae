"""Drop-in for the reference's `Interpolator` (main/interpolation/__init__.py:20-395) on the scipy path:
same constructor (`exec_ctx`, `mv_input`), same context keys, intertable naming / lookup / on-disk layout,
same `interpolate` / `interpolate_scipy` signatures and error codes - with the table build and the table
apply running on the GPU (pyg2p_b200.device), and `interpolate_batch` added so that callers can hand over
all messages (steps x members) of a run at once instead of looping per step (writers/__init__.py:55,96).
"""
import gzip
import logging
import os
import re
from functools import partial

import numpy as np
import numpy.ma as ma
import torch

from . import _lib as L
from . import device
from .exceptions import (MALFORMED_GRIB_LATLON, NO_INTERTABLE_CREATED, NOT_IMPLEMENTED, ApplicationException)
from .maps import LatLong
from .scipy_interpolation import ScipyInterpolation


_GZ_SUBFIELD = b'G2'            # gzip FEXTRA subfield of our members: uint64 length of the deflate payload


def save_npy_gz_parallel(path, array, level=1, chunk_bytes=16 << 20, threads=None):
    """`np.save` through gzip, compressed by all host cores: the `.npy` byte stream is cut into chunks, every
    chunk becomes one gzip member (zlib releases the GIL), the members are concatenated.  A multi-member file
    is a valid gzip stream - `gzip.GzipFile(path).read()` / `np.load` in the reference read it unchanged
    (interpolation/__init__.py:107-120) - and decompresses to exactly the bytes `np.save` writes.  The
    reference's single-threaded level-9 write costs ~40 s per million k=4 rows.
    Every member carries, in a gzip FEXTRA subfield (ignored by other readers, the way BGZF does it), the length of
    its deflate payload, so that `load_npy_gz_parallel` can find the members without inflating them."""
    import io
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    buf = io.BytesIO()
    np.save(buf, array)
    data = buf.getbuffer()

    def member(lo):
        raw = data[lo:lo + chunk_bytes]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)         # raw deflate: header and trailer are written here
        payload = c.compress(raw) + c.flush()
        extra = _GZ_SUBFIELD + struct.pack('<HQ', 8, len(payload))
        head = b'\x1f\x8b\x08\x04' + b'\x00\x00\x00\x00' + b'\x00\xff' + struct.pack('<H', len(extra)) + extra
        return head + payload + struct.pack('<II', zlib.crc32(raw) & 0xffffffff, len(raw) & 0xffffffff)

    with ThreadPoolExecutor(max_workers=threads or os.cpu_count() or 1) as pool, open(path, 'wb') as f:
        for part in pool.map(member, range(0, max(len(data), 1), chunk_bytes)):
            f.write(part)


def load_npy_gz_parallel(path, threads=None):
    """`np.load(gzip.GzipFile(path))` for files written by `save_npy_gz_parallel`, inflated by all host cores: the
    members are found by hopping over their payloads (length in the FEXTRA subfield, size of the content in the
    gzip trailer), every member is inflated straight into its slice of the result.  Any other gzip file (the
    reference's own single-member tables) is read the plain way (interpolation/__init__.py:107-120)."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    with open(path, 'rb') as f:
        blob = f.read()
    members, at, total = [], 0, 0
    while at < len(blob):
        ok = (blob[at:at + 4] == b'\x1f\x8b\x08\x04' and len(blob) >= at + 24 and blob[at + 12:at + 14] == _GZ_SUBFIELD
              and struct.unpack_from('<HH', blob, at + 10)[0] == 12 and struct.unpack_from('<H', blob, at + 14)[0] == 8)
        if not ok:
            members = None
            break
        n_payload = struct.unpack_from('<Q', blob, at + 16)[0]
        start = at + 24
        end = start + n_payload
        if end + 8 > len(blob):
            members = None
            break
        crc, size = struct.unpack_from('<II', blob, end)
        members.append((start, end, total, size, crc))
        total += size
        at = end + 8
    if not members:
        with gzip.GzipFile(path, 'r') as f:
            return np.load(f)
    out = np.empty(total, dtype=np.uint8)       # not zero-filled: the pages are first touched by the inflating threads
    view = memoryview(out)

    def inflate(m):
        start, end, lo, size, crc = m
        raw = zlib.decompressobj(-15).decompress(memoryview(blob)[start:end])
        if len(raw) != size or (zlib.crc32(raw) & 0xffffffff) != crc:
            raise OSError(f'{path}: corrupt gzip member at byte {start}')
        view[lo:lo + size] = raw

    with ThreadPoolExecutor(max_workers=threads or os.cpu_count() or 1) as pool:
        list(pool.map(inflate, members))
    return _npy_from_buffer(out)


def _npy_from_buffer(buf):
    """np.load of an in-memory `.npy` image without another copy of the data."""
    import io
    from numpy.lib import format as npf
    f = io.BytesIO(bytes(memoryview(buf)[:4096]))
    version = npf.read_magic(f)
    shape, fortran, dtype = npf.read_array_header_1_0(f) if version == (1, 0) else npf.read_array_header_2_0(f)
    arr = np.frombuffer(buf, dtype=dtype, offset=f.tell(), count=int(np.prod(shape, dtype=np.int64)))
    return arr.reshape(shape, order='F' if fortran else 'C')


def _normalize_filename(name):
    """util/files.py:98-110: lower-case stem without long digit runs and separators, first 10 characters."""
    stem = os.path.splitext(name)[0].lower()
    stem = re.sub(r'([0-9]{8,10})', '', stem) or stem
    stem = stem.replace('-', '').replace('_', '').replace('.', '')
    return name.replace('.', '_') if len(stem) < 3 else stem[:10]


class _RowStack:
    """a batch of pageable 1-D arrays presented as [b, n] rows to HostApplyPipeline._staged (which copies rows lo:hi into
    a page-locked slot through `copy_rows`); a None row is all zeros (an unmasked message)"""

    def __init__(self, rows, dtype, n=None):
        self.rows, self.dtype = rows, np.dtype(dtype)
        self.shape = (len(rows), n if n is not None else rows[0].size)

    def __len__(self):
        return self.shape[0]

    def copy_rows(self, lo, hi, dst):
        for i, r in enumerate(self.rows[lo:hi]):
            dst[i] = 0 if r is None else r


class Interpolator:
    _LOADED_INTERTABLES = {}      # path -> numpy record array (as the reference, :21)
    _DEVICE_TABLES = {}           # (path, device index) -> DeviceTable
    _prefix = 'I'
    scipy_modes_nnear = {'nearest': 1, 'invdist': 4, 'adw': 11, 'cdd': 11, 'bilinear': 4, 'triangulation': 3,
                         'bilinear_delaunay': 4}
    suffixes = {'grib_nearest': 'grib_nearest', 'grib_invdist': 'grib_invdist', 'nearest': 'scipy_nearest',
                'invdist': 'scipy_invdist', 'adw': 'scipy_adw', 'cdd': 'scipy_cdd', 'bilinear': 'scipy_bilinear',
                'triangulation': 'scipy_triangulation', 'bilinear_delaunay': 'scipy_bilinear_delaunay'}
    _format_intertable = 'tbl{prognum}_{source_file}_{target_size}_{suffix}.npy.gz'.format
    gzip_level = 1                # file content is identical at any level; level 9 costs ~40 s per million rows

    def __init__(self, exec_ctx, mv_input, mask_policy='as_reference'):
        self._logger = logging.getLogger()
        self._mv_grib = mv_input
        self.interpolate_with_grib = exec_ctx.is_with_grib_interpolation
        self._mode = exec_ctx.get('interpolation.mode')
        self._num_of_splits = exec_ctx.get('interpolation.num_of_splits')
        self._source_filename = os.path.basename(exec_ctx.get('input.file'))
        self._suffix = self.suffixes[self._mode]
        self._intertable_dirs = exec_ctx.get('interpolation.dirs')
        self._rotated_target_grid = exec_ctx.get('interpolation.rotated_target')
        lat_map, lon_map = exec_ctx.get('interpolation.latMap'), exec_ctx.get('interpolation.lonMap')
        self._target_coords = lat_map if isinstance(lat_map, LatLong) else LatLong(lat_map, lon_map)
        self.mv_out = self._target_coords.mv
        self.parallel = exec_ctx.get('interpolation.parallel')
        self.format_intertablename = partial(self._format_intertable,
                                             source_file=_normalize_filename(self._source_filename),
                                             target_size=self._target_coords.lats.size, suffix=self._suffix)
        self._aux_val = self._aux_gid = self._aux_2nd_res_gid = self._aux_2nd_res_val = None
        self.create_if_missing = exec_ctx.get('interpolation.create')
        self.intertables_config = exec_ctx.configuration.intertables
        # 'as_reference': masks of the input are ignored, as the reference does;
        # 'propagate': the result is a MaskedArray, masked where any of the k source points is masked
        self.mask_policy = mask_policy
        self.tie_rows = None

    # ------------------------------------------------------------------ reference API
    def interpolate(self, lats, longs, v, grid_id, geodetic_info, gid=-1, is_second_res=False):
        if self.interpolate_with_grib:
            return self.interpolate_grib(v, gid, grid_id, is_second_res=is_second_res)
        return self.interpolate_scipy(lats, longs, v, grid_id, geodetic_info)

    def interpolate_grib(self, v, gid, grid_id, is_second_res=False):
        """`grib_nearest` / `grib_invdist` (reference :268-371): existing intertables of both layouts are APPLIED
        on the GPU; creating one is an ecCodes-internal nearest-point search and stays with the reference."""
        return self.interpolate_grib_batch([v], gid, grid_id, is_second_res=is_second_res)[0]

    def _grib_table(self, v, grid_id):
        if self._mode not in ('grib_nearest', 'grib_invdist'):
            raise ApplicationException.get_exc(NOT_IMPLEMENTED, f'mode {self._mode} with GRIB interpolation')
        intertable_id, path = self._intertable_filename(grid_id)
        dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
        key = (path, dev)
        table = self._DEVICE_TABLES.get(key)
        if table is None:
            if not (os.path.exists(path) or os.path.exists(path + '.gz')):
                if self.create_if_missing:
                    raise ApplicationException.get_exc(
                        NOT_IMPLEMENTED, f'creating {self._mode} intertables needs ecCodes (codes_grib_find_nearest): '
                                         f'build {path} with pyg2p, or use nearest / invdist')
                raise ApplicationException.get_exc(NO_INTERTABLE_CREATED, details=path)
            self._read_intertable_raw(path)
            raw = self._LOADED_INTERTABLES[path]
            shape = self._target_coords.lons.shape
            n_src = int(np.asarray(ma.getdata(v)).size)
            table = (device.DeviceTable.from_grib_nearest(raw, n_src, shape) if self._mode == 'grib_nearest'
                     else device.DeviceTable.from_grib_invdist(raw, n_src, shape))
            self._DEVICE_TABLES[key] = table
        return table

    def interpolate_grib_batch(self, values, gid, grid_id, is_second_res=False):
        """all messages of a run through a `grib_nearest` / `grib_invdist` intertable on the batched host pipelines
        (pinned staging, H2D / kernel / D2H overlapped) -> [b, rows, cols].  A masked input gives `mv_output` where a
        contributing point is masked (`result_masked` fills, util/numeric.py:21-27 = MASK_FILL)."""
        values = list(values)
        if not values:
            return np.empty((0,) + self._target_coords.lons.shape)
        table = self._grib_table(values[0], grid_id)
        any_mask = any(isinstance(x, ma.MaskedArray) and x.mask is not ma.nomask for x in values)
        out, _ = self._run_pageable(table, values, L.MASK_FILL if any_mask else L.MASK_AS_REFERENCE)
        torch.cuda.current_stream(table.device).synchronize()
        return out.numpy().reshape((len(values),) + self._target_coords.lons.shape)

    def _run_pageable(self, table, values, policy, out=None, out_mask=None, corrector=None, writer_post=None):
        """pageable per-message arrays through the table's host pipeline: the touched values picked into a small pinned
        buffer (compact tables), or whole fields staged chunk by chunk through two page-locked slots"""
        masks_in = policy in (L.MASK_PROPAGATE, L.MASK_FILL)
        pipe = self._pipeline(table, masks_in)
        b, n = len(values), table.n_src
        masked_in = [isinstance(x, ma.MaskedArray) and x.mask is not ma.nomask for x in values]
        if isinstance(pipe, device.CompactHostPipeline):
            # a regional table: pick the touched values of every message straight into a small pinned buffer (2.2 MB of a
            # 52.8 MB O1280 field for EFAS) instead of staging whole fields
            nc = pipe.touched_host().size
            h_c = device.pinned_empty((b, nc))
            h_cm = device.pinned_empty((b, nc), torch.uint8) if masks_in else None
            for i, x in enumerate(values):
                pipe.pick(x, h_c.numpy()[i], h_cm.numpy()[i] if masks_in else None)
            return pipe.run_compact(h_c, float(self.mv_output), h_cm=h_cm, mask_policy=policy, out=out, out_mask=out_mask,
                                    correction=corrector, post=writer_post)
        # a table that reads most of the source grid: HostApplyPipeline stages the fields chunk by chunk through two
        # page-locked slots of ~256 MB (never the whole batch: a run of 383 O1280 messages is 20 GB)
        h_src = _RowStack([np.asarray(ma.getdata(x), dtype=np.float64).reshape(-1) for x in values], np.float64)
        h_msk = _RowStack([np.broadcast_to(ma.getmaskarray(x).reshape(-1), (n,)) if masked_in[i] else None
                           for i, x in enumerate(values)], np.uint8, n) if masks_in else None
        return pipe.run(h_src, float(self.mv_output), src_mask=h_msk, mask_policy=policy, out=out, out_mask=out_mask,
                        correction=corrector, post=writer_post)

    def _read_intertable_raw(self, tbl_fullpath):
        if tbl_fullpath not in self._LOADED_INTERTABLES:
            path = tbl_fullpath if os.path.exists(tbl_fullpath) else tbl_fullpath + '.gz'
            if path.endswith('.gz'):
                self._LOADED_INTERTABLES[tbl_fullpath] = load_npy_gz_parallel(path)
            else:
                self._LOADED_INTERTABLES[tbl_fullpath] = np.load(path)
        return self._LOADED_INTERTABLES[tbl_fullpath]

    def interpolate_scipy(self, latgrib, longrib, v, grid_id, grid_details=None):
        out = self.interpolate_batch(latgrib, longrib, [v], grid_id, grid_details)
        return out[0]

    def interpolate_batch(self, latgrib, longrib, values, grid_id, grid_details=None, masks=None, out=None,
                          out_mask=None, corrector=None, writer_post=None):
        """`values`: sequence of source fields [n] (ndarray / MaskedArray) or a 2-D array [b, n] - all messages
        of the run that share `grid_id`.  Returns an array [b, rows, cols] (MaskedArray under 'propagate').

        Fast path for decoders: `values` = the tensor from `pinned_source_buffer(b, n)` (and `masks` = a pinned
        uint8 tensor of the same shape, 1 = missing, when `mask_policy == 'propagate'`); `out` / `out_mask` =
        pinned tensors [b, rows*cols] from `pinned_output_buffer` to reuse between calls (the returned arrays
        are views of them).

        `corrector` (manipulation.Corrector) and `writer_post` (from `self.writer_post(...)`) hoist the rest of the
        writers' per-step loop (writers/__init__.py:62-67, 103-106) into the same kernel: correction, then
        `== mv_output -> NaN`, clone-map mask, value mask and NaN -> the file's missing value.  With `writer_post`
        the result is a plain ndarray ready for `values_nc[t] = ...` / `WriteArray`."""
        pinned = isinstance(values, torch.Tensor)
        first = values[0].numpy() if pinned else values[0]
        table = self._table_for(latgrib, longrib, first, grid_id, grid_details)
        shape = self._target_coords.lons.shape
        b = len(values)
        n = table.n_src
        if pinned:
            # messages decoded straight into `pinned_source_buffer(b, n)`: no staging copy on the host
            if not (values.is_pinned() and values.dtype == torch.float64 and values.dim() == 2 and values.shape[1] == n):
                raise ValueError('a tensor batch must come from Interpolator.pinned_source_buffer')
            masked_in = [masks is not None] * b
            propagate = self.mask_policy == 'propagate' and masks is not None
            pipe = self._pipeline(table, propagate)
            h_src, h_msk = values, (masks if propagate else None)
        else:
            masked_in = [isinstance(x, ma.MaskedArray) and x.mask is not ma.nomask for x in values]
            propagate = self.mask_policy == 'propagate' and any(masked_in)
            out, omask = self._run_pageable(table, values, L.MASK_PROPAGATE if propagate else L.MASK_AS_REFERENCE, out=out,
                                            out_mask=out_mask, corrector=corrector, writer_post=writer_post)
            return self._batch_result(table, out, omask, b, shape, propagate, pinned, values, writer_post)
        out, omask = pipe.run(h_src, float(self.mv_output), src_mask=h_msk,
                              mask_policy=L.MASK_PROPAGATE if propagate else L.MASK_AS_REFERENCE, out=out,
                              out_mask=out_mask, correction=corrector, post=writer_post)
        return self._batch_result(table, out, omask, b, shape, propagate, pinned, values, writer_post)

    def interpolate_stream(self, latgrib, longrib, values_iter, grid_id, grid_details=None, depth=4, corrector=None,
                           writer_post=None):
        """The per-step loop of the writers (writers/__init__.py:55-70, 96-109) as a generator: takes an iterator of
        source fields as the GRIB reader yields them - pageable numpy arrays / MaskedArrays, one per message
        (readers/grib.py:191-201) - and yields the interpolated fields in the same order.

        A host thread picks the touched values of the NEXT messages out of the pageable arrays into a ring of `depth`
        small pinned buffers (2.2 MB of a 52.8 MB O1280 field for the EFAS table) while the device copies, applies and
        returns the previous ones; every result lands in its own page-locked array (recycled by torch's pinned
        allocator once the caller drops it), so nothing is copied on the host after the D2H.  Tables that read most of
        the source grid fall back to one `interpolate_batch` call per message."""
        import collections
        import itertools
        import queue
        import threading
        it = iter(values_iter)
        try:
            first = next(it)
        except StopIteration:
            return
        table = self._table_for(latgrib, longrib, first, grid_id, grid_details)
        shape = self._target_coords.lons.shape
        want_mask = self.mask_policy == 'propagate'
        pipe = self._pipeline(table, want_mask)
        if not isinstance(pipe, device.CompactHostPipeline):
            for x in itertools.chain([first], it):
                yield self.interpolate_batch(latgrib, longrib, [x], grid_id, grid_details, corrector=corrector,
                                             writer_post=writer_post)[0]
            return
        n, m = table.n_src, table.m
        th = pipe.touched_host()
        nc = th.size
        nc_pad = -(-nc // 16) * 16
        dev = table.device
        ring_v = device.pinned_empty((depth, nc_pad))
        ring_m = device.pinned_empty((depth, nc_pad), torch.uint8) if want_mask else None
        ring_v.numpy()[:, nc:] = 0.0
        if want_mask:
            ring_m.numpy()[:, nc:] = 0
        free = [threading.Event() for _ in range(depth)]    # slot may be overwritten (its H2D has completed)
        for f in free:
            f.set()
        h2d_done = [None] * depth
        ready = queue.Queue(maxsize=depth)
        stop = threading.Event()

        def producer():
            try:
                for i, x in enumerate(itertools.chain([first], it)):
                    s = i % depth
                    while not free[s].wait(0.05):
                        if stop.is_set():
                            return
                    free[s].clear()
                    if h2d_done[s] is not None:
                        h2d_done[s].synchronize()
                    # (four threads: the helpers of a continuous stream never go to sleep)
                    pipe.pick(x, ring_v.numpy()[s, :nc], ring_m.numpy()[s, :nc] if want_mask else None, threads=4)
                    ready.put((s, isinstance(x, ma.MaskedArray)))
                ready.put(None)
            except BaseException as e:      # noqa: BLE001 - handed to the consumer
                ready.put(e)

        worker = threading.Thread(target=producer, daemon=True)
        worker.start()
        side = torch.cuda.Stream(device=dev)       # H2D + apply of message i ...
        s_out = torch.cuda.Stream(device=dev)      # ... while the results of message i-1 leave (D2H is the long leg)
        policy = L.MASK_PROPAGATE if want_mask else L.MASK_AS_REFERENCE
        d_v = [torch.zeros((1, nc_pad), dtype=torch.float64, device=dev) for _ in range(2)]
        d_m = [torch.zeros((1, nc_pad), dtype=torch.uint8, device=dev) for _ in range(2)] if want_mask else None
        o_dtype = torch.float32 if (writer_post is not None and writer_post.out_f32) else torch.float64
        d_o = [torch.empty((1, m), dtype=o_dtype, device=dev) for _ in range(2)]
        d_om = [torch.empty((1, m), dtype=torch.uint8, device=dev) for _ in range(2)] if want_mask else None
        pending = collections.deque()

        def finish(entry):
            h_o, h_om, ev, was_ma = entry
            ev.synchronize()
            res = h_o.numpy().reshape(shape)
            if writer_post is not None:
                return res
            if want_mask:
                return ma.masked_array(res, mask=h_om.numpy().view(np.bool_).reshape(shape), fill_value=self.mv_output,
                                       copy=False)
            return ma.masked_array(res) if (table.k == 1 and was_ma) else res

        try:
            i = 0
            with torch.cuda.device(dev), torch.cuda.stream(side):
                while True:
                    item = ready.get()
                    if item is None:
                        break
                    if isinstance(item, BaseException):
                        raise item
                    s, was_ma = item
                    b = i % 2
                    d_v[b].copy_(ring_v[s:s + 1], non_blocking=True)
                    if want_mask:
                        d_m[b].copy_(ring_m[s:s + 1], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(side)
                    h2d_done[s] = ev_in
                    free[s].set()
                    pipe.ct.apply(d_v[b][:, :nc], float(self.mv_output), src_mask=d_m[b][:, :nc] if want_mask else None,
                                  mask_policy=policy, out=d_o[b], out_mask=d_om[b] if want_mask else None,
                                  correction=corrector, post=writer_post)
                    ev_k = torch.cuda.Event()
                    ev_k.record(side)
                    h_o = device.pinned_empty((m,), o_dtype)
                    h_om = None
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev_k)
                        h_o.copy_(d_o[b][0], non_blocking=True)
                        if want_mask and writer_post is None:
                            h_om = device.pinned_empty((m,), torch.uint8)
                            h_om.copy_(d_om[b][0], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(s_out)
                    pending.append((h_o, h_om, ev, was_ma))
                    i += 1
                    if len(pending) >= 2:                  # the device buffers of message i-2 are free again after this
                        yield finish(pending.popleft())
                while pending:
                    yield finish(pending.popleft())
        finally:
            stop.set()

    def _batch_result(self, table, out, omask, b, shape, propagate, pinned, values, writer_post):
        torch.cuda.current_stream(table.device).synchronize()
        res = out.numpy().reshape((b,) + shape)
        if writer_post is not None:
            return res
        if propagate:
            return ma.masked_array(res, mask=omask.numpy().view(np.bool_).reshape((b,) + shape), fill_value=self.mv_output,
                                   copy=False)
        if table.k == 1 and not pinned and any(isinstance(x, ma.MaskedArray) for x in values):
            return ma.masked_array(res)          # `v[indexes]` of a MaskedArray stays one, with no mask (:153)
        return res

    def writer_post(self, clone_mask, mv_write, netcdf=True, out_dtype=None):
        """the writers' post-processing as an epilogue for `interpolate_batch`: `clone_mask` = missing cells of
        the clone map (`Writer._mask`), `mv_write` = the file's missing value (netCDF: `default_fillvals[fmt] *
        scale_factor + offset`, writers/netcdf.py:149; PCRaster: nodata of the clone map, writers/pcr.py:29);
        `netcdf` adds the `out_v[out_v == mv_output] = nan` of writers/__init__.py:66.  `out_dtype=np.float32`: the fields
        come back as float32 - what a PCRaster REAL4 band (writers/pcr.py:38) or a netCDF `f4` variable stores anyway -
        with half the bytes over PCIe (tables with a window plan)."""
        return device.WriterPost(clone_mask, mv_write, match=self.mv_output if netcdf else None, out_dtype=out_dtype)

    @staticmethod
    def pinned_source_buffer(b, n):
        """page-locked [b, n] float64 buffer (torch tensor; `.numpy()` is a zero-copy view) for the GRIB decoder
        to write message values into (`codes_get_double_array` results); `interpolate_batch` then reads it
        from the device without any host-side staging copy."""
        return device.pinned_empty((b, n))

    def pinned_output_buffer(self, b, dtype=torch.float64):
        """page-locked [b, rows*cols] result buffer to pass as `out=` (float64) / `out_mask=` (uint8)."""
        return device.pinned_empty((b, self._target_coords.lons.size), dtype)

    def aux_for_intertable_generation(self, aux_g, aux_v, aux_g2, aux_v2):
        self._aux_gid, self._aux_val, self._aux_2nd_res_gid, self._aux_2nd_res_val = aux_g, aux_v, aux_g2, aux_v2

    @property
    def mv_output(self):
        return self.mv_out

    # ------------------------------------------------------------------ intertable lookup / creation
    def _intertable_filename(self, grid_id):
        """id `I{grid_id}_{maps id}{suffix}` -> (id, path); lookup order user dir, global dir, `.gz` optional
        (:68-101)."""
        intertable_id = '{}{}_{}{}'.format(self._prefix, grid_id.replace('$', '_'), self._target_coords.identifier,
                                           self._suffix)
        user_dir, global_dir = self._intertable_dirs.get('user'), self._intertable_dirs.get('global')
        known = self.intertables_config.vars

        def present(path):
            return os.path.exists(path) or os.path.exists(path + '.gz')

        if intertable_id not in known:
            if not self.create_if_missing:
                raise ApplicationException.get_exc(NO_INTERTABLE_CREATED, details=f'Using {intertable_id}')
            self.intertables_config.check_write()
            i, name = 0, self.format_intertablename(prognum='')
            while os.path.exists(os.path.normpath(os.path.join(user_dir, name))):
                i += 1
                name = self.format_intertablename(prognum=f'_{i}')
            return intertable_id, os.path.normpath(os.path.join(user_dir, name))
        filename = known[intertable_id]['filename']
        path = os.path.normpath(os.path.join(user_dir, filename)) if user_dir else None
        if not path or not present(path):
            path = os.path.normpath(os.path.join(global_dir, filename))
            if not present(path):
                if not self.create_if_missing:
                    raise ApplicationException.get_exc(NO_INTERTABLE_CREATED, details=f'Tried to create {path}')
                self.intertables_config.check_write()
                path = os.path.normpath(os.path.join(user_dir, filename))
                path = path if path.endswith('.gz') else path + '.gz'
                self._logger.warning(f'An entry in configuration was found for {filename} but intertable does not exist.')
        return intertable_id, path

    def _read_intertable(self, tbl_fullpath):
        """-> (indexes, coeffs) strided field views of the record array, cached per path (:103-139)."""
        rec = self._LOADED_INTERTABLES.get(tbl_fullpath)
        if rec is None:
            path = tbl_fullpath
            if not path.endswith('.gz') and not os.path.exists(path):
                path += '.gz'
                if not os.path.exists(path):
                    raise ApplicationException.get_exc(
                        NO_INTERTABLE_CREATED, details=f'Tried to read both {tbl_fullpath} and {path} but none was found')
            if path.endswith('.gz'):
                rec = load_npy_gz_parallel(path)
            else:
                rec = np.load(path)
            self._LOADED_INTERTABLES[tbl_fullpath] = rec
            self._logger.info(f'Using interpolation table: {tbl_fullpath}')
        return rec['indexes'], rec['coeffs']

    def _table_for(self, latgrib, longrib, v, grid_id, grid_details):
        intertable_id, path = self._intertable_filename(grid_id)
        dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
        key = (path, dev)
        if key in self._DEVICE_TABLES:
            return self._DEVICE_TABLES[key]
        nnear = self.scipy_modes_nnear[self._mode]
        shape = self._target_coords.lons.shape
        n_src = int(np.asarray(ma.getdata(v)).size)
        if os.path.exists(path) or os.path.exists(path + '.gz'):
            self._read_intertable(path)
            table = device.DeviceTable.from_record(self._LOADED_INTERTABLES[path], n_src, grid_shape=shape)
        elif self.create_if_missing:
            self.intertables_config.check_write()
            if latgrib is None:
                self._logger.error('Trying to interpolate without grib lat/lons. Probably a malformed grib!')
                raise ApplicationException.get_exc(MALFORMED_GRIB_LATLON)
            self._logger.warning(f'\nInterpolating table not found\n Id: {intertable_id}\nWill create file: {path}')
            si = ScipyInterpolation(longrib, latgrib, grid_details, np.asarray(ma.getdata(v)).ravel(), nnear,
                                    self.mv_out, self._mv_grib, target_is_rotated=self._rotated_target_grid,
                                    parallel=self.parallel, mode=self._mode, num_of_splits=self._num_of_splits)
            table = si.build_table(self._target_coords.lons, self._target_coords.lats)
            self.tie_rows = si.tie_rows()
            rec = table.to_record()
            # under torch.distributed every rank holds the whole table: rank 0 alone writes the file (to a temporary
            # name, then an atomic rename) and dumps the configuration; the others wait for it
            from . import parallel
            rank, ws = parallel.world()
            if rank == 0:
                tmp = f'{path}.tmp{os.getpid()}'
                if path.endswith('.gz'):
                    save_npy_gz_parallel(tmp, rec, level=self.gzip_level)
                else:
                    with open(tmp, 'wb') as f:
                        np.save(f, rec)
                os.replace(tmp, path)
            self.update_intertable_conf(rec, intertable_id, path, np.asarray(v).shape, dump=rank == 0)
            if ws > 1:
                torch.distributed.barrier()
        else:
            raise ApplicationException.get_exc(NO_INTERTABLE_CREATED, details=path)
        assert (table.k == 1 and nnear == 1) or table.k == nnear, f'table has k={table.k}, mode {self._mode} needs {nnear}'
        self._DEVICE_TABLES[key] = table
        return table

    def _pipeline(self, table, with_mask):
        pipes = table.__dict__.setdefault('_host_pipes', {})
        if with_mask not in pipes:
            _, touched = table.compact()
            if table.k in (1, 4) and touched.numel() * 4 < table.n_src:
                # regional target: gather the touched values straight out of the pinned buffers (zero-copy)
                pipes[with_mask] = device.CompactHostPipeline(table, chunk_messages=16, with_mask=with_mask)
            else:
                chunk = max(1, min(16, (256 << 20) // (table.n_src * 8)))  # ~256 MB of pinned staging per slot
                pipes[with_mask] = device.HostApplyPipeline(table, chunk_messages=chunk, with_mask=with_mask)
        return pipes[with_mask]

    def update_intertable_conf(self, intertable, intertable_id, intertable_name, source_shape, dump=True):
        """:373-384 - remember the new table in the (user) intertables configuration."""
        self._LOADED_INTERTABLES[intertable_name] = intertable
        item = {'filename': os.path.basename(intertable_name), 'method': self._mode, 'source_shape': source_shape,
                'target_shape': self._target_coords.lons.shape}
        self.intertables_config.vars[intertable_id] = item
        self.intertables_config.user_vars[intertable_id] = item
        if dump:
            self.intertables_config.dump()
