"""The drop-in `Interpolator` / `ScipyInterpolation` on the GPU, driven the way the reference's own tests
drive them (reference tests/test_interpolation.py:13-46), but with numeric checks against the oracle."""
import gzip
import os
import shutil

import numpy as np
import numpy.ma as ma
import pytest

from oracle import interp_ref as R
from pyg2p_b200 import synth
from pyg2p_b200.exceptions import MALFORMED_GRIB_LATLON, NO_INTERTABLE_CREATED, ApplicationException
from tests.helpers import MockedExecutionContext, write_csf_real4

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def api():
    from pyg2p_b200 import device
    device.require_cuda()
    from pyg2p_b200.interpolator import Interpolator
    from pyg2p_b200.scipy_interpolation import ScipyInterpolation
    return Interpolator, ScipyInterpolation


def _target_maps(tmp_path, rows=60, cols=45):
    """PCRaster REAL4 maps of a small LAEA grid with an MV corner, like the EFAS lat.map / lon.map."""
    lat, lon = synth.efas_target(rows=rows, cols=cols, cell=25000.0, x_ul=3800000.0, y_ul=3500000.0, dtype=np.float32)
    lat[-8:, :10] = np.nan
    lon[-8:, :10] = np.nan
    write_csf_real4(tmp_path / 'lat.map', lat, 3800000.0, 3500000.0, 25000.0)
    write_csf_real4(tmp_path / 'lon.map', lon, 3800000.0, 3500000.0, 25000.0)
    return str(tmp_path / 'lat.map'), str(tmp_path / 'lon.map')


def _ctx(tmp_path, mode, create=True, entries=None):
    lat_map, lon_map = _target_maps(tmp_path)
    d = {'interpolation.dirs': {'user': str(tmp_path), 'global': str(tmp_path)}, 'interpolation.latMap': lat_map,
         'interpolation.lonMap': lon_map, 'interpolation.mode': mode, 'interpolation.create': create,
         'interpolation.parallel': True, 'input.file': 'tests/data/input.grib'}
    return MockedExecutionContext(d, False, entries)


def _source():
    lat, lon, det = synth.regular_ll_grid(62.0, 38.0, 97, -8.0, 32.0, 161)
    return lat, lon, det


@pytest.mark.parametrize('mode,k', [('nearest', 1), ('invdist', 4), ('adw', 11)])
def test_interpolation_create_and_reuse(api, tmp_path, mode, k):
    """build on first use (-B), file + configuration written in the reference's layout, reused afterwards."""
    Interpolator, _ = api
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    lat, lon, det = _source()
    v = synth.field_2t(lat, lon)
    ctx = _ctx(tmp_path, mode)
    it = Interpolator(ctx, synth.GRIB_MV)
    out = it.interpolate_scipy(lat, lon, v, '0.0$0.25$161$97$15617$regular_ll', det)
    assert out.shape == (60, 45)                                     # what the reference's test asserts
    path = tmp_path / f'tbl_input_2700_scipy_{mode}.npy.gz'
    assert path.exists() and ctx.configuration.intertables.dumps == 1
    (tid, item), = ctx.configuration.intertables.user_vars.items()
    assert item == {'filename': path.name, 'method': mode, 'source_shape': v.shape, 'target_shape': (60, 45)}
    with gzip.GzipFile(path) as f:
        rec = np.load(f)
    # float32 maps -> float32 distances (nearest) and float32 adw weights; invdist stores its weights into float64
    want_dt = [('indexes', '<i8'), ('coeffs', '<f8' if mode == 'invdist' else '<f4')]
    assert rec.dtype == np.dtype(want_dt) and rec.shape == ((2700,) if k == 1 else (2700, k))
    # numbers: the oracle's apply of the table that was written == what the call returned
    with np.errstate(all='ignore'):
        want = R.apply_as_reference(v, rec['coeffs'], rec['indexes'], k, it.mv_output).reshape(60, 45)
    np.testing.assert_array_equal(out, want)
    # and the table itself against the oracle's build (float32 maps: positions agree to ~1 m, see test_gpu_build)
    o = R.BuildOracle(lon, lat, det, v, k, it.mv_output, synth.GRIB_MV, mode=mode)
    _, w_ref, i_ref = o.interpolate(it._target_coords.lons, it._target_coords.lats)
    same = (rec['indexes'].reshape(2700, -1) == i_ref.reshape(2700, -1)).all(axis=1)
    assert same.mean() > 0.99
    # a second Interpolator finds the file through the configuration and applies it without rebuilding
    Interpolator._DEVICE_TABLES.clear()
    ctx2 = _ctx(tmp_path, mode, create=False, entries=ctx.configuration.intertables.vars)
    out2 = Interpolator(ctx2, synth.GRIB_MV).interpolate(lat, lon, v, '0.0$0.25$161$97$15617$regular_ll', det)
    np.testing.assert_array_equal(out2, out)
    assert ctx2.configuration.intertables.dumps == 0


def test_reference_golden_table_applies(api, tmp_path, golden_dir):
    """reference test_interpolation_use_scipy_nearest: the golden table of the reference's repository is found
    through an intertables.json entry and applied; shape (810, 680) as asserted there, values vs the oracle."""
    Interpolator, _ = api
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    shutil.copy(os.path.join(golden_dir, 'ref_tbl_pf10slhf_550800_scipy_nearest.npy.gz'),
                tmp_path / 'tbl_pf10slhf_550800_scipy_nearest.npy.gz')
    lat = np.full((810, 680), 50.0, np.float32)
    lon = np.full((810, 680), 10.0, np.float32)
    write_csf_real4(tmp_path / 'lat.map', lat, -1700000.0, 2700000.0, 5000.0)
    write_csf_real4(tmp_path / 'lon.map', lon, -1700000.0, 2700000.0, 5000.0)
    d = {'interpolation.dirs': {'user': str(tmp_path), 'global': str(tmp_path)},
         'interpolation.latMap': str(tmp_path / 'lat.map'), 'interpolation.lonMap': str(tmp_path / 'lon.map'),
         'interpolation.mode': 'nearest', 'input.file': 'tests/data/input.grib'}
    ctx = MockedExecutionContext(d, False)
    it = Interpolator(ctx, synth.GRIB_MV)
    gid = '-15.75$16.125$511$415$212065$rotated_ll'
    ctx.configuration.intertables.vars['I' + gid.replace('$', '_') + '_' + it._target_coords.identifier + 'scipy_nearest'] = {
        'filename': 'tbl_pf10slhf_550800_scipy_nearest.npy', 'method': 'nearest'}
    v = np.random.default_rng(0).standard_normal(212065) + 280
    out = it.interpolate_scipy(None, None, v, gid)
    assert out.shape == (810, 680)
    idx, coef = it._read_intertable(str(tmp_path / 'tbl_pf10slhf_550800_scipy_nearest.npy'))
    np.testing.assert_array_equal(out.ravel(), R.apply_as_reference(v, coef, idx, 1, it.mv_output))


def test_errors(api, tmp_path):
    Interpolator, _ = api
    Interpolator._DEVICE_TABLES.clear()
    lat, lon, det = _source()
    v = synth.field_2t(lat, lon)
    with pytest.raises(ApplicationException) as e:
        Interpolator(_ctx(tmp_path, 'nearest', create=False), synth.GRIB_MV).interpolate_scipy(lat, lon, v, 'g$1', det)
    assert e.value.code == NO_INTERTABLE_CREATED
    with pytest.raises(ApplicationException) as e:
        Interpolator(_ctx(tmp_path, 'invdist', create=True), synth.GRIB_MV).interpolate_scipy(None, None, v, 'g$2', det)
    assert e.value.code == MALFORMED_GRIB_LATLON


def test_batch_equals_per_message_and_mask_policies(api, tmp_path):
    Interpolator, _ = api
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    lat, lon, det = _source()
    rng = np.random.default_rng(1)
    fields = [synth.field_2t(lat, lon, member=i) for i in range(21)]
    mask = rng.random(lat.size) < 0.05
    masked = [ma.masked_where(mask, np.where(mask, synth.GRIB_MV, f)) for f in fields]
    it = Interpolator(_ctx(tmp_path, 'invdist'), synth.GRIB_MV)
    batch = it.interpolate_batch(lat, lon, fields, 'g$3', det)
    assert batch.shape == (21, 60, 45) and not isinstance(batch, ma.MaskedArray)
    for i in (0, 7, 20):
        np.testing.assert_array_equal(batch[i], it.interpolate_scipy(lat, lon, fields[i], 'g$3', det))
    # messages decoded into a pinned buffer go through without a host staging copy
    buf = it.pinned_source_buffer(len(fields), lat.size)
    buf.numpy()[:] = np.stack(fields)
    np.testing.assert_array_equal(it.interpolate_batch(lat, lon, buf, 'g$3', det), batch)
    # as_reference: the masks are ignored, payloads (9999) flow into the sums - exactly the reference's output
    idx, coef = it._read_intertable(it._intertable_filename('g$3')[1])
    ref = it.interpolate_batch(lat, lon, masked, 'g$3', det)
    with np.errstate(all='ignore'):
        want = R.apply_as_reference(masked[3], coef, idx, 4, it.mv_output)
    np.testing.assert_array_equal(ref[3].ravel(), np.asarray(want))
    # propagate: MaskedArray, masked where any neighbour is masked
    it.mask_policy = 'propagate'
    prop = it.interpolate_batch(lat, lon, masked, 'g$3', det)
    assert isinstance(prop, ma.MaskedArray)
    with np.errstate(all='ignore'):
        data, mk = R.apply_propagate(masked[3], coef, idx, 4, it.mv_output, mask)
    np.testing.assert_array_equal(ma.getmaskarray(prop[3]).ravel(), mk)
    np.testing.assert_array_equal(ma.getdata(prop[3]).ravel(), data)
    # the writers' per-step post-processing hoisted into the batch: fields arrive as the writer would store them
    clone = rng.random(prop.shape[1:]) < 0.2
    mv_nc = 9.969209968386869e36 * 0.1 + 5.0
    post = it.writer_post(clone, mv_nc, netcdf=True)
    ready = it.interpolate_batch(lat, lon, masked, 'g$3', det, writer_post=post)
    assert isinstance(ready, np.ndarray) and not isinstance(ready, ma.MaskedArray)
    np.testing.assert_array_equal(ready, R.writer_post(ma.getdata(prop), clone, mv_nc, mv_output=it.mv_output,
                                                       value_mask=ma.getmaskarray(prop)))
    # ... and as float32, the way a PCRaster REAL4 band / a netCDF f4 variable stores them: the float64 result cast
    # (round to nearest) element by element, half the bytes over PCIe
    post32 = it.writer_post(clone, -3.4028234663852886e38, netcdf=True, out_dtype=np.float32)
    ready32 = it.interpolate_batch(lat, lon, masked, 'g$3', det, writer_post=post32)
    assert ready32.dtype == np.float32
    want64 = R.writer_post(ma.getdata(prop), clone, -3.4028234663852886e38, mv_output=it.mv_output, value_mask=ma.getmaskarray(prop))
    np.testing.assert_array_equal(ready32, want64.astype(np.float32))
    got32 = list(it.interpolate_stream(lat, lon, iter(masked), 'g$3', det, writer_post=post32))
    np.testing.assert_array_equal(np.stack(got32), ready32)
    it.mask_policy = 'as_reference'
    ready = it.interpolate_batch(lat, lon, masked, 'g$3', det, writer_post=it.writer_post(clone, it.mv_output, netcdf=False))
    np.testing.assert_array_equal(ready, R.writer_post(ref, clone, it.mv_output))


def test_scipy_interpolation_signature_and_outputs(api):
    """`ScipyInterpolation(...).interpolate(lon, lat)` returns what the reference returns: shapes, dtypes,
    sentinel rows, NaN weights - compared with the oracle on float64 maps (bit-exact off near-ties)."""
    _, ScipyInterpolation = api
    lat, lon, det = _source()
    z = synth.field_2t(lat, lon)
    tl, tlo = synth.efas_target(rows=70, cols=50, cell=40000.0, x_ul=3000000.0, y_ul=4500000.0)
    for mode, k in (('nearest', 1), ('invdist', 4)):
        si = ScipyInterpolation(lon, lat, det, z, k, synth.NC_FILL_F64, synth.GRIB_MV, parallel=True, mode=mode)
        res, w, idx = si.interpolate(tlo, tl)
        o = R.BuildOracle(lon, lat, det, z, k, synth.NC_FILL_F64, synth.GRIB_MV, mode=mode)
        assert si.min_upper_bound == pytest.approx(o.min_upper_bound, rel=1e-12)
        r_ref, w_ref, i_ref = o.interpolate(tlo, tl)
        assert idx.shape == i_ref.shape and idx.dtype == i_ref.dtype
        assert w.shape == w_ref.shape and w.dtype == w_ref.dtype
        same = (idx.reshape(3500, -1) == i_ref.reshape(3500, -1)).all(axis=1)
        assert same.mean() > 0.995 and (idx == z.size).any() and (idx < z.size).any()
        np.testing.assert_allclose(w.reshape(3500, -1)[same], w_ref.reshape(3500, -1)[same], rtol=1e-11, atol=1e-7)
        ok = same & ((idx.reshape(3500, -1)[:, 0] < z.size) if k > 1 else np.ones(3500, bool))
        np.testing.assert_allclose(np.asarray(ma.getdata(res))[ok], np.asarray(ma.filled(r_ref))[ok], rtol=1e-12)
    with pytest.raises(ApplicationException):
        ScipyInterpolation(lon, lat, det, z, 11, 0.0, 0.0, mode='cdd')


def test_compact_zero_copy_pipeline_equals_full_copy(api):
    """the zero-copy host pipeline (gather of the touched values out of pinned memory) gives the bytes of the
    whole-field pipeline and of the device-resident apply, with and without masks, b > 2 chunks."""
    import torch
    from pyg2p_b200 import _lib as L
    from pyg2p_b200 import device
    lat, lon, det = synth.octahedral_grid(160)
    tl, tlo = synth.efas_target(rows=190, cols=200, cell=25000.0)
    ix = device.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    table = device.DeviceTable(out['idx'], out['coef'], lat.size, grid_shape=tl.shape)
    b = 37
    g = torch.Generator().manual_seed(4)
    h_src = device.pinned_empty((b, lat.size))
    h_src.copy_(torch.randn((b, lat.size), generator=g, dtype=torch.float64))
    h_msk = device.pinned_empty((b, lat.size), torch.uint8)
    h_msk.copy_((torch.rand((b, lat.size), generator=g) < 0.03).to(torch.uint8))
    full = device.HostApplyPipeline(table, chunk_messages=5, with_mask=True)
    comp = device.CompactHostPipeline(table, chunk_messages=16, with_mask=True)
    assert comp.ct.n_src < lat.size // 4
    for policy in (L.MASK_AS_REFERENCE, L.MASK_PROPAGATE, L.MASK_FILL):
        o1, m1 = full.run(h_src, 1e20, src_mask=h_msk, mask_policy=policy)
        o2, m2 = comp.run(h_src, 1e20, src_mask=h_msk, mask_policy=policy)
        torch.cuda.synchronize()
        assert torch.equal(o1.view(torch.int64), o2.view(torch.int64))
        assert (m1 is None and m2 is None) or torch.equal(m1, m2)
    ref = table.apply(h_src.cuda(), 1e20).cpu()
    o3, _ = comp.run(h_src, 1e20)
    torch.cuda.synchronize()
    assert torch.equal(ref.view(torch.int64), o3.view(torch.int64))
    assert comp.h2d_bytes < full.h2d_bytes / 4


@pytest.mark.parametrize('source', ['global', 'regional'])
def test_stream_of_pageable_messages_equals_batch(api, tmp_path, source):
    """`interpolate_stream` (the writers' per-step loop as a generator over pageable per-message arrays, the touched
    values picked by a host thread while the device works) gives the results of `interpolate_batch`, in order, under both
    mask policies and with the writers' post-processing; a global source takes the compact pinned-ring path, a regional
    one (the table reads most of the grid) the per-message fallback."""
    from pyg2p_b200 import device
    Interpolator, _ = api
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    lat, lon, det = synth.octahedral_grid(96) if source == 'global' else _source()
    rng = np.random.default_rng(2)
    n_msg = 11
    fields = [synth.field_2t(lat, lon, member=i) for i in range(n_msg)]
    masks = [rng.random(lat.size) < 0.02 * (i % 3) for i in range(n_msg)]
    msgs = [ma.masked_where(mk, np.where(mk, synth.GRIB_MV, f)) if i % 4 else f for i, (f, mk) in enumerate(zip(fields, masks))]
    it = Interpolator(_ctx(tmp_path, 'invdist'), synth.GRIB_MV)
    if source == 'regional':        # force the whole-field pipeline, as for a table that reads most of the source grid
        pipes = {}
        it._pipeline = lambda table, with_mask: pipes.setdefault(
            with_mask, device.HostApplyPipeline(table, chunk_messages=4, with_mask=with_mask))
    for policy in ('as_reference', 'propagate'):
        it.mask_policy = policy
        want = it.interpolate_batch(lat, lon, msgs, 'g$5', det)
        got = list(it.interpolate_stream(lat, lon, iter(msgs), 'g$5', det, depth=3))
        assert len(got) == n_msg
        table = it._table_for(lat, lon, fields[0], 'g$5', det)
        assert isinstance(it._pipeline(table, policy == 'propagate'), device.CompactHostPipeline) == (source == 'global')
        for i in range(n_msg):
            np.testing.assert_array_equal(ma.getdata(got[i]), ma.getdata(want[i]))
            if policy == 'propagate':
                np.testing.assert_array_equal(ma.getmaskarray(got[i]), ma.getmaskarray(want[i]))
    clone = rng.random(want.shape[1:]) < 0.2
    post = it.writer_post(clone, 1e20, netcdf=True)
    want = it.interpolate_batch(lat, lon, msgs, 'g$5', det, writer_post=post)
    got = list(it.interpolate_stream(lat, lon, iter(msgs), 'g$5', det, writer_post=post))
    np.testing.assert_array_equal(np.stack(got), want)
    assert list(it.interpolate_stream(lat, lon, iter([]), 'g$5', det)) == []
    with pytest.raises(ValueError):
        list(it.interpolate_stream(lat, lon, iter([fields[0], fields[1][:-3]]), 'g$5', det))


@pytest.mark.parametrize('mode', ['grib_nearest', 'grib_invdist'])
def test_grib_layout_tables_apply(api, tmp_path, mode):
    """intertables in the `grib_nearest` ([xs, ys, idxs]) and `grib_invdist` (record array (6, M'))
    layouts are applied on the GPU, against the literal restatement of the reference's three lines."""
    Interpolator, _ = api
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    rng = np.random.default_rng(6)
    n, rows, cols = 5000, 60, 45
    valid = rng.random((rows, cols)) < 0.8                         # targets outside the GRIB domain are not listed
    xs, ys = np.nonzero(valid)
    if mode == 'grib_nearest':
        tbl = np.asarray([xs, ys, rng.integers(0, n, xs.size)])
    else:
        ind = np.asarray([xs, ys] + [rng.integers(0, n, xs.size) for _ in range(4)])
        c = rng.random((4, xs.size))
        c /= c.sum(0)
        tbl = np.rec.fromarrays((ind, np.asarray(list(c) + [np.zeros(xs.size), np.zeros(xs.size)])),
                                names=('indexes', 'coeffs'))
    fname = f'tbl_input_2700_{mode}.npy'
    np.save(tmp_path / fname, tbl)
    lat_map, lon_map = _target_maps(tmp_path)
    d = {'interpolation.dirs': {'user': str(tmp_path), 'global': str(tmp_path)}, 'interpolation.latMap': lat_map,
         'interpolation.lonMap': lon_map, 'interpolation.mode': mode, 'input.file': 'tests/data/input.grib'}
    ctx = MockedExecutionContext(d, True)
    it = Interpolator(ctx, synth.GRIB_MV)
    ctx.configuration.intertables.vars['Ig_1_' + it._target_coords.identifier + mode] = {'filename': fname}
    v = rng.standard_normal(n) * 10 + 280
    out = it.interpolate(None, None, v, 'g$1', None, gid=-1)
    want = (R.apply_grib_nearest(v, xs, ys, tbl[2], (rows, cols), it.mv_output) if mode == 'grib_nearest'
            else R.apply_grib_invdist(v, tbl, (rows, cols), it.mv_output))
    np.testing.assert_array_equal(out, want)
    assert (out == it.mv_output).sum() == (~valid).sum()
    vm = ma.masked_where(rng.random(n) < 0.1, v)
    outm = it.interpolate(None, None, vm, 'g$1', None, gid=-1)
    wantm = (R.apply_grib_nearest(vm, xs, ys, tbl[2], (rows, cols), it.mv_output) if mode == 'grib_nearest'
             else R.apply_grib_invdist(vm, tbl, (rows, cols), it.mv_output))
    np.testing.assert_array_equal(outm, wantm)
    # all messages of a run at once, through the batched host pipeline (masked and plain ones mixed)
    batch = it.interpolate_grib_batch([v, vm, v * 2.0, vm + 1.0], -1, 'g$1')
    assert batch.shape == (4, rows, cols)
    np.testing.assert_array_equal(batch[0], want)
    np.testing.assert_array_equal(batch[1], wantm)
    want3 = (R.apply_grib_nearest(vm + 1.0, xs, ys, tbl[2], (rows, cols), it.mv_output) if mode == 'grib_nearest'
             else R.apply_grib_invdist(vm + 1.0, tbl, (rows, cols), it.mv_output))
    np.testing.assert_array_equal(batch[3], want3)
    assert it.interpolate_grib_batch([], -1, 'g$1').shape == (0, rows, cols)
