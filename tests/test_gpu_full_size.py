"""The benchmarked configurations at FULL size against the oracle, every row:

* BASELINE configs[3]: O1280 -> EFAS 5 km invdist (k=4) table built on the GPU (float64 maps and PCRaster-style float32
  maps with a missing-value region), the windowed apply kernel on 8 masked messages under every mask policy against
  `R.apply_propagate` / `R.apply_as_reference` on all 950 000 rows; the table itself against the oracle's cKDTree build.
* BASELINE configs[4]: x*1000 -> 24 h accumulation -> cut-off -> invdist apply -> p+gem-dem*0.0065 at O1280 -> EFAS, masked
  inputs, against the oracle's stage-by-stage chain (oracle/manip_ref.py + oracle/interp_ref.py).
"""
import numpy as np
import numpy.ma as ma
import pytest
import torch

from oracle import interp_ref as R
from oracle import manip_ref as MR
from pyg2p_b200 import _lib as L
from pyg2p_b200 import synth

pytestmark = pytest.mark.gpu

MV_OUT = synth.NC_FILL_F64


@pytest.fixture(scope='module')
def dev():
    from pyg2p_b200 import device
    device.require_cuda()
    return device


@pytest.fixture(scope='module')
def o1280():
    return synth.octahedral_grid(1280)


def _targets(kind):
    if kind == 'f64':
        return synth.efas_target()
    tl, tlo = synth.efas_target(dtype=np.float32)            # PCRaster REAL4 maps, 30 % missing-value cells
    return synth.with_mv_region(tl, tlo)


def _padded(v, dtype):
    b, n = v.shape
    n_pad = -(-n // 16) * 16
    buf = torch.zeros((b, n_pad), dtype=dtype, device='cuda')
    buf[:, :n] = torch.from_numpy(v).cuda()
    return buf[:, :n]


@pytest.mark.parametrize('maps', ['f64', 'f32'])
def test_headline_table_and_apply_all_rows(dev, o1280, maps):
    lat, lon, det = o1280
    n = lat.size
    tl, tlo = _targets(maps)
    m = tl.size
    ix = dev.SourceIndex(lon, lat, det['radius'])
    mub = ix.min_upper_bound(det['Nj'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=mub, want_flags=True)
    table = dev.DeviceTable(out['idx'], out['coef'], n, coef_f32=(maps == 'f32'), grid_shape=tl.shape)
    idx, w = table.indexes_coeffs()
    assert idx.shape == (m, 4) and w.dtype == (np.float32 if maps == 'f32' else np.float64)

    # ---- the table against the oracle's build (cKDTree, numpy weights) on all rows
    o = R.BuildOracle(lon, lat, det, np.zeros(n), 4, MV_OUT, synth.GRIB_MV, parallel=True, mode='invdist')
    assert mub == o.min_upper_bound
    _, w_ref, i_ref = o.interpolate(tlo, tl)
    tie = (out['flags'].cpu().numpy() & L.FLAG_TIE) != 0    # exact distance ties: cKDTree's order is traversal-dependent
    assert tie.sum() < 50
    same_idx = (idx == i_ref).all(axis=1)
    if maps == 'f32':
        # float32 maps: xyz bit-identical to the reference's (glibc sinf / cosf restated), garbage MV coordinates included
        assert same_idx[~tie].all()
        assert np.array_equal(w[~tie], w_ref[~tie], equal_nan=True)
    else:
        # float64 maps: the device's double-double trig differs from libm by 1 ulp in ~0.1 % of the values, i.e. target
        # positions move by up to a few ulp of the radius (1e-9 m).  A neighbour pair at nearly equal distance can swap;
        # everywhere else the indexes are identical and the weights agree within 1e-12 plus what that displacement
        # does to 1/d^2 for the (rare) targets that sit within metres of a source point: |dw/w| <= 4 delta / d_min
        assert same_idx.mean() > 0.9999
        ok = same_idx & ~tie
        assert np.array_equal(np.isnan(w[ok]), np.isnan(w_ref[ok]))
        fin = np.isfinite(w_ref[ok]).all(axis=1)
        rel = np.abs(w[ok][fin] - w_ref[ok][fin]) / np.abs(w_ref[ok][fin])
        delta = 4 * np.spacing(det['radius'])
        bound = 1e-12 + 4 * delta / o.raw_distances[ok][fin][:, :1]
        assert (rel <= bound).all(), float((rel / bound).max())
        assert (rel <= 1e-12).mean() > 0.9999
        assert (w[ok] == w_ref[ok]).mean() > 0.95

    # ---- the windowed apply on 8 masked messages, every mask policy, all rows, against the oracle on the GPU's table
    b = 8
    rng = np.random.default_rng(5)
    base_mask = synth.missing_mask(lat)
    fields = np.stack([synth.field_2t(lat, lon, member=i) for i in range(b)])
    masks = np.stack([base_mask | (rng.random(n) < 0.005 * i) for i in range(b)])
    fields[masks] = synth.GRIB_MV
    src = _padded(fields, torch.float64)
    msk = _padded(masks.view(np.uint8), torch.uint8)
    n_pad = src.stride(0)
    bits = torch.from_numpy(dev.pack_mask_bits(masks, n_pad)).cuda()
    marked = src.clone()
    dev.mark_masked(marked, msk)
    plain = table.apply(src, MV_OUT)
    assert table.last_path == 'window', table.plan_error
    p1, m1 = table.apply(src, MV_OUT, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
    p2 = table.apply(src, MV_OUT, src_mask=msk, mask_policy=L.MASK_FILL)
    p3, m3 = table.apply(src, MV_OUT, src_mask=bits, mask_policy=L.MASK_PROPAGATE_BITS)
    p4, m4 = table.apply(marked, MV_OUT, mask_policy=L.MASK_MARKED)
    assert table.last_path == 'window'
    m3, m4 = dev.unpack_mask_bits(m3, m), dev.unpack_mask_bits(m4, m)
    # the table as `_read_intertable` hands it to the apply: strided field views of the record array (with these numpy's
    # einsum evaluates ((w0*v0 + w1*v1) + w2*v2) + w3*v3), coefficients in the table's own dtype
    rec = table.to_record()
    w64, idx = rec['coeffs'], rec['indexes']
    for i in range(b):
        with np.errstate(all='ignore'):
            want_plain = R.apply_as_reference(ma.masked_array(fields[i], mask=masks[i]), w64, idx, 4, MV_OUT)
            want, want_mask = R.apply_propagate(fields[i], w64, idx, 4, MV_OUT, masks[i])
        np.testing.assert_array_equal(plain[i].cpu().numpy(), np.asarray(ma.getdata(want_plain)))
        for got, gm in ((p1, m1[i].cpu().numpy().astype(bool)), (p2, None), (p3, m3[i]), (p4, m4[i])):
            np.testing.assert_array_equal(got[i].cpu().numpy(), want)
            if gm is not None:
                np.testing.assert_array_equal(gm, want_mask)
    assert 0.02 < want_mask.mean() < 0.6


def test_configs4_chain_all_rows(dev, o1280):
    """eue_r24-style template at O1280 -> EFAS: `x*1000`, 24 h accumulation of steps 0..96 by 6 h, cut-off, invdist
    apply, lapse-rate correction with the geopotential interpolated through the same table - the two-launch pipeline
    (g2p_manip on the touched points + windowed apply with the Corrector epilogue) against the oracle chain on every
    row, without masks (corrected) and with masked inputs (propagated)."""
    from pyg2p_b200 import manipulation as M
    lat, lon, det = o1280
    n = lat.size
    tl, tlo = synth.efas_target()
    ix = dev.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    table = dev.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    idx, w = table.indexes_coeffs()
    steps = list(range(0, 97, 6))
    fields = np.cumsum(np.stack([synth.field_tp_increment(lat, lon, member=3, step=s) for s in steps]), axis=0)
    fields -= 1e-4                                            # some negative accumulations for the cut-off
    mask = synth.missing_mask(lat)
    conv = M.Converter(func='x=x*1000', cut_off=True)
    conv.set_missing_value(synth.GRIB_MV)
    agg = M.Aggregator(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6, step_type='accum',
                       start_step=0, end_step=96, unit_time=24, mv_grib=synth.GRIB_MV, force_zero_array=False)
    plan = agg.plan(steps)
    dem = synth.field_dem(tl, tlo, synth.PCR_MV_F32).astype(np.float32)
    z = synth.field_geopotential(lat, lon)
    z[::4099] = synth.GRIB_MV
    corr = M.Corrector.from_geopotential(table, z, synth.GRIB_MV, MV_OUT, dem, float(synth.PCR_MV_F32))
    pipe = M.FusedPipeline(table)
    src = _padded(fields, torch.float64)
    got = pipe.run(src, None, plan, converter=conv, cut_off=True, mv_out=MV_OUT, corrector=corr).cpu().numpy()
    assert pipe.last_path == 'two-kernel' and pipe.compact.last_path == 'window'
    one = pipe.run(src, None, plan, converter=conv, cut_off=True, mv_out=MV_OUT, corrector=corr, fused=True).cpu().numpy()
    assert pipe.last_path == 'fused'
    np.testing.assert_array_equal(one, got)                  # the single-launch variant gives the same bytes
    # ---- oracle chain
    converted = {s: MR.convert(fields[i].copy(), 'x=x*1000', synth.GRIB_MV) for i, s in enumerate(steps)}
    ref = MR.accumulation(converted, 0, 96, 24, 24)
    keys = sorted(ref, key=lambda k: (k[1], k[0]))
    assert [k for k, _ in plan.outputs] == keys and got.shape == (len(keys), tl.size) and len(keys) == 4
    gem_src = MR.gem_from_geopotential(z, synth.GRIB_MV)
    with np.errstate(all='ignore'):
        gem = R.apply_sequential(gem_src, w, idx, 4, MV_OUT)
        for o_, k in enumerate(keys):
            v = MR.cut_off_negative(np.asarray(ma.getdata(ref[k])))
            p = R.apply_sequential(v, w, idx, 4, MV_OUT)
            want = MR.correct(p, dem.ravel(), gem, float(synth.PCR_MV_F32), synth.GRIB_MV)
            np.testing.assert_array_equal(got[o_], np.asarray(ma.filled(want)))
    # ---- masked inputs: the output is masked wherever a used input is
    fields_m = fields.copy()
    fields_m[:, mask] = synth.GRIB_MV
    srcm = _padded(fields_m, torch.float64)
    msk = _padded(np.broadcast_to(mask, fields.shape).copy().view(np.uint8), torch.uint8)
    gotm, gmask = pipe.run(srcm, msk, plan, converter=conv, cut_off=True, mv_out=MV_OUT, mask_policy=L.MASK_PROPAGATE)
    gotm, gmask = gotm.cpu().numpy(), gmask.cpu().numpy().astype(bool)
    conv_m = {s: ma.masked_where(mask, MR.convert(fields_m[i].copy(), 'x=x*1000', synth.GRIB_MV)) for i, s in enumerate(steps)}
    refm = MR.accumulation(conv_m, 0, 96, 24, 24)
    with np.errstate(all='ignore'):
        for o_, k in enumerate(keys):
            v = MR.cut_off_negative(np.asarray(ma.getdata(refm[k])))
            mk = ma.getmaskarray(refm[k])
            data, dm = R.apply_propagate(v, w, idx, 4, MV_OUT, mk)
            np.testing.assert_array_equal(gotm[o_], data)
            np.testing.assert_array_equal(gmask[o_], dm)
    assert 0.02 < gmask.mean() < 0.6
