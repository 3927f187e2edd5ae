"""GPU parity of the fused manipulation stages (K8): Converter / Aggregator / Corrector against the golden
vectors produced by the reference's own classes (tests/golden/manip.npz) and the whole pipeline
convert -> aggregate -> cut-off -> apply -> correct against the oracle chain.  Bit-exact throughout."""
import os

import numpy as np
import numpy.ma as ma
import pytest
import torch

from oracle import interp_ref as R
from oracle import manip_ref as MR
from pyg2p_b200 import _lib as L
from pyg2p_b200 import synth

pytestmark = pytest.mark.gpu
RES = 640


@pytest.fixture(scope='module')
def M():
    from pyg2p_b200 import device
    device.require_cuda()
    from pyg2p_b200 import manipulation
    return manipulation


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'manip.npz'))


def test_converter_matches_reference(M, g):
    x = g['conv_x']
    for i, f in enumerate(['x=x*1000', 'x=x-273.15', 'x=x*(-0.0353)', 'x=-x*1000', 'x=x']):
        c = M.Converter(func=f, cut_off=True)
        c.set_missing_value(synth.GRIB_MV)
        y = c.convert(x.copy())
        np.testing.assert_array_equal(np.asarray(y), g[f'conv{i}'])
        np.testing.assert_array_equal(c.cut_off_negative(np.asarray(y).copy()), g[f'conv{i}_cut'])
    xm = ma.masked_where(x > 0.01, x)
    ym = M.Converter(func='x=x*1000').convert(xm)
    assert isinstance(ym, ma.MaskedArray) and np.array_equal(ym.mask, xm.mask)
    d = M.Converter(func='x=x*1000', cut_off=True).cut_off_negative({'a': x.copy(), 'b': -x})
    np.testing.assert_array_equal(d['b'], np.where(-x < 0, 0, -x))


def _check(g, tag, out):
    keys = sorted(out.keys(), key=lambda s: (s.end_step, s.start_step))
    np.testing.assert_array_equal(np.array([(s.start_step, s.end_step) for s in keys]), g[f'{tag}_keys'])
    np.testing.assert_array_equal(np.array([np.asarray(ma.getdata(out[k])) for k in keys]), g[f'{tag}_vals'])
    n = g['conv_x'].size
    masks = np.array([np.broadcast_to(ma.getmaskarray(out[k]) if isinstance(out[k], ma.MaskedArray) else False, (n,))
                      for k in keys])
    np.testing.assert_array_equal(masks, g[f'{tag}_masks'])


def test_aggregator_matches_reference(M, g):
    steps = list(g['agg_steps'])
    cum, mask, inst = g['agg_cum'], g['agg_mask'], g['inst_fields']

    def vals(stepset, src=cum, masked=False, start_is_end=False):
        d = {}
        for s in stepset:
            v = src[steps.index(s)].copy()
            if masked:
                v = ma.masked_where(mask, v, copy=True)
            d[M.Step(int(s) if start_is_end else 0, int(s), RES, 6, 0)] = v
        return d

    def agg(**kw):
        base = dict(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6, step_type='accum',
                    start_step=0, end_step=72, unit_time=24, mv_grib=synth.GRIB_MV, force_zero_array=False)
        base.update(kw)
        return M.Aggregator(**base)

    _check(g, 'acc_full', agg().do_manipulation(vals(steps)))
    _check(g, 'acc_nozero', agg().do_manipulation(vals(steps[1:])))
    _check(g, 'acc_gap', agg().do_manipulation(vals([s for s in steps if s != 48])))
    _check(g, 'acc_masked', agg().do_manipulation(vals(steps, masked=True)))
    _check(g, 'acc_forcezero', agg(force_zero_array=True).do_manipulation(vals(steps)))
    _check(g, 'acc_step12_unit6', agg(aggr_step=12, unit_time=6, start_step=12).do_manipulation(vals(steps)))
    iv = dict(aggr_type='average', step_type='instant')
    _check(g, 'avg_full', agg(**iv).do_manipulation(vals(steps, inst, start_is_end=True)))
    _check(g, 'avg_start24', agg(start_step=24, **iv).do_manipulation(vals(steps, inst, start_is_end=True)))
    _check(g, 'avg_masked', agg(**iv).do_manipulation(vals(steps, inst, masked=True, start_is_end=True)))
    # `@halfweights` (reference tests/test_aggregator.py:57-77): window incl. its first hour, half weights at both ends
    ih = dict(aggr_halfweights=True, **iv)
    _check(g, 'avgh_full', agg(**ih).do_manipulation(vals(steps, inst, start_is_end=True)))
    _check(g, 'avgh_start24', agg(start_step=24, **ih).do_manipulation(vals(steps, inst, start_is_end=True)))
    _check(g, 'avgh_gap', agg(**ih).do_manipulation(vals([s for s in steps if s not in (24, 54)], inst, start_is_end=True)))
    _check(g, 'avgh_masked', agg(**ih).do_manipulation(vals(steps, inst, masked=True, start_is_end=True)))
    _check(g, 'avgh_step12', agg(aggr_step=12, start_step=12, **ih).do_manipulation(vals(steps, inst, start_is_end=True)))
    # fewer than two steps inside a window: the reference divides 0 by 0
    lone = agg(aggr_step=3, start_step=3, end_step=6, **ih).do_manipulation(vals(steps, inst, start_is_end=True))
    assert all(np.isnan(np.asarray(v)).all() for v in lone.values()) and len(lone) == 2
    ii = dict(aggr_type='instantaneous', step_type='instant', aggr_step=12)
    _check(g, 'inst_full', agg(**ii).do_manipulation(vals(steps, inst, start_is_end=True)))
    _check(g, 'inst_gap', agg(**ii).do_manipulation(vals([s for s in steps if s not in (0, 36)], inst, start_is_end=True)))
    out = agg().do_manipulation(vals(steps))
    assert next(iter(out)) == M.Step(0, 24, RES, 24, 0)
    with pytest.raises(Exception):
        agg(aggr_type='average').do_manipulation(vals(steps))          # average of an accumulated parameter


def test_corrector_matches_reference(M, g):
    c = M.Corrector(g['corr_dem'], float(g['corr_dem_mv']), g['corr_gem'], float(g['corr_gem_mv']))
    np.testing.assert_array_equal(c.correct(g['corr_p'].copy()), g['corr_out'])
    c32 = M.Corrector(g['corr_dem'].astype(np.float32), float(g['corr_dem_mv']), g['corr_gem'], float(g['corr_gem_mv']))
    want = MR.correct(g['corr_p'].copy(), g['corr_dem'].astype(np.float32).astype(np.float64), g['corr_gem'],
                      float(g['corr_dem_mv']), float(g['corr_gem_mv']))
    np.testing.assert_array_equal(c32.correct(g['corr_p'].copy()), np.asarray(ma.filled(want)))
    # the README's pressure formula (README.md:968-978), against the vector the reference's own Corrector.correct made
    prs = 'p/gem*(10**((-0.159)*dem/1000))'
    cp = M.Corrector(g['corr_dem'], float(g['corr_dem_mv']), g['corr_gem'], float(g['corr_gem_mv']), formula=prs)
    np.testing.assert_array_equal(cp.correct(g['corr_p'].copy()), g['corr_out_pressure'])
    cp32 = M.Corrector(g['corr_dem'].astype(np.float32), float(g['corr_dem_mv']), g['corr_gem'], float(g['corr_gem_mv']),
                       formula=prs)
    want = MR.correct(g['corr_p'].copy(), g['corr_dem'].astype(np.float32), g['corr_gem'], float(g['corr_dem_mv']),
                      float(g['corr_gem_mv']), formula=prs)
    np.testing.assert_array_equal(cp32.correct(g['corr_p'].copy()), np.asarray(ma.filled(want)))
    with pytest.raises(Exception):
        M.Corrector(g['corr_dem'], 0.0, g['corr_gem'], 0.0, formula='p*gem+sqrt(dem)')


def test_pressure_correction_in_the_apply_epilogue(M):
    """`p/gem*(10**((-0.159)*dem/1000))` with gem = interpolate(`10**((-0.159)*(z/9.81)/1000)`): Corrector.from_geopotential
    + the windowed apply epilogue == the oracle's chain (gem formula on the source grid, apply, correct)."""
    from pyg2p_b200 import device
    lat, lon, det = synth.regular_ll_grid(62.0, 38.0, 97, -8.0, 32.0, 161)
    n = lat.size
    tl, tlo = synth.efas_target(rows=96, cols=70, cell=20000.0, x_ul=3500000.0, y_ul=4000000.0)
    ix = device.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    table = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    idx, w = table.indexes_coeffs()
    rng = np.random.default_rng(8)
    z = synth.field_geopotential(lat, lon)
    z[rng.random(n) < 0.01] = synth.GRIB_MV
    dem = synth.field_dem(tl, tlo, float(synth.PCR_MV_F32)).astype(np.float32)
    gem_f, prs = '(10**((-0.159)*(z/9.81)/1000))', 'p/gem*(10**((-0.159)*dem/1000))'
    corr = M.Corrector.from_geopotential(table, z, synth.GRIB_MV, synth.NC_FILL_F64, dem, float(synth.PCR_MV_F32),
                                         formula=prs, gem_formula=gem_f)
    fields = 1000.0 + 20 * rng.standard_normal((5, n))
    n_pad = -(-n // 16) * 16
    src = torch.zeros((5, n_pad), dtype=torch.float64, device='cuda')
    src[:, :n] = torch.from_numpy(fields).cuda()
    got = table.apply(src[:, :n], synth.NC_FILL_F64, correction=corr)
    assert table.last_path == 'window'
    with np.errstate(all='ignore'):
        gem_src = MR.gem_from_geopotential(z.copy(), synth.GRIB_MV, gem_f)
        gem_t = R.apply_sequential(gem_src, w, idx, 4, synth.NC_FILL_F64)
        for i in range(5):
            p = R.apply_sequential(fields[i], w, idx, 4, synth.NC_FILL_F64)
            want = MR.correct(p, dem.ravel(), gem_t, float(synth.PCR_MV_F32), synth.GRIB_MV, formula=prs)
            np.testing.assert_array_equal(got[i].cpu().numpy(), np.asarray(ma.filled(want)))


@pytest.mark.parametrize('kind', ['accumulation', 'average', 'instantaneous'])
def test_fused_pipeline_matches_oracle_chain(M, kind):
    """messages (steps 0..72 by 6 h) of one member -> convert -> aggregate -> cut-off -> invdist apply ->
    lapse-rate correction, two launches on the GPU, against the oracle's stage-by-stage chain on the same
    table; with and without masks."""
    from pyg2p_b200 import device
    lat, lon, det = synth.regular_ll_grid(62.0, 38.0, 97, -8.0, 32.0, 161)
    n = lat.size
    tl, tlo = synth.efas_target(rows=96, cols=70, cell=20000.0, x_ul=3500000.0, y_ul=4000000.0)
    ix = device.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    table = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    idx, w = table.indexes_coeffs()
    steps = list(range(0, 73, 6))
    rng = np.random.default_rng(3)
    if kind == 'accumulation':
        fields = np.cumsum(np.maximum(0, rng.standard_normal((len(steps), n))) * 2e-3 - 2e-4, axis=0)
        func, cut, agg = 'x=x*1000', True, dict(aggr_type='accumulation', step_type='accum', aggr_step=24, unit_time=24)
    else:
        fields = 280.0 + 5 * rng.standard_normal((len(steps), n))
        func, cut, agg = 'x=x-273.15', False, dict(aggr_type=kind, step_type='instant', aggr_step=24 if kind == 'average' else 12,
                                                   unit_time=24)
    mask = rng.random(n) < 0.03
    fields[:, mask] = synth.GRIB_MV
    conv = M.Converter(func=func, cut_off=cut)
    conv.set_missing_value(synth.GRIB_MV)
    a = M.Aggregator(input_step=6, start_step=0, end_step=72, mv_grib=synth.GRIB_MV, force_zero_array=False,
                     aggr_halfweights=False, **agg)
    plan = a.plan(steps)
    dem = synth.field_dem(tl, tlo, synth.PCR_MV_F32).astype(np.float32)
    z = synth.field_geopotential(lat, lon)
    z[::97] = synth.GRIB_MV
    mv_out = synth.NC_FILL_F64
    corr = M.Corrector.from_geopotential(table, z, synth.GRIB_MV, mv_out, dem, float(synth.PCR_MV_F32))
    pipe = M.FusedPipeline(table)
    assert pipe.compact.n_src == pipe.touched.numel() < n
    n_pad = -(-n // 16) * 16                      # rows padded to 16 values, as the windowed kernels need
    src = torch.zeros((len(steps), n_pad), dtype=torch.float64, device='cuda')[:, :n]
    src.copy_(torch.from_numpy(fields))
    F = kind != 'average'      # accumulation (2 inputs per output) and instantaneous (1) can run as ONE launch
    got = pipe.run(src, None, plan, converter=conv, cut_off=cut, mv_out=mv_out, corrector=corr, fused=F).cpu().numpy()
    assert pipe.last_path == ('fused' if F else 'two-kernel')
    two = pipe.run(src, None, plan, converter=conv, cut_off=cut, mv_out=mv_out, corrector=corr).cpu().numpy()
    assert pipe.last_path == 'two-kernel' and pipe.compact.last_path == 'window'
    np.testing.assert_array_equal(got, two)
    if not F:
        with pytest.raises(Exception):
            pipe.run(src, None, plan, converter=conv, cut_off=cut, mv_out=mv_out, corrector=corr, fused=True)
    # ---- oracle chain
    converted = {s: MR.convert(fields[i].copy(), func, synth.GRIB_MV) for i, s in enumerate(steps)}
    if kind == 'accumulation':
        ref = MR.accumulation(converted, 0, 72, 24, 24)
    elif kind == 'average':
        ref = MR.average(converted, 0, 72, 24)
    else:
        ref = MR.instantaneous(converted, 0, 72, 12, input_step=6)
    gem_src = MR.gem_from_geopotential(z, synth.GRIB_MV)
    with np.errstate(all='ignore'):
        gem = R.apply_sequential(gem_src, w, idx, 4, mv_out)
    keys = sorted(ref, key=lambda k: (k[1], k[0]))
    assert [k for k, _ in plan.outputs] == keys and got.shape == (len(keys), tl.size)
    for o, k in enumerate(keys):
        v = np.asarray(ma.getdata(ref[k]))
        if cut:
            v = MR.cut_off_negative(v)
        with np.errstate(all='ignore'):
            p = R.apply_sequential(v, w, idx, 4, mv_out)
            want = MR.correct(p, dem.ravel(), gem, float(synth.PCR_MV_F32), synth.GRIB_MV)
        np.testing.assert_array_equal(got[o], np.asarray(ma.filled(want)))
    # ---- the same without correction, masks propagated (fill): masked wherever a used input is masked
    msk = torch.zeros((len(steps), n_pad), dtype=torch.uint8, device='cuda')[:, :n]
    msk.copy_(torch.from_numpy(np.broadcast_to(mask, fields.shape).copy().view(np.uint8)))
    got2 = pipe.run(src, msk, plan, converter=conv, cut_off=cut, mv_out=mv_out, mask_policy=L.MASK_FILL, fused=F).cpu().numpy()
    two2 = pipe.run(src, msk, plan, converter=conv, cut_off=cut, mv_out=mv_out, mask_policy=L.MASK_FILL, fused=False)
    np.testing.assert_array_equal(got2, two2.cpu().numpy())
    p3, m3 = pipe.run(src, msk, plan, converter=conv, cut_off=cut, mv_out=mv_out, mask_policy=L.MASK_PROPAGATE, corrector=corr,
                      fused=F)
    q3, n3 = pipe.run(src, msk, plan, converter=conv, cut_off=cut, mv_out=mv_out, mask_policy=L.MASK_PROPAGATE, corrector=corr,
                      fused=False)
    np.testing.assert_array_equal(p3.cpu().numpy(), q3.cpu().numpy())
    np.testing.assert_array_equal(m3.cpu().numpy(), n3.cpu().numpy())
    masked_in = {s: ma.masked_where(mask, converted[s]) for s in steps}
    if kind == 'accumulation':
        refm = MR.accumulation(masked_in, 0, 72, 24, 24)
    elif kind == 'average':
        refm = MR.average(masked_in, 0, 72, 24)
    else:
        refm = MR.instantaneous(masked_in, 0, 72, 12, input_step=6)
    for o, k in enumerate(keys):
        v = np.asarray(ma.getdata(refm[k]))
        if cut:
            v = MR.cut_off_negative(v)
        mk = ma.getmaskarray(refm[k]) if isinstance(refm[k], ma.MaskedArray) else np.zeros(n, bool)
        if kind == 'accumulation' and k[0] == 0:
            mk = np.zeros(n, bool) | (ma.getmaskarray(refm[k]) if isinstance(refm[k], ma.MaskedArray) else False)
        with np.errstate(all='ignore'):
            data, _ = R.apply_propagate(v, w, idx, 4, mv_out, np.broadcast_to(mk, (n,)))
        np.testing.assert_array_equal(got2[o], data)


def test_pipeline_marked_and_host_buffers(M):
    """The marker form of the two-launch pipeline (byte masks in, marker NaNs between the kernels, bit-packed mask out)
    and the host-buffer form (`run_host`: the gather kernel reads the touched values out of pinned host memory, one member
    per pass, results back by D2H) give the bytes of the byte-mask pipeline that the oracle chain pins above."""
    from pyg2p_b200 import device
    lat, lon, det = synth.regular_ll_grid(62.0, 38.0, 97, -8.0, 32.0, 161)
    n = lat.size
    tl, tlo = synth.efas_target(rows=96, cols=70, cell=20000.0, x_ul=3500000.0, y_ul=4000000.0)
    m = tl.size
    ix = device.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    table = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    steps = list(range(0, 73, 6))
    members = 3
    rng = np.random.default_rng(11)
    fields = np.cumsum(np.maximum(0, rng.standard_normal((members, len(steps), n))) * 2e-3 - 2e-4, axis=1)
    masks = rng.random((members, len(steps), n)) < 0.01
    fields[masks] = synth.GRIB_MV
    fields = fields.reshape(members * len(steps), n)
    masks = masks.reshape(members * len(steps), n)
    conv = M.Converter(func='x=x*1000', cut_off=True)
    conv.set_missing_value(synth.GRIB_MV)
    a = M.Aggregator(input_step=6, start_step=0, end_step=72, mv_grib=synth.GRIB_MV, force_zero_array=False,
                     aggr_halfweights=False, aggr_type='accumulation', step_type='accum', aggr_step=24, unit_time=24)
    plan = a.plan(steps)
    dem = synth.field_dem(tl, tlo, synth.PCR_MV_F32).astype(np.float32)
    z = synth.field_geopotential(lat, lon)
    z[::97] = synth.GRIB_MV
    mv_out = synth.NC_FILL_F64
    corr = M.Corrector.from_geopotential(table, z, synth.GRIB_MV, mv_out, dem, float(synth.PCR_MV_F32))
    pipe = M.FusedPipeline(table)
    n_pad = -(-n // 16) * 16
    h_src = device.pinned_empty((fields.shape[0], n_pad))[:, :n]
    h_msk = device.pinned_empty((fields.shape[0], n_pad), torch.uint8)[:, :n]
    h_src.copy_(torch.from_numpy(fields))
    h_msk.copy_(torch.from_numpy(masks.view(np.uint8)))
    src, msk = h_src.cuda(), h_msk.cuda()
    src = torch.zeros((fields.shape[0], n_pad), dtype=torch.float64, device='cuda')[:, :n].copy_(src)
    msk = torch.zeros((fields.shape[0], n_pad), dtype=torch.uint8, device='cuda')[:, :n].copy_(msk)
    pipe.run(src, msk, plan, converter=conv, cut_off=True, mv_out=mv_out, mask_policy=L.MASK_PROPAGATE, corrector=corr,
             members=members)                         # first call: table compaction and window plan
    l0 = device.launch_count()
    want, want_mask = pipe.run(src, msk, plan, converter=conv, cut_off=True, mv_out=mv_out, mask_policy=L.MASK_PROPAGATE,
                               corrector=corr, members=members)
    assert device.launch_count() - l0 == 2            # one g2p_manip launch for all outputs + one apply
    want, want_mask = want.cpu().numpy(), want_mask.cpu().numpy().astype(bool)
    assert 0.005 < want_mask.mean() < 0.5
    got, bits = pipe.run(src, msk, plan, converter=conv, cut_off=True, mv_out=mv_out, mask_policy=L.MASK_MARKED,
                         corrector=corr, members=members)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(device.unpack_mask_bits(bits, m), want_mask)
    h_out, h_bits = pipe.run_host(h_src, h_msk, plan, converter=conv, cut_off=True, mv_out=mv_out, corrector=corr,
                                  members=members)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(h_out.numpy(), want)
    np.testing.assert_array_equal(device.unpack_mask_bits(h_bits, m), want_mask)
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes == want.size * 8 + h_bits.numel()
    # writer-ready variant: masked / NaN / clone-masked cells hold the file's missing value, no mask comes back
    clone = rng.random(tl.shape) < 0.2
    post = device.WriterPost(clone, 1e20, match=mv_out)
    h_out2, none = pipe.run_host(h_src, h_msk, plan, converter=conv, cut_off=True, mv_out=mv_out, corrector=corr,
                                 members=members, post=post)
    torch.cuda.synchronize()
    ready = want.copy()
    ready[want_mask | np.isnan(want) | (want == mv_out) | clone.ravel()[None, :]] = 1e20
    assert none is None
    np.testing.assert_array_equal(h_out2.numpy(), ready)
    # ... and as float32, what a PCRaster REAL4 band stores: the same fields cast element by element
    post32 = device.WriterPost(clone, 1e20, match=mv_out, out_dtype=np.float32)
    h_out32, _ = pipe.run_host(h_src, h_msk, plan, converter=conv, cut_off=True, mv_out=mv_out, corrector=corr,
                               members=members, post=post32)
    torch.cuda.synchronize()
    assert h_out32.dtype == torch.float32 and pipe.d2h_bytes == ready.size * 4
    np.testing.assert_array_equal(h_out32.numpy(), ready.astype(np.float32))
    # without masks
    plain = pipe.run(src, None, plan, converter=conv, cut_off=True, mv_out=mv_out, corrector=corr, members=members)
    h_out3, _ = pipe.run_host(h_src, None, plan, converter=conv, cut_off=True, mv_out=mv_out, corrector=corr, members=members)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(h_out3.numpy(), plain.cpu().numpy())


def test_division_by_message_constants_is_the_ieee_quotient(M):
    """the aggregation divisions (`/ aggr_step`, `/ count`) use a reciprocal + FMA candidate that is verified against its
    exact residual and falls back to the division routine: every quotient must be the correctly rounded IEEE one -
    zeros, signed zeros, subnormals, huge values, powers of two, NaN and infinities included (numpy as the oracle)."""
    rng = np.random.default_rng(123)
    n = 1 << 20
    bits = rng.integers(0, 1 << 63, size=n, dtype=np.int64) | (rng.integers(0, 2, size=n, dtype=np.int64) << 63)
    x = bits.view(np.float64).copy()                           # every exponent, both signs, NaNs and infinities
    x[:4096] = rng.standard_normal(4096) * 10.0 ** rng.integers(-12, 12, 4096)
    x[4096:8192] = np.ldexp(1.0, rng.integers(-1070, 1020, 4096)) * rng.choice([1.0, 3.0, 24.0, 6.0], 4096)
    x[8192:8200] = [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 1.7976931348623157e308]
    x[8200:12296] = np.arange(4096) * 24.0                     # exact quotients
    x[12296:16392] = np.nextafter(np.arange(1, 4097) * 24.0, np.inf)
    src = torch.from_numpy(x.reshape(1, n)).cuda()
    for d in (24.0, 6.0, 3.0, 7.0, 4.0, 0.1, 1e-3, 12.0, 1.0, -24.0, 1e120, 5e-324, np.inf):
        item = M.ManipItem(L.AGG_AVERAGE, [0], divisor=d, mask_all=-1)
        got, _ = M.manip_device(src, None, [item])
        with np.errstate(all='ignore'):
            want = (0.0 + x) / np.float64(d)
        np.testing.assert_array_equal(got.cpu().numpy()[0].view(np.int64), want.view(np.int64), err_msg=f'divisor {d}')
    # the accumulation form ((cur - prev) * unit) / step on a field with many dry cells
    cur = np.maximum(0.0, rng.standard_normal(n)) * 2e-3
    prev = np.where(rng.random(n) < 0.5, cur, cur * rng.random(n))
    both = torch.from_numpy(np.stack([cur, prev])).cuda()
    item = M.ManipItem(L.AGG_ACCUM, [0, 1], unit_time=24.0, divisor=24.0)
    got, _ = M.manip_device(both, None, [item])
    np.testing.assert_array_equal(got.cpu().numpy()[0], (((0.0 + cur) - prev) * 24.0) / 24.0)
