"""The reference's own test scenarios on the reference's own data, on the GPU.

Mirrors reference tests/test_interpolation.py (`test_interpolation_use_scipy_nearest`,
`test_interpolation_create_scipy_invdist`) and tests/test_correction.py (`test_correction`): the same inputs - the
first 2t message of tests/data/input.grib (COSMO rotated_ll 511x415), tests/data/geopotential.grib, the EFAS lat.map /
lon.map / dem.map and the packaged intertable tbl_pf10slhf_550800_scipy_nearest.npy.gz, all committed as fixtures
under tests/golden/ by make_golden.py -, the same calls, the same assertions; plus the numeric checks the reference's
tests leave out.  GRIB decoding is pyg2p_b200.grib1 here (ecCodes in the reference)."""
import os
import shutil

import numpy as np
import numpy.ma as ma
import pytest

from oracle import interp_ref as R
from pyg2p_b200 import grib1
from tests.helpers import MockedExecutionContext

pytestmark = pytest.mark.gpu
NC_FILL = 9.969209968386869e36          # netCDF4.default_fillvals['f8']
GOLDEN_TABLE = 'tbl_pf10slhf_550800_scipy_nearest.npy.gz'
PACKAGED_KEY = ('I-15.75_16.125_511_415_212065_rotated_ll_-1700000_2700000_5000_-5000_34.97_71.07_'
                '-1700000_2700000_5000_-5000_-24.35_51.35scipy_nearest')


@pytest.fixture(scope='module')
def data(golden_dir):
    with open(os.path.join(golden_dir, 'ref_input_2t_step0.grib'), 'rb') as f:
        msg = f.read()
    with open(os.path.join(golden_dir, 'ref_geopotential.grib'), 'rb') as f:
        geo = f.read()
    values, meta = grib1.read_values(msg)
    lats, lons, details = grib1.read_grid(msg)
    z, zmeta = grib1.read_values(geo)
    _, _, zdetails = grib1.read_grid(geo)
    assert (meta['indicatorOfParameter'], meta['level']) == (11, 2) and zmeta['indicatorOfParameter'] == 6   # 2t, z
    maps = np.load(os.path.join(golden_dir, 'ref_efas_latlon_maps.npz'))
    dem = np.load(os.path.join(golden_dir, 'ref_efas_dem_map.npz'))
    return dict(values=values, lats=lats, lons=lons, details=details, grid_id=details['grid_id'], z=z,
                z_grid_id=zdetails['grid_id'], missing=9999.0, lat_map=maps['lat'], lon_map=maps['lon'],
                map_id=str(maps['identifier']), map_mv=float(maps['mv']), dem=dem['dem'], dem_mv=float(dem['mv']),
                golden=golden_dir)


def _ctx(data, tmp_path, mode, create=False, with_golden_table=False):
    from pyg2p_b200.maps import LatLong
    target = LatLong(data['lat_map'], data['lon_map'], mv=data['map_mv'], identifier=data['map_id'])
    entries = {}
    if with_golden_table:
        shutil.copy(os.path.join(data['golden'], 'ref_' + GOLDEN_TABLE), tmp_path / GOLDEN_TABLE)
        # the entry the reference's packaged configuration/global/intertables.json:46-56 holds for this grid and these
        # maps, key and file name (without .gz: the loader tries both) copied literally
        entries[PACKAGED_KEY] = {'filename': 'tbl_pf10slhf_550800_scipy_nearest.npy', 'method': 'nearest',
                                 'source_shape': [212065], 'target_shape': [810, 680]}
    d = {'interpolation.dirs': {'user': str(tmp_path), 'global': str(tmp_path)}, 'interpolation.latMap': target,
         'interpolation.lonMap': None, 'interpolation.mode': mode, 'interpolation.create': create,
         'interpolation.parallel': True, 'input.file': 'tests/data/input.grib'}
    return MockedExecutionContext(d, False, entries)


@pytest.fixture()
def Interpolator():
    from pyg2p_b200 import device
    device.require_cuda()
    from pyg2p_b200.interpolator import Interpolator
    Interpolator._LOADED_INTERTABLES.clear()
    Interpolator._DEVICE_TABLES.clear()
    return Interpolator


def test_interpolation_use_scipy_nearest(Interpolator, data, tmp_path):
    """reference tests/test_interpolation.py:13-26: the packaged intertable is found and applied."""
    ctx = _ctx(data, tmp_path, 'nearest', with_golden_table=True)
    interpolator = Interpolator(ctx, data['missing'])
    values_resampled = interpolator.interpolate_scipy(data['lats'], data['lons'], data['values'], data['grid_id'],
                                                      data['details'])
    assert data['lat_map'].shape == values_resampled.shape                       # the reference's assertion
    # grid id from the GRIB message + identifier of the maps = the key of the reference's packaged configuration
    assert interpolator._intertable_filename(data['grid_id'])[0] == PACKAGED_KEY
    assert not list(tmp_path.glob('tbl*invdist*')) and ctx.configuration.intertables.dumps == 0   # nothing was built
    import gzip
    with gzip.GzipFile(tmp_path / GOLDEN_TABLE) as f:
        rec = np.load(f)
    want = R.apply_as_reference(data['values'], rec['coeffs'], rec['indexes'], 1, interpolator.mv_output)
    np.testing.assert_array_equal(values_resampled.ravel(), np.asarray(want))
    inside = rec['indexes'] < data['values'].size
    assert 0.3 < inside.mean() < 0.4 and 270.0 < values_resampled.ravel()[inside].mean() < 300.0   # 2 m temperatures


def test_interpolation_create_scipy_invdist(Interpolator, data, tmp_path):
    """reference tests/test_interpolation.py:29-46: `interpolation.create`, mode invdist - the table is built,
    written in the reference's layout and applied."""
    ctx = _ctx(data, tmp_path, 'invdist', create=True)
    interpolator = Interpolator(ctx, data['missing'])
    values_resampled = interpolator.interpolate_scipy(data['lats'], data['lons'], data['values'], data['grid_id'],
                                                      data['details'])
    assert data['lat_map'].shape == values_resampled.shape                       # the reference's assertion
    made = list(tmp_path.glob('tbl*_550800_scipy_invdist.npy.gz'))
    assert len(made) == 1 and ctx.configuration.intertables.dumps == 1
    import gzip
    with gzip.GzipFile(made[0]) as f:
        rec = np.load(f)
    assert rec.dtype == np.dtype([('indexes', '<i8'), ('coeffs', '<f8')]) and rec.shape == (550800, 4)
    with np.errstate(all='ignore'):
        want = R.apply_as_reference(data['values'], rec['coeffs'], rec['indexes'], 4, interpolator.mv_output)
    np.testing.assert_array_equal(values_resampled.ravel(), np.asarray(want))
    # rows the nearest table of the reference marks as outside are outside here too (same min_upper_bound rule)
    with gzip.GzipFile(os.path.join(data['golden'], 'ref_' + GOLDEN_TABLE)) as f:
        near = np.load(f)
    outside_ref = near['indexes'] == 212065
    outside = rec['indexes'][:, 0] == 212065
    assert (outside != outside_ref).sum() <= 10 and np.isnan(values_resampled.ravel()[outside]).all()


def test_correction(Interpolator, data, tmp_path):
    """reference tests/test_correction.py:15-58, same formula strings, same numeric restatement, same assertion."""
    from pyg2p_b200.manipulation import Corrector
    ctx = _ctx(data, tmp_path, 'nearest', with_golden_table=True)
    interpolator = Interpolator(ctx, data['missing'])
    lats, lons = data['lats'], data['lons']
    values_resampled = interpolator.interpolate_scipy(lats, lons, data['values'], data['grid_id'], data['details'])
    table = interpolator._table_for(lats, lons, data['values'], data['grid_id'], data['details'])
    corrector = Corrector.from_geopotential(table, data['z'], data['missing'], interpolator.mv_output, data['dem'],
                                            data['dem_mv'], formula='p+gem-dem*0.0065', gem_formula='(z/9.81)*0.0065')
    values_out = corrector.correct(values_resampled)

    z, mv_geopotential, dem, dem_mv = data['z'], data['missing'], data['dem'], data['dem_mv']
    gem = np.where(z != mv_geopotential, (z / 9.81) * 0.0065, mv_geopotential)
    gem_resampled = interpolator.interpolate_scipy(lats, lons, gem, data['z_grid_id'], data['details'])
    mv = NC_FILL
    reference = np.where((dem != dem_mv) & (values_resampled != mv) & (gem_resampled != mv_geopotential),
                         values_resampled + gem_resampled - dem * 0.0065,
                         mv)
    assert np.allclose(values_out, reference)                                     # the reference's assertion
    # and bit for bit where numexpr's typing is followed (float32 dem promoted to double before the product)
    exact = np.where((dem != dem_mv) & (values_resampled != mv) & (gem_resampled != mv_geopotential),
                     (values_resampled + gem_resampled) - dem.astype(np.float64) * 0.0065, mv)
    np.testing.assert_array_equal(np.asarray(ma.getdata(values_out)), exact)
    corrected = (dem != dem_mv) & (values_resampled != interpolator.mv_output)
    assert corrected.mean() > 0.2 and 230.0 < np.asarray(values_out)[corrected].mean() < 300.0


# ----------------------------------------------------------------------------- reference tests/test_aggregator.py
def _aggregator_case(data, **kw):
    """`messages.first_resolution_values()` of the five 2t messages of tests/data/input.grib (steps 0..24 by 6 h,
    instantaneous, Nj = 415, level 2).  The fixture holds the first message; the later steps are derived from it
    (the reference's assertions are on the keys)."""
    from pyg2p_b200 import manipulation as M
    values_orig = {M.Step(s, s, 415, 6, 2): data['values'] + 0.25 * s for s in (0, 6, 12, 18, 24)}
    base = dict(aggr_halfweights=False, input_step=6, step_type='instant', start_step=0, mv_grib=data['missing'], end_step=24,
                unit_time=24, force_zero_array=False)
    base.update(kw)
    return M, values_orig, M.Aggregator(**base).do_manipulation(values_orig)


def test_aggregator_instant(data):
    M, values_orig, values = _aggregator_case(data, aggr_step=6, aggr_type='instantaneous')
    assert len(values_orig) == len(values)                                      # the reference's assertions (:29-33)
    keys_orig, keys_res = list(values_orig.keys()), list(values.keys())
    assert keys_orig[0] == keys_res[0] and keys_orig[-1] == keys_res[-1]
    for k in keys_orig:
        np.testing.assert_array_equal(values[k], 0.0 + values_orig[k])


def test_aggregator_average(data):
    M, values_orig, values = _aggregator_case(data, aggr_step=24, aggr_type='average')
    assert len(values) == 1 and list(values.keys())[0] == M.Step(0, 24, 415, 24, 2)   # (:52-54)
    # hours 1..24, each taking the next available step: 6 x (6, 12, 18, 24)
    acc = np.zeros_like(data['values'])
    for h in range(1, 25):
        acc = acc + values_orig[M.Step(-(-h // 6) * 6, -(-h // 6) * 6, 415, 6, 2)]
    np.testing.assert_array_equal(np.asarray(list(values.values())[0]), acc / 24.0)


def test_aggregator_average_halfweights(data):
    M, values_orig, values = _aggregator_case(data, aggr_step=24, aggr_type='average', aggr_halfweights=True, start_step=24)
    assert len(values) == 1 and list(values.keys())[0] == M.Step(0, 24, 415, 24, 2)   # (:75-77)
    acc = np.zeros_like(data['values'])
    for s, w in ((0, 3.0), (6, 6.0), (12, 6.0), (18, 6.0), (24, 3.0)):
        acc = acc + values_orig[M.Step(s, s, 415, 6, 2)] * w
    np.testing.assert_array_equal(np.asarray(list(values.values())[0]), acc / 24.0)


def test_aggregator_accumulation(data):
    M, values_orig, values = _aggregator_case(data, aggr_step=6, aggr_type='accumulation')
    keys_res = list(values.keys())
    assert len(values) == 4 and keys_res[0] == M.Step(0, 6, 415, 6, 2) and keys_res[-1] == M.Step(18, 24, 415, 6, 2)   # (:98-101)
    for k in keys_res:
        cur, prev = values_orig[M.Step(k.end_step, k.end_step, 415, 6, 2)], values_orig[M.Step(k.start_step, k.start_step, 415, 6, 2)]
        np.testing.assert_array_equal(np.asarray(values[k]), ((0.0 + cur) - prev) * 24 / 6)


# ----------------------------------------------------------------------------- reference tests/test_conversion.py
class TestConversion:
    data = np.random.default_rng(0).uniform(low=-100.0, high=100.0, size=(100, 100))

    def test_identity(self):
        from pyg2p_b200.manipulation import Converter
        converter = Converter('x=x')
        data = self.data.copy()
        res = converter.convert(data)
        assert np.allclose(data, res)

    def test_conversion(self):
        from pyg2p_b200.manipulation import Converter
        converter = Converter('x=x-273.15')
        data = self.data.copy()
        res = converter.convert(data)
        assert np.allclose(data - 273.15, res)
        np.testing.assert_array_equal(res, data - 273.15)

    def test_cutoff(self):
        from pyg2p_b200.manipulation import Converter
        assert self.data[self.data < 0].size > 0
        converter = Converter('x=x', cut_off=True)
        data = self.data.copy()
        res = converter.convert(data)
        res = converter.cut_off_negative(res)
        assert res[res < 0].size == 0
        assert np.allclose(res[res > 0], data[data > 0])

    def test_messages_apply_conversion(self):
        """the Controller-side seam (controller.py:106-110 -> Messages.apply_conversion, src/pyg2p/__init__.py:199-207): all
        messages of both resolutions converted in one launch each, `where(x != mv, f(x), mv)` with the masks preserved."""
        import numpy.ma as ma

        from oracle import manip_ref as MR
        from pyg2p_b200 import synth
        from pyg2p_b200.manipulation import Converter, Messages, Step
        rng = np.random.default_rng(4)
        mv = 9999.0
        first = {Step(s, s, 415, 6, 2): np.where(rng.random((100, 100)) < 0.05, mv, self.data + s) for s in (0, 6, 12)}
        mask = rng.random((100, 100)) < 0.1
        second = {Step(s, s, 200, 12, 2): ma.masked_where(mask, self.data * 2 + s) for s in (24, 36)}
        msgs = Messages(dict(first), mv, 'K', 'surface', 'instant', 'h', synth.GridDetails(gridType='regular_ll'),
                        val_2nd=dict(second), data_date='20240101', data_time='00')
        assert len(msgs) == 5 and msgs.first_step_range == Step(0, 0, 415, 6, 2)
        conv = Converter('x=x-273.15')
        msgs.apply_conversion(conv)
        assert conv._mv == mv and conv._initial_unit == 'K' and 'x=x-273.15' in str(conv)
        for k, v in first.items():
            np.testing.assert_array_equal(msgs.first_resolution_values()[k], MR.convert(v.copy(), 'x=x-273.15', mv))
        for k, v in second.items():
            got = msgs.second_resolution_values()[k]
            assert isinstance(got, ma.MaskedArray)
            np.testing.assert_array_equal(ma.getmaskarray(got), mask)
            np.testing.assert_array_equal(ma.getdata(got), MR.convert(np.asarray(ma.getdata(v)).copy(), 'x=x-273.15', mv))
