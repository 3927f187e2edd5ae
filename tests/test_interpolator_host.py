"""CPU-side tests of the host mirror: PCRaster map reader, intertable id / filename / lookup rules and the
error codes of the boundary (reference main/interpolation/__init__.py:68-101, :221-260; exceptions.py)."""
import gzip
import os

import numpy as np
import pytest

from pyg2p_b200.exceptions import NO_INTERTABLE_CREATED, ApplicationException
from pyg2p_b200.interpolator import Interpolator, _normalize_filename
from pyg2p_b200.maps import LatLong, PCRasterMap
from tests.helpers import MockedExecutionContext, write_csf_real4

REF_DATA = '/root/reference/tests/data'


def _maps(tmp_path, rows=12, cols=9):
    lat = np.linspace(60, 40, rows, dtype=np.float32)[:, None] * np.ones((1, cols), np.float32)
    lon = np.ones((rows, 1), np.float32) * np.linspace(5, 15, cols, dtype=np.float32)[None, :]
    lat[-2:] = np.nan
    lon[-2:] = np.nan
    write_csf_real4(tmp_path / 'lat.map', lat, 100000.0, 900000.0, 5000.0)
    write_csf_real4(tmp_path / 'lon.map', lon, 100000.0, 900000.0, 5000.0)
    return str(tmp_path / 'lat.map'), str(tmp_path / 'lon.map')


def test_pcraster_map_roundtrip(tmp_path):
    lat_map, lon_map = _maps(tmp_path)
    ll = LatLong(lat_map, lon_map)
    assert ll.lats.dtype == np.float32 and ll.lats.shape == (12, 9)
    assert ll.mv == -3.4028234663852886e38
    assert (ll.lats[-2:] == np.float32(ll.mv)).all() and (ll.lats[:-2] > 39).all()
    assert ll.identifier == '100000_900000_5000_-5000_43.64_60.00_100000_900000_5000_-5000_5.00_15.00'


@pytest.mark.skipif(not os.path.exists(REF_DATA), reason='reference fixtures only exist in the build container')
def test_pcraster_reader_against_reference_fixture():
    """The identifier built from the reference's lat.map / lon.map is the key of the golden table's entry in
    the reference's intertables.json (configuration/global/intertables.json:46)."""
    ll = LatLong(f'{REF_DATA}/lat.map', f'{REF_DATA}/lon.map')
    assert ll.identifier == '-1700000_2700000_5000_-5000_34.97_71.07_-1700000_2700000_5000_-5000_-24.35_51.35'
    assert ll.lats.shape == (810, 680) and (ll.lats == np.float32(ll.mv)).sum() == 291776
    dem = PCRasterMap(f'{REF_DATA}/dem.map')
    assert dem.values.shape == (810, 680) and dem.mv == -3.4028234663852886e38


def test_normalize_filename():
    assert _normalize_filename('pf10slhf_20151225.grib') == 'pf10slhf'
    assert _normalize_filename('input.grib') == 'input'
    assert _normalize_filename('EUE_2t-2016.grb') == 'eue2t2016'
    assert _normalize_filename('a.b') == 'a_b'


def _ctx(tmp_path, mode='nearest', create=False, entries=None):
    lat_map, lon_map = _maps(tmp_path)
    d = {'interpolation.dirs': {'user': str(tmp_path), 'global': str(tmp_path / 'global')},
         'interpolation.latMap': lat_map, 'interpolation.lonMap': lon_map, 'interpolation.mode': mode,
         'interpolation.create': create, 'input.file': '/data/pf10tp_2020010100.grib'}
    return MockedExecutionContext(d, False, entries)


def test_intertable_id_and_new_filename(tmp_path):
    ctx = _ctx(tmp_path, mode='invdist', create=True)
    it = Interpolator(ctx, 9999.0)
    tid, path = it._intertable_filename('-15.75$16.125$511$415$212065$rotated_ll')
    assert tid == ('I-15.75_16.125_511_415_212065_rotated_ll_'
                   '100000_900000_5000_-5000_43.64_60.00_100000_900000_5000_-5000_5.00_15.00scipy_invdist')
    assert path == str(tmp_path / 'tbl_pf10tp_108_scipy_invdist.npy.gz')
    open(path, 'wb').close()                       # an existing file gets a progressive number
    _, path2 = it._intertable_filename('-15.75$16.125$511$415$212065$rotated_ll')
    assert path2 == str(tmp_path / 'tbl_1_pf10tp_108_scipy_invdist.npy.gz')
    assert it.mv_output == -3.4028234663852886e38
    assert it.scipy_modes_nnear['invdist'] == 4 and it.suffixes['nearest'] == 'scipy_nearest'


def test_missing_table_without_create_raises_8100(tmp_path):
    it = Interpolator(_ctx(tmp_path, create=False), 9999.0)
    with pytest.raises(ApplicationException) as e:
        it._intertable_filename('x$y')
    assert e.value.code == NO_INTERTABLE_CREATED and 'Using I' in str(e.value)
    # configured but the file is nowhere: same code
    it2 = Interpolator(_ctx(tmp_path, entries={'Ix_y_' + it._target_coords.identifier + 'scipy_nearest':
                                                {'filename': 'tbl_gone.npy'}}), 9999.0)
    with pytest.raises(ApplicationException) as e:
        it2._intertable_filename('x$y')
    assert e.value.code == NO_INTERTABLE_CREATED and 'Tried to create' in str(e.value)


def test_lookup_order_and_gz_reader(tmp_path):
    os.makedirs(tmp_path / 'global')
    rec = np.rec.fromarrays((np.arange(108), np.linspace(0, 1, 108)), names=('indexes', 'coeffs'))
    with gzip.GzipFile(tmp_path / 'global' / 'tbl_x.npy.gz', 'w') as f:
        np.save(f, rec)
    it = Interpolator(_ctx(tmp_path), 9999.0)
    tid = 'Ig_' + it._target_coords.identifier + 'scipy_nearest'
    it.intertables_config.vars[tid] = {'filename': 'tbl_x.npy'}
    _, path = it._intertable_filename('g')
    assert path == str(tmp_path / 'global' / 'tbl_x.npy')          # not in the user dir -> global dir
    idx, coef = it._read_intertable(path)                           # plain name missing -> `.gz` is tried
    assert idx.strides == (16,) and coef.strides == (16,) and idx[5] == 5
    assert path in Interpolator._LOADED_INTERTABLES
    np.save(tmp_path / 'tbl_x.npy', rec)                            # the user dir wins once the file is there
    assert it._intertable_filename('g')[1] == str(tmp_path / 'tbl_x.npy')


def test_parallel_gzip_table_is_read_by_the_reference_loader(tmp_path):
    """multi-member gzip written by all cores == what `gzip.GzipFile` + `np.load` (the reference's reader) expect,
    byte for byte the `.npy` stream of np.save."""
    import io
    from pyg2p_b200.interpolator import save_npy_gz_parallel
    rng = np.random.default_rng(0)
    m = 300_000
    rec = np.rec.fromarrays((rng.integers(0, 6_599_681, (m, 4)), rng.random((m, 4))), names=('indexes', 'coeffs'))
    path = tmp_path / 'tbl.npy.gz'
    save_npy_gz_parallel(str(path), rec, level=1, chunk_bytes=1 << 20)
    with gzip.GzipFile(path, 'r') as f:
        back = np.load(f)
    assert back.dtype == rec.dtype and back.shape == rec.shape and back.tobytes() == rec.tobytes()
    ref = io.BytesIO()
    np.save(ref, rec)
    assert gzip.open(path).read() == ref.getvalue()
    save_npy_gz_parallel(str(tmp_path / 'empty.npy.gz'), rec[:0])
    with gzip.GzipFile(tmp_path / 'empty.npy.gz', 'r') as f:
        assert np.load(f).shape == (0, 4)


def test_parallel_gzip_loader(tmp_path):
    """`load_npy_gz_parallel` inflates the members of our own files side by side (their lengths travel in a gzip FEXTRA
    subfield) and falls back to the plain reader for any other gzip file - e.g. the reference's own tables."""
    import zlib
    from pyg2p_b200.interpolator import load_npy_gz_parallel, save_npy_gz_parallel
    rng = np.random.default_rng(1)
    for shape in ((200_000, 4), (70_001,), (0, 4), (3, 4)):
        rec = np.zeros(shape, dtype=[('indexes', '<i8'), ('coeffs', '<f8' if len(shape) == 2 else '<f4')])
        rec['indexes'] = rng.integers(0, 6_599_681, shape)
        rec['coeffs'] = rng.random(shape)
        path = str(tmp_path / 'tbl.npy.gz')
        save_npy_gz_parallel(path, rec, chunk_bytes=1 << 19)
        back = load_npy_gz_parallel(path)
        assert back.dtype == rec.dtype and back.shape == rec.shape and back.tobytes() == rec.tobytes()
        with gzip.GzipFile(path, 'r') as f:                     # the reference's reader sees the same array
            assert np.load(f).tobytes() == rec.tobytes()
        plain = str(tmp_path / 'plain.npy.gz')                  # single member, no subfield: what the reference writes
        with gzip.GzipFile(plain, 'w', compresslevel=1) as f:
            np.save(f, rec)
        assert load_npy_gz_parallel(plain).tobytes() == rec.tobytes()
    # the reference's golden table goes through the fallback
    golden = os.path.join(os.path.dirname(__file__), 'golden', 'ref_tbl_pf10slhf_550800_scipy_nearest.npy.gz')
    assert load_npy_gz_parallel(golden).shape == (550800,)
    # a damaged member is noticed (CRC of the inflated bytes)
    blob = bytearray(open(path, 'rb').read())
    blob[len(blob) // 2] ^= 0x55
    bad = str(tmp_path / 'bad.npy.gz')
    open(bad, 'wb').write(bytes(blob))
    with pytest.raises((OSError, zlib.error, ValueError)):
        load_npy_gz_parallel(bad)
