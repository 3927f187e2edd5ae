"""GPU parity of the batched intertable apply (K7) and the table record pack/unpack (K6),
bit-exact against the oracle and the reference-generated golden vectors, through the C ABI."""
import gzip
import os

import numpy as np
import pytest
import torch

from oracle import interp_ref as R
from pyg2p_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from pyg2p_b200 import device
    device.require_cuda()
    return device


def _eq(a, b):
    """bit equality that treats NaN == NaN."""
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize('mode,k', [('nearest', 1), ('invdist', 4)])
def test_apply_matches_reference_golden(dev, golden_dir, mode, k):
    g = np.load(os.path.join(golden_dir, 'apply_regional.npz'))
    t = np.load(os.path.join(golden_dir, 'build_regional_global_f64.npz'))
    rec = R.to_record(t[f'{mode}_indexes'], t[f'{mode}_weights'])
    n = g['v0'].size
    tbl = dev.DeviceTable.from_record(rec, n)
    assert (tbl.k, tbl.m) == (k, 5241)
    src = torch.from_numpy(np.stack([g[f'v{b}'] for b in range(3)])).cuda()
    out = tbl.apply(src, float(g['mv_out'])).cpu().numpy()
    for b in range(3):
        _eq(out[b], g[f'{mode}_plain{b}'])            # what the reference's own code produced
        _eq(out[b], g[f'{mode}_masked{b}'])           # AS_REFERENCE: masks are dropped (App. C.1)
    # PROPAGATE: intended semantics
    mask = torch.from_numpy(np.stack([g[f'mask{b}'] for b in range(3)])).cuda()
    o2, m2 = tbl.apply(src, float(g['mv_out']), src_mask=mask, mask_policy=L.MASK_PROPAGATE)
    for b in range(3):
        data, mk = R.apply_propagate(g[f'v{b}'], rec['coeffs'], rec['indexes'], k, float(g['mv_out']), g[f'mask{b}'])
        _eq(o2[b].cpu().numpy(), data)
        _eq(m2[b].cpu().numpy().astype(bool), mk)


def test_record_roundtrip_and_reference_golden_table(dev, golden_dir):
    with gzip.GzipFile(os.path.join(golden_dir, 'ref_tbl_pf10slhf_550800_scipy_nearest.npy.gz')) as f:
        rec = np.load(f)
    n = 511 * 415
    tbl = dev.DeviceTable.from_record(rec, n)
    back = tbl.to_record()
    assert back.dtype == rec.dtype and back.shape == rec.shape
    assert back.tobytes() == rec.tobytes()
    mv = -3.4028234663852886e38
    rng = np.random.default_rng(3)
    v = rng.standard_normal((5, n))
    out = tbl.apply(torch.from_numpy(v).cuda(), mv).cpu().numpy()
    for b in range(5):
        _eq(out[b], R.apply_as_reference(v[b], rec['coeffs'], rec['indexes'], 1, mv))
    assert (out[0] == mv).sum() == 362052


@pytest.mark.parametrize('k', [1, 2, 3, 4, 11])
@pytest.mark.parametrize('m,b', [(1, 1), (2, 3), (1001, 7), (65537, 9), (300000, 2)])
def test_apply_random_tables(dev, k, m, b, monkeypatch):
    """ragged sizes (odd m -> unaligned rows), sentinels, NaN weights, all k the reference uses."""
    rng = np.random.default_rng(100 * k + m % 97 + b)
    n = 5000
    idx = rng.integers(0, n + 1, size=(m, k))                 # n = sentinel
    w = rng.random((m, k))
    w /= w.sum(axis=1, keepdims=True)
    w[rng.random(m) < 0.05] = np.nan
    if k == 1:
        idx, w = idx[:, 0], w[:, 0]
    rec = R.to_record(idx, w)
    tbl = dev.DeviceTable.from_record(rec, n)
    assert tbl.to_record().tobytes() == np.ascontiguousarray(rec).tobytes()
    v = rng.standard_normal((b, n)) * 100
    mask = rng.random((b, n)) < 0.1
    mv = 9.969209968386869e36
    out = tbl.apply(torch.from_numpy(v).cuda(), mv).cpu().numpy()
    o2, m2 = tbl.apply(torch.from_numpy(v).cuda(), mv, src_mask=torch.from_numpy(mask).cuda(),
                       mask_policy=L.MASK_PROPAGATE)
    for i in range(b):
        with np.errstate(all='ignore'):
            _eq(out[i], R.apply_sequential(v[i], rec['coeffs'], rec['indexes'], k, mv))
            data, mk = R.apply_propagate(v[i], rec['coeffs'], rec['indexes'], k, mv, mask[i])
        _eq(o2[i].cpu().numpy(), data)
        _eq(m2[i].cpu().numpy().astype(bool), mk)
    monkeypatch.setenv('G2P_APPLY_PATH', 'generic')
    _check_bits_and_marked(dev, tbl, v, mask, mv, rec, k, 'generic')


def _check_bits_and_marked(dev, tbl, v, mask, mv, rec, k, path):
    """PROPAGATE_BITS (bit-packed masks in and out) and MARKED (masked values carry the marker NaN) against the
    oracle's PROPAGATE result."""
    b, n = v.shape
    src = _padded(v)
    n_pad = src.stride(0)
    bits = dev.pack_mask_bits(mask, n_pad)
    assert bits.tobytes() == dev.pack_mask_bits(torch.from_numpy(mask).cuda(), n_pad).cpu().numpy().tobytes()
    o4, m4 = tbl.apply(src, mv, src_mask=torch.from_numpy(bits).cuda(), mask_policy=L.MASK_PROPAGATE_BITS)
    assert tbl.last_path == path, tbl.plan_error
    marked = src.clone() if src.is_contiguous() else _padded(v)
    dev.mark_masked(marked, _padded(mask.view(np.uint8)))
    marked2 = _padded(v)
    dev.mark_masked(marked2, torch.from_numpy(bits).cuda(), bits=True)
    assert torch.equal(marked.view(torch.int64), marked2.view(torch.int64))
    assert int((marked.view(torch.int64) == L.MASK_MARKER).sum()) == int(mask.sum())
    o5, m5 = tbl.apply(marked, mv, mask_policy=L.MASK_MARKED)
    o6 = tbl.apply(marked, mv, mask_policy=L.MASK_MARKED, out_mask=False)
    m4, m5 = dev.unpack_mask_bits(m4, tbl.m), dev.unpack_mask_bits(m5, tbl.m)
    for i in range(b):
        with np.errstate(all='ignore'):
            data, mk = R.apply_propagate(v[i], rec['coeffs'], rec['indexes'], k, mv, mask[i])
        for o, mm in ((o4, m4), (o5, m5), (o6, None)):
            _eq(o[i].cpu().numpy(), data)
            if mm is not None:
                _eq(mm[i], mk)


def test_apply_f32_coeff_table(dev):
    rng = np.random.default_rng(8)
    m, n = 777, 300
    idx = rng.integers(0, n + 1, size=m)
    d = (rng.random(m) * 1e5).astype(np.float32)
    rec = R.to_record(idx, d)
    assert rec.dtype.itemsize == 12
    tbl = dev.DeviceTable.from_record(rec, n)
    assert tbl.coef_f32 and tbl.to_record().tobytes() == np.ascontiguousarray(rec).tobytes()


@pytest.mark.parametrize('k', [2, 3, 5, 8, 11, 16])
def test_apply_f32_coeff_any_k_matches_einsum_order(dev, k):
    """float32-coefficient tables of any k (mode adw on PCRaster REAL4 maps: k = 11): numpy's einsum buffers the operands
    and sums with two SSE2 lanes (even / odd products, eight per trip in the order 6,7 | 4,5 | 2,3 | 0,1, then pairs, lane 0
    + lane 1); the float64 record views of a loaded table are summed left to right.  The generic kernel follows the
    table's dtype: bit for bit `_interpolate_scipy_append_mv`'s own einsum on the record's field views."""
    rng = np.random.default_rng(31 + k)
    m, n, b = 4099, 2000, 3
    idx = rng.integers(0, n + 1, size=(m, k))                  # n = the appended slot
    w = rng.random((m, k))
    w /= w.sum(axis=1, keepdims=True)
    v = 280.0 + 20 * rng.standard_normal((b, n))
    for dt in (np.float32, np.float64):
        rec = R.to_record(idx, w.astype(dt))
        tbl = dev.DeviceTable.from_record(rec, n)
        assert tbl.sum_pairwise == (dt == np.float32)
        out = tbl.apply(torch.from_numpy(v).cuda(), 1e20).cpu().numpy()
        for i in range(b):
            _eq(out[i], R.apply_as_reference(v[i], rec['coeffs'], rec['indexes'], k, 1e20))


@pytest.mark.parametrize('path', ['generic', 'window'])
def test_apply_f32_coeff_k4_matches_einsum_order(dev, path, monkeypatch):
    """float32 coefficients (tables built for PCRaster REAL4 maps): numpy's einsum buffers the operands and adds
    (w0*v0 + w2*v2) + (w1*v1 + w3*v3); float64 record views are summed left to right.  Both kernels follow the table's
    dtype and match `_interpolate_scipy_append_mv` (np.einsum on the record's field views) bit for bit."""
    monkeypatch.setenv('G2P_APPLY_PATH', path)
    rng = np.random.default_rng(21)
    rows, cols, n, k, b = 40, 130, 3000, 4, 5
    m = rows * cols
    centre = rng.integers(0, n, size=m)
    idx = np.clip(centre[:, None] + rng.integers(-30, 31, size=(m, k)), 0, n)
    w = rng.random((m, k))
    w /= w.sum(axis=1, keepdims=True)
    v = 280.0 + 20 * rng.standard_normal((b, n))
    mask = rng.random((b, n)) < 0.05
    for dt in (np.float32, np.float64):
        rec = R.to_record(idx, w.astype(dt))
        tbl = dev.DeviceTable.from_record(rec, n, grid_shape=(rows, cols))
        assert tbl.sum_pairwise == (dt == np.float32)
        out = tbl.apply(_padded(v), 1e20).cpu().numpy()
        assert tbl.last_path == path
        o2, m2 = tbl.apply(_padded(v), 1e20, src_mask=_padded(mask.view(np.uint8)), mask_policy=L.MASK_PROPAGATE)
        for i in range(b):
            _eq(out[i], R.apply_as_reference(v[i], rec['coeffs'], rec['indexes'], k, 1e20))      # the reference's own einsum
            data, mk = R.apply_propagate(v[i], rec['coeffs'], rec['indexes'], k, 1e20, mask[i])
            _eq(o2[i].cpu().numpy(), data)
            _eq(m2[i].cpu().numpy().astype(bool), mk)


def test_apply_strided_batch_and_empty(dev):
    rng = np.random.default_rng(9)
    m, n, k = 1000, 400, 4
    idx = rng.integers(0, n, size=(m, k))
    w = rng.random((m, k))
    rec = R.to_record(idx, w)
    tbl = dev.DeviceTable.from_record(rec, n)
    big = torch.from_numpy(rng.standard_normal((6, n + 40))).cuda()
    view = big[:, :n]                                         # src_stride > n
    out = tbl.apply(view, 0.0).cpu().numpy()
    for i in range(6):
        _eq(out[i], R.apply_sequential(big[i, :n].cpu().numpy(), w, idx, k, 0.0))
    empty = tbl.apply(torch.empty((0, n), dtype=torch.float64, device='cuda'), 0.0)
    assert empty.shape == (0, m)
    with pytest.raises(ValueError):
        tbl.apply(torch.zeros((1, n + 1), dtype=torch.float64, device='cuda'), 0.0)


def test_linearity_at_full_size(dev):
    """Size-independent property at BASELINE scale (O1280 -> EFAS, 16 messages): the apply is linear
    in the source field for finite weights, and a constant field maps to that constant."""
    m, n, k, b = 950000, 6599680, 4, 16
    g = torch.Generator(device='cuda').manual_seed(1)
    base = torch.randint(0, n - 4000, (1, m), generator=g, device='cuda', dtype=torch.int64)
    idx = (base + torch.randint(0, 4000, (k, m), generator=g, device='cuda')).to(torch.int32)
    w = torch.rand((k, m), generator=g, device='cuda', dtype=torch.float64)
    w = w / w.sum(0, keepdim=True)
    tbl = dev.DeviceTable(idx.contiguous(), w.contiguous(), n)
    src = torch.randn((b, n), generator=g, device='cuda', dtype=torch.float64)
    out = tbl.apply(src, 0.0)
    out2 = tbl.apply(src * 2.0, 0.0)
    assert torch.equal(out2, out * 2.0)                       # exact: scaling by 2 commutes with rounding
    ones = tbl.apply(torch.full((1, n), 3.0, device='cuda', dtype=torch.float64), 0.0)
    assert torch.allclose(ones, torch.full_like(ones, 3.0), rtol=1e-15, atol=0)
    # spot check rows against the oracle order
    rows = torch.randint(0, m, (2000,), generator=g, device='cuda')
    ii = idx[:, rows].long()
    ww = w[:, rows]
    acc = ww[0] * src[3][ii[0]]
    for j in range(1, k):
        acc = acc + ww[j] * src[3][ii[j]]
    assert torch.equal(out[3][rows], acc)


def _padded(v, align=16):
    """[b, n] ndarray -> device view [b, n] of a [b, ceil16(n)] buffer (what the windowed kernel needs)."""
    b, n = v.shape
    n_pad = -(-n // align) * align
    buf = torch.zeros((b, n_pad), dtype=torch.from_numpy(v[:0]).dtype, device='cuda')
    buf[:, :n] = torch.from_numpy(v).cuda()
    return buf[:, :n]


@pytest.mark.parametrize('mode,k', [('nearest', 1), ('invdist', 4)])
def test_window_path_matches_reference_golden(dev, golden_dir, mode, k, monkeypatch):
    monkeypatch.setenv('G2P_APPLY_PATH', 'window')          # k=1 tables default to the generic kernel
    g = np.load(os.path.join(golden_dir, 'apply_regional.npz'))
    t = np.load(os.path.join(golden_dir, 'build_regional_global_f64.npz'))
    rec = R.to_record(t[f'{mode}_indexes'], t[f'{mode}_weights'])
    n = g['v0'].size
    tbl = dev.DeviceTable.from_record(rec, n)
    assert tbl.plan_info()['max_window'] >= 16
    src = _padded(np.stack([g[f'v{b}'] for b in range(3)]))
    mask = _padded(np.stack([g[f'mask{b}'] for b in range(3)]).view(np.uint8))
    out = tbl.apply(src, float(g['mv_out'])).cpu().numpy()
    assert tbl.last_path == 'window'
    o2, m2 = tbl.apply(src, float(g['mv_out']), src_mask=mask, mask_policy=L.MASK_PROPAGATE)
    assert tbl.last_path == 'window'
    for b in range(3):
        _eq(out[b], g[f'{mode}_plain{b}'])
        data, mk = R.apply_propagate(g[f'v{b}'], rec['coeffs'], rec['indexes'], k, float(g['mv_out']), g[f'mask{b}'])
        _eq(o2[b].cpu().numpy(), data)
        _eq(m2[b].cpu().numpy().astype(bool), mk)


@pytest.mark.parametrize('k', [1, 4])
@pytest.mark.parametrize('shape,b', [((1, 1), 1), ((1, 2), 3), ((7, 143), 5), ((33, 257), 9), ((300, 1000), 2),
                                     ((1, 65537), 4), ((5, 4), 70)])
def test_window_path_random_tables(dev, k, shape, b, monkeypatch):
    """the windowed kernel against the oracle on ragged grids (odd columns, partial tiles), sentinels and
    NaN weights; batch sizes that exercise the stage ring (b > stages) and several message chunks."""
    monkeypatch.setenv('G2P_APPLY_PATH', 'window')
    rows, cols = shape
    m = rows * cols
    rng = np.random.default_rng(7 * k + m % 1013 + b)
    n = 4001                                                   # not a multiple of 16: rows are padded
    centre = rng.integers(0, n, size=m)
    idx = np.clip(centre[:, None] + rng.integers(-40, 41, size=(m, k)), 0, n)   # n = sentinel
    idx[rng.random(m) < 0.03] = n
    w = rng.random((m, k))
    w /= w.sum(axis=1, keepdims=True)
    w[rng.random(m) < 0.05] = np.nan
    if k == 1:
        idx, w = idx[:, 0], w[:, 0]
    rec = R.to_record(idx, w)
    v = rng.standard_normal((b, n)) * 100
    mask = rng.random((b, n)) < 0.1
    mv = 9.969209968386869e36
    for grid_shape in (None, shape):
        tbl = dev.DeviceTable.from_record(rec, n, grid_shape=grid_shape)
        src, msk = _padded(v), _padded(mask.view(np.uint8))
        out = tbl.apply(src, mv).cpu().numpy()
        assert tbl.last_path == 'window', tbl.plan_error
        o2, m2 = tbl.apply(src, mv, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
        o3 = tbl.apply(src, mv, src_mask=msk, mask_policy=L.MASK_FILL)
        assert tbl.last_path == 'window'
        for i in range(b):
            with np.errstate(all='ignore'):
                _eq(out[i], R.apply_sequential(v[i], rec['coeffs'], rec['indexes'], k, mv))
                data, mk = R.apply_propagate(v[i], rec['coeffs'], rec['indexes'], k, mv, mask[i])
            _eq(o2[i].cpu().numpy(), data)
            _eq(m2[i].cpu().numpy().astype(bool), mk)
            _eq(o3[i].cpu().numpy(), data)
        _check_bits_and_marked(dev, tbl, v, mask, mv, rec, k, 'window')
        # the same launches through the kernel that reads its switches (summation order, ballot / atomicOr mask
        # rows) at run time instead of the compile-time FAST variants
        monkeypatch.setenv('G2P_WIN_NO_FAST', '1')
        _check_bits_and_marked(dev, tbl, v, mask, mv, rec, k, 'window')
        monkeypatch.delenv('G2P_WIN_NO_FAST')
        # ... and the FAST variants on the __syncthreads loop instead of the lagged mbarrier rings
        monkeypatch.setenv('G2P_WIN_NO_LAG', '1')
        _check_bits_and_marked(dev, tbl, v, mask, mv, rec, k, 'window')
        monkeypatch.delenv('G2P_WIN_NO_LAG')


def test_scattered_table_falls_back_to_generic_gpu_kernel(dev):
    rng = np.random.default_rng(4)
    n, m, k = 4_000_000, 5000, 4
    idx = rng.integers(0, n, size=(m, k))
    w = rng.random((m, k))
    tbl = dev.DeviceTable.from_record(R.to_record(idx, w), n)
    v = rng.standard_normal((2, n))
    out = tbl.apply(torch.from_numpy(v).cuda(), 0.0).cpu().numpy()
    assert tbl.last_path == 'generic' and 'not windowable' in tbl.plan_error
    for i in range(2):
        _eq(out[i], R.apply_sequential(v[i], w, idx, k, 0.0))


def test_window_equals_generic_on_real_geometry(dev, monkeypatch):
    """O320 -> EFAS invdist table built on the GPU: both apply kernels give identical bytes."""
    from pyg2p_b200 import synth
    lat, lon, det = synth.octahedral_grid(320)
    tl, tlo = synth.efas_target()
    ix = dev.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
    tbl = dev.DeviceTable(out['idx'], out['coef'], lat.size, grid_shape=tl.shape)
    info = tbl.plan_info()
    assert info['tile_cols'] * info['tile_rows'] == 1024 and info['max_window'] < 4096
    g = torch.Generator(device='cuda').manual_seed(2)
    src = torch.randn((37, lat.size), generator=g, device='cuda', dtype=torch.float64)
    msk = (torch.rand((37, lat.size), generator=g, device='cuda') < 0.05).to(torch.uint8)
    a, am = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
    af = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_FILL)
    ar = tbl.apply(src, 1e20)
    assert tbl.last_path == 'window'
    monkeypatch.setenv('G2P_APPLY_PATH', 'generic')
    b_, bm = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
    bf = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_FILL)
    br = tbl.apply(src, 1e20)
    assert tbl.last_path == 'generic'
    assert torch.equal(a.view(torch.int64), b_.view(torch.int64)) and torch.equal(am, bm)
    assert torch.equal(af.view(torch.int64), bf.view(torch.int64)) and torch.equal(af, a)
    assert torch.equal(ar.view(torch.int64), br.view(torch.int64))
    assert 0.05 < am.float().mean() < 0.5 and torch.equal(a == 1e20, am.bool())


@pytest.mark.parametrize('res,k', [(0.1, 1), (0.05, 4)])
def test_full_size_glofas_tables_window_equals_generic(dev, monkeypatch, res, k):
    """BASELINE configs[1] / [2] tables at full size (O1280 -> 0.1 deg k=1, 5.4 M targets; 0.05 deg k=4, 21.6 M):
    the windowed kernel (plan on a global lat/lon grid) and the generic kernel give identical bytes, with masks,
    for a single message and for a short batch; a constant field maps to that constant."""
    from pyg2p_b200 import synth
    lat, lon, det = synth.octahedral_grid(1280)
    tl, tlo = synth.latlon_target(res)
    ix = dev.SourceIndex(lon, lat, det['radius'])
    out = ix.knn(tlo, tl, k=k, mode=L.MODE_NEAREST if k == 1 else L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']),
                 want_flags=False)
    tbl = dev.DeviceTable(out['idx'], out['coef'], lat.size, grid_shape=tl.shape)
    info = tbl.plan_info()
    assert info is not None and info['tile_rows'] * info['tile_cols'] == 1024, tbl.plan_error
    g = torch.Generator(device='cuda').manual_seed(5)
    for b in (1, 5):
        src = torch.randn((b, lat.size), generator=g, device='cuda', dtype=torch.float64)
        msk = (torch.rand((b, lat.size), generator=g, device='cuda') < 0.02).to(torch.uint8)
        monkeypatch.setenv('G2P_APPLY_PATH', 'window')
        a, am = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
        ar = tbl.apply(src, 1e20)
        assert tbl.last_path == 'window'
        monkeypatch.setenv('G2P_APPLY_PATH', 'generic')
        b_, bm = tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE)
        br = tbl.apply(src, 1e20)
        assert tbl.last_path == 'generic'
        monkeypatch.delenv('G2P_APPLY_PATH')
        assert torch.equal(a.view(torch.int64), b_.view(torch.int64)) and torch.equal(am, bm)
        assert torch.equal(ar.view(torch.int64), br.view(torch.int64))
    const = tbl.apply(torch.full((1, lat.size), 7.25, device='cuda', dtype=torch.float64), 0.0)
    assert torch.allclose(const, torch.full_like(const, 7.25), rtol=2e-15, atol=0)


def test_writer_post_matches_reference_golden(dev, golden_dir):
    """g2p_writer_post_apply against the outputs of the reference writers' own `_mask_values` (netCDF with and
    without scale/offset, PCRaster, plain and masked values)."""
    g = np.load(os.path.join(golden_dir, 'writer_post.npz'))
    v, vmask, clone, mvo = g['values'], g['vmask'], g['clone'], float(g['mv_output'])
    t = v.shape[0]

    def run(post, masked):
        d = torch.from_numpy(v.reshape(t, -1).copy()).cuda()
        dm = torch.from_numpy(vmask.reshape(t, -1).astype(np.uint8)).cuda() if masked else None
        return post.apply_(d, dm).cpu().numpy().reshape(v.shape)

    np.testing.assert_array_equal(run(dev.WriterPost(clone, float(g['nc_mv']), match=mvo), False), g['nc_out'])
    np.testing.assert_array_equal(run(dev.WriterPost(clone, float(g['nc_scaled_mv']), match=mvo), False), g['nc_scaled_out'])
    np.testing.assert_array_equal(run(dev.WriterPost(clone, float(g['nc_mv']), match=mvo), True), g['nc_masked_out'])
    np.testing.assert_array_equal(run(dev.WriterPost(clone, mvo), False), g['pcr_out'])
    np.testing.assert_array_equal(run(dev.WriterPost(clone, mvo), True), g['pcr_masked_out'])
    # no clone map: only NaN / == mv_output / masked cells change
    np.testing.assert_array_equal(run(dev.WriterPost(None, 7.0, match=mvo), True),
                                  R.writer_post(v, np.zeros_like(clone), 7.0, mv_output=mvo, value_mask=vmask))


def test_apply_with_writer_post_epilogue(dev, monkeypatch):
    """interpolate (+ correct) + the writers' post-processing in ONE windowed launch == generic apply, then the
    Corrector pass, then the oracle's restatement of the writers; O320 -> EFAS invdist and nearest."""
    from pyg2p_b200 import manipulation as M
    from pyg2p_b200 import synth
    lat, lon, det = synth.octahedral_grid(320)
    tl, tlo = synth.efas_target()
    ix = dev.SourceIndex(lon, lat, det['radius'])
    mub = ix.min_upper_bound(det['Nj'])
    rng = np.random.default_rng(3)
    clone = rng.random(tl.shape) < 0.3
    mv_out = float(synth.PCR_MV_F32)
    dem = (500.0 + 300 * rng.random(tl.shape)).astype(np.float32)
    dem[rng.random(tl.shape) < 0.1] = synth.PCR_MV_F32
    gem = 3.0 * rng.random(tl.shape)
    gem[rng.random(tl.shape) < 0.05] = synth.GRIB_MV
    corr = M.Corrector(dem, float(synth.PCR_MV_F32), gem, synth.GRIB_MV)
    g = torch.Generator(device='cuda').manual_seed(4)
    src = 280.0 + torch.randn((9, lat.size), generator=g, device='cuda', dtype=torch.float64)
    msk = (torch.rand((9, lat.size), generator=g, device='cuda') < 0.05).to(torch.uint8)
    for k, mode in ((4, L.MODE_INVDIST), (1, L.MODE_NEAREST)):
        # a regional source subset would leave rows beyond mub; shrink mub instead so that both NaN rows (invdist)
        # and mv_out rows (nearest) occur
        out = ix.knn(tlo, tl, k=k, mode=mode, mub=0.6 * mub, want_flags=False)
        tbl = dev.DeviceTable(out['idx'], out['coef'], lat.size, grid_shape=tl.shape)
        for post, mvw, match in ((dev.WriterPost(clone, 9.969209968386869e36 * 0.01 + 273.15, match=mv_out), None, mv_out),
                                 (dev.WriterPost(clone, mv_out), None, None)):
            for policy in (L.MASK_AS_REFERENCE, L.MASK_PROPAGATE, L.MASK_FILL):
                for c in (None, corr):
                    kw = dict(src_mask=msk if policy != L.MASK_AS_REFERENCE else None, mask_policy=policy)
                    monkeypatch.setenv('G2P_APPLY_PATH', 'window')
                    got = tbl.apply(src, mv_out, correction=c, post=post, **kw)
                    assert tbl.last_path == 'window'
                    got = got[0] if policy == L.MASK_PROPAGATE else got
                    monkeypatch.setenv('G2P_APPLY_PATH', 'generic')
                    two = tbl.apply(src, mv_out, correction=c, post=post, **kw)
                    two = two[0] if policy == L.MASK_PROPAGATE else two
                    ref = tbl.apply(src, mv_out, correction=c, **kw)
                    assert tbl.last_path == 'generic'
                    ref, rmask = (ref if policy == L.MASK_PROPAGATE else (ref, None))
                    want = R.writer_post(ref.cpu().numpy().reshape((-1,) + tl.shape), clone, post.mv_write,
                                         mv_output=post.match,
                                         value_mask=rmask.cpu().numpy().reshape((-1,) + tl.shape) if rmask is not None else None)
                    np.testing.assert_array_equal(got.cpu().numpy().reshape(want.shape), want)
                    np.testing.assert_array_equal(two.cpu().numpy().reshape(want.shape), want)
        assert np.isnan(ref.cpu().numpy()).any() or (ref == mv_out).any()



@pytest.mark.parametrize('cols', [1003, 1000])
def test_windowed_kernel_repeats_bit_identically(dev, monkeypatch, cols):
    """Stand-in for a racecheck run (compute-sanitizer is closed on the GPU pool, profiles/r02_sanitizer_closed.log): the
    4-stage cp.async ring + fold pass of the windowed kernel is the race-prone code of the tree.  A race between the
    staging copies, the fold pass and the gathers shows up as run-to-run differences: the same launch is repeated 40 times
    for every mask form on a ragged table with several chunks per tile (b > stages, odd sizes), each result compared bit
    for bit with the first and with the generic gather kernel, which has no shared-memory staging at all.  cols = 1000
    (rows of whole mask bytes) puts the marked forms on the lagged loop - five stages handed over through mbarrier rings
    instead of __syncthreads - with ballot-assembled mask rows; cols = 1003 on the atomicOr rows."""
    rng = np.random.default_rng(99)
    rows, k, b, n = 61, 4, 37, 30011
    m = rows * cols
    centre = (np.arange(m) * (n / m)).astype(np.int64)
    idx = np.clip(centre[:, None] + rng.integers(-60, 61, size=(m, k)), 0, n)
    idx[rng.random(m) < 0.02] = n
    w = rng.random((m, k))
    w /= w.sum(axis=1, keepdims=True)
    tbl = dev.DeviceTable.from_record(R.to_record(idx, w), n, grid_shape=(rows, cols))
    v = rng.standard_normal((b, n)) * 100
    mask = rng.random((b, n)) < 0.05
    src, msk = _padded(v), _padded(mask.view(np.uint8))
    bits = torch.from_numpy(dev.pack_mask_bits(mask, src.stride(0))).cuda()
    marked = _padded(v)
    dev.mark_masked(marked, msk)
    forms = {'plain': lambda: (tbl.apply(src, 1e20),),
             'propagate': lambda: tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE),
             'bits': lambda: tbl.apply(src, 1e20, src_mask=bits, mask_policy=L.MASK_PROPAGATE_BITS),
             'marked': lambda: tbl.apply(marked, 1e20, mask_policy=L.MASK_MARKED),
             'marked, no mask rows': lambda: (tbl.apply(marked, 1e20, mask_policy=L.MASK_MARKED, out_mask=False),)}
    monkeypatch.setenv('G2P_APPLY_PATH', 'generic')
    ref_plain = tbl.apply(src, 1e20).clone()
    ref_o, ref_m = [t.clone() for t in tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE)]
    monkeypatch.setenv('G2P_APPLY_PATH', 'window')
    for name, fn in forms.items():
        first = [t.clone() for t in fn()]
        assert tbl.last_path == 'window', tbl.plan_error
        want = ref_plain if name == 'plain' else ref_o
        assert torch.equal(first[0].view(torch.int64), want.view(torch.int64)), name
        if name == 'propagate':
            assert torch.equal(first[1], ref_m)
        if name in ('bits', 'marked'):
            assert np.array_equal(dev.unpack_mask_bits(first[1], m), ref_m.cpu().numpy().astype(bool)), name
        for _ in range(40):
            again = fn()
            for a_, f_ in zip(again, first):
                assert torch.equal(a_.view(torch.uint8), f_.view(torch.uint8)), f'{name}: a repeated launch differs'
