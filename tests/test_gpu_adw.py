"""GPU parity of the ADW (Shepard, k = 11) intertable build - SURVEY 8(f)-4 - against vectors produced by the reference's
own code (`_build_weights_invdist(adw_type='Shepard')` + its numba kernel, tests/golden/make_golden.py adw) and against
the oracle.  Everything goes through the C ABI (g2p_knn RAW, g2p_adw_distance_sum, g2p_weights_adw).

Contract:
  * the weight stage is BIT-EXACT given the same neighbours, distances and reference radius (float64 and float32 maps);
  * the reference radius (a mean over all targets) is the exactly rounded mean; the reference's own value comes from a
    numba parallel reduction whose rounding depends on its thread count (a few ulp);
  * end to end from lat/lon: indexes equal off flagged ties; weights equal the oracle's on the GPU's own distances bit for
    bit, and the reference's own output within 1e-12 relative except where a station sits on the cut-off radius r' (there
    `d/r' - 1` cancels and one ulp of d or r_ref moves the weight by ~1e-16 / |d/r' - 1| relative).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import interp_ref as R
from pyg2p_b200 import _lib as L
from pyg2p_b200 import synth

pytestmark = pytest.mark.gpu

ADW_CASES = ['o24_box_f64', 'o24_laea_f32mv', 'o24_coincident_f64']


@pytest.fixture(scope='module')
def dev():
    from pyg2p_b200 import device
    device.require_cuda()
    return device


def _oracle(g):
    det = synth.GridDetails(radius=float(g['radius']), Nj=int(g['nj']), gridType='reduced_gg')
    o = R.BuildOracle(g['src_lon'], g['src_lat'], det, g['z'], 11, float(g['mv_target']), synth.GRIB_MV, mode='adw')
    return o, det


@pytest.mark.parametrize('case', ADW_CASES)
def test_adw_weights_bit_exact_given_same_neighbours(dev, golden_dir, case):
    """K5b in isolation: the oracle's (idx, dist) and its reference radius through g2p_weights_adw."""
    g = np.load(os.path.join(golden_dir, f'build_adw_{case}.npz'))
    o, _ = _oracle(g)
    _, w_ref, i_ref = o.interpolate(g['tgt_lon'], g['tgt_lat'])
    np.testing.assert_array_equal(w_ref, g['adw_weights'])
    d = o.raw_distances
    f32 = d.dtype == np.float32
    _, _, r_ref = R.adw_cutoff_weights(d)
    idx = torch.from_numpy(np.ascontiguousarray(o.raw_indexes.T).astype(np.int32)).cuda()
    coef = torch.from_numpy(np.ascontiguousarray(d.T).astype(np.float64)).cuda()
    tdt = torch.float32 if f32 else torch.float64
    t_lat = dev.to_device(g['tgt_lat'].ravel(), tdt, 'cuda')
    t_lon = dev.to_device(g['tgt_lon'].ravel(), tdt, 'cuda')
    s_lat = dev.to_device(g['src_lat'].ravel(), torch.float64, 'cuda')
    s_lon = dev.to_device(g['src_lon'].ravel(), torch.float64, 'cuda')
    # the device's exact mean against the sequential one of the vectors: the same number to a few ulp
    got_ref = dev.adw_reference_radius(coef, L.F32 if f32 else L.F64)
    assert got_ref == math.fsum(d[:, 6].astype(np.float64).tolist()) / len(d)
    assert abs(got_ref - r_ref) <= 8 * np.spacing(r_ref)
    dev.weights_adw(idx, coef, t_lat, t_lon, s_lat, s_lon, float(r_ref), L.F32 if f32 else L.F64)
    w = coef.cpu().numpy().T
    np.testing.assert_array_equal(w, g['adw_weights'].astype(np.float64))         # == the reference's own output
    np.testing.assert_array_equal(idx.cpu().numpy().T.astype(np.int64), g['adw_indexes'])


@pytest.mark.parametrize('case', ADW_CASES + ['o24_box_f64_splits'])
def test_adw_build_from_latlon(dev, golden_dir, case):
    """ScipyInterpolation(mode='adw') end to end on the GPU against the reference's output."""
    from pyg2p_b200.scipy_interpolation import ScipyInterpolation
    g = np.load(os.path.join(golden_dir, f'build_adw_{case}.npz'))
    o, det = _oracle(g)
    si = ScipyInterpolation(g['src_lon'], g['src_lat'], det, g['z'], 11, float(g['mv_target']), synth.GRIB_MV,
                            mode='adw', num_of_splits=int(g['num_of_splits']) or None)
    assert si.min_upper_bound is None
    result, w, idx = si.interpolate(g['tgt_lon'], g['tgt_lat'])
    gw, gi = g['adw_weights'], g['adw_indexes']
    assert w.dtype == gw.dtype and w.shape == gw.shape and idx.dtype == gi.dtype and idx.shape == gi.shape
    tie = (si.flags.cpu().numpy() & (L.FLAG_TIE | L.FLAG_NEAR_TIE)) != 0
    same = (idx == gi).all(axis=1)
    assert same[~tie].all()
    f32 = gw.dtype == np.float32
    # (a) the oracle on the GPU's own neighbours and radius: bit for bit
    raw = si.index.knn(g['tgt_lon'].ravel(), g['tgt_lat'].ravel(), k=11, mode=L.MODE_RAW)
    d_gpu = raw['coef'].cpu().numpy().T
    i_gpu = raw['idx'].cpu().numpy().T.astype(np.int64)
    if f32:
        d_gpu = d_gpu.astype(np.float32)
    _, w_o, _ = R.build_weights_adw(d_gpu, i_gpu, g['z'], g['tgt_lat'], g['tgt_lon'], g['src_lat'], g['src_lon'],
                                    ref_radius=si.ref_radius)
    np.testing.assert_array_equal(w, w_o)
    # (b) the reference's own output: identical for float32 maps (xyz and distances are bit-identical there), else
    # within 1e-12 except for stations on the cut-off radius
    ok = same & ~tie
    if f32 and not int(g['num_of_splits']):
        _, _, r_seq = R.adw_cutoff_weights(d_gpu)
        if si.ref_radius == r_seq:
            np.testing.assert_array_equal(w[ok], gw[ok])
    rel = np.abs(w[ok].astype(np.float64) - gw[ok]) / np.maximum(np.abs(gw[ok]), 1e-300)
    rel[gw[ok] == w[ok]] = 0
    tol = 1e-6 if f32 else 1e-12
    close = rel <= tol
    assert close.mean() > 0.995, close.mean()
    # the rest: a station within a hair of the cut-off radius
    rows = np.nonzero(ok)[0][np.nonzero(~close.all(axis=1))[0]]
    d64 = d_gpu.astype(np.float64)
    for r in rows:
        _, s_r, _ = R.adw_cutoff_weights(d64[r:r + 1], ref_radius=si.ref_radius)
        rr = [d64[r, 10], d64[r, 4], si.ref_radius]
        edge = min(abs(d64[r, j] / x - 1) for j in range(11) for x in rr)
        assert edge < 1e-3, (r, edge)
    np.testing.assert_allclose(w[ok].astype(np.float64), gw[ok], rtol=2e-3 if f32 else 1e-6, atol=1e-12)
    # the result field: the table applied to z
    res_ref = np.einsum('ij,ij->i', w.astype(np.float64), g['z'][idx])
    np.testing.assert_allclose(np.asarray(result), res_ref, rtol=1e-13)


def test_adw_o320_to_efas_sample(dev):
    """BASELINE configs[0] grid shapes in mode adw: O320 -> EFAS 5 km, all 950 000 rows against the oracle evaluated on
    the GPU's own neighbours (bit for bit), neighbours against cKDTree on the same xyz."""
    from pyg2p_b200.scipy_interpolation import ScipyInterpolation
    lat, lon, det = synth.octahedral_grid(320)
    tl, tlo = synth.efas_target()
    z = synth.field_2t(lat, lon)
    si = ScipyInterpolation(lon, lat, det, z, 11, synth.NC_FILL_F64, synth.GRIB_MV, mode='adw')
    table = si.build_table(tlo, tl)
    idx, w = table.indexes_coeffs()
    assert idx.shape == (tl.size, 11) and w.dtype == np.float64
    raw = si.index.knn(tlo.ravel(), tl.ravel(), k=11, mode=L.MODE_RAW, want_xyz=True)
    d_gpu = raw['coef'].cpu().numpy().T
    i_gpu = raw['idx'].cpu().numpy().T.astype(np.int64)
    tree = R.build_tree(si.index.xyz().cpu().numpy())
    d, i = tree.query(raw['xyz'].cpu().numpy(), k=11, workers=-1)
    tie = (raw['flags'].cpu().numpy() & L.FLAG_TIE) != 0
    np.testing.assert_array_equal(d_gpu, d)
    np.testing.assert_array_equal(i_gpu[~tie], i[~tie])
    assert si.ref_radius == math.fsum(d[:, 6].tolist()) / len(d)
    _, w_o, _ = R.build_weights_adw(d_gpu, i_gpu, z, tl, tlo, lat, lon, ref_radius=si.ref_radius)
    np.testing.assert_array_equal(w, w_o)
    np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=1e-14)
    # the k = 11 table applies through the generic kernel, bit-exact against the oracle's sequential einsum order
    fields = np.stack([synth.field_2t(lat, lon, member=b) for b in range(2)])
    got = table.apply(torch.from_numpy(fields).cuda(), synth.NC_FILL_F64).cpu().numpy()
    rec = table.to_record()
    for b in range(2):
        want = R.apply_sequential(fields[b], rec['coeffs'], rec['indexes'], 11, synth.NC_FILL_F64)
        np.testing.assert_array_equal(got[b], want)
