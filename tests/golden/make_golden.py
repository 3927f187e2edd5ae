#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (pyg2p 3.2.8) cannot be imported as is: eccodes, numexpr, GDAL, netCDF4,
dask, ujson are absent and `np.NaN` is gone from numpy 2.  This script installs import
shims for exactly those third-party names, then imports the UNMODIFIED reference modules
from /root/reference/src and calls, on small synthetic inputs:

  * `ScipyInterpolation(...).interpolate(...)`           (modes nearest, invdist, adw; f64 and f32 targets)
  * `Interpolator._interpolate_scipy_append_mv(...)`     (k=1, k=4, plain and masked input)
  * `Converter.convert / cut_off_negative`, `Aggregator.do_manipulation` (3 modes, `@halfweights`
    averages, masked inputs), `Corrector.correct` (both formulas of the README)
  * `NetCDFWriter._mask_values`, `PCRasterWriter._mask_values` around the writers' inline statements

The only arithmetic not executed by reference code is inside the `numexpr.evaluate` shim,
which evaluates the reference's expression strings with numpy under numexpr's typing rules
(Python float constants are doubles, masks ignored).  numpy float64 sin/cos are bit-equal
to libm's (which numexpr calls) in this image.  Outputs: tests/golden/*.npz.
"""
import ast
import contextlib
import importlib.abc
import importlib.machinery
import io
import os
import sys
import zlib
from unittest.mock import MagicMock

import numpy as np
import numpy.ma as ma

sys.dont_write_bytecode = True
# numba's parallel reductions (the ADW reference radius, a mean) round differently for every thread count: one thread
# = a sequential sum, which the oracle restates
os.environ.setdefault('NUMBA_NUM_THREADS', '1')
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
np.NaN = np.nan  # removed in numpy 2; reference util/numeric.py:15 uses it as a default

# ----------------------------------------------------------------------------- shims
STUBBED = {'eccodes', 'osgeo', 'dask', 'ujson', 'lisfloodutilities', 'netCDF4', 'numexpr', 'gribapi'}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in STUBBED:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


sys.meta_path.append(_StubFinder())


class _Const(ast.NodeTransformer):
    """numexpr typing: float literals are doubles, int literals are (long) ints."""

    def visit_Constant(self, node):
        if isinstance(node.value, bool):
            return node
        if isinstance(node.value, float):
            return ast.copy_location(ast.Call(ast.Name('_f64', ast.Load()), [node], []), node)
        if isinstance(node.value, int):
            return ast.copy_location(ast.Call(ast.Name('_i64', ast.Load()), [node], []), node)
        return node


def _ne_evaluate(expr, local_dict=None, global_dict=None, out=None, **kw):
    frame = sys._getframe(1)
    names = dict(frame.f_globals if global_dict is None else global_dict)
    names.update(frame.f_locals if local_dict is None else local_dict)
    tree = ast.fix_missing_locations(_Const().visit(ast.parse(expr, mode='eval')))
    # float32 sin/cos: numexpr calls libm sinf/cosf (oracle/libm32.py), not numpy's SIMD kernels
    env = {'where': np.where, 'cos': libm32.cos, 'sin': libm32.sin, 'sqrt': np.sqrt, 'exp': np.exp,
           'log': np.log, 'abs': np.abs, '_f64': np.float64, '_i64': np.int64}
    for n in {x.id for x in ast.walk(tree) if isinstance(x, ast.Name)}:
        if n in env:
            continue
        v = names[n]
        if isinstance(v, np.ndarray):
            v = np.asarray(ma.getdata(v))       # numexpr ignores masks
        elif isinstance(v, bool):
            pass
        elif isinstance(v, float):
            v = np.float64(v)
        elif isinstance(v, int):
            v = np.int64(v)
        env[n] = v
    with np.errstate(all='ignore'):
        res = eval(compile(tree, '<numexpr-shim>', 'eval'), {'__builtins__': {}}, env)
    if out is not None:
        out[...] = res
        return out
    return np.asarray(res)


import numexpr  # noqa: E402  (the stub)
numexpr.evaluate = _ne_evaluate
import eccodes  # noqa: E402
eccodes.codes_get_api_version = lambda: '2.30.0'
import netCDF4  # noqa: E402
netCDF4.default_fillvals = {'f4': 9.969209968386869e+36, 'f8': 9.969209968386869e+36,
                            'i4': -2147483647, 'i8': -9223372036854775806}

sys.path.insert(0, '/root/reference/src')
import warnings  # noqa: E402
warnings.simplefilter('ignore')
from pyg2p import Step  # noqa: E402
from pyg2p.main.interpolation import Interpolator  # noqa: E402
from pyg2p.main.interpolation.scipy_interpolation_lib import ScipyInterpolation  # noqa: E402
from pyg2p.main.manipulation.aggregator import Aggregator  # noqa: E402
from pyg2p.main.manipulation.conversion import Converter  # noqa: E402
from pyg2p.main.manipulation.correction import Corrector  # noqa: E402

from pyg2p_b200 import synth  # noqa: E402  (input generators only)
from oracle import libm32  # noqa: E402  (ctypes binding of glibc sinf/cosf)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def filled(a):
    return np.asarray(ma.filled(a)) if isinstance(a, ma.MaskedArray) else np.asarray(a)


# ----------------------------------------------------------------------------- build cases
def build_case(name, slat, slon, details, tlat, tlon, mv_target, rotated_target=False):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    z = 280.0 + 10.0 * rng.standard_normal(slat.shape)
    out = dict(src_lat=slat, src_lon=slon, tgt_lat=tlat, tgt_lon=tlon, z=z,
               radius=np.float64(details.get('radius')), nj=np.int64(details.get('Nj')),
               mv_target=np.float64(mv_target))
    for mode, k in (('nearest', 1), ('invdist', 4)):
        si = quiet(ScipyInterpolation, slon, slat, details, z, k, mv_target, synth.GRIB_MV,
                   target_is_rotated=rotated_target, parallel=False, mode=mode)
        result, weights, indexes = quiet(si.interpolate, tlon, tlat)
        out[f'{mode}_mub'] = np.float64(si.min_upper_bound)
        out[f'{mode}_result'] = filled(result)
        out[f'{mode}_weights'] = np.asarray(weights)
        out[f'{mode}_indexes'] = np.asarray(indexes)
    np.savez_compressed(os.path.join(HERE, f'build_{name}.npz'), **out)
    print(name, 'N', slat.size, 'M', tlat.size, 'mub', out['nearest_mub'],
          'sentinel rows', int((out['nearest_indexes'] == slat.size).sum()),
          'coeff dtypes', out['nearest_weights'].dtype, out['invdist_weights'].dtype)


def make_build_cases():
    slat, slon, det = synth.octahedral_grid(24)
    # A: global O24 -> regular f64 box (netCDF-style target maps)
    lon = np.arange(-30.0, 40.0, 1.0) + 0.5
    lat = np.arange(70.0, 20.0, -1.0) - 0.5
    tlon, tlat = np.meshgrid(lon, lat)
    build_case('o24_box_f64', slat, slon, det, tlat, tlon, synth.NC_FILL_F64)
    # B: global O24 -> coarse LAEA float32 maps with an MV block (PCRaster-style)
    tlat, tlon = synth.efas_target(rows=48, cols=50, cell=100000.0, dtype=np.float32)
    tlat, tlon = synth.with_mv_region(tlat, tlon)
    build_case('o24_laea_f32mv', slat, slon, det, tlat, tlon, float(synth.PCR_MV_F32))
    # C: regional source -> coarse global target: most rows are beyond min_upper_bound
    rlat, rlon, rdet = synth.regular_ll_grid(50.0, 35.0, 31, 0.0, 20.0, 41)
    lon = np.arange(-180.0, 180.0, 4.0) + 1.0
    lat = np.arange(88.0, -60.0, -4.0)
    tlon, tlat = np.meshgrid(lon, lat)
    # densify around the source so that a good share of rows is inside
    lon2, lat2 = np.meshgrid(np.linspace(-2.0, 22.0, 49), np.linspace(52.0, 33.0, 39))
    tlat = np.concatenate([tlat.ravel(), lat2.ravel()]).reshape(1, -1)
    tlon = np.concatenate([tlon.ravel(), lon2.ravel()]).reshape(1, -1)
    build_case('regional_global_f64', rlat, rlon, rdet, tlat, tlon, synth.NC_FILL_F64)
    # D: targets that coincide with source points (d0 <= 1e-10 branch) mixed with offsets
    sel = np.arange(0, slat.size, 7)
    tlat = np.concatenate([slat[sel], slat[sel] + 0.123]).reshape(2, -1)
    tlon = np.concatenate([slon[sel], slon[sel] - 0.321]).reshape(2, -1)
    build_case('o24_coincident_f64', slat, slon, det, tlat, tlon, synth.NC_FILL_F64)


def adw_case(name, slat, slon, details, tlat, tlon, mv_target, num_of_splits=None):
    """mode adw (Shepard, k = 11): `_build_weights_invdist(adw_type='Shepard')` incl. the numba kernel
    `adw_compute_weights_from_cutoff_distances`, run by the reference itself (numba is importable here)."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    z = 280.0 + 10.0 * rng.standard_normal(slat.shape)
    si = quiet(ScipyInterpolation, slon, slat, details, z, 11, mv_target, synth.GRIB_MV, parallel=False, mode='adw',
               num_of_splits=num_of_splits)
    assert si.min_upper_bound is None
    result, weights, indexes = quiet(si.interpolate, tlon, tlat)
    out = dict(src_lat=slat, src_lon=slon, tgt_lat=tlat, tgt_lon=tlon, z=z, radius=np.float64(details.get('radius')),
               nj=np.int64(details.get('Nj')), mv_target=np.float64(mv_target), adw_result=filled(result),
               adw_weights=np.asarray(weights), adw_indexes=np.asarray(indexes),
               num_of_splits=np.int64(num_of_splits or 0))
    np.savez_compressed(os.path.join(HERE, f'build_adw_{name}.npz'), **out)
    print('adw', name, 'N', slat.size, 'M', tlat.size, 'weights', out['adw_weights'].dtype, 'exact rows',
          int((out['adw_weights'][:, 0] == 1).sum()), 'NaN rows', int(np.isnan(out['adw_weights']).any(axis=1).sum()))


def make_adw_cases():
    slat, slon, det = synth.octahedral_grid(24)
    lon = np.arange(-30.0, 40.0, 1.0) + 0.5
    lat = np.arange(70.0, 20.0, -1.0) - 0.5
    tlon, tlat = np.meshgrid(lon, lat)
    adw_case('o24_box_f64', slat, slon, det, tlat, tlon, synth.NC_FILL_F64)
    # the same target in 5 row blocks: the reference radius comes from a separate k=7 query over all targets (:576-585)
    adw_case('o24_box_f64_splits', slat, slon, det, tlat, tlon, synth.NC_FILL_F64, num_of_splits=5)
    tlat, tlon = synth.efas_target(rows=48, cols=50, cell=100000.0, dtype=np.float32)
    tlat, tlon = synth.with_mv_region(tlat, tlon)
    adw_case('o24_laea_f32mv', slat, slon, det, tlat, tlon, float(synth.PCR_MV_F32))
    # targets on source points (d0 <= 1e-10 -> [1, 0, ...]; the angular term of such a row is 0/0) mixed with offsets
    sel = np.arange(0, slat.size, 7)
    tlat = np.concatenate([slat[sel], slat[sel] + 0.123]).reshape(2, -1)
    tlon = np.concatenate([np.where(slon[sel] > 180, slon[sel] - 360, slon[sel]), slon[sel] - 0.321]).reshape(2, -1)
    adw_case('o24_coincident_f64', slat, slon, det, tlat, tlon, synth.NC_FILL_F64)


# ----------------------------------------------------------------------------- apply cases
def make_apply_cases():
    g = np.load(os.path.join(HERE, 'build_regional_global_f64.npz'))
    n = g['z'].size
    rng = np.random.default_rng(5)
    interp = object.__new__(Interpolator)
    interp.mv_out = float(g['mv_target'])
    out = {'mv_out': np.float64(interp.mv_out)}
    for b in range(3):
        v = 250.0 + 50.0 * rng.random(n)
        mask = rng.random(n) < 0.1
        vm = ma.masked_where(mask, np.where(mask, synth.GRIB_MV, v), copy=True)
        out[f'v{b}'] = np.asarray(ma.getdata(vm))
        out[f'mask{b}'] = mask
        for mode, k in (('nearest', 1), ('invdist', 4)):
            idx, w = g[f'{mode}_indexes'], g[f'{mode}_weights']
            # strided field views, as `_read_intertable` hands them over (interpolation/__init__.py:137-139)
            rec = np.rec.fromarrays((idx, w), names=('indexes', 'coeffs'))
            for tag, vin in (('plain', np.asarray(ma.getdata(vm))), ('masked', vm)):
                with np.errstate(all='ignore'):
                    r = interp._interpolate_scipy_append_mv(vin, rec['coeffs'], rec['indexes'], k)
                out[f'{mode}_{tag}{b}'] = filled(r)
                out[f'{mode}_{tag}{b}_ismasked'] = np.bool_(isinstance(r, ma.MaskedArray))
    np.savez_compressed(os.path.join(HERE, 'apply_regional.npz'), **out)
    print('apply cases written; masked results are MaskedArray:',
          {k: bool(v) for k, v in out.items() if k.endswith('_ismasked')})


# ----------------------------------------------------------------------------- manipulation cases
def make_manip_cases():
    rng = np.random.default_rng(9)
    n = 257
    res = 640
    out = {}
    # Converter
    x = 0.01 * rng.standard_normal(n)
    x[::17] = synth.GRIB_MV
    for i, f in enumerate(['x=x*1000', 'x=x-273.15', 'x=x*(-0.0353)', 'x=-x*1000', 'x=x']):
        c = Converter(func=f, cut_off=True)
        c.set_missing_value(synth.GRIB_MV)
        y = c.convert(x.copy())
        out[f'conv{i}'] = np.asarray(y)
        out[f'conv{i}_cut'] = np.asarray(quiet(c.cut_off_negative, np.asarray(y).copy()))
    out['conv_x'] = x
    # Aggregator: cumulated field every 6 h, steps 0..72 (+ variants without step 0 and with a gap)
    steps = list(range(0, 73, 6))
    cum = np.cumsum(np.maximum(0, rng.standard_normal((len(steps), n))) * 2e-3, axis=0)
    mask = rng.random(n) < 0.05
    out['agg_steps'] = np.array(steps)
    out['agg_cum'] = cum
    out['agg_mask'] = mask

    def run(tag, vals, **kw):
        base = dict(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6,
                    step_type='accum', start_step=0, end_step=72, unit_time=24, mv_grib=synth.GRIB_MV,
                    force_zero_array=False)
        base.update(kw)
        a = Aggregator(**base)
        r = quiet(a.do_manipulation, vals)
        ks = sorted(r.keys(), key=lambda s: (s.end_step, s.start_step))
        out[f'{tag}_keys'] = np.array([(s.start_step, s.end_step) for s in ks])
        out[f'{tag}_vals'] = np.array([np.asarray(ma.getdata(r[s])) for s in ks])
        out[f'{tag}_masks'] = np.array([np.broadcast_to(ma.getmaskarray(r[s]) if isinstance(r[s], ma.MaskedArray)
                                                         else False, (n,)) for s in ks])

    def vals(stepset, masked=False):
        d = {}
        for s in stepset:
            v = cum[steps.index(s)].copy()
            if masked:
                v = ma.masked_where(mask, v, copy=True)
            d[Step(0, s, res, 6, 0)] = v
        return d

    run('acc_full', vals(steps))
    run('acc_nozero', vals(steps[1:]))
    run('acc_gap', vals([s for s in steps if s != 48]))
    run('acc_masked', vals(steps, masked=True))
    run('acc_forcezero', vals(steps), force_zero_array=True)
    run('acc_step12_unit6', vals(steps), aggr_step=12, unit_time=6, start_step=12)
    inst = 280.0 + rng.standard_normal((len(steps), n))
    out['inst_fields'] = inst

    def ivals(stepset):
        return {Step(s, s, res, 6, 0): inst[steps.index(s)].copy() for s in stepset}

    run('avg_full', ivals(steps), aggr_type='average', step_type='instant', start_step=0, end_step=72)
    run('avg_start24', ivals(steps), aggr_type='average', step_type='instant', start_step=24, end_step=72)
    # `@halfweights` averages (reference tests/test_aggregator.py:57): full, later start, a missing step, masked inputs
    half = dict(aggr_type='average', step_type='instant', aggr_halfweights=True)

    def mvals(stepset):
        return {Step(s, s, res, 6, 0): ma.masked_where(mask, inst[steps.index(s)].copy(), copy=True) for s in stepset}

    run('avgh_full', ivals(steps), start_step=0, end_step=72, **half)
    run('avgh_start24', ivals(steps), start_step=24, end_step=72, **half)
    run('avgh_gap', ivals([s for s in steps if s not in (24, 54)]), start_step=0, end_step=72, **half)
    run('avgh_masked', mvals(steps), start_step=0, end_step=72, **half)
    run('avgh_step12', ivals(steps), start_step=12, end_step=72, aggr_step=12, **half)
    run('avg_masked', mvals(steps), aggr_type='average', step_type='instant', start_step=0, end_step=72)
    run('inst_full', ivals(steps), aggr_type='instantaneous', step_type='instant', start_step=0, end_step=72,
        aggr_step=12)
    run('inst_gap', ivals([s for s in steps if s not in (0, 36)]), aggr_type='instantaneous',
        step_type='instant', start_step=0, end_step=72, aggr_step=12)
    # Corrector.correct on a target grid
    m = 300
    p = 280.0 + 5 * rng.standard_normal(m)
    p[::13] = np.nan                              # invdist rows beyond mub stay NaN
    p[5::29] = 9.969209968386869e36               # already-missing output values
    dem = 500.0 + 300 * rng.random(m)
    dem_mv = float(synth.PCR_MV_F32)
    dem[::11] = dem_mv
    gem = 3.0 * rng.random(m)
    gem_mv = synth.GRIB_MV
    gem[3::19] = gem_mv
    c = object.__new__(Corrector)
    Corrector.__mro__[1].__init__(c)              # Loggable.__init__
    c._dem_values, c._dem_missing_value = dem, dem_mv
    c._gem_values, c._gem_missing_value = gem, gem_mv
    c._formula = 'p+gem-dem*0.0065'
    c._numexpr_eval = f'where((dem!=dem_mv) & (p!=mv) & (gem!=gem_mv), {c._formula}, mv)'
    out.update(corr_p=p, corr_dem=dem, corr_gem=gem, corr_dem_mv=np.float64(dem_mv), corr_gem_mv=np.float64(gem_mv))
    out['corr_out'] = filled(c.correct(p.copy()))
    c._formula = 'p/gem*(10**((-0.159)*dem/1000))'
    c._numexpr_eval = f'where((dem!=dem_mv) & (p!=mv) & (gem!=gem_mv), {c._formula}, mv)'
    out['corr_out_pressure'] = filled(c.correct(p.copy()))
    np.savez_compressed(os.path.join(HERE, 'manip.npz'), **out)
    print('manip cases written:', sorted(k for k in out if k.endswith('_keys')))


# ----------------------------------------------------------------------------- writer post-processing cases
def make_writer_cases():
    """The writers' treatment of an interpolated field, by the reference's own `_mask_values` methods
    (writers/netcdf.py:170-183, writers/pcr.py:44-50) around the inline statements of writers/__init__.py:66,
    writers/netcdf.py:149-151 and writers/pcr.py:33 (copied here as they are not callable on their own)."""
    from pyg2p.main.writers.netcdf import NetCDFWriter
    from pyg2p.main.writers.pcr import PCRasterWriter
    rng = np.random.default_rng(21)
    t, rows, cols = 3, 7, 9
    mv_output = float(synth.PCR_MV_F32)
    clone = rng.random((rows, cols)) < 0.25
    v = 280.0 + rng.standard_normal((t, rows, cols))
    v[rng.random(v.shape) < 0.1] = np.nan          # invdist rows beyond min_upper_bound
    v[rng.random(v.shape) < 0.1] = mv_output       # nearest rows beyond min_upper_bound / masked under FILL
    vmask = rng.random(v.shape) < 0.15
    out = dict(values=v, vmask=vmask, clone=clone, mv_output=np.float64(mv_output))
    # netCDF: per step `out_v[out_v == mv_output] = nan`; np.array(list) drops masks; scaled missing value
    for tag, scale, offset in (('nc', 1.0, 0.0), ('nc_scaled', 0.01, 273.15)):
        w = object.__new__(NetCDFWriter)
        w._mask = clone
        steps = []
        for i in range(t):
            out_v = v[i].copy()
            out_v[out_v == mv_output] = np.nan
            steps.append(out_v)
        values = np.array(steps, dtype=np.float64)
        mv_scaled = np.float64(netCDF4.default_fillvals['f8']) * np.float64(scale) + np.float64(offset)
        values = w._mask_values(values, mv_scaled)
        values[np.isnan(values)] = mv_scaled
        out[f'{tag}_out'] = np.asarray(values)
        out[f'{tag}_mv'] = np.float64(mv_scaled)
    # the same with MaskedArray steps kept masked (what a mask-propagating caller hands to `_mask_values`)
    w = object.__new__(NetCDFWriter)
    w._mask = clone
    mvals = ma.masked_array(np.where(v == mv_output, np.nan, v), mask=vmask)
    mv_scaled = np.float64(netCDF4.default_fillvals['f8'])
    values = w._mask_values(mvals, mv_scaled)
    values[np.isnan(values)] = mv_scaled
    out['nc_masked_out'] = np.asarray(values)
    # PCRaster: one map at a time, `values[isnan(values)] = mv` then `_mask_values`
    pw = object.__new__(PCRasterWriter)
    pw._mask = clone
    pw.mv = mv_output
    res_plain, res_masked = [], []
    for i in range(t):
        a = v[i].copy()
        a[np.isnan(a)] = pw.mv
        res_plain.append(np.asarray(pw._mask_values(a)))
        b = ma.masked_array(v[i].copy(), mask=vmask[i])
        b[np.isnan(b)] = pw.mv
        res_masked.append(np.asarray(pw._mask_values(b)))
    out['pcr_out'] = np.array(res_plain)
    out['pcr_masked_out'] = np.array(res_masked)
    np.savez_compressed(os.path.join(HERE, 'writer_post.npz'), **out)
    print('writer cases written')


def make_reference_fixture_copies():
    """Inputs of the reference's own golden intertable (tests/data/tbl_pf10slhf_550800_scipy_nearest.npy.gz):
    the EFAS lat/lon PCRaster maps (as float32 arrays, MV cells = -3.4028235e38 as GDAL delivers them) and the
    first 400 bytes of tests/data/input.grib (indicator + product + grid description sections of message 1)."""
    from pyg2p_b200.maps import LatLong
    ll = LatLong('/root/reference/tests/data/lat.map', '/root/reference/tests/data/lon.map')
    np.savez_compressed(os.path.join(HERE, 'ref_efas_latlon_maps.npz'), lat=ll.lats, lon=ll.lons,
                        identifier=np.array(ll.identifier), mv=np.float64(ll.mv))
    with open('/root/reference/tests/data/input.grib', 'rb') as f:
        head = f.read(400)
    with open(os.path.join(HERE, 'ref_input_grib_header.bin'), 'wb') as f:
        f.write(head)
    # inputs of the reference's own test scenarios (tests/test_interpolation.py, tests/test_correction.py): the first
    # 2t message of tests/data/input.grib (step 0 = `first_step_range`), tests/data/geopotential.grib, tests/data/dem.map
    from pyg2p_b200 import grib1
    from pyg2p_b200.maps import Dem
    with open('/root/reference/tests/data/input.grib', 'rb') as f:
        first = grib1.split_messages(f.read())[0]
    with open(os.path.join(HERE, 'ref_input_2t_step0.grib'), 'wb') as f:
        f.write(first)
    with open('/root/reference/tests/data/geopotential.grib', 'rb') as f:
        geo = grib1.split_messages(f.read())[0]
    with open(os.path.join(HERE, 'ref_geopotential.grib'), 'wb') as f:
        f.write(geo)
    dem = Dem('/root/reference/tests/data/dem.map')
    np.savez_compressed(os.path.join(HERE, 'ref_efas_dem_map.npz'), dem=dem.values, mv=np.float64(dem.mv))
    print('reference fixture copies written:', ll.identifier)


if __name__ == '__main__':
    only = set(sys.argv[1:])            # e.g. `make_golden.py adw`: regenerate one group
    for tag, fn in (('fixtures', make_reference_fixture_copies), ('build', make_build_cases), ('adw', make_adw_cases),
                    ('apply', make_apply_cases), ('manip', make_manip_cases), ('writer', make_writer_cases)):
        if not only or tag in only:
            fn()
