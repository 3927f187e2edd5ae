"""Oracle (test infrastructure): numpy restatement of the source-grid manipulation stages
that the CUDA pipeline fuses with the intertable apply.  Citations are to
/root/reference/src/pyg2p/main/manipulation/.  numexpr evaluates left to right in
float64 and ignores masks, which are re-attached afterwards; the same is done here.

Not imported by the product package.  See oracle/__init__.py.
"""
import bisect

import numpy as np
import numpy.ma as ma

NC_FILL_F64 = 9.969209968386869e36  # netCDF4.default_fillvals['f8'] (correction.py:78)


def _masks(*arrays):
    """util/numeric.py:29-36 `get_masks`."""
    mask = None
    for a in arrays:
        if not isinstance(a, ma.core.MaskedArray):
            continue
        mask = a.mask if mask is None else mask | a.mask
    return mask


def _eval(expr, **names):
    env = {'where': np.where, 'abs': np.abs, 'sqrt': np.sqrt, 'exp': np.exp, 'log': np.log}
    env.update({k: (ma.getdata(v) if isinstance(v, np.ndarray) else v) for k, v in names.items()})
    return eval(expr, {'__builtins__': {}}, env)


# --------------------------------------------------------------------------- Converter
def convert(x, function_str, mv):
    """conversion.py:21,33-42: `where(x!=mv, f(x), mv)`, mask preserved; 'x=x' is identity."""
    if function_str == 'x=x':
        return x
    res = _eval('where(x!=mv, {}, mv)'.format(function_str.replace('x=', '')), x=x, mv=mv)
    if isinstance(x, ma.core.MaskedArray):
        res = ma.masked_where(x.mask, res, copy=False)
    return res


def cut_off_negative(values):
    """conversion.py:44-52: `where(values<0, 0, values)` on plain data (mask dropped)."""
    if isinstance(values, dict):
        return {k: cut_off_negative(v) for k, v in values.items()}
    values = ma.getdata(values)
    return np.where(values < 0, 0, values)


# --------------------------------------------------------------------------- Aggregator
def _synth(v_ord, keys, t):
    """Linear-in-time synthesis of a missing step (aggregator.py:88-106, 118-129)."""
    i_next = bisect.bisect_left(keys, t)
    nxt = keys[i_next]
    i_orig = bisect.bisect_right(keys, t)
    if i_orig == i_next:
        i_orig -= 1
    org = keys[i_orig]
    a, b = v_ord[org], v_ord[nxt]
    out = ma.getdata(a) + (ma.getdata(b) - ma.getdata(a)) * ((t - org) / (nxt - org))
    return ma.masked_where(_masks(a, b), out, copy=False)


def accumulation(values, start, end, aggr_step, unit_time, force_zero=False):
    """aggregator.py:72-153.  `values`: {end_step:int -> array}.  Returns {(t-D, t) -> array}."""
    v_ord = dict(sorted(values.items()))
    shape = next(iter(v_ord.values())).shape
    if start == 0:
        start = aggr_step
    out = {}
    created_zero = False
    for t in range(start, end + 1, aggr_step):
        keys = list(v_ord.keys())
        if t not in keys:
            v_ord[t] = _synth(v_ord, keys, t)
        if t - aggr_step >= 0 and (t - aggr_step) not in keys and not created_zero:
            if t - aggr_step == 0:
                v_ord[0] = np.zeros(shape)
                created_zero = True
            else:
                # NB the reference interpolates with `iter_` (=t), not t-D, in the time fraction (:128)
                tt = t - aggr_step
                i_next = bisect.bisect_left(keys, tt)
                nxt = keys[i_next]
                i_orig = bisect.bisect_right(keys, tt)
                if i_orig == i_next:
                    i_orig -= 1
                org = keys[i_orig]
                a, b = v_ord[org], v_ord[nxt]
                vo = ma.getdata(a) + (ma.getdata(b) - ma.getdata(a)) * ((t - org) / (nxt - org))
                v_ord[tt] = ma.masked_where(_masks(a, b), vo, copy=False)
        if t - aggr_step == 0 and force_zero:
            prev = np.zeros(shape)
        else:
            prev = v_ord[t - aggr_step]
        cur = np.zeros(shape)
        cur += v_ord[t]          # plain ndarray: `v_iter_ma` never carries a mask (:142-143)
        val = (cur - ma.getdata(prev)) * unit_time / aggr_step
        out[(t - aggr_step, t)] = ma.masked_where(_masks(cur, prev), val, copy=False)
    return out


def average(values, start, end, aggr_step, second_t_res=False):
    """aggregator.py:155-236 (no half weights): one add per hourly iterator, next available
    step when an hour is absent, divide by the count; the result is NOT masked (see the note at the end)."""
    v_ord = dict(sorted(values.items()))
    shape = next(iter(v_ord.values())).shape
    if start == 0 or second_t_res:
        it0, it1 = start, end - aggr_step + 2
    else:
        it0, it1 = start - aggr_step, end - aggr_step + 1
    out = {}
    keys = list(v_ord.keys())
    for t in range(it0, it1, aggr_step):
        acc = np.zeros(shape)
        cnt = 0.0
        for h in range(t + 1, t + aggr_step + 1):
            if h in keys:
                v = v_ord[h]
            else:
                v = v_ord[keys[bisect.bisect_left(keys, h)]]
            cnt += 1
            acc = acc + ma.getdata(v)
        res = acc / cnt
        # `numeric.get_masks(v_ord.values())` (:232): ONE argument, the dict view, no MaskedArray -> mask None
        out[(t, t + aggr_step)] = ma.masked_where(None, res, copy=False)
    return out


def average_halfweights(values, start, end, aggr_step, second_t_res=False):
    """aggregator.py:155-236 with `aggr_halfweights` (:176-207): window [t, t+D] inclusive, only the steps present
    in the input, weights dt1 = D / (present - 1), dt1 / 2 for the window's first and last hour; `v * dt` on a
    MaskedArray leaves masked elements unmultiplied (numpy.ma binary operations copy the first operand's data back
    under the mask); sum / sum of the weights (0/0 = NaN when fewer than two steps are present); not masked."""
    v_ord = dict(sorted(values.items()))
    shape = next(iter(v_ord.values())).shape
    if start == 0 or second_t_res:
        it0, it1 = start, end - aggr_step + 2
    else:
        it0, it1 = start - aggr_step, end - aggr_step + 1
    out = {}
    keys = list(v_ord.keys())
    for t in range(it0, it1, aggr_step):
        iter_from, iter_to = t, t + aggr_step + 1
        acc = np.zeros(shape)
        cnt = 0.0
        count_steps = sum(1 for h in range(iter_from, iter_to) if h in keys) - 1
        if count_steps > 0:
            dt1 = aggr_step / count_steps
            dt2 = dt1 / 2
            for h in range(iter_from, iter_to):
                if h not in keys:
                    continue
                w = dt2 if (h == iter_from or h == iter_to - 1) else dt1
                v = v_ord[h]
                data = np.asarray(ma.getdata(v), dtype=np.float64)
                vm = data * w
                if isinstance(v, ma.MaskedArray) and v.mask is not ma.nomask:
                    vm = np.where(ma.getmaskarray(v), data, vm)
                cnt += w
                acc = acc + vm
        with np.errstate(all='ignore'):
            res = acc / cnt if cnt != 0.0 else acc / np.float64(0.0)
        out[(t, t + aggr_step)] = ma.masked_where(None, res, copy=False)   # unmasked, see average()
    return out


def instantaneous(values, start, end, aggr_step, input_step=0, step_type='instant'):
    """aggregator.py:238-279."""
    v_ord = dict(sorted(values.items()))
    keys = list(v_ord.keys())
    shape = next(iter(v_ord.values())).shape
    s = start - aggr_step if start - aggr_step > 0 else start
    if step_type == 'avg':
        s = start + input_step
    out = {}
    for t in range(s, end + 1, aggr_step):
        res = np.zeros(shape)
        if t in keys:
            res += v_ord[t]
        elif t != 0:
            res += v_ord[keys[bisect.bisect_right(keys, t)]]
        out[(t, t)] = res
    return out


# --------------------------------------------------------------------------- Corrector
def gem_from_geopotential(z, mv, gem_formula='(z/9.81)*0.0065'):
    """correction.py:47,95-99: `where(z != mv, gem_formula, mv)` on the source grid."""
    return _eval(f'where(z != mv, {gem_formula}, mv)', z=z, mv=mv)


def correct(values, dem, gem, dem_mv, gem_mv, formula='p+gem-dem*0.0065'):
    """correction.py:46,59-82: guarded formula on the target grid, `mv` = netCDF f8 fill."""
    p = values
    # numexpr typing: the literal 0.0065 is a double, so a float32 dem (PCRaster REAL4 map) is promoted to
    # double before the product; numpy >= 2 would keep `dem*0.0065` in float32 (weak Python scalars)
    dem = np.asarray(dem, dtype=np.float64) if np.asarray(dem).dtype == np.float32 else dem
    mv = NC_FILL_F64 if ma.getdata(values).dtype == np.float64 else 9.969209968386869e36
    with np.errstate(over='ignore', invalid='ignore'):
        res = _eval(f'where((dem!=dem_mv) & (p!=mv) & (gem!=gem_mv), {formula}, mv)',
                    p=p, dem=dem, gem=gem, dem_mv=dem_mv, gem_mv=gem_mv, mv=mv)
    return ma.masked_where(_masks(p), res)
