"""Host logic of the aggregation stage on the CPU: the planners of pyg2p_b200.manipulation turn an
`Aggregator` configuration into g2p_manip items; here the items are evaluated by a few lines of numpy (the
arithmetic g2p_manip.cu documents for each item kind) and compared, bit for bit, with the vectors the reference's own
`Aggregator.do_manipulation` produced (tests/golden/manip.npz).  No GPU, no library call."""
import os

import numpy as np
import pytest

from pyg2p_b200 import _lib as L
from pyg2p_b200 import manipulation as M
from pyg2p_b200 import synth


def _run_plan(plan, fields):
    vals = [np.asarray(f, dtype=np.float64) for f in fields]
    for _, a, b, frac in plan.synth:                      # g2p_time_interp: V[a] + (V[b] - V[a]) * frac
        va, vb = vals[plan.slot(a)], vals[plan.slot(b)]
        vals.append(va + (vb - va) * frac)
    outs = []
    for key, it in plan.outputs:
        ins = it.inputs
        zero = np.zeros_like(vals[0])
        if it.kind == L.AGG_PICK:
            res = 0.0 + vals[ins[0]] if ins[0] >= 0 else zero
        elif it.kind == L.AGG_ACCUM:
            prev = vals[ins[1]] if ins[1] >= 0 else 0.0
            res = ((0.0 + vals[ins[0]]) - prev) * it.unit_time / it.divisor
        elif it.kind == L.AGG_AVERAGE:
            s = zero.copy()
            for i in ins:
                s = s + vals[i]
            res = s / it.divisor
        elif it.kind == L.AGG_WAVERAGE:
            s = zero.copy()
            for q, i in enumerate(ins):
                edge = (q == 0 and it.edge_mask & 1) or (q == len(ins) - 1 and it.edge_mask & 2)
                s = s + vals[i] * (it.w_edge if edge else it.unit_time)
            with np.errstate(all='ignore'):
                res = s / np.float64(it.divisor)
        else:
            raise AssertionError(it.kind)
        outs.append((key, res))
    return outs


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'manip.npz'))


def _check(g, tag, plan, fields):
    outs = sorted(_run_plan(plan, fields), key=lambda kv: (kv[0][1], kv[0][0]))
    np.testing.assert_array_equal(np.array([k for k, _ in outs]), g[f'{tag}_keys'])
    np.testing.assert_array_equal(np.array([v for _, v in outs]), g[f'{tag}_vals'])


def test_planners_reproduce_the_reference_aggregator(g):
    steps = [int(s) for s in g['agg_steps']]
    cum, inst = g['agg_cum'], g['inst_fields']

    def agg(**kw):
        base = dict(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6, step_type='accum',
                    start_step=0, end_step=72, unit_time=24, mv_grib=synth.GRIB_MV, force_zero_array=False)
        base.update(kw)
        return M.Aggregator(**base)

    def run(tag, a, stepset, src):
        plan = a.plan(list(stepset))
        _check(g, tag, plan, [src[steps.index(s)] for s in plan.slots[:len(stepset)]])

    run('acc_full', agg(), steps, cum)
    run('acc_nozero', agg(), steps[1:], cum)
    run('acc_gap', agg(), [s for s in steps if s != 48], cum)
    run('acc_forcezero', agg(force_zero_array=True), steps, cum)
    run('acc_step12_unit6', agg(aggr_step=12, unit_time=6, start_step=12), steps, cum)
    iv = dict(aggr_type='average', step_type='instant')
    run('avg_full', agg(**iv), steps, inst)
    run('avg_start24', agg(start_step=24, **iv), steps, inst)
    ih = dict(aggr_halfweights=True, **iv)            # reference tests/test_aggregator.py:57-77
    run('avgh_full', agg(**ih), steps, inst)
    run('avgh_start24', agg(start_step=24, **ih), steps, inst)
    run('avgh_gap', agg(**ih), [s for s in steps if s not in (24, 54)], inst)
    run('avgh_step12', agg(aggr_step=12, start_step=12, **ih), steps, inst)
    ii = dict(aggr_type='instantaneous', step_type='instant', aggr_step=12)
    run('inst_full', agg(**ii), steps, inst)
    run('inst_gap', agg(**ii), [s for s in steps if s not in (0, 36)], inst)


def test_halfweight_plan_structure():
    """steps every 6 h, 24 h windows: 5 steps per window, dt = 6, both ends at half weight, divisor 24."""
    plan = M.plan_average_halfweights(list(range(0, 49, 6)), 0, 48, 24)
    assert [k for k, _ in plan.outputs] == [(0, 24), (24, 48)]
    for (lo, hi), it in plan.outputs:
        assert it.kind == L.AGG_WAVERAGE and [plan.slots[i] for i in it.inputs] == list(range(lo, hi + 1, 6))
        assert (it.unit_time, it.w_edge, it.edge_mask, it.divisor, it.mask_all) == (6.0, 3.0, 3, 24.0, -1)
    # a window whose last hour is absent: only the first input is an edge; one with a single step adds nothing
    plan = M.plan_average_halfweights([0, 6, 12, 18, 30], 0, 24, 24)
    it = plan.outputs[0][1]
    assert (it.edge_mask, it.unit_time, it.w_edge) == (1, 8.0, 4.0) and it.divisor == 4.0 + 8.0 * 3
    lone = M.plan_average_halfweights([0, 30], 0, 24, 24).outputs[0][1]
    assert lone.inputs == [] and lone.divisor == 0.0


def test_formula_parsers():
    assert M.parse_function('x=x*1000') and M.parse_function('x=x-273.15')
    assert M.Corrector._LINEAR.match('p+gem-dem*0.0065').group(1) == '0.0065'
    prs = M.Corrector._PRESSURE.match('p/gem*(10**((-0.159)*dem/1000))')
    assert (prs.group(1), prs.group(2)) == ('-0.159', '1000')
    gp = M.Corrector._GEM_PRESSURE.match('(10**((-0.159)*(z/9.81)/1000))')
    assert (gp.group(1), gp.group(2), gp.group(3)) == ('-0.159', '9.81', '1000')
    assert M.Corrector._PRESSURE.match('p*gem') is None
