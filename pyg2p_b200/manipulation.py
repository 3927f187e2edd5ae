"""Host mirror of the reference's source-grid manipulation stages - `Converter`
(main/manipulation/conversion.py), `Aggregator` (main/manipulation/aggregator.py), `Corrector`
(main/manipulation/correction.py) - as planners for the fused CUDA stages (g2p_manip,
g2p_apply_planned_corrected).  The classes keep the reference's constructor arguments and method names;
the arithmetic runs on the GPU in the reference's evaluation order.  No CPU fallback.
"""
import ast
import bisect
import ctypes as C
import re

import numpy as np
import numpy.ma as ma
import torch

from . import _lib as L
from . import device
from .exceptions import NOT_IMPLEMENTED, ApplicationException

AVERAGE, ACCUMULATION, INSTANTANEOUS = 'average', 'accumulation', 'instantaneous'
PARAM_INSTANT, PARAM_AVG, PARAM_CUM = 'instant', 'avg', 'accum'
NC_FILL_F64 = 9.969209968386869e36


class Step:
    """hashable message key, as the reference's pyg2p.Step (src/pyg2p/__init__.py:30-54)."""

    def __init__(self, start_step, end_step, points_meridian, input_step, level):
        self.start_step, self.end_step = start_step, end_step
        self.resolution, self.input_step, self.level = points_meridian, input_step, level

    def _key(self):
        return self.start_step, self.end_step, self.resolution, self.input_step, self.level

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return self._key() == other._key()

    def __lt__(self, other):
        return self.start_step < other.start_step

    def __le__(self, other):
        return self.start_step <= other.start_step

    def __repr__(self):
        return (f's:{self.start_step} e:{self.end_step} res:{self.resolution} step-lenght:{self.input_step} '
                f'level:{self.level}')


class Messages:
    """The decoded messages of one GRIB parameter as the reference's readers hand them to the Controller:
    `pyg2p.Messages` (src/pyg2p/__init__.py:146-209) - `{Step: values}` for the first (or single) resolution and
    optionally the second, with their missing value, unit, step type and grid details.  `apply_conversion` is the
    Controller-side seam of the unit conversion (controller.py:106-110): here all messages go through one kernel
    launch per resolution instead of one numexpr call each."""

    def __init__(self, values, mv, unit, type_of_level, type_of_step, step_units, grid_details, val_2nd=None,
                 data_date=None, data_time='0'):
        from datetime import datetime
        self.values_first_or_single_res = values
        self.values_second_res = val_2nd or {}
        self.step_type = type_of_step
        self.step_units = step_units
        self.type_of_level = type_of_level
        self.unit = unit
        self.missing_value = mv
        self.data_date = datetime.strptime(f'{data_date}{data_time}', '%Y%m%d%H') if data_date is not None else None
        self.grid_details = grid_details
        # order key list to get first step
        self.first_step_range = sorted(self.values_first_or_single_res.keys(), key=lambda k: (int(k.end_step)))[0]

    def append_2nd_res_messages(self, messages):
        self.grid_details.set_2nd_resolution(messages.grid_details, messages.first_step_range)
        self.values_second_res = messages.first_resolution_values()

    def first_resolution_values(self):
        return self.values_first_or_single_res

    def second_resolution_values(self):
        return self.values_second_res

    @property
    def grid_id(self):
        return self.grid_details.grid_id

    @property
    def grid2_id(self):
        return self.grid_details.get_2nd_resolution().grid_id

    @property
    def latlons(self):
        return self.grid_details.latlons

    @property
    def latlons_2nd(self):
        second = self.grid_details.get_2nd_resolution()
        return second.latlons if second else (None, None)

    def have_resolution_change(self):
        return self.grid_details.get_2nd_resolution() is not None

    def change_resolution_step(self):
        return self.grid_details.get_change_res_step()

    def apply_conversion(self, converter):
        converter.set_unit_to_convert(self.unit)
        converter.set_missing_value(self.missing_value)
        self.values_first_or_single_res = converter.convert_messages(self.values_first_or_single_res)
        self.values_second_res = converter.convert_messages(self.values_second_res)

    def __len__(self):
        return len(self.values_first_or_single_res) + len(self.values_second_res)


# --------------------------------------------------------------------------- conversion functions
def parse_function(function_str, var='x'):
    """'x=x*1000' / 'x=x-273.15' / 'x=-x*1000' / '(z/9.81)*0.0065' ... -> [(op, constant), ...] (<= 2 steps).

    Only chains of `var (*|/|+|-) constant` are accepted - every function of the reference's
    parameters.json and the gem formulas of its templates have this shape.  Each step is exact:
    x-c == x+(-c), (-x)*c == x*(-c), c*x == x*c in IEEE arithmetic."""
    expr = function_str.split('=', 1)[1] if '=' in function_str else function_str
    tree = ast.parse(expr.strip(), mode='eval').body

    def const(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = const(node.operand)
            return None if v is None else (-v if isinstance(node.op, ast.USub) else v)
        return None

    def walk(node):
        """-> (ops, negated) for a sub-expression that contains the variable"""
        if isinstance(node, ast.Name) and node.id == var:
            return [], False
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            ops, neg = walk(node.operand)
            return ops, not neg
        if isinstance(node, ast.BinOp):
            lc, rc = const(node.left), const(node.right)
            if rc is not None and lc is None:
                ops, neg = walk(node.left)
                if isinstance(node.op, ast.Mult):
                    return ops + [(L.OP_MUL, -rc if neg else rc)], False
                if isinstance(node.op, ast.Div):
                    return ops + [(L.OP_DIV, -rc if neg else rc)], False
                if not neg and isinstance(node.op, ast.Add):
                    return ops + [(L.OP_ADD, rc)], False
                if not neg and isinstance(node.op, ast.Sub):
                    return ops + [(L.OP_ADD, -rc)], False
            if lc is not None and rc is None:
                ops, neg = walk(node.right)
                if isinstance(node.op, ast.Mult):
                    return ops + [(L.OP_MUL, -lc if neg else lc)], False
                if not neg and isinstance(node.op, ast.Add):
                    return ops + [(L.OP_ADD, lc)], False
        raise ApplicationException.get_exc(NOT_IMPLEMENTED, f'function {function_str!r} is not a chain of '
                                                             f'{var} (*|/|+|-) constant steps')

    ops, neg = walk(tree)
    if neg:
        ops = ops + [(L.OP_MUL, -1.0)]
    if len(ops) > 2:
        raise ApplicationException.get_exc(NOT_IMPLEMENTED, f'function {function_str!r} has more than two steps')
    return ops


class Converter:
    """`Converter(func='x=x*1000', cut_off=False)`; `.set_missing_value(mv)`; `.convert(x)`;
    `.cut_off_negative(values)` (conversion.py:10-59)."""

    def __init__(self, func=None, cut_off=False):
        self._mv = -1
        self._initial_unit = None
        self._function_str = func
        self._do_cut_off = cut_off
        self.ops = [] if func in (None, 'x=x') else parse_function(func)
        self.identity = func in (None, 'x=x')

    def set_unit_to_convert(self, unit):
        self._initial_unit = unit

    def set_missing_value(self, mv):
        self._mv = mv

    @property
    def must_cut_off(self):
        return self._do_cut_off

    def __str__(self):
        return ('\nConverting values from units {}. \nFunction: {}\nMissing value: {:.2f}'
                .format(getattr(self, '_initial_unit', None), self._function_str, self._mv))

    def convert_messages(self, values):
        """`{key: converter.convert(v)}` for all messages of a dict in ONE launch (what Messages.apply_conversion
        loops over, src/pyg2p/__init__.py:205-206); masks preserved per message."""
        if self.identity or not values:
            return dict(values)
        keys = list(values)
        outs = run_manip([values[k] for k in keys], [ManipItem.pick(i) for i in range(len(keys))], converter=self)
        return {k: (ma.masked_where(values[k].mask, o, copy=False) if isinstance(values[k], ma.MaskedArray) else o)
                for k, o in zip(keys, outs)}

    def desc_fields(self):
        ops = self.ops + [(L.OP_NONE, 0.0)] * (2 - len(self.ops))
        return dict(conv_op1=ops[0][0], conv_c1=ops[0][1], conv_op2=ops[1][0], conv_c2=ops[1][1], mv_grib=float(self._mv))

    def convert(self, x):
        """where(x != mv, f(x), mv), mask preserved (conversion.py:33-42) - on the GPU."""
        if self.identity:
            return x
        out = run_manip([x], [ManipItem.pick(0)], converter=self)[0]
        return ma.masked_where(x.mask, out, copy=False) if isinstance(x, ma.MaskedArray) else out

    def cut_off_negative(self, values):
        """where(v < 0, 0, v) on every message of a dict (or on one array), mask dropped (conversion.py:44-52)."""
        if isinstance(values, dict):
            keys = list(values)
            outs = run_manip([values[k] for k in keys], [ManipItem.pick(i) for i in range(len(keys))], cut_off=True)
            return dict(zip(keys, outs))
        return run_manip([values], [ManipItem.pick(0)], cut_off=True)[0]


# --------------------------------------------------------------------------- aggregation plans
class ManipItem:
    """one output message of g2p_manip."""

    def __init__(self, kind, inputs, unit_time=0.0, divisor=1.0, mask_all=0, w_edge=0.0, edge_mask=0):
        self.kind, self.inputs, self.unit_time, self.divisor, self.mask_all = kind, list(inputs), unit_time, divisor, mask_all
        self.w_edge, self.edge_mask = w_edge, edge_mask

    @classmethod
    def pick(cls, i):
        return cls(L.AGG_PICK, [i])

    def to_c(self):
        if len(self.inputs) > L.MANIP_MAX_IN:
            raise ApplicationException.get_exc(NOT_IMPLEMENTED, f'average over {len(self.inputs)} iterator steps '
                                                                f'(max {L.MANIP_MAX_IN})')
        it = L.ManipItem()
        it.kind, it.n_in, it.mask_all = self.kind, len(self.inputs), self.mask_all
        for q, v in enumerate(self.inputs):
            it.inputs[q] = v
        it.unit_time, it.divisor = float(self.unit_time), float(self.divisor)
        it.w_edge, it.edge_mask = float(self.w_edge), int(self.edge_mask)
        return it


class AggregationPlan:
    """What `Aggregator.do_manipulation` would compute, as data: the input slots (end_step of the messages,
    in slot order), the steps to synthesise first - (new_step, step_a, step_b, fraction) -> slot appended
    at the end - and one ManipItem per output key (start, end)."""

    def __init__(self, slots):
        self.slots = list(slots)
        self.synth = []
        self.outputs = []        # [((start, end), ManipItem)]

    def slot(self, step):
        return self.slots.index(step)

    def add_synth(self, step, a, b, frac):
        self.synth.append((step, a, b, frac))
        self.slots.append(step)


def plan_accumulation(steps, start, end, aggr_step, unit_time, force_zero=False):
    """aggregator.py:72-153, including its quirks: step 0 absent -> zeros, a missing step t is
    V[a] + (V[b]-V[a])*((t-a)/(b-a)), a missing PREVIOUS step t-D is interpolated with the fraction of t."""
    keys = sorted(steps)
    plan = AggregationPlan(keys)
    known = list(keys)            # v_ord keys; grows with the synthesised steps
    zero_created = False
    if start == 0:
        start = aggr_step
    for t in range(start, end + 1, aggr_step):
        vkeys = list(known)       # insertion order, like the reference's OrderedDict scan (synthesised steps at the end)
        if t not in vkeys:
            i_next = bisect.bisect_left(vkeys, t)
            nxt = vkeys[i_next]
            i_org = bisect.bisect_right(vkeys, t)
            if i_org == i_next:
                i_org -= 1
            org = vkeys[i_org]
            plan.add_synth(t, org, nxt, (t - org) / (nxt - org))
            known.append(t)
        prev = t - aggr_step
        if prev >= 0 and prev not in vkeys and not zero_created:
            if prev == 0:
                zero_created = True
                known.append(0)
            else:
                i_next = bisect.bisect_left(vkeys, prev)
                nxt = vkeys[i_next]
                i_org = bisect.bisect_right(vkeys, prev)
                if i_org == i_next:
                    i_org -= 1
                org = vkeys[i_org]
                plan.add_synth(prev, org, nxt, (t - org) / (nxt - org))
                known.append(prev)
        zero_prev = (prev == 0 and (force_zero or 0 not in plan.slots))
        item = ManipItem(L.AGG_ACCUM, [plan.slot(t), -1 if zero_prev else plan.slot(prev)], unit_time=unit_time,
                         divisor=aggr_step)
        plan.outputs.append(((prev, t), item))
    return plan


def plan_average(steps, start, end, aggr_step, second_t_res=False, mask_all=-1):
    """aggregator.py:155-236 without half weights: one add per hourly iterator, the next available step when an
    hour is absent, divided by the count.  `mask_all` = -1: unmasked, as the reference's averages are (its
    `get_masks(v_ord.values())`, :232, receives the dict view as one argument and returns None); 1: the OR of all
    inputs that line intends."""
    keys = sorted(steps)
    plan = AggregationPlan(keys)
    if start == 0 or second_t_res:
        it0, it1 = start, end - aggr_step + 2
    else:
        it0, it1 = start - aggr_step, end - aggr_step + 1
    for t in range(it0, it1, aggr_step):
        ins = []
        for h in range(t + 1, t + aggr_step + 1):
            ins.append(plan.slot(h if h in keys else keys[bisect.bisect_left(keys, h)]))
        plan.outputs.append(((t, t + aggr_step), ManipItem(L.AGG_AVERAGE, ins, divisor=float(len(ins)), mask_all=mask_all)))
    return plan


def plan_average_halfweights(steps, start, end, aggr_step, second_t_res=False, mask_all=-1):
    """aggregator.py:155-236 with `@halfweights` (:176-207): the window includes its first hour, only the steps
    PRESENT in the input are added (no next-available substitution), each weighted dt1 = aggr_step / (present - 1),
    the window's first and last hour dt1 / 2; divided by the sum of the weights; mask as in plan_average.
    Fewer than two present steps: nothing is added and the reference's 0/0 gives a NaN field."""
    keys = sorted(steps)
    plan = AggregationPlan(keys)
    if start == 0 or second_t_res:
        it0, it1 = start, end - aggr_step + 2
    else:
        it0, it1 = start - aggr_step, end - aggr_step + 1
    for t in range(it0, it1, aggr_step):
        iter_from, iter_to = t, t + aggr_step + 1
        present = [h for h in range(iter_from, iter_to) if h in keys]
        count_steps = len(present) - 1
        ins, dt1, dt2, edge, sum_counter = [], 0.0, 0.0, 0, 0.0
        if count_steps > 0:
            dt1 = aggr_step / count_steps
            dt2 = dt1 / 2
            for h in present:
                is_edge = h == iter_from or h == iter_to - 1
                sum_counter += dt2 if is_edge else dt1
                ins.append(plan.slot(h))
            edge = (1 if present[0] == iter_from else 0) | (2 if present[-1] == iter_to - 1 else 0)
        item = ManipItem(L.AGG_WAVERAGE, ins, unit_time=dt1, divisor=sum_counter, mask_all=mask_all, w_edge=dt2, edge_mask=edge)
        plan.outputs.append(((t, t + aggr_step), item))
    return plan


def plan_instantaneous(steps, start, end, aggr_step, input_step=0, step_type=PARAM_INSTANT):
    """aggregator.py:238-279."""
    keys = sorted(steps)
    plan = AggregationPlan(keys)
    s = start - aggr_step if start - aggr_step > 0 else start
    if step_type == PARAM_AVG:
        s = start + input_step
    for t in range(s, end + 1, aggr_step):
        if t in keys:
            src = plan.slot(t)
        elif t == 0:
            src = -1
        else:
            src = plan.slot(keys[bisect.bisect_right(keys, t)])
        plan.outputs.append(((t, t), ManipItem(L.AGG_PICK, [src])))
    return plan


class Aggregator:
    """Same keyword arguments as the reference's Aggregator (aggregator.py:26-60); `do_manipulation(values)`
    takes / returns `{Step: array}` and runs conversion-free aggregation on the GPU."""

    def __init__(self, **kwargs):
        self._mv_grib = kwargs.get('mv_grib')
        self._step_type = kwargs.get('step_type')
        self._input_step = int(kwargs.get('input_step'))
        self._unit_time = int(kwargs.get('unit_time'))
        self._aggregation_step = int(kwargs.get('aggr_step'))
        self._aggregation = kwargs.get('aggr_type')
        self._aggregation_halfweights = kwargs.get('aggr_halfweights')
        self._force_zero = kwargs.get('force_zero_array')
        self._start = int(kwargs.get('start_step'))
        self._end = int(kwargs.get('end_step'))
        self._second_t_res = kwargs.get('sec_temp_res')
        self._usable_start = max(self._start - self._aggregation_step, 0) if self._input_step != 0 else self._start

    def change_end_step(self, end_first_res):
        self._end = end_first_res

    def get_real_start_end_steps(self):
        return self._usable_start, self._end

    def plan(self, end_steps):
        if self._aggregation == ACCUMULATION:
            return plan_accumulation(end_steps, self._start, self._end, self._aggregation_step, self._unit_time,
                                     bool(self._force_zero))
        if self._step_type == PARAM_CUM:
            raise ApplicationException.get_exc(
                NOT_IMPLEMENTED, details=f'Manipulation {self._aggregation} for parameter type: {self._step_type}')
        if self._aggregation == AVERAGE:
            if self._aggregation_halfweights:
                return plan_average_halfweights(end_steps, self._start, self._end, self._aggregation_step,
                                                bool(self._second_t_res))
            return plan_average(end_steps, self._start, self._end, self._aggregation_step, bool(self._second_t_res))
        return plan_instantaneous(end_steps, self._start, self._end, self._aggregation_step, self._input_step,
                                  self._step_type)

    def do_manipulation(self, values):
        first = next(iter(values))
        by_end = dict(sorted((k.end_step, v) for k, v in values.items()))
        plan = self.plan(list(by_end))
        outs = run_plan(plan, [by_end[s] for s in plan.slots[:len(by_end)]])
        res = {}
        for ((s, e), _), o in zip(plan.outputs, outs):
            res[Step(s, e, first.resolution, self._aggregation_step, first.level)] = o
        return res


# --------------------------------------------------------------------------- running plans on the device
def _upload(fields):
    """list of [n] arrays (some MaskedArrays) -> (src [b, n] device, mask [b, n] uint8 device or None)."""
    device.require_cuda()
    n = np.asarray(ma.getdata(fields[0])).size
    src = torch.empty((len(fields), n), dtype=torch.float64, device='cuda')
    any_mask = any(isinstance(f, ma.MaskedArray) and f.mask is not ma.nomask for f in fields)
    msk = torch.zeros((len(fields), n), dtype=torch.uint8, device='cuda') if any_mask else None
    for i, f in enumerate(fields):
        src[i] = torch.from_numpy(np.ascontiguousarray(ma.getdata(f), dtype=np.float64).ravel())
        if any_mask and isinstance(f, ma.MaskedArray) and f.mask is not ma.nomask:
            msk[i] = torch.from_numpy(np.array(np.broadcast_to(ma.getmaskarray(f).ravel(), (n,))).view(np.uint8))
    return src, msk


class ItemArray(list):
    """a list of ManipItems with its ctypes array built once (a pass over 90 outputs rebuilds nothing per call)"""

    def c_array(self):
        if getattr(self, '_c', None) is None or self._c_len != len(self):
            self._c = (L.ManipItem * max(len(self), 1))(*[it.to_c() for it in self])
            self._c_len = len(self)
        return self._c


def manip_device(src, src_mask, items, converter=None, cut_off=False, touched=None, marked=False, src_ptrs=None,
                 scratch=None):
    """g2p_manip on device tensors: src [b, n] -> out [n_out, nc] (+ mask).  `touched`: int32 device tensor.
    `marked`: no mask array comes out, masked outputs hold G2P_MASK_MARKER instead (the input form of the apply
    kernels' MARKED policy).  `src_ptrs` = (values address, mask address or 0, b_in, n, row stride, device): raw
    addresses instead of tensors - PINNED HOST memory works, the kernel then reads the touched values over PCIe.
    `scratch`: a dict the caller keeps between calls - the output buffers (padding zeroed once) are then reused."""
    lib = device.require_cuda()
    if src_ptrs is not None:
        v_ptr, m_ptr, b_in, n, stride, dev = src_ptrs
    else:
        b_in, n = src.shape
        v_ptr, m_ptr, stride, dev = src.data_ptr(), (src_mask.data_ptr() if src_mask is not None else 0), src.stride(0), src.device
    has_mask = bool(m_ptr)
    nc = n if touched is None else touched.numel()
    n_out = len(items)
    carr = items.c_array() if isinstance(items, ItemArray) else (L.ManipItem * max(n_out, 1))(*[it.to_c() for it in items])
    desc = L.ManipDesc()
    fields = converter.desc_fields() if converter is not None and not converter.identity else dict(
        conv_op1=L.OP_NONE, conv_c1=0.0, conv_op2=L.OP_NONE, conv_c2=0.0, mv_grib=0.0)
    for k, v in fields.items():
        setattr(desc, k, v)
    desc.cut_off_negative = 1 if cut_off else 0
    desc.n_out = n_out
    desc.items = C.cast(carr, C.POINTER(L.ManipItem))
    nc_pad = -(-nc // 16) * 16                       # rows padded to 16 values: what the windowed apply needs
    want_omask = has_mask and not marked
    key = (n_out, nc_pad, want_omask, str(dev))
    if scratch is not None and key in scratch:
        out, omask = scratch[key]
    else:
        out = torch.empty((n_out, nc_pad), dtype=torch.float64, device=dev)
        omask = torch.empty((n_out, nc_pad), dtype=torch.uint8, device=dev) if want_omask else None
        if nc_pad != nc:
            out[:, nc:] = 0.0
            if omask is not None:
                omask[:, nc:] = 0
        if scratch is not None:
            scratch.clear()              # one shape at a time: a pass over other outputs replaces the buffers
            scratch[key] = (out, omask)
    L.check(lib.g2p_set_device(dev.index))
    L.check(lib.g2p_manip(C.byref(desc), C.c_void_p(v_ptr), C.c_void_p(m_ptr) if has_mask else None, n, b_in, stride,
                          C.c_void_p(touched.data_ptr()) if touched is not None else None, nc, out.stride(0),
                          C.c_void_p(out.data_ptr()), C.c_void_p(omask.data_ptr()) if omask is not None else None,
                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out[:, :nc], (omask[:, :nc] if omask is not None else None)


def synthesise_steps(plan, src, src_mask):
    """append the linearly-in-time interpolated messages a plan needs (g2p_time_interp); masks are OR-ed."""
    if not plan.synth:
        return src, src_mask
    lib = device.require_cuda()
    b0, n = src.shape
    full = torch.empty((b0 + len(plan.synth), n), dtype=torch.float64, device=src.device)
    full[:b0] = src
    fmask = None
    if src_mask is not None:
        fmask = torch.zeros((b0 + len(plan.synth), n), dtype=torch.uint8, device=src.device)
        fmask[:b0] = src_mask
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i, (step, a, b, frac) in enumerate(plan.synth):
        ia, ib = plan.slot(a), plan.slot(b)
        L.check(lib.g2p_time_interp(C.c_void_p(full[ia].data_ptr()), C.c_void_p(full[ib].data_ptr()), float(frac), n,
                                    C.c_void_p(full[b0 + i].data_ptr()), st))
        if fmask is not None:
            fmask[b0 + i] = fmask[ia] | fmask[ib]
    return full, fmask


def run_plan(plan, fields, converter=None, cut_off=False):
    src, msk = _upload(fields)
    src, msk = synthesise_steps(plan, src, msk)
    out, omask = manip_device(src, msk, [it for _, it in plan.outputs], converter=converter, cut_off=cut_off)
    return _download(out, omask, np.asarray(ma.getdata(fields[0])).shape)


def run_manip(fields, items, converter=None, cut_off=False):
    src, msk = _upload(fields)
    out, omask = manip_device(src, msk if not cut_off else None, items, converter=converter, cut_off=cut_off)
    return _download(out, omask, np.asarray(ma.getdata(fields[0])).shape)


def _download(out, omask, shape):
    o = out.cpu().numpy()
    res = [o[i].reshape(shape) for i in range(o.shape[0])]
    if omask is not None:
        m = omask.cpu().numpy().astype(bool)
        res = [ma.masked_where(m[i].reshape(shape), r, copy=False) if m[i].any() else r for i, r in enumerate(res)]
    return res


# --------------------------------------------------------------------------- correction
class Corrector:
    """DEM / geopotential lapse-rate correction (manipulation/correction.py:13-113) as per-target terms on the
    device.  `gem_values` is the geopotential field ALREADY interpolated to the target grid with the same
    intertable, after `gem_formula` was applied on the source grid (correction.py:84-108) - use
    `Corrector.from_geopotential` to do both steps."""

    # the two formula families of the reference's templates and README (README.md:958-978)
    _NUM = r'\(?\s*(-?[0-9.]+(?:[eE][-+]?[0-9]+)?)\s*\)?'
    _LINEAR = re.compile(r'^p\+gem-dem\*' + _NUM + r'$')
    _PRESSURE = re.compile(r'^p/gem\*\(10\*\*\(' + _NUM + r'\*dem/' + _NUM + r'\)\)$')
    _GEM_PRESSURE = re.compile(r'^\(?10\*\*\(' + _NUM + r'\*\(z/' + _NUM + r'\)/' + _NUM + r'\)\)?$')

    def __init__(self, dem_values, dem_mv, gem_values, gem_mv, formula='p+gem-dem*0.0065', dev=None):
        compact = formula.replace(' ', '')
        lib = device.require_cuda()
        dev = torch.device('cuda', torch.cuda.current_device()) if dev is None else torch.device(dev)   # the table's GPU
        dem = np.ascontiguousarray(dem_values)
        if dem.dtype not in (np.float32, np.float64):
            dem = dem.astype(np.float64)
        self.shape = dem.shape
        m = dem.size
        self.gem = device.to_device(np.asarray(ma.getdata(gem_values), dtype=np.float64).ravel(), torch.float64, dev)
        self.dem_term = torch.empty(m, dtype=torch.float64, device=dev)
        self.valid = torch.empty(m, dtype=torch.uint8, device=dev)
        lin, prs = self._LINEAR.match(compact), self._PRESSURE.match(compact)
        if lin:
            # `p+gem-dem*c`: dem_term = dem*c and the guards, per target, on the device
            self.mode, self.coeff = 0, float(lin.group(1))
            d_dem = torch.from_numpy(dem.ravel()).to(dev)
            L.check(lib.g2p_set_device(dev.index))
            L.check(lib.g2p_correction_prepare(C.c_void_p(d_dem.data_ptr()), L.F32 if dem.dtype == np.float32 else L.F64,
                                               float(dem_mv), C.c_void_p(self.gem.data_ptr()), float(gem_mv), self.coeff, m,
                                               C.c_void_p(self.dem_term.data_ptr()), C.c_void_p(self.valid.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        elif prs:
            # `p/gem*(10**((c)*dem/d))`: the per-target factor 10**((c*dem)/d) is a constant of the target grid; it is
            # evaluated once, here, with numpy's float64 power - the libm `pow` numexpr calls in the reference - so that
            # it carries the reference's bits; (p/gem)*factor runs per message in the kernel epilogue
            self.mode, self.coeff = 1, float(prs.group(1))
            dem64 = dem.ravel().astype(np.float64)          # numexpr promotes a float32 dem before `c*dem`
            with np.errstate(all='ignore'):
                term = np.power(10.0, (self.coeff * dem64) / float(prs.group(2)))
            gem_h = np.asarray(ma.getdata(gem_values), dtype=np.float64).ravel()
            valid = (dem64 != float(dem_mv)) & (gem_h != gem_mv)    # numexpr compares a float32 dem in double
            self.dem_term.copy_(torch.from_numpy(term))
            self.valid.copy_(torch.from_numpy(valid.astype(np.uint8)))
        else:
            raise ApplicationException.get_exc(NOT_IMPLEMENTED, f'correction formula {formula!r} (p+gem-dem*c and '
                                                                f'p/gem*(10**((c)*dem/d)) are built)')
        self.mv = NC_FILL_F64
        self._c = L.Correction(self.gem.data_ptr(), self.dem_term.data_ptr(), self.valid.data_ptr(), self.mv, self.mode)

    @classmethod
    def from_geopotential(cls, table, z_values, z_mv, mv_out, dem_values, dem_mv, formula='p+gem-dem*0.0065',
                          gem_formula='(z/9.81)*0.0065'):
        """`_read_geo` (correction.py:84-108): gem = interpolate(where(z != mv, gem_formula(z), mv))."""
        prs = cls._GEM_PRESSURE.match(gem_formula.replace(' ', ''))
        if prs:
            # `10**((c)*(z/g)/d)`: one field per run, evaluated on the host with the reference's libm `pow`
            # (see __init__), then interpolated on the device like any message
            c_, g_, d_ = (float(prs.group(i)) for i in (1, 2, 3))
            z = np.asarray(ma.getdata(z_values), dtype=np.float64).ravel()
            with np.errstate(all='ignore'):
                gz = np.where(z != z_mv, np.power(10.0, (c_ * (z / g_)) / d_), z_mv)
            g_src = torch.from_numpy(gz).cuda().unsqueeze(0)
        else:
            conv = Converter.__new__(Converter)
            conv._mv, conv._do_cut_off, conv.identity = z_mv, False, False
            conv.ops = parse_function(gem_formula, var='z')
            src, _ = _upload([z_values])
            g_src, _ = manip_device(src, None, [ManipItem.pick(0)], converter=conv)
        # `0.0 + x` of PICK is exact except for -0.0, which the formula cannot produce from a geopotential
        n_pad = -(-table.n_src // 16) * 16
        buf = torch.zeros((1, n_pad), dtype=torch.float64, device='cuda')
        buf[0, :table.n_src] = g_src[0]
        gem = table.apply(buf[:, :table.n_src], float(mv_out))[0]
        return cls(dem_values, dem_mv, gem.cpu().numpy(), z_mv, formula, dev=table.device)

    def c_struct(self):
        return self._c

    def correct(self, values):
        """standalone `Corrector.correct(values)` on a target-grid array (correction.py:59-82)."""
        lib = device.require_cuda()
        v = torch.from_numpy(np.ascontiguousarray(ma.getdata(values), dtype=np.float64).reshape(1, -1)).cuda()
        L.check(lib.g2p_correct(C.c_void_p(v.data_ptr()), None, v.shape[1], 1, v.stride(0), C.byref(self._c),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out = v.cpu().numpy().reshape(np.asarray(values).shape)
        return ma.masked_where(values.mask, out, copy=False) if isinstance(values, ma.MaskedArray) else out


# --------------------------------------------------------------------------- the fused pipeline
class FusedPipeline:
    """convert -> aggregate -> cut-off (source grid, compact) -> intertable apply -> correction, for all output
    messages of a run: two kernel launches (g2p_manip on the touched source points, then the windowed apply
    with the Corrector epilogue).  Stage order as the reference's controller / writers (controller.py:106-132,
    writers/__init__.py:55-70)."""

    def __init__(self, table):
        self.table = table
        self.compact, self.touched = table.compact()
        self._scratch = {}         # compact fields between the two kernels, reused from pass to pass

    def run(self, src, src_mask, plan, converter=None, cut_off=False, mv_out=NC_FILL_F64,
            mask_policy=L.MASK_AS_REFERENCE, corrector=None, members=1, fused=None):
        """src: device [b, n] float64 (messages in `plan.slots` order; with `members` > 1 the rows are
        member-major: all steps of member 0, then member 1, ... and the plan is applied to each member),
        src_mask: device uint8 [b, n] or None.  -> out [members * n_out, m] device (+ out_mask under PROPAGATE).

        Default (`fused=None` / False): two launches - g2p_manip evaluates conversion / aggregation / cut-off once
        per touched source point into compact fields, the windowed apply (with the Corrector epilogue) reads them.
        `fused=True`: ONE launch (g2p_apply_planned_fused), the source-grid stages run inside the staging step of
        the apply kernel; possible when every output combines at most two inputs (accumulation, instantaneous,
        1-2 step averages) and the buffers are aligned for the windowed kernel.  Both give identical bytes; the
        single launch is the slower one (measured 1.22 ms against 0.57 ms for 90 accumulation outputs of O1280 ->
        EFAS: inside the kernel every source value is aggregated once per tile window that contains it - 2.4x on
        average - with a float64 division each, and the two staged inputs halve the occupancy)."""
        cache = plan.__dict__.setdefault('_item_arrays', {})
        if members == 1:
            src, src_mask = synthesise_steps(plan, src, src_mask)
            items = cache.get(1) or cache.setdefault(1, ItemArray(it for _, it in plan.outputs))
        else:
            if plan.synth:
                raise ApplicationException.get_exc(NOT_IMPLEMENTED, 'synthesised steps with members > 1')
            if src_mask is not None and any(it.mask_all == 1 for _, it in plan.outputs):
                # mask_all = 1 ORs the masks of ALL input rows of the launch: with the members stacked in one launch one
                # member's mask would leak into the others
                raise ApplicationException.get_exc(NOT_IMPLEMENTED, 'mask_all averages with members > 1 (run the members one by one)')
            per = len(plan.slots)
            items = cache.get(members) or cache.setdefault(members, ItemArray(
                ManipItem(it.kind, [i + mb * per if i >= 0 else i for i in it.inputs], it.unit_time, it.divisor,
                          it.mask_all, it.w_edge, it.edge_mask) for mb in range(members) for _, it in plan.outputs))
        use_mask = src_mask is not None and mask_policy != L.MASK_AS_REFERENCE
        policy = mask_policy if use_mask else L.MASK_AS_REFERENCE
        can_fuse = self._can_fuse(src, src_mask if use_mask else None, items, use_mask)
        if fused is True and not can_fuse:
            raise ApplicationException.get_exc(NOT_IMPLEMENTED, 'this plan / these buffers cannot use the fused kernel')
        self.last_path = 'fused' if (can_fuse and fused is True) else 'two-kernel'
        if self.last_path == 'fused':
            return self._run_fused(src, src_mask if use_mask else None, items, converter, cut_off, mv_out, policy,
                                   corrector)
        if policy == L.MASK_MARKED:
            # masks in (bytes), markers between the two kernels, bit-packed mask out: the apply kernel stages no mask
            vals, _ = manip_device(src, src_mask, items, converter=converter, cut_off=cut_off, touched=self.touched,
                                   marked=True, scratch=self._scratch)
            return self.compact.apply(vals, float(mv_out), mask_policy=L.MASK_MARKED, correction=corrector)
        vals, vmask = manip_device(src, src_mask if use_mask else None, items,
                                   converter=converter, cut_off=cut_off, touched=self.touched, scratch=self._scratch)
        return self.compact.apply(vals, float(mv_out), src_mask=vmask, mask_policy=policy, correction=corrector)

    def run_host(self, h_src, h_mask, plan, converter=None, cut_off=False, mv_out=NC_FILL_F64, corrector=None,
                 members=1, out=None, out_mask=None, post=None):
        """The same chain from messages that lie in PINNED HOST memory, one member at a time: h_src [members * steps, n]
        float64 pinned tensor (rows member-major, `plan.slots` order), h_mask uint8 [., n] pinned or None ->
        (out [members * n_out, m] pinned float64, out_mask bit-packed pinned uint8 or None).
        Per member: g2p_manip reads exactly the touched values of the member's messages out of the host buffers over PCIe
        (zero-copy) and evaluates convert / aggregate / cut-off on them (masked values -> markers), the windowed apply with
        the Corrector / writer epilogue follows, and the results go back by D2H while the next member computes.  With
        `post` (a WriterPost) masked cells already hold the file's missing value and no mask is returned."""
        if plan.synth:
            raise ApplicationException.get_exc(NOT_IMPLEMENTED, 'synthesised steps in the host pipeline')
        t = self.table
        per, n_out = len(plan.slots), len(plan.outputs)
        if h_src.shape[0] != members * per or not h_src.is_pinned() or h_src.dtype != torch.float64:
            raise ValueError('h_src must be a pinned float64 tensor [members * steps, n]')
        if h_mask is not None and (not h_mask.is_pinned() or h_mask.stride(0) != h_src.stride(0)):
            raise ValueError('h_mask must be a pinned uint8 tensor with the row stride of h_src')
        cache = plan.__dict__.setdefault('_item_arrays', {})
        items = cache.get(1) or cache.setdefault(1, ItemArray(it for _, it in plan.outputs))
        want_mask = h_mask is not None and post is None
        o_dtype = torch.float32 if (post is not None and post.out_f32) else torch.float64
        if out is None:
            out = device.pinned_empty((members * n_out, t.m), o_dtype)
        brow = device.bits_row_bytes(t.m)
        if want_mask and out_mask is None:
            out_mask = device.pinned_empty((members * n_out, brow), torch.uint8)
        dev = t.device
        cur = torch.cuda.current_stream(dev)
        if getattr(self, '_s_out', None) is None:
            self._s_out = torch.cuda.Stream(device=dev)
            self._slots = None
        if self._slots is None or self._slots[0][0].shape[0] != n_out or self._slots[0][0].dtype != o_dtype:
            self._slots = [(torch.empty((n_out, t.m), dtype=o_dtype, device=dev),
                            torch.empty((n_out, brow), dtype=torch.uint8, device=dev)) for _ in range(2)]
        ev_out = [None, None]
        self.h2d_bytes = self.d2h_bytes = 0
        nc = self.compact.n_src
        with torch.cuda.device(dev):
            for mb in range(members):
                s = mb % 2
                d_out, d_bits = self._slots[s]
                if ev_out[s] is not None:
                    cur.wait_event(ev_out[s])
                ptrs = (h_src.data_ptr() + mb * per * h_src.stride(0) * 8,
                        (h_mask.data_ptr() + mb * per * h_src.stride(0)) if h_mask is not None else 0,
                        per, t.n_src, h_src.stride(0), dev)
                vals, _ = manip_device(None, None, items, converter=converter, cut_off=cut_off, touched=self.touched,
                                       marked=True, src_ptrs=ptrs, scratch=self._scratch)
                # every output reads its inputs' touched values (an accumulation: two steps), masks as bytes
                self.h2d_bytes += sum(len([i for i in it.inputs if i >= 0]) for it in items) * nc * (9 if h_mask is not None else 8)
                if h_mask is not None:
                    self.compact.apply(vals, float(mv_out), mask_policy=L.MASK_MARKED, out=d_out,
                                       out_mask=d_bits if want_mask else False, correction=corrector, post=post)
                else:
                    self.compact.apply(vals, float(mv_out), out=d_out, correction=corrector, post=post)
                ev_k = torch.cuda.Event()
                ev_k.record(cur)
                with torch.cuda.stream(self._s_out):
                    self._s_out.wait_event(ev_k)
                    out[mb * n_out:(mb + 1) * n_out].copy_(d_out, non_blocking=True)
                    self.d2h_bytes += n_out * t.m * (4 if o_dtype == torch.float32 else 8)
                    if want_mask:
                        out_mask[mb * n_out:(mb + 1) * n_out].copy_(d_bits, non_blocking=True)
                        self.d2h_bytes += n_out * brow
                    ev_out[s] = torch.cuda.Event()
                    ev_out[s].record(self._s_out)
            for e in ev_out:
                if e is not None:
                    cur.wait_event(e)
        return out, (out_mask if want_mask else None)

    def _can_fuse(self, src, src_mask, items, use_mask):
        t = self.table
        if t.k not in (1, 4) or t.plan() is None or t.plan_info()['max_window'] > 2048:
            return False
        n_pad = -(-t.n_src // 16) * 16
        if src.data_ptr() % 16 or src.stride(0) % 2 or src.stride(0) < n_pad or src.stride(1) != 1:
            return False
        if use_mask and (src_mask.data_ptr() % 16 or src.stride(0) % 16 or src_mask.stride(0) != src.stride(0)):
            return False
        for it in items:
            if len(it.inputs) > 2 or it.kind == L.AGG_WAVERAGE or (use_mask and it.mask_all):
                return False
        return True

    def _run_fused(self, src, src_mask, items, converter, cut_off, mv_out, policy, corrector):
        lib = device.require_cuda()
        t = self.table
        n_out = len(items)
        carr = items.c_array() if isinstance(items, ItemArray) else (L.ManipItem * max(n_out, 1))(*[it.to_c() for it in items])
        desc = L.ManipDesc()
        fields = converter.desc_fields() if converter is not None and not converter.identity else dict(
            conv_op1=L.OP_NONE, conv_c1=0.0, conv_op2=L.OP_NONE, conv_c2=0.0, mv_grib=0.0)
        for k, v in fields.items():
            setattr(desc, k, v)
        desc.cut_off_negative = 1 if cut_off else 0
        desc.n_out = n_out
        desc.items = C.cast(carr, C.POINTER(L.ManipItem))
        out = torch.empty((n_out, t.m), dtype=torch.float64, device=src.device)
        omask = torch.empty((n_out, t.m), dtype=torch.uint8, device=src.device) if policy == L.MASK_PROPAGATE else None
        L.check(lib.g2p_set_device(src.device.index))
        L.check(lib.g2p_apply_planned_fused(
            t.plan(), C.c_void_p(t.coef.data_ptr()) if t.k > 1 else None, C.byref(desc), C.c_void_p(src.data_ptr()),
            C.c_void_p(src_mask.data_ptr()) if src_mask is not None else None, src.shape[0], src.stride(0), out.stride(0),
            float(mv_out), int(policy) | (L.SUM_PAIRWISE if t.sum_pairwise else 0),
            C.byref(corrector.c_struct()) if corrector is not None else None,
            C.c_void_p(out.data_ptr()), C.c_void_p(omask.data_ptr()) if omask is not None else None,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return (out, omask) if policy == L.MASK_PROPAGATE else out
