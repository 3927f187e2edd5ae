#!/usr/bin/env python
"""g2p_manip alone on the configs[4] shape: 6 members x 61 O1280 steps -> 90 accumulation outputs on the 276 589 source points
the EFAS table touches; with and without masks (marker form)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402
from pyg2p_b200 import manipulation as MP  # noqa: E402

lat, lon, det = synth.octahedral_grid(1280)
n = lat.size
tl, tlo = synth.efas_target()
ix = device.SourceIndex(lon, lat, det['radius'])
out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
table = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
ct, touched = table.compact()
members, steps_h = 6, list(range(0, 361, 6))
conv = MP.Converter(func='x=x*1000', cut_off=True)
conv.set_missing_value(synth.GRIB_MV)
agg = MP.Aggregator(aggr_step=24, aggr_type='accumulation', aggr_halfweights=False, input_step=6, step_type='accum',
                    start_step=0, end_step=360, unit_time=24, mv_grib=synth.GRIB_MV, force_zero_array=False)
plan = agg.plan(steps_h)
per = len(plan.slots)
items = MP.ItemArray(MP.ManipItem(it.kind, [i + mb * per if i >= 0 else i for i in it.inputs], it.unit_time, it.divisor, it.mask_all,
                      it.w_edge, it.edge_mask) for mb in range(members) for _, it in plan.outputs)
g = torch.Generator(device='cuda').manual_seed(1)
src = torch.empty((members * per, n), dtype=torch.float64, device='cuda')
src.normal_(0.0, 1.0, generator=g).clamp_(min=0.0).mul_(2e-3)
msk = torch.from_numpy(synth.missing_mask(lat).view(np.uint8)).cuda().unsqueeze(0).repeat(src.shape[0], 1).contiguous()
nc = touched.numel()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


alg = len(items) * nc * 24
for name, fn in (('no masks', lambda: MP.manip_device(src, None, items, converter=conv, cut_off=True, touched=touched)),
                 ('masks -> markers', lambda: MP.manip_device(src, msk, items, converter=conv, cut_off=True, touched=touched, marked=True)),
                 ('masks -> byte masks', lambda: MP.manip_device(src, msk, items, converter=conv, cut_off=True, touched=touched))):
    ms = timed(fn)
    print(f'{name:22s} {ms:7.4f} ms  {alg / ms / 1e6:7.0f} GB/s algorithmic ({len(items)} outputs x {nc} points x 24 B)', flush=True)
vals, _ = MP.manip_device(src, None, items, converter=conv, cut_off=True, touched=touched)
ms = timed(lambda: ct.apply(vals, synth.NC_FILL_F64))
print(f'apply of the 90 compact fields {ms:7.4f} ms')
dem = synth.field_dem(tl, tlo, synth.PCR_MV_F32).astype(np.float32)
corr = MP.Corrector.from_geopotential(table, synth.field_geopotential(lat, lon), synth.GRIB_MV, synth.NC_FILL_F64, dem,
                                      float(synth.PCR_MV_F32))
ms = timed(lambda: ct.apply(vals, synth.NC_FILL_F64, correction=corr))
print(f'... with the Corrector epilogue   {ms:7.4f} ms')
