#!/usr/bin/env python
"""Windowed apply kernel, O1280 -> EFAS invdist k=4, 383 messages: every mask policy timed and checked against the
byte-mask PROPAGATE result of the generic kernel.
    python tools/mask_variants.py [messages]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402

PEAK = 6553.0
b = int(sys.argv[1]) if len(sys.argv) > 1 else 383
lat, lon, det = synth.octahedral_grid(1280)
n = lat.size
tl, tlo = synth.efas_target()
m, k = tl.size, 4
ix = device.SourceIndex(lon, lat, det['radius'])
out = ix.knn(tlo, tl, k=k, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
n_touched = int(torch.unique(out['idx'][out['idx'] < n]).numel())
tbl = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
print('plan', tbl.plan_info(), flush=True)


def timed(fn, reps=10):
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


g = torch.Generator(device='cuda').manual_seed(1)
src = torch.randn((b, n), device='cuda', dtype=torch.float64, generator=g)
mask1 = torch.from_numpy(synth.missing_mask(lat).view(np.uint8)).cuda()
msk = mask1.unsqueeze(0).repeat(b, 1).contiguous()
msk[1::2] |= (torch.rand((b // 2, n), device='cuda', generator=g) < 0.01).to(torch.uint8)
bits = device.pack_mask_bits(msk, n)
marked = src.clone()
device.mark_masked(marked, msk)
o = torch.empty((b, m), device='cuda', dtype=torch.float64)
om = torch.empty((b, m), device='cuda', dtype=torch.uint8)
ob = torch.empty((b, device.bits_row_bytes(m)), device='cuda', dtype=torch.uint8)
os.environ['G2P_APPLY_PATH'] = 'generic'
ro, rm = torch.empty_like(o), torch.empty_like(om)
tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE, out=ro, out_mask=rm)
rbits = device.pack_mask_bits(rm, ob.shape[1] * 8)
os.environ['G2P_APPLY_PATH'] = 'window'
alg = lambda masks: 8 * b * (n_touched + m) + m * k * 12 + (b * ((n_touched + 7) // 8) if masks else 0)
cases = [('as_reference', lambda: tbl.apply(src, 1e20, out=o), False, None),
         ('propagate (byte masks)', lambda: tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE, out=o, out_mask=om), True, 'bytes'),
         ('fill (byte masks in)', lambda: tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_FILL, out=o), True, 'none'),
         ('propagate_bits', lambda: tbl.apply(src, 1e20, src_mask=bits, mask_policy=L.MASK_PROPAGATE_BITS, out=o, out_mask=ob), True, 'bits'),
         ('marked + bit out_mask', lambda: tbl.apply(marked, 1e20, mask_policy=L.MASK_MARKED, out=o, out_mask=ob), True, 'bits'),
         ('marked, no out_mask', lambda: tbl.apply(marked, 1e20, mask_policy=L.MASK_MARKED, out=o, out_mask=False), True, 'none')]
only = os.environ.get('G2P_ONLY')          # e.g. G2P_ONLY=propagate_bits for an ncu capture of one policy
for name, fn, masks, kind in cases:
    if only and not name.startswith(only):
        continue
    o.zero_(); om.zero_(); ob.fill_(255)
    ms = timed(fn)
    assert tbl.last_path == 'window', tbl.plan_error
    same = ''
    if masks:
        ok = bool((o.view(torch.int64) == ro.view(torch.int64)).all())
        if kind == 'bytes':
            ok = ok and bool((om == rm).all())
        if kind == 'bits':
            ok = ok and bool((ob == rbits).all())
        same = f'  bit-equal to generic propagate: {ok}'
    print(f'{name:28s} {ms:7.4f} ms  {alg(masks) / ms / 1e6:7.0f} GB/s  {alg(masks) / ms / 1e6 / PEAK:6.1%}{same}', flush=True)
