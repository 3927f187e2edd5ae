#!/usr/bin/env python
"""Windowed apply kernel: half-warp block shapes / skews / packed mask stores, timed on the BASELINE tables.
    python tools/win_variants.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402

PEAK = 6553.0
lat, lon, det = synth.octahedral_grid(1280)
n = lat.size
ix = device.SourceIndex(lon, lat, det['radius'])
mub = ix.min_upper_bound(det['Nj'])


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = [('O1280->EFAS invdist k=4', synth.efas_target(), 4, 383), ('O1280->0.05deg invdist k=4', synth.latlon_target(0.05), 4, 16)]
for name, (tl, tlo), k, b in cases:
    out = ix.knn(tlo, tl, k=k, mode=L.MODE_INVDIST, mub=mub, want_flags=False)
    m = tl.size
    n_touched = int(torch.unique(out['idx'][out['idx'] < n]).numel())
    src = torch.randn((b, n), device='cuda', dtype=torch.float64)
    msk = (torch.rand((b, n), device='cuda') < 0.02).to(torch.uint8)
    o = torch.empty((b, m), device='cuda', dtype=torch.float64)
    om = torch.empty((b, m), device='cuda', dtype=torch.uint8)
    ref = None
    print(f'## {name}: m={m} n_touched={n_touched} b={b}', flush=True)
    for rows in ('',):
        for skew in ('',):
            for byte_masks in ('',):
                for key, val in (('G2P_WIN_BLOCK_ROWS', rows), ('G2P_WIN_SKEW_MOD', skew), ('G2P_WIN_BYTE_MASKS', byte_masks)):
                    if val:
                        os.environ[key] = val
                    else:
                        os.environ.pop(key, None)
                os.environ['G2P_APPLY_PATH'] = 'window'
                tbl = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
                pi = tbl.plan_info()
                line = (f'  block_rows={rows or "auto"} skew={skew or "auto"} byte_masks={byte_masks or 0} plan={pi["tile_rows"]}x{pi["tile_cols"]} '
                        f'win {pi["mean_window"]:.0f} block {pi["block_rows"]}x{pi["block_cols"]} waves {pi["gather_wavefronts"]:.3f}:')
                for policy, pname in ((L.MASK_AS_REFERENCE, 'nomask'), (L.MASK_PROPAGATE, 'propagate')):
                    if byte_masks and not policy:
                        continue
                    bytes_alg = 8 * b * (n_touched + m) + m * k * 12 + (b * n_touched // 8 if policy else 0)
                    ms = timed(lambda: tbl.apply(src, 1e20, src_mask=msk if policy else None, mask_policy=policy, out=o,
                                                 out_mask=om if policy else None), 10)
                    assert tbl.last_path == 'window'
                    line += f'  {pname} {ms:7.3f} ms ({bytes_alg / ms / 1e6 / PEAK:5.1%})'
                    if policy:
                        if ref is None:
                            os.environ['G2P_APPLY_PATH'] = 'generic'
                            ro, rm = torch.empty_like(o), torch.empty_like(om)
                            tbl.apply(src, 1e20, src_mask=msk, mask_policy=policy, out=ro, out_mask=rm)
                            ref = (ro, rm)
                            os.environ['G2P_APPLY_PATH'] = 'window'
                        same = bool((o.view(torch.int64) == ref[0].view(torch.int64)).all()) and bool((om == ref[1]).all())
                        line += f' bit-equal to generic: {same}'
                print(line, flush=True)
                del tbl
    del src, msk, o, om, ref
