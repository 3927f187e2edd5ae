#!/usr/bin/env python
"""Windowed apply: items per CTA (message chunk length) against batch size, O1280 -> EFAS.
    python tools/chunk_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402

lat, lon, det = synth.octahedral_grid(1280)
n = lat.size
ix = device.SourceIndex(lon, lat, det['radius'])
tl, tlo = synth.efas_target()
out = ix.knn(tlo, tl, k=4, mode=L.MODE_INVDIST, mub=ix.min_upper_bound(det['Nj']), want_flags=False)
tbl = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
m = tl.size
nt = int(torch.unique(tbl.idx[tbl.idx < n]).numel())


def timed(fn, reps=10):
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


for b in (8, 16, 32, 64, 128, 383):
    src = torch.randn((b, n), device='cuda', dtype=torch.float64)
    msk = (torch.rand((b, n), device='cuda') < 0.02).to(torch.uint8)
    o = torch.empty((b, m), device='cuda', dtype=torch.float64)
    om = torch.empty((b, m), device='cuda', dtype=torch.uint8)
    line = f'b={b:4d}'
    for per in ('1', '2', '3', '4', '6', '8', '12'):
        os.environ['G2P_WIN_ITEMS_PER_CTA'] = per
        t0 = timed(lambda: tbl.apply(src, 1e20, out=o))
        t1 = timed(lambda: tbl.apply(src, 1e20, src_mask=msk, mask_policy=L.MASK_PROPAGATE, out=o, out_mask=om))
        alg = 8 * b * (nt + m) + m * 4 * 12
        line += f'  | {per}: {t0:.3f} ({alg / t0 / 1e6 / 6553:4.0%}) {t1:.3f} ({(alg + b * nt // 8) / t1 / 1e6 / 6553:4.0%})'
    print(line, flush=True)
    del src, msk, o, om
