#!/usr/bin/env python
"""Apply-kernel timings over the BASELINE table shapes and batch sizes (windowed vs generic kernel).
    python tools/apply_sweep.py"""
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
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = [('O1280->EFAS invdist k=4', synth.efas_target(), 4), ('O1280->0.1deg nearest k=1', synth.latlon_target(0.1), 1),
         ('O1280->0.05deg invdist k=4', synth.latlon_target(0.05), 4)]
for name, (tl, tlo), k in cases:
    out = ix.knn(tlo, tl, k=k, mode=L.MODE_NEAREST if k == 1 else L.MODE_INVDIST, mub=mub, want_flags=False)
    tbl = device.DeviceTable(out['idx'], out['coef'], n, grid_shape=tl.shape)
    m = tl.size
    n_touched = int(torch.unique(tbl.idx[tbl.idx < n]).numel())
    print(f'## {name}: m={m} n_touched={n_touched} plan={tbl.plan_info()}', flush=True)
    for b in (1, 4, 16, 64):
        if b * (n + m) * 8 > 60e9:
            continue
        src = torch.randn((b, n), device='cuda', dtype=torch.float64)
        msk = (torch.rand((b, n), device='cuda') < 0.02).to(torch.uint8)
        o = torch.empty((b, m), device='cuda', dtype=torch.float64)
        om = torch.empty((b, m), device='cuda', dtype=torch.uint8)
        for policy, pname in ((L.MASK_AS_REFERENCE, 'nomask'), (L.MASK_PROPAGATE, 'propagate')):
            bytes_alg = 8 * b * (n_touched + m) + m * k * (12 if k > 1 else 4) + (b * n_touched // 8 if policy else 0)
            res = {}
            for path in ('window', 'generic'):
                os.environ['G2P_APPLY_PATH'] = path
                ms = timed(lambda: tbl.apply(src, 1e20, src_mask=msk if policy else None, mask_policy=policy, out=o,
                                             out_mask=om if policy else None), 3 if b >= 16 else 10)
                assert tbl.last_path == path
                res[path] = ms
            os.environ['G2P_APPLY_PATH'] = 'auto'
            print(f'  b={b:3d} {pname:9s} window {res["window"]:8.3f} ms ({bytes_alg / res["window"] / 1e6 / PEAK:5.1%} of peak)'
                  f'   generic {res["generic"]:8.3f} ms ({bytes_alg / res["generic"] / 1e6 / PEAK:5.1%})', flush=True)
        del src, msk, o, om
