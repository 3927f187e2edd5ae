#!/usr/bin/env python
"""One intertable build of BASELINE configs[2] (O1280 -> GloFAS 0.05 deg, invdist k=4) or configs[1]
(O1280 -> 0.1 deg nearest), for `ncu` launch lists / captures of the build kernels.
    python tools/profile_build.py [invdist|nearest] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'invdist'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lat, lon, det = synth.octahedral_grid(1280)
tl, tlo = synth.latlon_target(0.05 if mode == 'invdist' else 0.1)
k = 4 if mode == 'invdist' else 1
d_lon, d_lat = torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda()
d_tlo, d_tla = torch.from_numpy(tlo.ravel().copy()).cuda(), torch.from_numpy(tl.ravel().copy()).cuda()
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix = device.SourceIndex(d_lon, d_lat, det['radius'])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mub = ix.min_upper_bound(det['Nj'])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = ix.knn(d_tlo, d_tla, k=k, mode=L.MODE_INVDIST if mode == 'invdist' else L.MODE_NEAREST, mub=mub)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    best = min(best, t3 - t0) if r else t3 - t0
    print(f'rep {r}: index {1e3 * (t1 - t0):.2f} ms, self-query+mub {1e3 * (t2 - t1):.2f} ms, knn+weights '
          f'{1e3 * (t3 - t2):.2f} ms, total {1e3 * (t3 - t0):.2f} ms for {tl.size} targets k={k}', flush=True)
print(f'best total {1e3 * best:.2f} ms'); print(ix.info())
