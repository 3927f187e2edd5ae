#!/usr/bin/env python
"""Regional source -> global target: most target rows lie beyond min_upper_bound.  invdist rows end as soon as the
nearest source point is known to be farther than the bound; `nearest` stores the true distance of every row (reference
quirk C.4) and needs the wide search.  Times both and checks a sample against the oracle."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import interp_ref as R  # noqa: E402
from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402

lat, lon, det = synth.regular_ll_grid(70.0, 30.0, 401, -20.0, 40.0, 601)        # a COSMO-like limited-area grid
res = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
tl, tlo = synth.latlon_target(res)
ix = device.SourceIndex(lon, lat, det['radius'])
mub = ix.min_upper_bound(det['Nj'])
d_lo, d_la = torch.from_numpy(tlo.ravel()).cuda(), torch.from_numpy(tl.ravel()).cuda()
for mode, k, name in ((L.MODE_INVDIST, 4, 'invdist k=4'), (L.MODE_NEAREST, 1, 'nearest')):
    ix.knn(d_lo, d_la, k=k, mode=mode, mub=mub, want_flags=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ix.knn(d_lo, d_la, k=k, mode=mode, mub=mub, want_flags=False)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    inside = float((out['idx'][0] < lat.size).float().mean())
    print(f'{name:12s} {lat.size} sources -> {tl.size} targets ({inside:.1%} inside): {ms:9.2f} ms', flush=True)
    if mode == L.MODE_INVDIST:
        o = R.BuildOracle(lon, lat, det, np.zeros(lat.size), 4, synth.NC_FILL_F64, synth.GRIB_MV, parallel=True, mode='invdist')
        sel = np.random.default_rng(0).choice(tl.size, 200000, replace=False)
        _, w, idx = o.interpolate(tlo.ravel()[sel].reshape(1, -1), tl.ravel()[sel].reshape(1, -1))
        gi = out['idx'].cpu().numpy().T[sel]
        gw = out['coef'].cpu().numpy().T[sel]
        same = (gi == idx).all(axis=1)
        print('   sample of 200000 rows vs oracle: identical index rows', same.mean(), 'outside rows agree',
              np.array_equal(np.isnan(gw[:, 0]), np.isnan(w[:, 0])), flush=True)
