#!/usr/bin/env python
"""g2p_host_pick alone (host-side staging of the ingest path): the 276 589 source points the O1280 -> EFAS table reads,
picked out of a pageable 52.8 MB field with its mask, for 1 / 2 / 4 / 8 threads."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402

lib = L.lib()
n, nc = 6599680, 276589
rng = np.random.default_rng(0)
start = int(rng.integers(0, n - nc * 2))
th = np.sort(rng.choice(np.arange(start, start + 2 * nc), nc, replace=False)).astype(np.int32)
fields = [rng.random(n) for _ in range(4)]
mv = (rng.random(n) < 0.02).view(np.uint8)
out, outm = np.empty(nc), np.empty(nc, np.uint8)
for thr in (1, 2, 4, 8, 1, 4):
    t0 = time.perf_counter()
    for r in range(400):
        x = fields[r % 4]
        lib.g2p_host_pick(x.ctypes.data, mv.ctypes.data, th.ctypes.data, nc, out.ctypes.data, outm.ctypes.data, 0, thr)
    ms = (time.perf_counter() - t0) / 400 * 1e3
    assert np.array_equal(out, x[th]) and np.array_equal(outm, mv[th])
    print(f'threads {thr}: {ms:.3f} ms per message', flush=True)
