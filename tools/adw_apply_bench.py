#!/usr/bin/env python
"""Apply of a mode-adw table (k = 11, O320 -> EFAS): the batched generic kernel (table of a target pair in registers, read once
per chunk of messages) against the per-message any-k kernel (G2P_APPLY_ANYK=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, synth  # noqa: E402
from pyg2p_b200.scipy_interpolation import ScipyInterpolation  # noqa: E402

lat, lon, det = synth.octahedral_grid(320)
tl, tlo = synth.efas_target()
si = ScipyInterpolation(lon, lat, det, synth.field_2t(lat, lon), 11, synth.NC_FILL_F64, synth.GRIB_MV, mode='adw')
table = si.build_table(tlo, tl)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device='cuda').manual_seed(1)
src = torch.randn((b, lat.size), device='cuda', dtype=torch.float64, generator=g)
msk = (torch.rand((b, lat.size), device='cuda', generator=g) < 0.02).to(torch.uint8)
out = torch.empty((b, tl.size), device='cuda', dtype=torch.float64)
om = torch.empty((b, tl.size), device='cuda', dtype=torch.uint8)
n_touched = int(torch.unique(table.idx[table.idx < lat.size]).numel())
alg = 8 * b * (n_touched + tl.size) + tl.size * 11 * 12


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


ms = timed(lambda: table.apply(src, synth.NC_FILL_F64, out=out))
ref = out.clone()
print(f'k=11, {b} messages, no masks : {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic ({table.last_path})', flush=True)
ms = timed(lambda: table.apply(src, synth.NC_FILL_F64, src_mask=msk, mask_policy=L.MASK_PROPAGATE, out=out, out_mask=om))
print(f'k=11, {b} messages, byte masks: {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic', flush=True)
np.save(f'/tmp/adw_apply_{"anyk" if os.environ.get("G2P_APPLY_ANYK") else "batched"}.npy', ref[:2].cpu().numpy())
