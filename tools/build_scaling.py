#!/usr/bin/env python
"""Multi-GPU intertable build (BASELINE configs[2]: O1280 -> GloFAS 0.05 deg invdist k=4) under torchrun, the table
assembled on every GPU: times the stages and the assembly strategies of pyg2p_b200.parallel.TableAssembler.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/build_scaling.py [blocks ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyg2p_b200 import _lib as L  # noqa: E402
from pyg2p_b200 import device, parallel, synth  # noqa: E402

rank, lr, ws = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
if ws > 1:
    dist.init_process_group('nccl', device_id=dev)
lat, lon, det = synth.octahedral_grid(1280)
gl, glo = synth.latlon_target(0.05)
M, cols = gl.size, gl.shape[1]
lo, hi = parallel.shard_bounds(M, rank, ws, align=cols)
d_lon, d_lat = torch.from_numpy(lon).to(dev), torch.from_numpy(lat).to(dev)
d_tlo, d_tla = torch.from_numpy(glo.ravel()[lo:hi].copy()).to(dev), torch.from_numpy(gl.ravel()[lo:hi].copy()).to(dev)


def sync():
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()


def median_ms(fn, reps=7):
    for _ in range(3):
        fn()
    out = []
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        keep = fn()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
        del keep
    return parallel.max_over_ranks(float(np.median(out)), dev)


def stage_index():
    return device.SourceIndex(d_lon, d_lat, det['radius'])


ix0 = stage_index()
mub0 = ix0.min_upper_bound(det['Nj'])
res = {'n_gpus': ws, 'index_ms': median_ms(stage_index), 'self_query_ms': median_ms(lambda: ix0.min_upper_bound(det['Nj'])),
       'knn_shard_ms': median_ms(lambda: ix0.knn(d_tlo, d_tla, k=4, mode=L.MODE_INVDIST, mub=mub0, want_flags=True))}
for spec in (sys.argv[1:] or ['4']):
    force = spec.startswith('nccl')
    blocks = int(spec.replace('nccl', '') or 4)
    if ws == 1:
        break
    asm = parallel.TableAssembler(M, 4, cols, dev, blocks=blocks, force_nccl=force)

    def full():
        ix = device.SourceIndex(d_lon, d_lat, det['radius'])
        return asm.build(ix, d_tlo, d_tla, L.MODE_INVDIST, ix.min_upper_bound(det['Nj']))

    res[f'{"nccl" if not asm.p2p else "p2p"}_blocks{blocks}_build_ms'] = median_ms(full)
    res[f'{"nccl" if not asm.p2p else "p2p"}_blocks{blocks}_assemble_only_ms'] = median_ms(
        lambda: asm.build(ix0, d_tlo, d_tla, L.MODE_INVDIST, mub0))
    # correctness: the assembled table against a single-GPU style build of two probe row blocks
    _, idx_f, coef_f = full()
    torch.cuda.synchronize()
    for r in (0, ws - 1):
        plo, phi = parallel.shard_bounds(M, r, ws, align=cols)
        phi = min(phi, plo + 4 * cols)
        t_lo = torch.from_numpy(glo.ravel()[plo:phi].copy()).to(dev)
        t_la = torch.from_numpy(gl.ravel()[plo:phi].copy()).to(dev)
        want = ix0.knn(t_lo, t_la, k=4, mode=L.MODE_INVDIST, mub=mub0, want_flags=False)
        assert torch.equal(idx_f[:, plo:phi], want['idx']), f'rank {rank}: assembled indexes differ in block of rank {r}'
        assert torch.equal(coef_f[:, plo:phi].view(torch.int64), want['coef'].view(torch.int64))
    res[f'{spec}_how'] = asm.how
    del asm
if rank == 0:
    import json
    print(json.dumps(res), flush=True)
if ws > 1:
    dist.destroy_process_group()
