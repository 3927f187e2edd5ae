"""Two-GPU test of the build + table assembly over NVLink peer memory (pyg2p_b200.parallel.TableAssembler: g2p_knn_block,
g2p_peer_push / g2p_multicast_push, symmetric-memory barrier) and of the NCCL exchange used without peer access: every
rank must end up with the table a single GPU builds.  Skipped on a box with one GPU (the world_size-2 gloo tests in
test_parallel_gloo.py cover the host logic there)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from pyg2p_b200 import _lib as L
    from pyg2p_b200 import device, parallel, synth
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=dev)
    try:
        lat, lon, det = synth.octahedral_grid(160)
        tl, tlo = synth.latlon_target(0.5)
        m, cols = tl.size, tl.shape[1]
        lo, hi = parallel.shard_bounds(m, rank, ws, align=cols)
        ix = device.SourceIndex(lon, lat, det['radius'])
        mub = ix.min_upper_bound(det['Nj'])                      # sharded self query + all-reduce
        d_lo = torch.from_numpy(tlo.ravel()[lo:hi].copy()).to(dev)
        d_la = torch.from_numpy(tl.ravel()[lo:hi].copy()).to(dev)
        whole = ix.knn(tlo.ravel(), tl.ravel(), k=4, mode=L.MODE_INVDIST, mub=mub, want_flags=True)   # the single-GPU table
        how = []
        ok = True
        for mode in ('auto', 'mc', 'nccl'):
            os.environ['G2P_ASSEMBLE'] = '' if mode == 'auto' else mode
            asm = parallel.TableAssembler(m, 4, cols, dev, blocks=3)
            for _ in range(2):                                     # twice: the barrier before a rebuild
                flags, idx, coef = asm.build(ix, d_lo, d_la, L.MODE_INVDIST, mub)
            torch.cuda.synchronize()
            ok = ok and torch.equal(idx, whole['idx']) and torch.equal(coef.view(torch.int64), whole['coef'].view(torch.int64))
            ok = ok and torch.equal(flags, whole['flags'][lo:hi])
            how.append((mode, asm.p2p, bool(getattr(asm, '_mc_ptr', 0))))
            del asm
        q.put((rank, ok, how, mub))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_table_assembly_on_two_gpus():
    ws = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][3] == res[1][3]                                   # the same min_upper_bound on every rank
