"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: target-row sharding, table all-gather
(equal and ragged shards), member->rank assignment and max-over-ranks timing.  The kernels themselves need
a GPU; what is exercised here is exactly the code that runs between them when WORLD_SIZE > 1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyg2p_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, rows, cols, k, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        m = rows * cols
        full_idx = torch.arange(k * m, dtype=torch.int32).reshape(k, m)
        full_w = torch.arange(k * m, dtype=torch.float64).reshape(k, m) / 7.0
        lo, hi = parallel.shard_bounds(m, rank, ws, align=cols)
        assert lo % cols == 0 and (hi % cols == 0 or hi == m)
        idx, w = parallel.all_gather_table(full_idx[:, lo:hi].contiguous(), full_w[:, lo:hi].contiguous(), m, align=cols)
        ok = torch.equal(idx, full_idx) and torch.equal(w, full_w)
        t = parallel.max_over_ranks(1.0 + rank, torch.device('cpu'))
        s = parallel.sum_over_ranks(1.0 + rank, torch.device('cpu'))
        # the partial double-double sums of the ADW reference radius, in rank order on every rank
        pairs = parallel.gather_pairs((10.0 + rank, 1e-20 * (rank + 1)), torch.device('cpu'))
        ok = ok and pairs == [(10.0, 1e-20), (11.0, 2e-20)]
        # build + assembly in target blocks (the exchange used without peer memory): a stand-in index whose "k-NN" of
        # target j is a function of j alone, so the assembled table is known
        asm = parallel.TableAssembler(m, k, cols, torch.device('cpu'), blocks=3)
        assert not asm.p2p and asm.bounds[rank] == (lo, hi)
        bb = asm._block_bounds()
        assert bb[0][0] == lo and bb[-1][1] == hi and all(a[1] == b[0] for a, b in zip(bb, bb[1:]))
        assert all((b0 - lo) % cols == 0 for b0, _ in bb)

        class FakeIndex:
            def knn(self, tlon, tlat, k, mode, mub, radius=None, rotation=None, want_flags=True):
                j = tlon.to(torch.int64)                       # the targets carry their own number
                idx_b = (j[None, :] * k + torch.arange(k)[:, None]).to(torch.int32)
                return {'idx': idx_b, 'coef': idx_b.double() / 7.0, 'flags': (j % 3).to(torch.uint8)}

        tl = torch.arange(lo, hi, dtype=torch.float64)
        flags, idx_a, w_a = asm.build(FakeIndex(), tl, tl, 2, 1.0)
        want_idx = (torch.arange(m)[None, :] * k + torch.arange(k)[:, None]).to(torch.int32)
        ok = ok and torch.equal(idx_a, want_idx) and torch.equal(w_a, want_idx.double() / 7.0)
        ok = ok and torch.equal(flags, (torch.arange(lo, hi) % 3).to(torch.uint8))
        q.put((rank, ok, lo, hi, t, s))
    except Exception as e:  # noqa: BLE001 - report instead of letting the parent wait for the queue
        q.put((rank, False, -1, -1, repr(e), 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('rows,cols', [(8, 5), (7, 5), (3, 11)])
def test_sharded_table_assembles_on_every_rank(rows, cols):
    ws, k = 2, 4
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, rows, cols, k, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == rows * cols      # contiguous cover
    assert all(r[4] == 2.0 and r[5] == 3.0 for r in res)


def test_shard_bounds_properties():
    for total, ws, align in [(21_600_000, 8, 7200), (950_000, 3, 1000), (10, 4, 1), (5, 8, 1), (0, 2, 1)]:
        b = [parallel.shard_bounds(total, r, ws, align) for r in range(ws)]
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= align
    assert parallel.shard_bounds(21_600_000, 3, 8, 7200) == (8_100_000, 10_800_000)


def test_members_round_robin():
    got = [parallel.members_of_rank(51, r, 8) for r in range(8)]
    assert sorted(sum(got, [])) == list(range(51))
    assert max(len(g) for g in got) - min(len(g) for g in got) == 1
    assert parallel.world() == (0, 1)
    idx = torch.zeros((4, 3), dtype=torch.int32)
    assert parallel.all_gather_table(idx, idx.double(), 3)[0] is idx       # single process: no-op
