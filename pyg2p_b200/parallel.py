"""Multi-GPU sharding of the interpolation path (one process per GPU, torch.distributed).

* intertable BUILD: the source index is replicated, target points are split into contiguous
  blocks (the reference's `num_of_splits` row chunks, scipy_interpolation_lib.py:588-607, become
  the rank partition); the only exchange is one all-gather of the finished table shards - and only
  when every rank needs the whole table (NCCL over NVLink on GPUs, gloo in the CPU tests).
* intertable APPLY: messages (step x ensemble member) are split by MEMBER so that the accumulation
  V[t] - V[t-D] of one member stays on one GPU; no collective at all.

Nothing here computes on the CPU: the functions take and return tensors on whatever device the
process group works on.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total, rank, world_size, align=1):
    """[lo, hi) of `rank`'s contiguous block of `total` items; block starts are multiples of `align`
    (e.g. the number of columns, to split whole target rows).  Blocks differ by at most one unit."""
    units = -(-total // align)
    base, rem = divmod(units, world_size)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, total), min(hi_u * align, total)


def members_of_rank(n_members, rank, world_size):
    """round-robin ensemble members -> rank (all steps of one member stay together)."""
    return list(range(rank, n_members, world_size))


def all_gather_table(idx_shard, coef_shard, m_total, align=1, group=None):
    """Assemble SoA table shards ([k, m_r] per rank, the contiguous target blocks of `shard_bounds(m_total, r,
    world, align)`) into the full [k, m_total] table on every rank.  The shard sizes follow from the arguments,
    so nothing but the table itself crosses the wire: with equal shards each neighbour row is gathered straight
    into its place (k collectives per array, no staging copy); ragged shards are padded to the largest."""
    rank, ws = world()
    if ws == 1:
        return idx_shard, coef_shard
    k, m_r = idx_shard.shape
    bounds = [shard_bounds(m_total, r, ws, align) for r in range(ws)]
    sizes = [hi - lo for lo, hi in bounds]
    if sizes[rank] != m_r:
        raise ValueError(f'rank {rank} holds {m_r} targets, its block has {sizes[rank]}')
    m_max = max(sizes)

    def gather(shard):
        if all(s == m_max for s in sizes):
            full = shard.new_empty((k, m_total))
            for kk in range(k):
                dist.all_gather_into_tensor(full[kk], shard[kk].contiguous(), group=group)
            return full
        pad = shard.new_zeros((k, m_max))
        pad[:, :m_r] = shard
        flat = shard.new_empty((ws * k, m_max))          # concatenation along dim 0 (the form gloo and nccl share)
        dist.all_gather_into_tensor(flat, pad, group=group)
        buf = flat.view(ws, k, m_max)
        return torch.cat([buf[r, :, :sizes[r]] for r in range(ws)], dim=1).contiguous()

    return gather(idx_shard), gather(coef_shard)


def max_over_ranks(value, device):
    """device-timed duration -> max over ranks (the contract for every multi-GPU number)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_pairs(pair, device):
    """(hi, lo) of every rank, in rank order (the partial double-double sums of the ADW reference radius)."""
    rank, ws = world()
    if ws == 1:
        return [tuple(pair)]
    mine = torch.tensor([float(pair[0]), float(pair[1])], dtype=torch.float64, device=device)
    full = torch.empty((ws * 2,), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, mine)
    h = full.cpu().tolist()
    return [(h[2 * r], h[2 * r + 1]) for r in range(ws)]


class TableAssembler:
    """Build + assembly of a sharded intertable on every GPU of the node, the exchange fused with the build.

    The full SoA table (idx int32 [k, M], coef float64 [k, M]) is allocated ONCE in symmetric memory
    (torch.distributed._symmetric_memory: the same buffer of every rank is mapped into every process over NVLink).
    `build` runs the rank's k-NN in `blocks` target blocks; each block is written straight into the rank's own copy
    (g2p_knn_block) and then pushed into the 7 peers' copies by g2p_peer_push (plain 16-byte stores through the peer
    mapping, our own kernel) on a second stream, so the NVLink transfer of block i overlaps the k-NN of block i+1.  One
    barrier at the end tells every rank that all pushes have landed.
    Where symmetric memory cannot be set up (no peer access, gloo / CPU groups), the same blocks are exchanged with ONE
    `all_gather_into_tensor` per block of the packed {idx, coef} bytes (NCCL over NVLink), still overlapped with the
    k-NN of the next block.  `how` says which of the two ran."""

    def __init__(self, m_total, k, align, device, blocks=None, force_nccl=None):
        self.m, self.k, self.align, self.device = int(m_total), int(k), int(align), device
        self.rank, self.ws = world()
        self.bounds = [shard_bounds(self.m, r, self.ws, align) for r in range(self.ws)]
        self.lo, self.hi = self.bounds[self.rank]
        # target blocks per rank: 4 at 2 GPUs, 8 from 4 GPUs up (8-GPU build: 1 block 3.10 ms, 2: 2.87, 4: 2.75, 8: 2.69)
        self.blocks = max(1, int(blocks)) if blocks else (8 if self.ws >= 4 else 4)
        # high priority: the push kernel needs only a few SMs, and it gets them as soon as CTAs of the running k-NN retire
        self.side = torch.cuda.Stream(device=device, priority=-1) if device.type == 'cuda' else None
        self.p2p = False
        self.why_not = None
        if force_nccl is None:
            force_nccl = os.environ.get('G2P_ASSEMBLE', '') == 'nccl'
        nbytes_i, nbytes_c = self.k * self.m * 4, self.k * self.m * 8
        self._off_c = -(-nbytes_i // 256) * 256
        if self.ws > 1 and device.type == 'cuda' and not force_nccl:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self._off_c + nbytes_c, dtype=torch.uint8, device=device)
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                self._buf, self._hdl = buf, hdl
                self._peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
                # NVLS: one store to the multicast address lands in every GPU's copy (G2P_ASSEMBLE=p2p keeps the per-peer stores)
                # (default from 4 GPUs up: with one peer the per-peer stores are faster - 2.66 against 2.82 ms per 2-GPU
                # assembly - with seven the block leaves the GPU once instead of seven times; G2P_ASSEMBLE=mc|p2p forces one)
                mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
                mode = os.environ.get('G2P_ASSEMBLE', '')
                want_mc = mode == 'mc' or (mode != 'p2p' and self.ws >= 4)
                self._mc_ptr = mc if (mc and want_mc and self._off_c % 16 == 0 and (self.m * 4) % 16 == 0) else 0
                self.p2p = True
            except Exception as e:           # noqa: BLE001 - any failure of the optional fast path selects the NCCL path
                self.why_not = f'{type(e).__name__}: {e}'
        if not self.p2p:
            self._buf = torch.empty(self._off_c + nbytes_c, dtype=torch.uint8, device=device)
        self.idx = self._buf[:nbytes_i].view(torch.int32).view(self.k, self.m)
        self.coef = self._buf[self._off_c:self._off_c + nbytes_c].view(torch.float64).view(self.k, self.m)
        self.how = (('multicast stores through the NVSwitch into symmetric memory (g2p_multicast_push, NVLS), '
                     if getattr(self, '_mc_ptr', 0) else 'peer stores over NVLink into symmetric memory (g2p_peer_push), ') +
                    f'overlapped with the k-NN in {self.blocks} target blocks, one barrier' if self.p2p else
                    f'one NCCL all_gather_into_tensor of the packed {{idx, coef}} bytes per target block ({self.blocks} blocks), '
                    f'overlapped with the k-NN' + (f' [symmetric memory unavailable: {self.why_not}]' if self.why_not else ''))

    def _block_bounds(self):
        units = -(-(self.hi - self.lo) // self.align)
        out = []
        for b in range(self.blocks):
            lo_u, hi_u = shard_bounds(units, b, self.blocks)
            lo, hi = self.lo + lo_u * self.align, min(self.hi, self.lo + hi_u * self.align)
            if hi > lo:
                out.append((lo, hi))
        return out

    def build(self, index, tlon, tlat, mode, mub, radius=None, rotation=None):
        """tlon / tlat: this rank's targets (device tensors, columns [lo, hi) of the table).  Returns (flags of the rank's
        targets, idx [k, M], coef [k, M]) - the assembled table, valid on the current stream when this returns."""
        flags = []
        if self.p2p:
            from . import _lib as L
            from . import device as D
            lib = D.require_cuda()
            cur = torch.cuda.current_stream(self.device)
            peers = [p for r, p in enumerate(self._peer_ptrs) if r != self.rank]
            arr_i = (C.c_void_p * len(peers))(*peers)
            arr_c = (C.c_void_p * len(peers))(*[p + self._off_c for p in peers])
            self._hdl.barrier()                         # nobody still reads the previous table
            for lo, hi in self._block_bounds():
                o = index.knn(tlon[lo - self.lo:hi - self.lo], tlat[lo - self.lo:hi - self.lo], k=self.k, mode=mode, mub=mub,
                              radius=radius, rotation=rotation, want_flags=True, into=(self.idx, self.coef, lo))
                flags.append(o['flags'])
                ev = torch.cuda.Event()
                ev.record(cur)
                self.side.wait_event(ev)
                st = C.c_void_p(self.side.cuda_stream)
                if self._mc_ptr and (lo * 4) % 16 == 0 and ((hi - lo) * 4) % 16 == 0:
                    L.check(lib.g2p_multicast_push(C.c_void_p(self.idx.data_ptr()), C.c_void_p(self._mc_ptr), self.m * 4,
                                                   lo * 4, (hi - lo) * 4, self.k, st))
                    L.check(lib.g2p_multicast_push(C.c_void_p(self.coef.data_ptr()), C.c_void_p(self._mc_ptr + self._off_c),
                                                   self.m * 8, lo * 8, (hi - lo) * 8, self.k, st))
                    continue
                L.check(lib.g2p_peer_push(C.c_void_p(self.idx.data_ptr()), arr_i, len(peers), self.m * 4, lo * 4, (hi - lo) * 4,
                                          self.k, st))
                L.check(lib.g2p_peer_push(C.c_void_p(self.coef.data_ptr()), arr_c, len(peers), self.m * 8, lo * 8, (hi - lo) * 8,
                                          self.k, st))
            cur.wait_stream(self.side)
            self._hdl.barrier()                         # every rank's pushes have landed
        else:
            # packed per-block exchange: [ws][k * mb * 12 bytes]; blocks of equal size on every rank are required for the
            # one-call form, ragged tails go through all_gather_table
            works = []
            sizes = {hi - lo for lo, hi in self.bounds}
            if len(sizes) != 1:
                o = index.knn(tlon, tlat, k=self.k, mode=mode, mub=mub, radius=radius, rotation=rotation, want_flags=True)
                idx_f, coef_f = all_gather_table(o['idx'], o['coef'], self.m, align=self.align)
                self.idx.copy_(idx_f)
                self.coef.copy_(coef_f)
                return o['flags'], self.idx, self.coef
            for lo, hi in self._block_bounds():
                mb = hi - lo
                o = index.knn(tlon[lo - self.lo:hi - self.lo], tlat[lo - self.lo:hi - self.lo], k=self.k, mode=mode, mub=mub,
                              radius=radius, rotation=rotation, want_flags=True)
                flags.append(o['flags'])
                pack = torch.empty(self.k * mb * 12, dtype=torch.uint8, device=self.device)
                pack[:self.k * mb * 4].view(torch.int32).view(self.k, mb).copy_(o['idx'])
                pack[self.k * mb * 4:].view(torch.float64).view(self.k, mb).copy_(o['coef'])
                full = torch.empty(self.ws * pack.numel(), dtype=torch.uint8, device=self.device)
                works.append((dist.all_gather_into_tensor(full, pack, async_op=True), full, lo - self.lo, mb))
            for w, full, rel, mb in works:
                w.wait()
                parts = full.view(self.ws, -1)
                for r in range(self.ws):
                    at = self.bounds[r][0] + rel
                    self.idx[:, at:at + mb].copy_(parts[r, :self.k * mb * 4].view(torch.int32).view(self.k, mb))
                    self.coef[:, at:at + mb].copy_(parts[r, self.k * mb * 4:].view(torch.float64).view(self.k, mb))
        return torch.cat(flags) if len(flags) > 1 else flags[0], self.idx, self.coef
