"""CPU model of the shared-memory gathers of apply_window_kernel (g2p_apply_win.cu) on the O1280 -> EFAS table.

Replays the tile plan (16-element ranges, skewed offsets) with numpy and counts shared-memory wavefronts for
alternative lane -> target mappings, skews and load widths, so that a kernel variant is only built when the
model says it pays.  Bank model: 32 banks x 4 B; a 64-bit load is served per half-warp (16 lanes, 16 8-byte
banks), a 128-bit load per quarter-warp (8 lanes, 8 16-byte banks); lanes reading the same word share one
access; wavefronts = max number of distinct words on one bank.

    python tools/bank_sim.py [n_tiles_sampled]
"""
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, '.')
from pyg2p_b200 import synth  # noqa: E402


def table():
    lats, lons, det = synth.octahedral_grid(1280)
    r = det['radius']

    def xyz(lo, la):
        lo, la = np.radians(lo), np.radians(la)
        return np.stack([r * np.cos(lo) * np.cos(la), r * np.sin(lo) * np.cos(la), r * np.sin(la)], axis=-1)

    tree = cKDTree(xyz(lons, lats), leafsize=30)
    tlat, tlon = synth.efas_target()
    d, idx = tree.query(xyz(tlon.ravel(), tlat.ravel()), k=4, workers=-1)
    return idx.reshape(tlat.shape + (4,)), lats.size


def plan_tile(idx_tile, skew_tab):
    """ranges + offsets of one tile as the plan kernel makes them; returns local offsets (same shape as idx_tile)."""
    v = idx_tile.ravel()
    blocks = np.unique(v // 16)
    # runs of consecutive blocks
    brk = np.flatnonzero(np.diff(blocks) != 1) + 1
    starts = np.concatenate([[0], brk])
    ends = np.concatenate([brk, [blocks.size]])
    r_start = blocks[starts] * 16
    r_len = (blocks[ends - 1] - blocks[starts] + 1) * 16
    nr = r_start.size
    rr = np.searchsorted(r_start, idx_tile, side='right') - 1           # range of every reference
    pos = idx_tile - r_start[rr]
    # alignment statistics between consecutive ranges (per target: refs in range r and r+1)
    dsum = np.zeros(nr)
    dcnt = np.zeros(nr)
    k = idx_tile.shape[-1]
    flat_rr = rr.reshape(-1, k)
    flat_pos = pos.reshape(-1, k)
    for a in range(k):
        for b in range(k):
            sel = flat_rr[:, b] == flat_rr[:, a] + 1
            np.add.at(dsum, flat_rr[sel, a], flat_pos[sel, a] - flat_pos[sel, b])
            np.add.at(dcnt, flat_rr[sel, a], 1)
    off = np.zeros(nr, dtype=np.int64)
    u = 0
    cur = 0
    for i in range(nr):
        if i > 0 and dcnt[i - 1] > 0:
            u -= int(np.round(dsum[i - 1] / dcnt[i - 1]))
        want = ((skew_tab[i % len(skew_tab)] - u) % 16) & ~1
        o = cur + ((want - cur) % 16)
        off[i] = o
        cur = o + r_len[i]
    return off[rr] + pos, cur


def wavefronts(loc, group, bank_elems):
    """loc: [n_requests, lanes] local element offsets; lanes are split in groups of `group`; banks of `bank_elems`
    8-byte elements (1 for 64-bit loads, 2 for 128-bit loads over message pairs)."""
    n, lanes = loc.shape
    word = loc // 1                      # the element (per message pair: the pair) addressed
    g = word.reshape(n, lanes // group, group)
    nb = 16 // bank_elems
    total = 0
    # distinct words per bank: sort by (bank, word)
    bank = g % nb
    key = bank * (1 << 20) + g
    key = np.sort(key, axis=-1)
    new = np.ones_like(key, dtype=bool)
    new[..., 1:] = key[..., 1:] != key[..., :-1]
    kb = key >> 20
    cnt = np.zeros(g.shape[:2] + (nb,), dtype=np.int64)
    for b in range(nb):
        cnt[..., b] = ((kb == b) & new).sum(-1)
    total = cnt.max(-1).sum()
    return total / (n * (lanes // group))  # mean wavefronts per group


def lane_map_pairs(shape_rows, shape_lanes, parity):
    """half-warp = shape_rows x shape_lanes lanes, every lane owns the column pair (2x, 2x+1); slot parity picks one."""
    out = []
    for br in range(0, 16, shape_rows):
        for bc in range(0, 64, 2 * shape_lanes):
            rr, cc = np.meshgrid(np.arange(shape_rows), np.arange(shape_lanes), indexing='ij')
            out.append((br + rr.ravel(), bc + 2 * cc.ravel() + parity))
    return out


def lane_map(shape_rows, shape_cols):
    """tile 16 x 64 -> list of warps; each warp 32 lanes (r, c) built from two half-warp blocks of shape
    (shape_rows x shape_cols) = 16 lanes."""
    out = []
    for br in range(0, 16, shape_rows):
        for bc in range(0, 64, shape_cols):
            rr, cc = np.meshgrid(np.arange(shape_rows), np.arange(shape_cols), indexing='ij')
            out.append((br + rr.ravel(), bc + cc.ravel()))
    return out  # list of half-warps


def main():
    n_sample = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    idx, n = table()
    rows, cols, k = idx.shape
    rng = np.random.default_rng(1)
    tiles = [(ty, tx) for ty in range(rows // 16) for tx in range(cols // 64)]
    pick = rng.choice(len(tiles), n_sample, replace=False)
    for skew_tab in [(0, 8), (0, 4, 8, 12), (0, 8, 4, 12)]:
        res = {}
        for t in pick:
            ty, tx = tiles[t]
            it = idx[ty * 16:(ty + 1) * 16, tx * 64:(tx + 1) * 64]
            loc, win = plan_tile(it, skew_tab)
            for name, (sr, sc) in {'1x16': (1, 16), '2x8': (2, 8), '4x4': (4, 4), '8x2': (8, 2), '16x1': (16, 1)}.items():
                hw = lane_map(sr, sc)
                reqs = np.stack([loc[r, c, :] for r, c in hw])          # [half-warps, 16, k]
                w64 = np.mean([wavefronts(reqs[:, :, j], 16, 1) for j in range(k)])
                # 128-bit loads over message pairs: quarter-warps = halves of the half-warp block
                w128 = np.mean([wavefronts(reqs[:, :, j], 8, 2) for j in range(k)])   # 8 lanes, 16 B banks -> 8 banks of pairs
                res.setdefault(name, []).append((w64, w128, win))
            for name, (sr, sl) in {'2x8 pairs': (2, 8), '4x4 pairs': (4, 4), '8x2 pairs': (8, 2)}.items():
                w = []
                for parity in (0, 1):
                    hw = lane_map_pairs(sr, sl, parity)
                    reqs = np.stack([loc[r, c, :] for r, c in hw])
                    w.append(np.mean([wavefronts(reqs[:, :, j], 16, 1) for j in range(k)]))
                res.setdefault(name, []).append((np.mean(w), 0.0, win))
        print(f'skew table {skew_tab}')
        for name, v in res.items():
            v = np.array(v)
            print(f'  half-warp {name:5s}: LDS.64 {v[:, 0].mean():.3f} wavefronts/half-warp ({2 * v[:, 0].mean():.2f}/request)   '
                  f'LDS.128 pairs {v[:, 1].mean():.3f}/quarter-warp ({4 * v[:, 1].mean() / 2:.2f}/request/message)   window {v[:, 2].mean():.0f}')


if __name__ == '__main__':
    main()
