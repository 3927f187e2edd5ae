#!/usr/bin/env python
"""CPU model of the neighbour-order search of plan_kernel (pick_pair_order, g2p_apply_win.cu) on O1280 -> EFAS.

The sequential sum ((w0 v0 + w1 v1) + w2 v2) + w3 v3 lets a lane load neighbours 0 / 1 in either order for free and
2 / 3 in either order for one select.  This script replays the tile plan of tools/bank_sim.py, applies the same greedy
(+ flip sweeps) per half-warp of 2x8 targets and prints the shared-memory wavefronts per half-warp LDS.64 of the four
gathers before and after - the numbers quoted in DESIGN.md and profiles/r02_apply_lag.md; the plan kernel's own count
(`g2p_plan_info.gather_wavefronts`: 1.414 -> 1.065 on the whole table) is the measured counterpart.

    python tools/pair_order_sim.py [n_tiles_sampled]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bank_sim as B  # noqa: E402


def hw_cost(words):
    """wavefronts of one half-warp load: distinct words on the busiest of the 16 8-byte banks"""
    banks = {}
    for w in set(words.tolist()):
        banks[w % 16] = banks.get(w % 16, 0) + 1
    return max(banks.values())


def greedy(a, b):
    """lane q takes the orientation that adds fewer wavefronts to the two loads given lanes 0..q-1"""
    a, b = a.copy(), b.copy()
    sa, sb = {}, {}

    def add(s, w):
        s.setdefault(w % 16, set()).add(w)

    def inc(s, w):
        st = s.get(w % 16, set())
        return 0 if w in st else len(st) + 1

    cura = curb = 0
    for q in range(16):
        keep = max(cura, inc(sa, a[q])) + max(curb, inc(sb, b[q]))
        swap = max(cura, inc(sa, b[q])) + max(curb, inc(sb, a[q]))
        if swap < keep:
            a[q], b[q] = b[q], a[q]
        add(sa, a[q])
        add(sb, b[q])
        cura = max(len(v) for v in sa.values())
        curb = max(len(v) for v in sb.values())
    return a, b


def sweeps(a, b, n):
    """flip single lanes while the pair of loads gets cheaper"""
    best = hw_cost(a) + hw_cost(b)
    for _ in range(n):
        improved = False
        for q in range(16):
            a[q], b[q] = b[q], a[q]
            c = hw_cost(a) + hw_cost(b)
            if c < best:
                best, improved = c, True
            else:
                a[q], b[q] = b[q], a[q]
        if not improved:
            break
    return a, b


def main():
    n_sample = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    idx, _ = B.table()
    rows, cols, _ = idx.shape
    rng = np.random.default_rng(1)
    tiles = [(ty, tx) for ty in range(rows // 16) for tx in range(cols // 64)]
    pick = rng.choice(len(tiles), n_sample, replace=False)
    plain, g_only, g_sw = np.zeros(4), np.zeros(4), np.zeros(4)
    cnt = 0
    for t in pick:
        ty, tx = tiles[t]
        loc, _ = B.plan_tile(idx[ty * 16:(ty + 1) * 16, tx * 64:(tx + 1) * 64], (0, 4, 8, 12))
        for r, c in B.lane_map(2, 8):
            w = loc[r, c, :]
            plain += [hw_cost(w[:, j]) for j in range(4)]
            for pr in (0, 1):
                a, b = greedy(w[:, 2 * pr], w[:, 2 * pr + 1])
                g_only[2 * pr:2 * pr + 2] += (hw_cost(a), hw_cost(b))
                a, b = sweeps(a, b, 2)
                g_sw[2 * pr:2 * pr + 2] += (hw_cost(a), hw_cost(b))
            cnt += 1
    for name, v in (('distance order', plain), ('greedy', g_only), ('greedy + 2 sweeps', g_sw)):
        print(f'{name:18s} per gather {np.round(v / cnt, 3)}  mean {v.mean() / cnt:.3f}   '
              f'(0/1 only: {(v[:2].sum() + plain[2:].sum()) / 4 / cnt:.3f})')


if __name__ == '__main__':
    main()
