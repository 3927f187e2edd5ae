#!/usr/bin/env python
"""Top stall lines of an `ncu --page source --csv` dump (SASS view): address, samples, executed, instruction.
    python tools/ncu_source_top.py gpurun_out/x.source.csv.gz [N] [lo-hi address window in hex offsets]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, ism, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = rows[2:]
base = int(body[0][ia], 16)
tot = sum(int(r[ism]) for r in body)
print(f'{len(body)} instructions, {tot} samples')
mode = sys.argv[3] if len(sys.argv) > 3 else 'top'
if mode == 'all':
    sel = body
else:
    sel = sorted(body, key=lambda r: -int(r[ism]))[:n]
for r in sel:
    why = sorted(((int(r[i]), hdr[i][6:]) for i in stalls if int(r[i]) > 0), reverse=True)[:3]
    print(f'{int(r[ia], 16) - base:6x} {int(r[ism]):6d} {100 * int(r[ism]) / tot:5.1f}% ex {int(r[iex]):9d}  {r[isrc].strip()[:70]:70s} {" ".join(f"{w}:{c}" for c, w in why)}')
