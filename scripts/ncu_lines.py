#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[2] == '-' and r[0].isdigit():
        ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); ns = hdr.index('# Samples')
        try: inst = int(r[ie]); tinst = int(r[te]); smp = int(r[ns])
        except ValueError: continue
        if inst > 0 or smp > 0: agg[(cur, int(r[0]))] = (inst, tinst, smp, r[1][:100])
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values()); tott = sum(v[1] for v in agg.values())
print(f'total warp-inst {tot}  thread-inst {tott}  avg threads/inst {tott/max(tot,1):.2f}  samples {tots}')
for (f, l), (i, t, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print(f'{f}:{l:4d} smp {100*s/max(tots,1):4.1f}% inst {100*i/max(tot,1):4.1f}% thr {t/max(i,1):5.1f} | {src}')
