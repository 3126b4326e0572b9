#!/usr/bin/env python
"""Hot CUDA source lines of an ncu source-page CSV (ncu -i rep --page source --csv --print-source cuda,sass)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "# Samples" in r)
iSamp, iE, iLsb, iW = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('stall_long_sb'), hdr.index('stall_wait')
agg = {}
for r in rows:
    if len(r) <= iLsb or not r[0] or r[2] != '-':
        continue
    try:
        k = int(r[0])
        v = agg.setdefault(k, [r[1], 0, 0, 0, 0])
        v[1] += int(r[iSamp]); v[2] += int(r[iE]); v[3] += int(r[iLsb]); v[4] += int(r[iW])
    except ValueError:
        pass
tot = sum(v[1] for v in agg.values()) or 1
totE = sum(v[2] for v in agg.values()) or 1
print('source lines', len(agg), 'samples', tot, 'warp-instr', totE)
for ln, (t, s, e, l, w) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5d %5.1f%% samp %5.1f%% instr  long_sb %3.0f%% wait %3.0f%% | %s' % (ln, 100 * s / tot, 100 * e / totE, 100 * l / max(s, 1), 100 * w / max(s, 1), t.strip()[:105]))
