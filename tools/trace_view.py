#!/usr/bin/env python
"""Summarise a YA_TRACE timeline of yaha_b200_host: for one pass, a per-phase table, per-batch lifelines and worker utilisation."""
import sys
from collections import defaultdict
ev = [l.split() for l in open(sys.argv[1])]
ev = [(k, int(w), int(b), float(t0), float(t1), int(x)) for k, w, b, t0, t1, x in ev]
passes = [e for e in ev if e[0] == 'A']
a = passes[int(sys.argv[2]) if len(sys.argv) > 2 else -1]
T0, T1 = a[3], a[4]
print(f"pass {a[1]}: {T1 - T0:.0f} us")
sel = [e for e in ev if e[0] != 'A' and e[3] >= T0 - 1 and e[4] <= T1 + 1]
tot = defaultdict(float); cnt = defaultdict(int)
for k, w, b, t0, t1, x in sel:
    tot[k] += t1 - t0; cnt[k] += 1
for k in sorted(tot): print(f"  {k}: n={cnt[k]} total={tot[k]:.0f} us  mean={tot[k]/cnt[k]:.0f} us")
byb = defaultdict(list)
for e in sel: byb[e[2]].append(e)
for b in sorted(byb):
    es = sorted(byb[b], key=lambda e: e[3])
    line = []
    for k, w, bb, t0, t1, x in es:
        if k == 'H': continue
        line.append(f"{k}[{t0-T0:.0f}-{t1-T0:.0f}{'/'+str(x) if k=='D' else ''}{'x'+str(w) if k=='D' else ''}]")
    hs = [e for e in es if e[0] == 'H']
    print(f"batch {b}: H passes={len(hs)} Hbusy={sum(e[4]-e[3] for e in hs):.0f}us  " + " ".join(line))
nb = int((T1 - T0) / 1000) + 1
util = [0.0] * nb
for k, w, b, t0, t1, x in sel:
    if k != 'H': continue
    s, e = t0 - T0, t1 - T0
    i = int(s // 1000)
    while s < e and i < nb:
        seg = min(e, (i + 1) * 1000) - s
        util[i] += seg; s += seg; i += 1
print("worker-busy threads per ms bucket:", " ".join(f"{u/1000:.1f}" for u in util))
