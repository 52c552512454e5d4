"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
cur, agg, hdr = None, {}, None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) <= si or r[0] == '': continue
    try: agg[(cur, int(r[0]))] = (int(r[si] or 0), int(r[ii] or 0), r[1][:95])
    except ValueError: pass
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print('samples', tot, 'warp-inst', toti)
for (f, ln), (s, ins, src) in sorted(agg.items()):
    if s > tot * thr / 100 or ins > toti * thr / 100:
        print(f"{f[:16]:16s} {ln:4d} {100*s/tot:5.1f}%s {100*ins/toti:5.1f}%i  {src}")
if len(sys.argv) > 3:
    # ranges: name:file_prefix:a-b,...
    for spec in sys.argv[3].split(','):
        name, f, ab = spec.split(':'); a, b = map(int, ab.split('-'))
        s = sum(v[0] for (ff, l), v in agg.items() if ff.startswith(f) and a <= l <= b)
        i = sum(v[1] for (ff, l), v in agg.items() if ff.startswith(f) and a <= l <= b)
        print(f"{name:14s} {100*s/tot:5.1f}%s {100*i/toti:5.1f}%i")
