"""Per-region stall breakdown of an `ncu --page source --print-source cuda,sass --csv` dump.
usage: ncu_regions.py src.csv name:fileprefix:a-b,..."""
import csv, sys
def I(x):
    try: return int(x or 0)
    except ValueError: return 0
rows = list(csv.reader(open(sys.argv[1])))
cur, hdr = None, None
regions = []
for spec in sys.argv[2].split(','):
    name, f, ab = spec.split(':'); a, b = map(int, ab.split('-')); regions.append((name, f, a, b))
stalls = ["stall_barrier", "stall_branch_resolving", "stall_long_sb", "stall_math", "stall_membar", "stall_mio",
          "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_dispatch", "stall_sleep"]
agg = {r[0]: dict(samples=0, inst=0, **{s: 0 for s in stalls}) for r in regions}
tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or r[0] == '': continue
    try: ln = int(r[0])
    except ValueError: continue
    s = I(r[idx['# Samples']]); tot += s
    for name, f, a, b in regions:
        if cur.startswith(f) and a <= ln <= b:
            d = agg[name]; d['samples'] += s; d['inst'] += I(r[idx['Instructions Executed']])
            for st in stalls:
                d[st] += I(r[idx[st]])
print(f"{'region':10s} {'samp%':>6s} {'Minst':>7s} " + ' '.join(f"{s[6:12]:>6s}" for s in stalls))
for name, d in agg.items():
    sm = max(d['samples'], 1)
    print(f"{name:10s} {100*d['samples']/tot:6.1f} {d['inst']/1e6:7.0f} " + ' '.join(f"{100*d[s]/sm:6.1f}" for s in stalls))
