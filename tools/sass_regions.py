"""Static SASS instruction count per source region from `nvdisasm -g -c cubin` output.
usage: sass_regions.py lines.txt name:fileprefix:a-b,...   (innermost inlined-at frame attribution: the last
'//## File' marker before an instruction)."""
import re, sys
regions = []
for spec in sys.argv[2].split(','):
    name, f, ab = spec.split(':'); a, b = map(int, ab.split('-')); regions.append((name, f, a, b))
cnt = {r[0]: 0 for r in regions}; other = {}; total = 0
cur = (None, 0)
for line in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        if 'inlined at' in m.group(3) and False: pass
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        total += 1
        hit = False
        for name, f, a, b in regions:
            if cur[0] and cur[0].startswith(f) and a <= cur[1] <= b:
                cnt[name] += 1; hit = True; break
        if not hit: other[cur[0]] = other.get(cur[0], 0) + 1
print('total', total)
for k, v in cnt.items(): print(f"{k:12s} {v:6d}")
print('other', other)
