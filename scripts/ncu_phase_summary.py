"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump by source line / kernel phase."""
import collections, csv, re, sys
path, srcfile = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
cur = None
agg = collections.defaultdict(float)
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Function Name'):
        continue
    if r[0] and r[0].isdigit():
        nums = [x for x in r[1:] if re.fullmatch(r'[\d\.]+', x or '')]
        if len(nums) > 2:
            agg[(cur, int(r[0]))] += float(nums[2])
tot = sum(agg.values())
src = open(srcfile).read().split('\n')
name = srcfile.split('/')[-1]
marks = [i + 1 for i, l in enumerate(src) if '========' in l or 'for (int stage' in l or '---- outputs' in l or '---- load walkers' in l]
def rng(a, b): return sum(v for (f, l), v in agg.items() if f == name and a <= l <= b)
prev = None
for m in marks + [len(src)]:
    if prev is not None:
        print("lines %d-%d: %5.1f%%  %s" % (prev, m - 1, 100 * rng(prev, m - 1) / tot, src[prev - 1].strip()[:70]))
    prev = m
files = collections.defaultdict(float)
for (f, l), v in agg.items(): files[f] += v
for f, v in files.items(): print("%s: %.1f%%" % (f, 100 * v / tot))
print("top lines:")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
    print("  %-14s %4d %5.1f%%" % (f, l, 100 * v / tot))
