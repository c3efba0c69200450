import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', '')); u = r[mu]
    ms = v/1e6 if u in ('ns','nsecond') else v/1e3 if u in ('us','usecond') else v if u in ('ms','msecond') else v*1e3
    a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += ms
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-90s %4d %9.3f ms  %8.4f ms/launch" % (k[:90], c, ms, ms/c))
print("total", sum(v[1] for v in agg.values()), "ms over", sum(v[0] for v in agg.values()), "launches")
