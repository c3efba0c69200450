"""Per-source-line digest of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
samples, warp instructions executed, shared-memory wavefronts (total / excessive), local-memory sectors, top stalls.
usage: ncu_cs_lines.py dump.csv [top-N] [file-substring]"""
import collections, csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50; want = sys.argv[3] if len(sys.argv) > 3 else ""
cur, hdr = None, None
agg = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if not r[0].isdigit() or hdr is None: continue
    d = {}
    for k, v in zip(hdr[4:], r[4:]): d[k] = v
    def f(k):
        try: return float(d.get(k, "") or 0)
        except ValueError: return 0.0
    a = agg.setdefault((cur, int(r[0])), collections.Counter())
    a["samples"] += f("# Samples"); a["inst"] += f("Instructions Executed")
    a["wave"] += f("L1 Wavefronts Shared"); a["excess"] += f("L1 Wavefronts Shared Excessive"); a["local"] += f("L2 Theoretical Sectors Local")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a[k] += f(k)
    a["_src"] = r[1][:70]
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
print("total samples %d, warp instructions %.4g, shared wavefronts %.4g (excessive %.4g), local sectors %.4g" % (
    tot, toti, sum(a["wave"] for a in agg.values()), sum(a["excess"] for a in agg.values()), sum(a["local"] for a in agg.values())))
files = collections.Counter(); filei = collections.Counter()
for (f_, l), a in agg.items(): files[f_] += a["samples"]; filei[f_] += a["inst"]
for f_, v in files.most_common(): print("  %-22s %5.1f%% of samples %5.1f%% of instructions" % (f_, 100 * v / tot, 100 * filei[f_] / toti))
rows = [(k, a) for k, a in agg.items() if want in (k[0] or "")]
rows.sort(key=lambda kv: -kv[1]["samples"])
for (f_, l), a in rows[:topn]:
    st = sorted(((k[6:], v) for k, v in a.items() if isinstance(k, str) and k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print("%-20s %4d smp %5.2f%% inst %5.2f%% wave %9d exc %9d loc %8d  %-28s | %s" % (f_, l, 100 * a["samples"] / tot, 100 * a["inst"] / toti, a["wave"], a["excess"], a["local"],
          " ".join("%s:%.0f%%" % (k, 100 * v / max(a["samples"], 1)) for k, v in st), a["_src"].strip()[:60]))
