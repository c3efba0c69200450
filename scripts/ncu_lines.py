"""Per-source-line digest of `ncu -i X.ncu-rep --page source --csv --print-source cuda` :
samples, instructions executed, dominant stall reasons, shared-memory excess wavefronts.
usage: ncu_lines.py dump.csv [file-substring] [top-N]"""
import collections, csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = None; hdr = None
agg = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if not r[0].isdigit() or hdr is None: continue
    d = dict(zip(hdr, r))
    def f(k):
        try: return float(d.get(k, "") or 0)
        except ValueError: return 0.0
    key = (cur, int(r[0]))
    a = agg.setdefault(key, collections.Counter())
    a["samples"] += f("# Samples"); a["inst"] += f("Instructions Executed")
    a["excess"] += f("L1 Wavefronts Shared Excessive"); a["local"] += f("L2 Theoretical Sectors Local")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a[k] += f(k)
    a["src"] = 0
    agg[key]["_src"] = d.get("Source", "")[:60] if isinstance(d.get("Source"), str) else ""
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
print("total samples %d, instructions %.3g" % (tot, toti))
files = collections.Counter()
for (f_, l), a in agg.items(): files[f_] += a["samples"]
for f_, v in files.most_common(): print("  %-22s %5.1f%% of samples" % (f_, 100 * v / tot))
rows = [(k, a) for k, a in agg.items() if want in (k[0] or "")]
rows.sort(key=lambda kv: -kv[1]["samples"])
for (f_, l), a in rows[:topn]:
    st = sorted(((k[6:], v) for k, v in a.items() if isinstance(k, str) and k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print("%-16s %4d  smp %5.2f%%  inst %5.2f%%  exc %8d loc %8d  %s" % (f_, l, 100 * a["samples"] / tot, 100 * a["inst"] / toti, a["excess"], a["local"],
          " ".join("%s:%.0f%%" % (k, 100 * v / max(a["samples"], 1)) for k, v in st)))
