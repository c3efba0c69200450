"""Per-opcode digest of the SASS page of an ncu report: warp instructions, shared-memory wavefronts, global tag requests
per unit of work.  usage: ncu_sass_ops.py report.ncu-rep units [kernel-regex]"""
import csv, collections, io, subprocess, sys
rep, n = sys.argv[1], float(sys.argv[2])
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + (["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else [])
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
hdr = None
agg = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    t = d.get('Source', '').split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    parts = op.split('.')
    op = parts[0] + ''.join('.' + x for x in parts[1:] if x in ('64', '128', 'U16', 'E', '8x8x4'))
    def f(k):
        try: return float(d.get(k, '') or 0)
        except ValueError: return 0.0
    a = agg[op]
    a['inst'] += f('Instructions Executed'); a['wave'] += f('L1 Wavefronts Shared'); a['exc'] += f('L1 Wavefronts Shared Excessive')
    a['tag'] += f('L1 Tag Requests Global'); a['smp'] += f('# Samples')
ts = sum(a['smp'] for a in agg.values()) or 1
print("per unit: warp instructions %.0f, shared wavefronts %.0f (excessive %.0f), global tag requests %.0f" % (
    sum(a['inst'] for a in agg.values()) / n, sum(a['wave'] for a in agg.values()) / n, sum(a['exc'] for a in agg.values()) / n, sum(a['tag'] for a in agg.values()) / n))
for op, a in sorted(agg.items(), key=lambda kv: -kv[1]['inst'])[:28]:
    print("%-12s inst %7.1f  shared wavefronts %7.1f (exc %6.1f)  global tags %7.1f  samples %5.2f%%" % (op, a['inst'] / n, a['wave'] / n, a['exc'] / n, a['tag'] / n, 100 * a['smp'] / ts))
