"""Where does a kernel spill?  nvdisasm --print-line-info listing -> local loads/stores per source line.
usage: sass_spills.py <lib.so> <kernel-name-substring>"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
on, cur = False, None
cnt = collections.Counter()
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = pat in m.group(1); continue
    if re.match(r"\s*\.section", l):
        on = False
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r"\b(STL|LDL)\b", l)
    if m:
        cnt[(cur, m.group(1))] += 1
for (k, op), v in sorted(cnt.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(k, op, v)
