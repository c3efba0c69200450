"""Run bench.py with the given extra args and print a one-line digest."""
import json, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline"] + sys.argv[1:], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print("bench failed:", out.stderr[-800:]); sys.exit(1)
d = json.loads(line[-1])
print("tag=%s walkers/s=%.0f ms/step=%.1f eloc_ms=%.1f frac=%.3f e2e=%.0f clocks=%s" % (
    os.environ.get("TAG", ""), d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
