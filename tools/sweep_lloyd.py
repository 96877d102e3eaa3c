"""Variant sweep of the Lloyd step (development tool). usage: sweep_lloyd.py N K "v1,v2,..." [flags...]"""
import subprocess, sys
n, K = sys.argv[1], sys.argv[2]
variants = sys.argv[3].split(",")
flagsets = sys.argv[4:] or ["0", "1"]
for v in variants:
	for f in flagsets:
		r = subprocess.run([sys.executable, "tools/prof_lloyd.py", n, K, v, f, "20"], capture_output=True, text=True)
		print(r.stdout.strip() or r.stderr[-300:], flush=True)
