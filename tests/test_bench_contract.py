"""bench.py's reference arm runs on the CPU: check the JSON contract of the line it prints."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
	r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
	                   capture_output=True, text=True, timeout=600)
	assert r.returncode == 0, r.stderr[-2000:]
	lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
	assert len(lines) == 1
	d = json.loads(lines[0])
	for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
	          "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
		assert k in d, k
	assert d["impl"] == "reference" and d["unit"] == "MPix/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
	assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
	assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_own_arm_refuses_to_run_without_a_gpu():
	import torch

	if torch.cuda.is_available():
		return
	r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
	                   text=True, timeout=600)
	assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_bench_grid_path_thresholds_follow_the_library():
	"""bench.py decides how many kernels a step launches (`gpu_launches`) from the shard sizes at which the library
	takes the grid-filtered path (csrc/lloyd.cu grid_eligible): the two must name the same thresholds."""
	import re

	cu = (ROOT / "image_segmenter_b200" / "csrc" / "lloyd.cu").read_text()
	m = re.search(r"p\.K <= 16 \? (\d+)LL : p\.K <= 32 \? \(1LL << (\d+)\) : \(1LL << (\d+)\)", cu)
	assert m, "grid_eligible thresholds not found"
	lib = (int(m.group(1)), 1 << int(m.group(2)), 1 << int(m.group(3)))
	py = (ROOT / "bench.py").read_text()
	found = re.findall(r"\(([\d_]+) if (?:K|kk) <= 16 else 1 << (\d+) if (?:K|kk) <= 32 else 1 << (\d+)\)", py)
	assert len(found) == 2
	for a, b, c in found:
		assert (int(a.replace("_", "")), 1 << int(b), 1 << int(c)) == lib
