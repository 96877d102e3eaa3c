"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the fused P2P-exchange Lloyd
kernel == the NCCL all_reduce path == an unsharded run, bit-identical centres on every rank."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_rank_p2p_matches_nccl_and_unsharded():
	import torch

	if torch.cuda.device_count() < 2:
		pytest.skip("needs 2 GPUs")
	cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
	       "127.0.0.1", "--master-port", "29533", str(ROOT / "tools" / "mg_check.py")]
	r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
	# keep the evidence: gpurun_out/ travels back from the GPU box, profiles/ is where it is committed
	for d in ("gpurun_out", "profiles"):
		if (ROOT / d).is_dir():
			(ROOT / d / "r2_multigpu_parity_2gpu.log").write_text(
				"$ " + " ".join(cmd[1:]) + "\n" + r.stdout[-6000:] + ("\n[stderr tail]\n" + r.stderr[-1500:] if r.returncode else ""))
	assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
	assert "ok=True" in r.stdout and "ok=False" not in r.stdout
