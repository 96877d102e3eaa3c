"""Exchange latency of the sharded Lloyd iteration (run under torchrun): per-iteration time for shards from a
few thousand pixels (pure exchange + fixed launch cost) up to 64 MP per rank, P2P-fused and NCCL paths, next
to the same shard iterated without any exchange."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200.engine import KMeansGPU, get_engine
from image_segmenter_b200.sharded import make_gpu_lloyd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
K = 16
sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 1 << 20, 1 << 23, 1 << 26]
for n in sizes:
	g = torch.Generator(device=eng.dev); g.manual_seed(1 + rank)
	rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
	planes = eng.rgba_to_lab(rgba)
	C0 = np.ascontiguousarray(planes[:, :K].T.double().cpu().numpy() + 0.01)
	t0 = torch.from_numpy(C0).to(eng.dev); dist.broadcast(t0, 0); C0 = t0.cpu().numpy()
	out = {}
	for ex in ("p2p", "nccl"):
		drv = make_gpu_lloyd(eng, planes, n, K, labels=None, exact=False, exchange=ex)
		drv.set_centers(C0)
		for _ in range(20):
			drv.iterate()
		dist.barrier(); torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(200):
			drv.iterate()
		e1.record(); torch.cuda.synchronize()
		t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], dtype=torch.float64, device=eng.dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		out[ex] = float(t.item())
	km = KMeansGPU(eng, "f32", n, planes=planes, exact=False)
	a, b = torch.from_numpy(C0.copy()).to(eng.dev), torch.zeros((K, 3), dtype=torch.float64, device=eng.dev)
	s, c, st = (torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev),
	            torch.zeros(4, dtype=torch.float64, device=eng.dev))
	ctl = torch.tensor([0.0, 0.0, -1.0, 0.0], dtype=torch.float64, device=eng.dev)
	km._run(a, b, K, s, c, st, 10, ctl)
	dist.barrier(); torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); km._run(a, b, K, s, c, st, 100, ctl); e1.record(); torch.cuda.synchronize()
	t = torch.tensor([e0.elapsed_time(e1) / 100 * 1e3], dtype=torch.float64, device=eng.dev)
	dist.all_reduce(t, op=dist.ReduceOp.MAX)
	if rank == 0:
		print(f"world={world} n/rank={n:9d}: no exchange {float(t.item()):7.2f} us | p2p {out['p2p']:7.2f} us | nccl {out['nccl']:7.2f} us per iteration", flush=True)
dist.barrier()
dist.destroy_process_group()
