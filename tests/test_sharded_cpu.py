"""The N>1 path on CPU: world_size-2 gloo run of image_segmenter_b200.sharded.ShardedLloyd with the
local step supplied by the oracle — checks the row sharding, the all-reduce plumbing, the loop /
convergence control and that both ranks end with bit-identical centres equal to the unsharded run."""
import os
import socket

import numpy as np
import pytest

from image_segmenter_b200.sharded import shard_rows


def test_shard_rows_cover_exactly():
	for h in (1, 7, 8, 1080, 8192):
		for world in (1, 2, 3, 8):
			blocks = [shard_rows(h, world, r) for r in range(world)]
			assert blocks[0][0] == 0 and blocks[-1][1] == h
			assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
			sizes = [b - a for a, b in blocks]
			assert max(sizes) - min(sizes) <= 1


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank, world, port, H, W, K, iters, tol, out_dir):
	import torch
	import torch.distributed as dist

	from image_segmenter_b200.sharded import ShardedLloyd
	from oracle import kmeans as okm

	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		rng = np.random.default_rng(5)
		X = rng.uniform(0, 100, (H, W, 3))
		C0 = X.reshape(-1, 3)[rng.choice(H * W, K, replace=False)].copy()
		r0, r1 = shard_rows(H, world, rank)
		Xl = X[r0:r1].reshape(-1, 3)

		def local_step(c_in, acc):
			C = c_in.numpy()
			lab = okm.assign_labels(Xl, C) if len(Xl) else np.zeros(0, np.int64)
			s, c = okm.accumulate(Xl, lab, K) if len(Xl) else (np.zeros((K, 3)), np.zeros(K))
			acc[:3 * K] = torch.from_numpy(s.reshape(-1))
			acc[3 * K:] = torch.from_numpy(c)

		def finalize(acc, c_in, c_out, stats):
			a = acc.numpy()
			new = okm.average_centers(a[:3 * K].reshape(K, 3), a[3 * K:])
			stats[0] = okm.center_shift_total(c_in.numpy(), new)
			stats[1] = float((a[3 * K:] == 0).sum())
			c_out.copy_(torch.from_numpy(new))

		drv = ShardedLloyd(K, local_step, finalize, device="cpu")
		res = drv.run(C0, iters, tol)
		np.savez(os.path.join(out_dir, f"r{rank}.npz"), centers=res.centers, n_iter=res.n_iter, shift2=res.shift2)
	finally:
		dist.destroy_process_group()


@pytest.mark.parametrize("H,tol", [(37, 0.0), (64, 1e-3)])
def test_two_rank_gloo_matches_unsharded(tmp_path, H, tol):
	import torch.multiprocessing as mp

	from oracle import kmeans as okm

	W, K, iters = 50, 5, 12
	mp.spawn(_worker, args=(2, _free_port(), H, W, K, iters, tol, str(tmp_path)), nprocs=2, join=True)
	r = [np.load(tmp_path / f"r{i}.npz") for i in range(2)]
	assert np.array_equal(r[0]["centers"], r[1]["centers"])  # bit-identical on both ranks
	assert int(r[0]["n_iter"]) == int(r[1]["n_iter"])
	rng = np.random.default_rng(5)
	X = rng.uniform(0, 100, (H, W, 3)).reshape(-1, 3)
	C = X[rng.choice(H * W, K, replace=False)].copy()
	n_ref = 0
	for _ in range(iters):
		_, _, _, C_new, sh = okm.lloyd_iter(X, C, relocate=False)
		C = C_new
		n_ref += 1
		if sh <= tol:
			break
	assert int(r[0]["n_iter"]) == n_ref
	assert np.allclose(r[0]["centers"], C, rtol=1e-12, atol=1e-12)


def _reloc_worker(rank, world, port, H, W, K, iters, check_every, out_dir):
	"""Initial centres with a far-away duplicate pair: the first iteration leaves a cluster empty."""
	import torch
	import torch.distributed as dist

	from image_segmenter_b200.sharded import ShardedLloyd
	from oracle import kmeans as okm

	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		X, C0 = _reloc_problem(H, W, K)
		r0, r1 = shard_rows(H, world, rank)
		Xl = X.reshape(H, W, 3)[r0:r1].reshape(-1, 3)
		base = r0 * W
		state = {}

		def local_step(c_in, acc):
			C = c_in.numpy()
			lab = okm.assign_labels(Xl, C)
			state["labels"] = lab
			s, c = okm.accumulate(Xl, lab, K)
			acc[:3 * K] = torch.from_numpy(s.reshape(-1))
			acc[3 * K:] = torch.from_numpy(c)

		def finalize(acc, c_in, c_out, stats):
			a = acc.numpy()
			new = okm.average_centers(a[:3 * K].reshape(K, 3), a[3 * K:])
			stats[0] = okm.center_shift_total(c_in.numpy(), new)
			stats[1] = float((a[3 * K:] == 0).sum())
			c_out.copy_(torch.from_numpy(new))

		def local_farthest(c_in, prev):
			lab = state["labels"]
			d = ((Xl - c_in.numpy()[lab]) ** 2).sum(axis=1)
			gi = base + np.arange(len(d))
			ok = np.ones(len(d), bool) if prev is None else (d < prev[0]) | ((d == prev[0]) & (gi > prev[1]))
			if not ok.any():
				return (0.0, -1.0, 0.0, 0.0, 0.0, 0.0)
			cand = np.nonzero(ok)[0]
			j = cand[np.lexsort((gi[cand], -d[cand]))[0]]
			return (float(d[j]), float(gi[j]), *map(float, Xl[j]), float(lab[j]))

		drv = ShardedLloyd(K, local_step, finalize, device="cpu", check_every=check_every, local_farthest=local_farthest)
		res = drv.run(C0, iters, 0.0)
		np.savez(os.path.join(out_dir, f"rl{rank}.npz"), centers=res.centers, n_iter=res.n_iter, n_reloc=drv.n_relocated)
	finally:
		dist.destroy_process_group()


def _reloc_problem(H, W, K):
	rng = np.random.default_rng(11)
	X = rng.uniform(0, 100, (H * W, 3))
	C0 = X[rng.choice(H * W, K, replace=False)].copy()
	C0[1] = [900.0, 900.0, 900.0]   # nobody is nearest to these two: two clusters come out empty
	C0[3] = [950.0, 900.0, 900.0]
	return X, C0


@pytest.mark.parametrize("check_every", [1, 3])
def test_two_rank_gloo_relocates_empty_clusters_like_unsharded(tmp_path, check_every):
	"""the sharded loop must relocate empty clusters (sklearn/cluster/_k_means_common.pyx:167-211)
	exactly as the unsharded loop does — same picks (distance descending, global index ascending)."""
	import torch.multiprocessing as mp

	from oracle import kmeans as okm

	H, W, K, iters = 21, 30, 6, 7
	mp.spawn(_reloc_worker, args=(2, _free_port(), H, W, K, iters, check_every, str(tmp_path)), nprocs=2, join=True)
	r = [np.load(tmp_path / f"rl{i}.npz") for i in range(2)]
	assert np.array_equal(r[0]["centers"], r[1]["centers"])
	assert int(r[0]["n_reloc"]) == 2 and int(r[1]["n_reloc"]) == 2
	X, C = _reloc_problem(H, W, K)
	for _ in range(iters):
		_, _, _, C, sh = okm.lloyd_iter(X, C, relocate=True)
		if sh <= 0.0:
			break
	assert np.allclose(r[0]["centers"], C, rtol=1e-12, atol=1e-12)


def test_sharded_loop_raises_without_relocation_hook():
	import torch

	from image_segmenter_b200.sharded import ShardedLloyd

	def local_step(c_in, acc):
		acc.zero_()
		acc[3 * 2] = 5.0  # cluster 0 has 5 pixels, cluster 1 none

	def finalize(acc, c_in, c_out, stats):
		stats[0], stats[1] = 1.0, 1.0

	drv = ShardedLloyd(2, local_step, finalize, device="cpu")
	with pytest.raises(RuntimeError, match="empty cluster"):
		drv.run(np.zeros((2, 3)), 3, 0.0)


def _mc_worker(rank, world, port, H, W, k, out_dir):
	import torch
	import torch.distributed as dist

	from image_segmenter_b200.sharded import ShardedMedianCut
	from oracle import mediancut as omc

	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		rng = np.random.default_rng(9)
		cent = rng.integers(0, 256, (7, 3))
		img = np.clip(cent[rng.integers(0, 7, (H, W))] + rng.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)
		r0, r1 = shard_rows(H, world, rank)
		loc = img[r0:r1].reshape(-1, 3)
		key = lambda p: (p[:, 0].astype(np.int64) << 16) | (p[:, 1].astype(np.int64) << 8) | p[:, 2].astype(np.int64)

		def local_hist():
			return torch.from_numpy(np.bincount(key(loc), minlength=1 << 24).astype(np.int32))

		def palette_from_hist(hist, n_colors):
			h = hist.numpy()
			keys = np.nonzero(h)[0]
			rgb = np.stack([(keys >> 16) & 0xFF, (keys >> 8) & 0xFF, keys & 0xFF], axis=1).astype(np.uint8)
			expanded = np.repeat(rgb, h[keys], axis=0)  # same multiset of pixels as the whole image
			pal, idx = omc.quantize(expanded, n_colors)
			first = np.cumsum(h[keys]) - h[keys]
			return {"palette": pal, "keys": keys, "index_of_key": idx[first]}

		def local_map(plan):
			pos = np.searchsorted(plan["keys"], key(loc))
			return plan["index_of_key"][pos]

		idx, plan = ShardedMedianCut(local_hist, palette_from_hist, local_map).run(k)
		np.savez(os.path.join(out_dir, f"mc{rank}.npz"), idx=idx, pal=plan["palette"])
	finally:
		dist.destroy_process_group()


def test_two_rank_gloo_median_cut_matches_unsharded(tmp_path):
	import torch.multiprocessing as mp

	from oracle import mediancut as omc

	H, W, k = 45, 40, 12
	mp.spawn(_mc_worker, args=(2, _free_port(), H, W, k, str(tmp_path)), nprocs=2, join=True)
	r = [np.load(tmp_path / f"mc{i}.npz") for i in range(2)]
	rng = np.random.default_rng(9)
	cent = rng.integers(0, 256, (7, 3))
	img = np.clip(cent[rng.integers(0, 7, (H, W))] + rng.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)
	pal, idx = omc.quantize(img, k)
	assert np.array_equal(r[0]["pal"], pal) and np.array_equal(r[1]["pal"], pal)
	assert np.array_equal(np.concatenate([r[0]["idx"], r[1]["idx"]]), idx.reshape(-1))
