"""Batched k-means (config 4): one launch for all images == the per-image path, image by image."""
import numpy as np
import pytest

from oracle import kmeans as okm

from gpu_util import engine, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,h,w,K", [(1, 40, 52, 8), (5, 64, 48, 8), (37, 30, 44, 5), (3, 1080, 1920, 8)])
def test_batched_equals_per_image_and_oracle(B, h, w, K):
	from image_segmenter_b200.batch import kmeans_rgb_batch
	from image_segmenter_b200.engine import KMeansGPU

	rng = np.random.default_rng(B * 100 + K)
	imgs = rng.integers(0, 256, (B, h, w, 4), dtype=np.uint8)
	imgs[..., 3] = 255
	imgs[0, :3, :, 3] = 0  # a few transparent rows in the first image
	n = h * w
	inits = np.stack([imgs[i].reshape(-1, 4)[rng.choice(n, K, replace=False) + 0, :3].astype(np.float64) + 0.125
	                  for i in range(B)])
	iters = 6
	labels, centers, counts, n_empty = kmeans_rgb_batch(imgs, K, inits, iters, exact=True)
	lab = labels.cpu().numpy()
	e = engine()
	for i in range(B):
		if B > 8 and i % 6:
			continue
		d = to_dev(imgs[i].reshape(-1, 4))
		fit = KMeansGPU(e, "rgba8", n, px=d, mask_mode=0, min_bright=-1).fit_single(inits[i], max_iter=iters, tol=-1.0)
		assert np.array_equal(fit.centers, centers[i])  # exact integer sums: bit-identical
		assert np.array_equal(fit.labels[:n].cpu().numpy(), lab[i])
		assert np.array_equal(fit.counts, counts[i])
	# oracle on a small image
	i = 0
	px = imgs[i].reshape(-1, 4)
	keep = px[:, 3] > 0
	X = px[keep][:, :3].astype(np.float64)
	C = inits[i].copy()
	for _ in range(iters):
		_, _, _, C, _ = okm.lloyd_iter(X, C, relocate=False)
	if n <= 100000:
		assert np.allclose(centers[i], C, rtol=1e-12, atol=1e-9)
		ref = okm.assign_labels(X, C)
		got = lab[i][keep]
		assert (got != ref).sum() <= 2 and (lab[i][~keep] == 255).all()


def test_partition_batch():
	from image_segmenter_b200.batch import partition_batch

	for n in (1, 7, 1024):
		for world in (1, 2, 8):
			parts = [partition_batch(n, world, r) for r in range(world)]
			assert parts[0][0] == 0 and parts[-1][1] == n
			assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
