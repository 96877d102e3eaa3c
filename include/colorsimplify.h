/*
 * colorsimplify.h — C ABI of libcolorsimplify.so (B200 / sm_100a).
 *
 * Drop-in boundary for the colour-simplification hot path of
 * jeffreyperez1620/image_segmenter (app/processing/color_simplify.py).  Every entry point
 * below replaces one native third-party routine that the reference reaches from that file;
 * the "replaces" line of each declaration cites the reference call site (file:line relative
 * to the reference checkout, or sklearn/… PIL/… for the wheel the reference calls into).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions across the boundary;
 *   - every function returns int: 0 = ok, >0 = cudaError_t, <0 = cs_status (argument /
 *     library error); the text of the last failure of the calling thread is
 *     cs_last_error();
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. a torch tensor's
 *     data_ptr()), h_* are HOST pointers; `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream); device entry points are asynchronous on `stream`;
 *   - the library owns nothing persistent except the opaque cs_ctx (per-device scratch:
 *     per-block partials, the "last block" counter).  A cs_ctx is single-threaded, like the
 *     single GUI-thread caller of the reference (app/ui/main_window.py:596-601).
 *   - pixel counts are int64_t; planar float buffers must be 16-byte aligned.
 */
#ifndef COLORSIMPLIFY_H_
#define COLORSIMPLIFY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_ABI_VERSION 1
#define CS_MAX_K 256 /* UI range: K in [2,256], app/ui/color_processing_panel.py:110-114 */

typedef enum cs_status {
	CS_OK = 0,
	CS_ERR_ARG = -1,      /* bad argument (null pointer, K out of range, misaligned plane) */
	CS_ERR_NO_DEVICE = -2,/* no CUDA device / not an sm_100 part */
	CS_ERR_ALLOC = -3,
	CS_ERR_UNSUPPORTED = -4
} cs_status;

typedef struct cs_ctx cs_ctx;

/* ---- library / context ------------------------------------------------------------- */
int cs_abi_version(void);
const char *cs_last_error(void);
/* creates the per-device scratch on CUDA device `device` (cudaSetDevice is called). */
int cs_ctx_create(int device, cs_ctx **out);
int cs_ctx_destroy(cs_ctx *ctx);
/* SM count of the context's device (grid sizing is a multiple of it). */
int cs_ctx_sm_count(const cs_ctx *ctx);
/* development aid: the context's 64 x u64 scratch block (relocation keys; phase time stamps of
 * CS_PHASE_TIMING builds) copied to the host. */
int cs_debug_scratch(cs_ctx *ctx, unsigned long long *h_out64);

/* ---- K1: sRGB u8 -> CIELAB ----------------------------------------------------------
 * replaces skimage.color.rgb2lab over every opaque pixel
 *   (color_simplify.py:470, 540, 658, 688, 757, 1090-1091).
 * d_rgba: n x 4 u8 (RGBA8888, alpha ignored).  Output planar fp32 L,a,b (n each).  The
 * conversion is evaluated in fp64 (exact 256-entry sRGB-linearisation table, fp64 matrix,
 * fp64 cbrt) and rounded once to fp32.  d_lut256: 256 doubles, the linearised value of each
 * u8 level as the host computes it (so the curve is bit-identical to the reference's). */
int cs_rgba8_to_lab(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, const double *d_lut256,
                    float *d_L, float *d_a, float *d_b, void *stream);

/* Same conversion, not rounded: d_lab = n x 3 fp64 rows {L, a, b} — the array
 * simplify_colors_adaptive_distance hands to StandardScaler / DBSCAN (color_simplify.py:757). */
int cs_rgba8_to_lab_f64(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, const double *d_lut256,
                        double *d_lab, void *stream);

/* ---- K2/K3: one Lloyd iteration (assign + update) ----------------------------------
 * replaces sklearn lloyd_iter_chunked_dense + _update_chunk_dense
 *   (sklearn/cluster/_k_means_lloyd.pyx:23-218), reached from KMeans.fit at
 *   color_simplify.py:79-80, 669-675, 811-812, 992-993.
 * Features are three planar fp32 arrays (CIELAB L,a,b or any 3-feature space).
 * d_centers: K x 3 fp64 row-major.  d_labels (nullable): n x u8, label = first argmin_j
 * of ||x - c_j||^2 (lowest j on exact ties, as _k_means_lloyd.pyx:205-213).
 * d_sums: K x 3 fp64, d_counts: K fp64 (= sklearn's centers_new before averaging and
 * weight_in_clusters), overwritten.  d_inertia (nullable): 1 fp64, sum of squared distances
 * to the assigned centre.
 * feat_norm2_max: an upper bound on f0^2+f1^2+f2^2 over all n pixels (CS_LAB_NORM2_MAX for
 * planes written by cs_rgba8_to_lab, or the value cs_feature_norm2_max_f32 measures); the
 * fp32 keys are offset by it so that they stay positive.  A bound that is too small makes
 * the labels of the offending pixels undefined; too large only costs key precision.
 * flags: CS_LLOYD_EXACT_TIES re-evaluates every pixel whose two best fp32 distances are
 * within the fp32 error bound in fp64 (labels then equal the fp64 argmin); without it the
 * label of such a pixel may be either of the two (documented near-tie).
 * CS_LLOYD_CHAINED (fused / batched / multi-GPU iteration calls): the operation queued on `stream`
 * immediately before this call was a Lloyd launch of this library and the pixel buffers (planes,
 * packed pixels, feature table) have not been written since.  The kernel is then launched as a
 * programmatic dependent launch: its barrier set-up, accumulator zeroing and first tile loads overlap
 * the previous launch's combine + M-step tail, and it orders itself behind that launch before it reads
 * the centres (griddepcontrol.wait).  Results are identical with and without the flag.
 */
/* Exactness limit of the integer (RGBA8 / per-byte-table) paths: every consumer lane accumulates its pixels in
 * fp32 {sum, count} slots that are folded to fp64 once per CTA, so a slot stays exact while its sum is below
 * 2^24 — with 148 CTAs x 16 warps x (32 .. 2) lane copies per cluster that is ~19 G pixel-values per cluster and
 * GPU at K <= 16 (e.g. 75 MP of value 255 in ONE cluster) and ~1.2 G at K = 256.  The entry points refuse
 * (CS_ERR_ARG) packed-pixel launches whose worst case could exceed it: n * 255 > 2^24 * slots. */
#define CS_LLOYD_EXACT_TIES 1
#define CS_LLOYD_CHAINED 2
/* L in [0,100], a in [-86.2,98.3], b in [-107.9,94.5] over the sRGB gamut */
#define CS_LAB_NORM2_MAX 31400.0
int cs_lloyd_step_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                      int64_t n, const double *d_centers, int K, uint8_t *d_labels,
                      double *d_sums, double *d_counts, double *d_inertia, double feat_norm2_max,
                      int flags, void *stream);

/* max over pixels of f0^2+f1^2+f2^2 -> *d_out (1 fp64, overwritten). */
int cs_feature_norm2_max_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                             int64_t n, double *d_out, void *stream);

/* Same step on packed RGBA8 pixels with RGB as the three features (u8 -> exact integers).
 * replaces the same sklearn kernel as reached from simplify_colors_kmeans
 *   (color_simplify.py:44-80): pixels with alpha == 0 or r+g+b <= min_rgb_sum are skipped
 *   (label 255, no contribution) — the reference's `alpha > 0` and `mean(rgb) > 30|10` masks
 *   (color_simplify.py:44, 56-64; mean>30 <=> r+g+b>90).  min_rgb_sum < 0 keeps every
 *   opaque pixel.  Sums are exact integers. */
int cs_lloyd_step_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                        const double *d_centers, int K, uint8_t *d_labels, double *d_sums,
                        double *d_counts, double *d_inertia, int flags, void *stream);

/* fused step + finalize on RGBA8 pixels (see cs_lloyd_iter_f32). */
int cs_lloyd_iter_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                        const double *d_centers_in, int K, uint8_t *d_labels, double *d_sums,
                        double *d_counts, double *d_centers_out, double *d_stats, int flags,
                        void *stream);

/* Batched fused iteration: n_images independent k-means problems on packed RGBA8 images of
 * n_per_image pixels each (contiguous, n_per_image % 4 == 0) in one launch — one K x 3 centre table,
 * label map, sums / counts / stats block per image, laid out image-major:
 * d_centers_in/out [n_images][K][3], d_labels [n_images][n_per_image], d_sums [n_images][K][3],
 * d_counts [n_images][K], d_stats [n_images][4].  Same per-image semantics as cs_lloyd_iter_rgba8
 * (BASELINE config 4: a batch of images partitioned across GPUs, no collective). */
int cs_lloyd_iter_rgba8_batched(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n_per_image, int n_images,
                                int min_rgb_sum, const double *d_centers_in, int K, uint8_t *d_labels,
                                double *d_sums, double *d_counts, double *d_centers_out, double *d_stats,
                                int flags, void *stream);

/* General packed-pixel step: the three features of a pixel are d_lut3[b0], d_lut3[256+b1],
 * d_lut3[512+b2] (3 x 256 fp32 tables) of its first three bytes; byte 3 is alpha.  Used for
 * simplify_colors_hsv_clustering's weighted HSV features (color_simplify.py:969-981), each an
 * injective function of one u8.  mask_mode 0: skip when b0+b1+b2 <= min_bright; 1: skip when
 * b2 <= min_bright (the V > 30 | 10 filter, :956-963); alpha == 0 always skips.
 * d_inertia, d_centers_out/d_stats nullable (fused finalize when d_centers_out is given). */
int cs_lloyd_step_px8lut(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3, int mask_mode,
                         int min_bright, double feat_norm2_max, const double *d_centers, int K,
                         uint8_t *d_labels, double *d_sums, double *d_counts, double *d_inertia,
                         double *d_centers_out, double *d_stats, int flags, void *stream);

/* Optional hint for the planar-fp32 Lloyd entry points (cs_lloyd_step_f32 / iter / run / iter_f32_mg):
 * an axis-aligned box [h_lo3, h_hi3] that contains (most of) the features, e.g. CS_LAB_BOX_* for the output
 * of cs_rgba8_to_lab.  With a box set, CS_LLOYD_EXACT_TIES launches on >= 2^18 pixels may take the
 * grid-filtered assignment: a table of the <= 4 centres that can be nearest anywhere in each cell of a
 * ~10 000-cell grid over the box is built once per iteration (a second, tiny kernel), and the Lloyd kernel
 * evaluates four distances per pixel instead of K.  Its cost hardly depends on K (it is bound by
 * shared-memory bandwidth): slower than the full walk at K <= 8, faster from K = 9 up (1.17x at K = 16, 1.8x at
 * K = 32, 2.2x at K = 64) on shards large enough to pay for the table (>= 10^7 pixels at K <= 16, 2^22 at K <= 32, 2^21 above).
 * Policy (cs_lloyd_set_grid_policy): 0 (default) = use it where it is faster (those bounds), 1 = for
 * 4 <= K <= 64 on >= 2^18 pixels, -1 = never.
 * Labels are unchanged — the fp64 first minimum (sklearn/cluster/_k_means_lloyd.pyx:205-213) — for ANY input:
 * pixels outside the box fall into border cells that extend to infinity, so a wrong box costs speed, never
 * correctness.  NULL pointers clear the box. */
#define CS_LAB_BOX_LO0 0.0
#define CS_LAB_BOX_LO1 -87.0
#define CS_LAB_BOX_LO2 -108.5
#define CS_LAB_BOX_HI0 100.5
#define CS_LAB_BOX_HI1 99.0
#define CS_LAB_BOX_HI2 95.0
int cs_lloyd_set_feature_box(cs_ctx *ctx, const double *h_lo3, const double *h_hi3);
int cs_lloyd_set_grid_policy(cs_ctx *ctx, int policy);

/* M-step tail: replaces _relocate_empty_clusters_dense (detection only), _average_centers
 * and _center_shift (sklearn/cluster/_k_means_common.pyx:167-311) and the tolerance sum of
 * _kmeans_single_lloyd (sklearn/cluster/_kmeans.py:731-738).
 * d_centers_new[k] = d_sums[k] * (1/d_counts[k]); an empty cluster takes the new centre of
 * the heaviest cluster (first argmax).  d_stats: 4 fp64 = { sum_k shift_k^2, n_empty,
 * argmax_weight, total_weight }. */
int cs_lloyd_finalize(cs_ctx *ctx, const double *d_sums, const double *d_counts,
                      const double *d_centers_old, int K, double *d_centers_new,
                      double *d_stats, void *stream);

/* Fused single-GPU iteration: step + finalize in ONE kernel (the last block to finish does
 * the global combine and the M-step tail).  Equivalent to cs_lloyd_step_f32 followed by
 * cs_lloyd_finalize; d_centers_in and d_centers_out must not alias. */
int cs_lloyd_iter_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                      int64_t n, const double *d_centers_in, int K, uint8_t *d_labels,
                      double *d_sums, double *d_counts, double *d_centers_out,
                      double *d_stats, double feat_norm2_max, int flags, void *stream);

/* Device-side loop control: queue n_launch fused iterations back to back, ping-ponging the two centre
 * buffers (launch i reads a for even i, b for odd i, and writes the other), without a host round trip
 * between iterations.  d_ctl = 4 fp64 owned by the caller: [0] halt (0 = run; set to 1 by the iteration
 * whose sum of squared centre shifts is <= tol, to 2 by an iteration that found an empty cluster and was
 * therefore NOT counted — the host redoes it with cs_lloyd_relocate_*), [1] iterations completed
 * (incremented by every counted iteration), [2] tol, [3] unused.  A launch that finds [0] != 0 returns at
 * once.  After the batch the current centres are in a if (completed - completed_before) is even, else in b;
 * d_sums / d_counts / d_stats are those of the last iteration that ran.
 * replaces the per-iteration convergence test of _kmeans_single_lloyd (sklearn/cluster/_kmeans.py:705-738). */
int cs_lloyd_run_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2, int64_t n,
                     double *d_centers_a, double *d_centers_b, int K, double *d_sums, double *d_counts,
                     double *d_stats, double feat_norm2_max, int flags, int n_launch, double *d_ctl,
                     void *stream);
int cs_lloyd_run_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3, int mask_mode,
                     int min_bright, double feat_norm2_max, double *d_centers_a, double *d_centers_b, int K,
                     double *d_sums, double *d_counts, double *d_stats, int flags, int n_launch,
                     double *d_ctl, void *stream);

/* ---- multi-GPU: fused compute + exchange over NVLink peer memory -------------------------
 * One process per GPU.  Each rank creates a mailbox in its own HBM (cs_mg_create returns its 64-byte
 * cudaIpc handle), the host exchanges the handles (e.g. torch.distributed.all_gather) and every rank
 * maps its peers' mailboxes (cs_mg_connect; h_handles = world x 64 bytes in rank order).  After that
 * cs_lloyd_iter_f32_mg is cs_lloyd_iter_f32 over the UNION of all ranks' pixels: the last CTA of each
 * rank's kernel stores its (4K [+1])-double partial into every peer's mailbox, waits (bounded, 20 s)
 * for the peers' partials of the same iteration and sums them in rank order, so every rank ends with
 * bit-identical d_sums / d_counts / d_centers_out / d_stats — the per-iteration all-reduce of
 * SURVEY.md §8e without a collective launch.  All ranks must call it the same number of times, each
 * on its own GPU.  A timed-out wait poisons the outputs with NaN and is reported by cs_mg_error. */
int cs_mg_create(cs_ctx *ctx, int world, int rank, void *h_handle64);
int cs_mg_connect(cs_ctx *ctx, const void *h_handles);
int cs_mg_error(cs_ctx *ctx, unsigned long long *h_epoch);
int cs_mg_destroy(cs_ctx *ctx);
int cs_lloyd_iter_f32_mg(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                         int64_t n, const double *d_centers_in, int K, uint8_t *d_labels,
                         double *d_sums, double *d_counts, double *d_centers_out, double *d_stats,
                         double feat_norm2_max, int flags, void *stream);

/* Empty-cluster relocation: replaces _relocate_empty_clusters_dense
 * (sklearn/cluster/_k_means_common.pyx:167-211).  Finds, for every empty cluster in index
 * order, the sample farthest from its assigned (old) centre — descending distance, lowest
 * index on equal distances — and moves it, adjusting d_sums/d_counts in place.  Needs the
 * labels of the step that produced d_sums.  No-op when no cluster is empty or when every
 * distance is 0. */
int cs_lloyd_relocate_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                          int64_t n, const uint8_t *d_labels, const double *d_centers_old,
                          int K, double *d_sums, double *d_counts, void *stream);

/* One step of the same relocation for a row-SHARDED run (one process per GPU): this shard's farthest
 * labelled pixel that comes after the previous GLOBAL pick h_prev2 = {distance bits, global index}
 * (index ~0 = no previous pick) in the order (distance descending, global index ascending);
 * global index = index_base + local index.  h_out6 = {distance bits, global index (~0 = none here),
 * x, y, z as double bits, label}.  The ranks all-gather these records, take the first in that order
 * and apply it to the all-reduced sums / counts — the picks equal the unsharded run's.  Synchronous. */
int cs_lloyd_farthest_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                          int64_t n, const uint8_t *d_labels, const double *d_centers_old, int K,
                          uint64_t index_base, const uint64_t *h_prev2, uint64_t *h_out6, void *stream);

/* packed 4 x u8 pixels (RGBA or HSVA); d_lut3 nullable (identity) as in cs_lloyd_step_px8lut */
int cs_lloyd_relocate_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3,
                          const uint8_t *d_labels, const double *d_centers_old, int K, double *d_sums,
                          double *d_counts, void *stream);

/* ---- fp64 feature rows (rows64.cu): the per-pixel steps of simplify_colors_adaptive_distance -----------
 * cs_nn_argmin_rows64     replaces pairwise_distances_argmin_min(lab_flat[dark], lab_filtered) (color_simplify.py:861-867):
 *                         d_index[i] = first minimum over the n_ref reference rows of the direct fp64 squared distance.
 * cs_lloyd_step_rows64    E-step of KMeans.fit on n x 3 fp64 rows (the full-N fallback fit, :809-814): int32 labels
 *                         (fp64 first minimum) and the inertia (per-block partials added in block order).
 * cs_sum_by_label_rows64  the matching M-step sums: per-cluster fp64 sums and counts in a fixed order; feed them to
 *                         cs_lloyd_finalize. */
/* cs_kmeans_fit_rows64_small replaces KMeans(n_clusters=K, random_state=42, n_init=10, max_iter=100).fit(lab) on the
 *                         <= 5000 distinct sampled colours of simplify_colors_perceptual_fast (color_simplify.py:669-675;
 *                         sklearn/cluster/_kmeans.py:705-758): n_init complete Lloyd loops in ONE launch, one CTA per
 *                         initialisation, the n <= 5120 fp64 rows resident in shared memory — E-step (fp64 first
 *                         minimum), fixed-order sums, relocation of empty clusters, M-step tail, stop rule (labels repeat,
 *                         or sum(shift^2) <= tol), final E-step and inertia.  d_inits: n_init x K x 3 starting centres
 *                         (the k-means++ seeds of cs_kpp_*); outputs per initialisation: K x 3 centres, n u8 labels,
 *                         4 doubles {inertia, iterations run, 2 = labels repeated / 1 = tol / 0 = max_iter, empty
 *                         clusters seen in the last iteration}.  The caller keeps the run of least inertia as KMeans.fit does. */
int cs_kmeans_fit_rows64_small(cs_ctx *ctx, const double *d_rows, int64_t n, const double *d_inits, int n_init, int K,
                               int max_iter, double tol, double *d_centers, uint8_t *d_labels, double *d_stats, void *stream);
int cs_nn_argmin_rows64(cs_ctx *ctx, const double *d_query, int64_t n_query, const double *d_ref, int64_t n_ref,
                        int64_t *d_index, void *stream);
int cs_lloyd_step_rows64(cs_ctx *ctx, const double *d_rows, int64_t n, const double *d_centers, int K,
                         int32_t *d_labels, double *d_inertia, void *stream);
int cs_sum_by_label_rows64(cs_ctx *ctx, const double *d_rows, int64_t n, const int32_t *d_labels, int K,
                           double *d_sums, double *d_counts, void *stream);

/* ---- k-means++ seeding passes -------------------------------------------------------------
 * replaces the O(N) steps of sklearn's _kmeans_plusplus (sklearn/cluster/_kmeans.py:180-278), which
 * KMeans.fit runs before each of its n_init Lloyd runs (color_simplify.py:79-80, 992-993, 811-812).  The
 * RandomState stream and the per-round decisions stay on the host; the random numbers a fit consumes are
 * fixed in count, so the host advances the n_batch = n_init initialisations in LOCKSTEP and every pass is one
 * launch for all of them (n_batch = 1 is the plain case).  Samples: d_px = the COMPACTED selected pixels
 * (cs_select_compact_px8: row i = row i of the reference's filtered array) with the features d_lut768[b0],
 * d_lut768[256+b1], d_lut768[512+b2] (fp64) — or, with d_rows given (d_px / d_lut768 NULL), n x 3 fp64
 * feature ROWS (the standardised CIELAB rows of simplify_colors_adaptive_distance).  Distances are direct fp64.
 * Stacked arrays: d_closest [n_batch][n], d_tile_sums [n_batch][ntiles], d_block_pots [n_batch][pot_stride][8],
 * d_cands [n_batch][8][3] candidate features (on the DEVICE), d_pick [n_batch], d_tile / d_index [n_batch][8],
 * d_prefix_val [n_batch][16] (prefix in 0..7, value in 8..15), d_index_px [n_batch][8] packed pixels.
 * cs_kpp_eval_batched:   d_block_pots[b][r][t] = partial sum over block r of min(closest_i, |x_i - cand_t|^2)
 *                        for t < n_cand <= 8; *h_n_blocks (<= pot_stride, <= 4 * SM count) rows are written per
 *                        initialisation — `candidates_pot`, :247-253.
 * cs_kpp_pick_batched:   the potentials (the partial rows added in block order) -> d_pick[b] = first minimum
 *                        (np.argmin, :254-256), d_pot[b] = its potential.
 * cs_kpp_update_batched: closest_i = first ? d_i : min(closest_i, d_i) for the centre = candidate d_pick[b]
 *                        (NULL: candidate 0), and d_tile_sums[b][j] = sum of the new closest values of samples
 *                        [4096 j, 4096 (j+1)) — :255-258.
 * cs_kpp_locate_batched: for each of n_query values v: walk tile d_tile[b][q] in index order starting from the
 *                        running prefix; d_index[b][q] = first index whose cumulative sum is >= v
 *                        (np.searchsorted(np.cumsum(closest), v), :241-243, with the O(N) scan reduced to one tile:
 *                        the host picks the tile from the cumulative tile sums); optional d_index_px = the pixel there. */
int cs_kpp_eval_batched(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const double *d_lut768, const double *d_rows,
                        const double *d_cands, int n_cand, const double *d_closest, double *d_block_pots,
                        int pot_stride, int n_batch, int *h_n_blocks, void *stream);
int cs_kpp_update_batched(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const double *d_lut768, const double *d_rows,
                          const double *d_cands, const int *d_pick, int first, double *d_closest,
                          double *d_tile_sums, int n_batch, void *stream);
int cs_kpp_locate_batched(cs_ctx *ctx, const double *d_closest, int64_t n, const int64_t *d_tile,
                          const double *d_prefix_val, int n_query, const uint8_t *d_px, int64_t *d_index,
                          uint8_t *d_index_px, int n_batch, void *stream);
int cs_kpp_pick_batched(cs_ctx *ctx, const double *d_block_pots, int pot_stride, int n_blocks, int n_cand,
                        int n_batch, int *d_pick, double *d_pot, void *stream);
/* A whole round without the host (the round's uniforms are known in advance): cs_kpp_draw_batched forms
 * rand_vals = d_uniforms[b][q] * d_pot[b] (:239), finds each value's tile from the sequential cumulative sum of
 * d_tile_sums[b][.] and its crossing index inside the tile (np.searchsorted(np.cumsum(closest), v), :241-243, clipped
 * to n - 1, :244), and writes d_cand_index[b][q] and the candidate's features d_cands[b][q][3] — bit-identical to
 * the host computation + cs_kpp_locate_batched.  cs_kpp_record_batched stores the round's winner (d_pick[b] of
 * cs_kpp_pick_batched) as centre `slot` of initialisation b: d_index_out [n_batch][K], d_centers_out [n_batch][K][3]. */
int cs_kpp_draw_batched(cs_ctx *ctx, const double *d_closest, int64_t n, const double *d_tile_sums, const double *d_pot,
                        const double *d_uniforms, int n_query, const uint8_t *d_px, const double *d_lut768,
                        const double *d_rows, int64_t *d_cand_index, double *d_cands, int n_batch, void *stream);
int cs_kpp_record_batched(cs_ctx *ctx, const int64_t *d_cand_index, const double *d_cands, const int *d_pick, int slot,
                          int K, int n_batch, int64_t *d_index_out, double *d_centers_out, void *stream);

/* ---- K4: nearest centre + palette remap -----------------------------------------
 * replaces sklearn pairwise_distances_argmin_min + `quantized_rgb[mask] = centres[idx]` +
 * the alpha epilogue + np.dstack
 *   (color_simplify.py:543-557, 691-705, 1106-1121; sklearn/metrics/pairwise.py:711-845).
 * space: which feature space the K x 3 fp64 `d_centers` live in; the pixel is converted
 * from RGBA8 in registers.  Ties resolve to the lowest index (sklearn/utils/_heap.pyx:46-47).
 * d_palette_rgb: K x 3 u8 colours written for the winning index.  Pixels with alpha == 0
 * get RGB 0 (zero-initialised `quantized_rgb`, color_simplify.py:537, 685, 1110).
 * alpha_out = alpha, or (alpha > 128) * 255 when preserve_alpha == 0
 *   (color_simplify.py:93-97).  d_labels (nullable): n x u8 winning index (255 = skipped). */
#define CS_SPACE_RGB 0
#define CS_SPACE_LAB 1
#define CS_SPACE_HSV 2 /* OpenCV 8-bit HSV, H in [0,179] (color_simplify.py:1097-1098) */
int cs_assign_remap_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int space,
                          const double *d_lut256, const double *d_centers,
                          const uint8_t *d_palette_rgb, int K, int preserve_alpha,
                          uint8_t *d_rgba_out, uint8_t *d_labels, void *stream);
/* Which kernels cs_assign_remap_rgba8 uses for the RGB and LAB metrics (the labels are the same fp64 first
 * minimum on every path; HSV and K < 4 always take the direct kernel, and so do buffers that are not 16-byte
 * aligned): 0 (default) = by image size — direct below 2^22 pixels (2^21 for K >= 32), candidate table over 32^3 RGB
 * cells + three-phase tiles below 2^24, per-colour label table of the mixed cells + streaming pass from 2^24 up;
 * 1 = the three-phase tiles, 2 = the colour table, at any size; -1 = the direct kernel. */
int cs_remap_set_policy(cs_ctx *ctx, int policy);

/* "Valid pixel" convention of the label helpers below: when d_selpx (n packed 4 x u8 pixels,
 * nullable) is given, pixel i is valid iff alpha(d_selpx[i]) > 0 and its brightness passes
 * (mask_mode 0: b0+b1+b2 > min_bright; 1: b2 > min_bright; min_bright < 0 keeps all opaque) —
 * the same selection the masked Lloyd step applied; when d_selpx is NULL a pixel is valid iff
 * its label is not the 255 sentinel (ambiguous only for K = 256, where d_selpx must be given). */

/* Gather remap from precomputed labels: out[i] = palette[label[i]] for valid pixels with
 * alpha > 0, RGB 0 otherwise; the intended behaviour of color_simplify.py:90 and the gathers
 * at :1024, :870.  Alpha epilogue as cs_assign_remap_rgba8. */
int cs_remap_labels_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, const uint8_t *d_labels,
                          int64_t n, const uint8_t *d_selpx, int mask_mode, int min_bright,
                          const uint8_t *d_palette_rgb, int K, int preserve_alpha,
                          uint8_t *d_rgba_out, void *stream);

/* per-label sums: d_acc = K x 4 u64 {sum_r, sum_g, sum_b, count} over valid pixels,
 * overwritten — the "cluster centres in RGB space" of color_simplify.py:996-1000, 842-846. */
int cs_sum_by_label_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, const uint8_t *d_labels, int64_t n,
                          const uint8_t *d_selpx, int mask_mode, int min_bright, int K,
                          unsigned long long *d_acc, void *stream);
/* out[i] = valid(i) ? primary[i] : fallback[i]  (color_simplify.py:1009-1021). */
int cs_merge_labels_u8(cs_ctx *ctx, const uint8_t *d_primary, const uint8_t *d_fallback, int64_t n,
                       const uint8_t *d_selpx, int mask_mode, int min_bright, uint8_t *d_out,
                       void *stream);
/* d_matrix: 256 x 256 u32, overwritten: [a*256+b] = 1 iff some valid pixel has labels (a, b) in
 * the two maps — the evidence _is_same_clustering needs
 * (sklearn/cluster/_k_means_common.pyx:314-328, used by KMeans.fit's best-of-n_init). */
int cs_label_cooccurrence_u8(cs_ctx *ctx, const uint8_t *d_labels_a, const uint8_t *d_labels_b,
                             int64_t n, const uint8_t *d_selpx, int mask_mode, int min_bright,
                             uint32_t *d_matrix, void *stream);

/* ---- selection / sampling plumbing ---------------------------------------------------
 * replaces the NumPy boolean-index compactions `rgb[non_transparent]`, `rgb_flat[non_black_mask]`
 * (color_simplify.py:49-66, 439-466, 599-655, 946-967).  Order-preserving: the i-th selected
 * pixel is row i of the reference's compacted array.  d_out_px (capacity x 4 u8) and
 * d_out_index (capacity x i64 source positions) are each nullable; *d_count (u64) = number
 * selected (counting continues past capacity). */
int cs_select_compact_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, int mask_mode, int min_bright,
                          uint8_t *d_out_px, int64_t *d_out_index, int64_t capacity,
                          unsigned long long *d_count, void *stream);
/* per-byte histograms of the selected pixels, d_hist768[c*256 + v] (u64, overwritten): every
 * feature on this path is a function of one byte, so mean / variance of the feature columns
 * (KMeans.fit's centring and `tol`, sklearn/cluster/_kmeans.py:285-293, 1487-1490) follow
 * exactly from these counts. */
int cs_channel_hist_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, int mask_mode, int min_bright,
                        unsigned long long *d_hist768, void *stream);
/* out[i] = px[index[i]]: `rgb_flat[np.random.choice(...)]` (color_simplify.py:443-445, 633-635). */
int cs_gather_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const int64_t *d_index, int64_t m,
                  uint8_t *d_out_px, void *stream);

/* ---- K5: 24-bit colour histogram ---------------------------------------------------
 * replaces Pillow's create_pixel_hash (PIL/_imaging: Quant.c, reached from
 * Image.quantize at color_simplify.py:145, 201) and the counting half of np.unique.
 * d_hist: 2^24 u32 bins keyed (r<<16)|(g<<8)|b, ACCUMULATED into (caller zeroes).
 * alpha is ignored (Image.fromarray(rgb), color_simplify.py:144). */
int cs_hist_rgb24(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint32_t *d_hist,
                  void *stream);
/* Fold the 2^24 histogram to cells (r>>s, g>>s, b>>s): d_cells has 2^(3*(8-s)) u32 bins,
 * overwritten; d_ncells: 1 u32 = number of non-empty cells (Pillow's hash-table size at
 * scale s). */
int cs_hist_fold(cs_ctx *ctx, const uint32_t *d_hist, int shift, uint32_t *d_cells,
                 uint32_t *d_ncells, void *stream);
/* Compact the non-empty cells: d_keys/d_counts receive (cell key, pixel count) in ascending
 * key order; capacity = number reported by cs_hist_fold. */
int cs_hist_compact(cs_ctx *ctx, const uint32_t *d_cells, int64_t nbins, uint32_t *d_keys,
                    uint32_t *d_counts, uint32_t capacity, uint32_t *d_n, void *stream);

/* ---- host: Pillow MEDIANCUT box tree ---------------------------------------------
 * replaces Pillow's median_cut + the array heap of QuantHeap.c on the (<= 65536-cell)
 * histogram.  h_keys/h_counts: n cells at scale `shift` (key = (r<<2b)|(g<<b)|b, b = 8-shift
 * bits per channel).  Writes h_cell_box[i] = palette index (DFS leaf order, high side first)
 * of the box holding cell i, and *n_boxes.  Pure host code. */
int cs_median_cut_boxes(const uint32_t *h_keys, const uint32_t *h_counts, uint32_t n,
                        int shift, int n_colors, uint16_t *h_cell_box, int *n_boxes);

/* ---- K6: box means + nearest-palette map --------------------------------------------
 * replaces Pillow's compute_palette_from_median_cut and map_image_pixels_from_median_box.
 * cs_box_sums: from the 2^24 histogram and a 2^(3*(8-shift))-entry cell->box LUT (0xFFFF =
 * empty cell) accumulate per-box sums of the UNSCALED r,g,b and the pixel count:
 * d_box_acc = n_boxes x 4 u64 {sum_r, sum_g, sum_b, count}, overwritten. */
int cs_box_sums(cs_ctx *ctx, const uint32_t *d_hist, const uint16_t *d_cell_box, int shift,
                int n_boxes, unsigned long long *d_box_acc, void *stream);
/* pixel -> palette index with Pillow's rule: minimum integer squared RGB distance; the
 * pixel's own box wins a tie, otherwise the tied entry with the smallest (squared palette
 * distance from the own entry, index).  Writes RGBA8 (palette colour + alpha epilogue;
 * alpha is NOT used as a mask here, color_simplify.py:144-162) and optional u8 indices. */
int cs_palette_map_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n,
                         const uint16_t *d_cell_box, int shift, const uint8_t *d_palette_rgb,
                         int n_pal, int preserve_alpha, uint8_t *d_rgba_out,
                         uint8_t *d_index, void *stream);

/* ---- K7: posterize -------------------------------------------------------------------
 * replaces `(c // step) * step` per channel (color_simplify.py:255-261) + alpha epilogue.
 * Also marks the quantised colours present in d_present (2^24-bit bitmap, 2 MiB, caller
 * zeroes; nullable) — the set np.unique returns at color_simplify.py:274. */
int cs_posterize_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int step,
                       int preserve_alpha, uint8_t *d_rgba_out, uint32_t *d_present,
                       void *stream);

/* ---- K8: statistics ------------------------------------------------------------------
 * replaces np.unique(rgba rows) / alpha>0 count / mean / std of get_color_statistics
 * (color_simplify.py:363-376).  d_bitmap: 2^32-bit presence bitmap (512 MiB, caller zeroes)
 * keyed by the little-endian RGBA word.  d_acc: 8 u64, overwritten =
 * {n_opaque, sum_r, sum_g, sum_b, sum_r2, sum_g2, sum_b2, 0} over alpha > 0. */
int cs_stats_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint32_t *d_bitmap,
                   unsigned long long *d_acc, void *stream);
/* population count of a bitmap of n_words u32 -> *d_count (u64, overwritten). */
int cs_bitmap_popcount(cs_ctx *ctx, const uint32_t *d_bitmap, int64_t n_words,
                       unsigned long long *d_count, void *stream);
/* mask counts for simplify_colors_kmeans (color_simplify.py:44-64): d_acc = 4 u64,
 * overwritten = {alpha>0, alpha>0 && r+g+b>90, alpha>0 && r+g+b>30, 0}; marks the RGB
 * colours of pixels passing `r+g+b > min_rgb_sum && alpha>0` in d_present (2^24-bit bitmap,
 * nullable) for the unique-colour cap at color_simplify.py:69-70. */
int cs_mask_stats_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                        uint32_t *d_present, unsigned long long *d_acc, void *stream);

/* same counts on HSVA pixels (output of cs_rgba8_to_hsv8) for simplify_colors_hsv_clustering
 * (color_simplify.py:956-963, 984-985): d_acc = {alpha>0, alpha>0 && v>30, alpha>0 && v>10, 0};
 * marks the (h,s,v) triples of pixels with alpha>0 && v > min_v in d_present — the weighted
 * HSV feature rows np.unique counts are an injective function of (h,s,v). */
int cs_mask_stats_hsv8(cs_ctx *ctx, const uint8_t *d_hsva, int64_t n, int min_v, uint32_t *d_present,
                       unsigned long long *d_acc, void *stream);

/* ---- K9: RGB -> HSV (OpenCV 8-bit) -------------------------------------------------
 * replaces cv2.cvtColor(COLOR_RGB2HSV) on u8 (color_simplify.py:947, 1097-1098): integer
 * exact, H in [0,179].  Output n x 4 u8 {h, s, v, alpha}. */
int cs_rgba8_to_hsv8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint8_t *d_hsva,
                     void *stream);

/* ---- connected components of the simplified image ------------------
 * replaces the per-colour cv.connectedComponentsWithStats loop of analyze_regions
 * (app/processing/region_cleanup.py:48-88) with ONE labelling of all colours.
 * cs_ccl_label:   d_labels[i] = smallest linear index of the component of pixel i — pixels are joined
 *                 when both are opaque (alpha > 0), have equal RGB and are 4- / 8-neighbours; -1 for
 *                 transparent pixels.  width * height < 2^31.
 * cs_ccl_roots:   the component roots (d_labels[i] == i) in raster order: *d_count = number of
 *                 components; d_rank (nullable, n x i32) = component id at root pixels, -1 elsewhere;
 *                 d_roots (nullable, capacity x i32) = root index of each component.
 * cs_ccl_stats:   per component: pixel count, bbox {min x, min y, max x, max y} and the order key that
 *                 reproduces OpenCV's component numbering inside one colour mask — raster index of the
 *                 first pixel (connectivity 4) or of the first 2x2 block (connectivity 8).
 * cs_ccl_extract: the per-colour arrays analyze_regions returns: d_out_labels[i] = d_comp_local[c] and
 *                 d_out_mask[i] = 255 where pixel i belongs to a component c with d_comp_color[c] == color,
 *                 0 elsewhere (either output nullable). */
int cs_ccl_label(cs_ctx *ctx, const uint8_t *d_rgba, int width, int height, int connectivity,
                 int32_t *d_labels, void *stream);
int cs_ccl_roots(cs_ctx *ctx, const int32_t *d_labels, int64_t n, int32_t *d_rank, int32_t *d_roots,
                 int64_t capacity, unsigned long long *d_count, void *stream);
int cs_ccl_stats(cs_ctx *ctx, const int32_t *d_labels, const int32_t *d_rank, int width, int height,
                 int connectivity, int n_comp, uint32_t *d_area, int32_t *d_bbox,
                 unsigned long long *d_order_key, void *stream);
int cs_ccl_extract(cs_ctx *ctx, const int32_t *d_labels, const int32_t *d_rank, int64_t n,
                   const int32_t *d_comp_color, const int32_t *d_comp_local, int color,
                   int32_t *d_out_labels, uint8_t *d_out_mask, void *stream);

/* ---- host memory at the caller's boundary --------------------------
 * The reference hands images over as NumPy views of QImage bits (app/utils/qt_image.py:9-32,
 * app/ui/main_window.py:546-553, 604, 612).  cs_host_register page-locks such a caller-owned
 * buffer IN PLACE (cudaHostRegister, portable), so that every later upload from / download into
 * it is a single DMA at PCIe rate instead of the driver's staged pageable copy; the buffer is
 * not copied or moved.  cs_host_unregister must be called before the buffer is freed. */
int cs_host_register(void *h_ptr, size_t bytes);
int cs_host_unregister(void *h_ptr);
/* Upload from PAGEABLE caller memory (an ordinary NumPy array, as the colour panel passes it,
 * app/ui/main_window.py:596-601) without page-locking it: host threads of the library copy `bytes` from h_src,
 * 8 MB at a time, into a ring of page-locked buffers owned by the context, and every filled buffer leaves as one
 * asynchronous DMA to d_dst on `stream` while the next one is filled.  Returns when h_src has been read completely
 * (the caller may reuse it); the last DMAs may still be in flight on `stream`.  CS_HOST_THREADS sets the number of
 * copy threads (default: three quarters of the CPUs the process may run on, 2..12). */
int cs_host_upload(cs_ctx *ctx, const void *h_src, size_t bytes, void *d_dst, void *stream);

/* ---- host-buffer convenience (the e2e path: H2D + kernels + D2H inside) --------------
 * One call = what the colour panel's "process" click needs for LAB k-means from given
 * initial centres: uploads h_rgba (n x 4 u8, pinned or pageable), converts to LAB, runs
 * `n_iter` fused Lloyd iterations (stops early when sum shift^2 <= tol), writes labels and
 * final centres back to host.  h_centers: K x 3 fp64 in/out.  h_labels nullable.
 * Synchronous: waits for work queued earlier on the device, uploads in chunks on a copy stream
 * while the LAB conversion of the previous chunk runs on a compute stream, and returns when the
 * results are in host memory. */
int cs_host_lab_kmeans(cs_ctx *ctx, const uint8_t *h_rgba, int64_t n, const double *h_lut256,
                       double *h_centers, int K, int n_iter, double tol, int flags,
                       uint8_t *h_labels, int *n_iter_done, double *h_inertia);

#ifdef __cplusplus
}
#endif
#endif /* COLORSIMPLIFY_H_ */
