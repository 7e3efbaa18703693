/*
 * lens_b200.h -- C ABI of liblens_b200.so, the B200 (sm_100a) implementation of
 * the LENS inference hot path (event binning -> IAF spiking network ->
 * similarity matrix -> sequence matching -> top-N / Recall@N).
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = bad argument (see lens_last_error()),
 *     >0 = the cudaError_t that was raised;
 *   - every pointer is a DEVICE pointer unless it says "host"; the caller owns
 *     all buffers, the library owns only what lives inside an opaque handle
 *     (plus stream-ordered scratch, cudaMallocAsync / cudaFreeAsync on the given
 *     stream, inside lens_seqmatch_topk when it splits the places of a query);
 *   - every function is asynchronous and ordered on the cudaStream_t it is given
 *     (pass as void*: 0 = legacy default stream); none synchronises the device;
 *   - handles are not thread-safe (one per stream / rank);
 *   - no C++ types, no torch types, no exceptions cross this boundary;
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Hard limits (a violation is reported as a negative return code, never silently):
 *   lens_snn_create        I <= 1024, F <= 992; the tensor-core path needs 6 * 128 * Fp + 3 * 64 * Fp bytes of
 *                          shared memory (Fp = F rounded up to 32: F <= 224 today), larger F runs on the CUDA-core path
 *   hidden spikes          <= LENS_MAX_SPIKE (127) per neuron and timestep (int8 transport); larger counts are
 *                          clipped and counted (lens_snn_get_overflow)
 *   lens_seqmatch_topk     N <= 64, 1 <= L <= min(Q, P)
 *   lens_recall / _bounds  at most 8 values of N per call
 *   lens_pr_counts         Qo <= 8192 (one CTA, shared-memory resident)
 *   lens_sad_matrix        npix <= 65 793 (sums stay exact in fp32)
 *   lens_topn_merge        W * N <= 512
 *   lens_bin_events        t_us ascending (lens_check_sorted_u32), x / y 16-byte aligned
 *   lens_snn_forward_float leaves IAF#0 possibly away from rest: the raster fast paths (and the tensor-core hidden
 *                          layer) stay disabled for the handle until lens_snn_reset
 *
 * Reference sites (relative to the reference repo root) each entry replaces are
 * cited per function.
 */
#ifndef LENS_B200_H
#define LENS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LENS_B200_VERSION_MAJOR 0
#define LENS_B200_VERSION_MINOR 1

/* Largest per-step spike count a neuron may emit and stay exact (int8 operand). */
#define LENS_MAX_SPIKE 127

/* ---- diagnostics -------------------------------------------------------- */
int lens_version(int *major, int *minor);
/* thread-local, never NULL, valid until the next failing call on this thread */
const char *lens_last_error(void);
/* number of SMs of the current device (grid sizing helper for callers) */
int lens_device_sm_count(int *n_sm);
/* kernels launched by this library since it was loaded (host counter; bench evidence) */
int lens_launch_count(int64_t *n_launches);

/* ---- K1: event binning + pooling ----------------------------------------
 * Replaces lens/collect_data.py:186-202 (event_collector + create_images; same
 * loop in lens/tools/manual_eventframe_generator.py:6-14) and the one-hot
 * strided pooling conv of lens/run_model.py:130-137.
 *
 * Events are SoA, sorted by time: t_us[n], x[n], y[n] (polarity already merged,
 * lens/run_speck.py:266).  Window w covers [t0 + w*window_us, t0 + (w+1)*window_us).
 * An event is kept when (x - roi_x0, y - roi_y0) lies inside the roi x roi crop;
 * it increments frame[(yr - index_shift) mod roi][(xr - index_shift) mod roi]
 * (index_shift = 1 is the reference's `frame[y-1, x-1]`).  Counts are stored as
 * uint8: wrap_u8 = 1 wraps modulo 256 (`astype(np.uint8)`), 0 saturates at 255.
 * pooled[w][i*dims + j] = frame[k*i + c][k*j + c], dims = roi / k, c = (k/2) - 1
 * (c = 0 when k == 1).
 *   frames      [n_win][roi][roi] u8   (nullable)
 *   pooled      [n_win][dims*dims] u8  (nullable)
 *   win_events  [n_win] i32 in-ROI events per window (nullable; the reference
 *               skips empty windows, collect_data.py:194,201-202)
 *   win_offsets [n_win + 1] i64 scratch/out: first event index of every window
 */
int lens_bin_events(const uint32_t *t_us, const uint16_t *x, const uint16_t *y, int64_t n_events,
                    uint32_t t0_us, uint32_t window_us, int roi_x0, int roi_y0, int roi, int k,
                    int index_shift, int wrap_u8, uint8_t *frames, uint8_t *pooled,
                    int32_t *win_events, int64_t *win_offsets, int64_t n_win, void *stream);

/* Precondition of lens_bin_events: t_us ascending (the window ranges are found by binary search, like the
 * reference's drain-every-timebin loop sees the events in arrival order, lens/collect_data.py:186-191).
 * Sets *unsorted (device int) to 1 if some t_us[i+1] < t_us[i], else 0.  One streaming pass over t_us
 * (4 B/event), therefore separate from the binning call; lens_b200.ops.bin_events(check_sorted=True)
 * runs it and raises.                                                                                  */
int lens_check_sorted_u32(const uint32_t *t_us, int64_t n, int *unsorted, void *stream);

/* Pooling alone (frames already exist, e.g. decoded PNGs):
 * frames [n][roi][roi] u8 -> pooled [n][dims*dims] u8.  lens/run_model.py:130-137. */
int lens_pool_frames(const uint8_t *frames, int64_t n, int roi, int k, uint8_t *pooled,
                     void *stream);

/* ---- K2 + K3: the spiking network --------------------------------------
 * Stands where `self.sinabs_model` stands (lens/run_model.py:139-156,238):
 *   IAF#0 -> Linear(I->F, no bias) -> IAF#1 -> Linear(F->P, no bias) -> IAF#2
 * with sinabs IAFSqueeze semantics (spike_threshold = thr, MultiSpike,
 * MembraneSubtract, min_v_mem = v_min), lens/src/blitnet.py:59-64 weights.
 * The handle owns fixed-point copies of the weights and the membrane potentials
 * v0[B][I], v1[B][F], v2[B][P], which persist across calls exactly like the
 * reference's (it never calls reset_states()).
 *   W_feat [F][I] f32, W_out [P][F] f32, U [T][I] f32 (the sub-sampled columns of
 *   dataset.py:120-121's seed-50 `torch.rand(T, roi*roi)`; nullable if only
 *   lens_snn_forward_float is used).
 * *n_inexact (host, nullable) receives the number of weights that needed rounding
 * to the 46-bit-per-row fixed-point grid (0 => every contraction is exact).      */
int lens_snn_create(int I, int F, int P, int T, float thr, float v_min, const float *W_feat,
                    const float *W_out, const float *U, int max_streams, void **handle,
                    int64_t *n_inexact, void *stream);
int lens_snn_destroy(void *handle);
/* sinabs Network.reset_states(): zero all membrane potentials. */
int lens_snn_reset(void *handle, void *stream);
/* copy membrane potentials out (any pointer nullable): v0 [B][I], v1 [B][F], v2 [B][P] */
int lens_snn_get_state(void *handle, float *v0, float *v1, float *v2, void *stream);
/* overflow[0] (device, i64) counts per-step spike counts that exceeded LENS_MAX_SPIKE */
int lens_snn_get_overflow(void *handle, int64_t *overflow, void *stream);

/* Optional per-kernel timing (bench.py's roofline): when enabled every feature- / output-layer
 * launch is bracketed by cudaEvents on its own stream.  lens_snn_get_timing waits for the recorded
 * events (host pointers), returns the accumulated milliseconds and launch counts, and clears them. */
int lens_snn_set_timing(void *handle, int enable);
int lens_snn_get_timing(void *handle, float *feature_ms, float *output_ms, int64_t *n_feature,
                        int64_t *n_output);

/* mode for lens_snn_forward: which output-layer kernel runs */
#define LENS_SNN_AUTO 0   /* pick by problem size                      */
#define LENS_SNN_SIMT 1   /* event-driven CUDA-core kernel             */
#define LENS_SNN_TC   2   /* tcgen05 digit-plane kernel (batched)      */

/* The per-query loop of lens/run_model.py:229-246 for B independent streams of Q
 * queries each, fed by the raster of lens/src/dataset.py:14-51,118-125:
 *   p = pooled / 255;  input spike at step t = (U[t][i] < p[i]);  T steps per query;
 *   counts[b][q][p] = number of output spikes of place p during query q.
 * Continues from the handle's current state (carry-over across queries and calls).
 *   pooled        [B][Q][I] u8
 *   counts        [B][Q][P] f32
 *   hidden_steps  [B][Q*T][F]  u8, nullable (debug / tests)
 *   out_steps     [B][Q*T][P]  u8, nullable (debug / tests)                      */
int lens_snn_forward(void *handle, const uint8_t *pooled, int B, int Q, float *counts,
                     uint8_t *hidden_steps, uint8_t *out_steps, int mode, void *stream);

/* Same for the sub-range [b0, b0 + nb) of the handle's streams (pooled / counts / debug buffers start
 * at stream b0): lets a caller pipeline host->device copies of one group of streams with the
 * computation of the previous group.  b0 must be even (streams are tiled in pairs). */
int lens_snn_forward_range(void *handle, const uint8_t *pooled, int b0, int nb, int Q, float *counts,
                           uint8_t *hidden_steps, uint8_t *out_steps, int mode, void *stream);

/* Operator seam `sinabs_model(x)` (lens/run_model.py:238) for an arbitrary float
 * raster: x [B][steps][I] f32 (already pooled), spikes_out [B][steps][P] f32.     */
int lens_snn_forward_float(void *handle, const float *x, int B, int steps, float *spikes_out,
                           void *stream);

/* ---- K4: sequence matching + top-N --------------------------------------
 * Replaces lens/run_model.py:248-254 (conv2d with eye(L), / L, transpose) and the
 * selection half of lens/src/metrics.py:183-226 (argsort(0)[-K:]).
 *   S [B][Q][P] f32 similarity (spike counts).  Qo = Q - L + 1, Po = P - L + 1.
 *   D[b][r][q] = (sum_{j<L} S[b][q+j][r+j]) / L
 *   D_out   [B][Po][Qo] f32, nullable (rows = database, as the reference returns it)
 *   top_val [B][Qo][N] f32, top_idx [B][Qo][N] i32: the N largest D[b][.][q],
 *           ordered by (value desc, index desc) == np.argsort(kind='stable')[-N:][::-1];
 *           missing entries (N > Po) are -inf / -1.
 * L >= 1 (the reference's L == 0 branch returns S itself and is handled by the caller). */
int lens_seqmatch_topk(const float *S, int B, int Q, int P, int L, int N, float *D_out,
                       float *top_val, int32_t *top_idx, void *stream);

/* Recall@N counters (lens/src/metrics.py:213-224 + lens/run_model.py:301-302).
 * Ground truth is either dense, gt_dense [B or 1][Po][Qo] u8 (gt_stream_stride = Po*Qo
 * or 0 to share one matrix), or a band: gt_center [B][Qo] i32 (-1 = query has no
 * match) with |idx - center| <= gt_tol counting as a hit.  Exactly one is non-NULL.
 *   ns      [n_ns] host ints, ascending, each <= N (e.g. 1,5,10,15,20,25)
 *   hits    [n_ns] i64 device, ACCUMULATED (zero it first): queries with a hit in top-n
 *   n_valid [1]    i64 device, ACCUMULATED: queries that have at least one positive  */
int lens_recall(const int32_t *top_idx, int B, int Qo, int Po, int N, const uint8_t *gt_dense,
                int64_t gt_stream_stride, const int32_t *gt_center, int gt_tol, const int *ns,
                int n_ns, int64_t *hits, int64_t *n_valid, void *stream);

/* Final top-N merge for a PLACE-sharded database (BASELINE config 5: "NCCL all-gather of database
 * representations + top-N merge"): every shard ranks its own range of places with lens_seqmatch_topk, the
 * W per-shard lists of a query are gathered, and the global list is the N best of their union under the
 * same order as lens/src/metrics.py:218 under the deterministic rule (value desc, place index desc).
 *   val, idx [W][M][N] device (idx = GLOBAL place index, -1 = empty slot), M = streams * queries,
 *   out_val, out_idx [M][N] device.  W * N <= 512.                                                       */
int lens_topn_merge(const float *val, const int32_t *idx, int W, int64_t M, int N, float *out_val,
                    int32_t *out_idx, void *stream);

/* Tie-aware bounds of Recall@N (lens/src/metrics.py:218 ranks with numpy's default, unstable argsort, so the
 * reference's number is defined only up to the order of equal similarities): over every such order,
 *   lo [n_ns] i64 device, ACCUMULATED: queries that hit in the top ns[i] whatever the order,
 *   hi [n_ns] i64 device, ACCUMULATED: queries that hit for some order,
 *   n_valid [1] i64 device, ACCUMULATED: queries with at least one positive (metrics.py:214-216).
 * D, GT [Po][Qo] (rows = database, the matrices recallAtK receives), ns host ints ascending.          */
int lens_recall_bounds(const float *D, const uint8_t *GT, int Po, int Qo, const int *ns, int n_ns,
                       int64_t *lo, int64_t *hi, int64_t *n_valid, void *stream);

/* Precision-recall counters of lens/src/metrics.py:21-139 (createPR, matching='single'), used by
 * lens/run_model.py:319-327 with n_thresh = 100.  S and GT are [Po][Qo] (rows = database), exactly the
 * matrices the reference passes (after its transposes).  Per query column the best match is the FIRST
 * row attaining the maximum (np.argmax); thresholds are np.linspace(max, min, n_thresh) over the best
 * similarities, evaluated in float64 like numpy.
 *   tp, fp [n_thresh] i64 device: true / false positives with best >= threshold
 *   gtp    [1] i64 device: number of queries that have a positive at all (GT.any(0))             */
int lens_pr_counts(const float *S, const uint8_t *GT, int Po, int Qo, int n_thresh, int64_t *tp,
                   int64_t *fp, int64_t *gtp, void *stream);
/* The same for matching='multi' (lens/src/metrics.py:63-91): every entry of S counts, thresholds are
 * np.linspace(S.max(), S.min(), n_thresh), gtp = count_nonzero(GT).  tp, fp, gtp are overwritten.          */
int lens_pr_counts_multi(const float *S, const uint8_t *GT, int Po, int Qo, int n_thresh, int64_t *tp,
                         int64_t *fp, int64_t *gtp, void *stream);

/* Sum-of-absolute-differences baseline, lens/src/sad.py:25-42: dist[q][r] = sum_p |a[q][p] - b[r][p]|
 * (torch.cdist(a, b, p=1) on float32 copies of uint8 frames; the sums are integers < 2^24, hence exact
 * in fp32 in any order).  a [Q][npix] u8 (query frames), b [R][npix] u8 (reference frames),
 * dist [Q][R] f32.  lens_reciprocal: out[i] = 1.0f / in[i] (IEEE, numpy's `1 / dist`).             */
int lens_sad_matrix(const uint8_t *a, const uint8_t *b, int Q, int R, int npix, float *dist, void *stream);
int lens_reciprocal(const float *in, int64_t n, float *out, void *stream);

/* Online matcher of the event-driven deployment, lens/run_speck.py:155-226.
 *
 * lens_online_accumulate: one readout (run_speck.py:159-164, 188-198): sum[i] += (int32)counts[i]; when
 * row_out != NULL also row_out[i] = floor(sum[i] / div) (`vector // 4`, the averaged sequence row).
 *   sum [P] i32 device (running spike counts, NOT cleared between rows - the reference clears it only
 *   after a match), counts [P] f32 device (integer-valued spike counts of one readout interval).
 * lens_online_match: run_speck.py:200-204: result = scipy.signal.convolve2d(seq.T, eye(L), mode='same') / L
 * and its per-column first-maximum argmax:
 *   result[p][r] = (1/L) * sum_{a<L} seq[r + o - a][p + o - a],  o = (L-1)/2, zero outside the matrix.
 *   seq [R][P] i32 device (sequence rows, oldest first), result [P][R] f64 device (integer sums, one
 *   IEEE float64 division), argmax [R] i32 device.  1 <= L, R <= 64.                                */
int lens_online_accumulate(int32_t *sum, const float *counts, int P, int div, int32_t *row_out, void *stream);
int lens_online_match(const int32_t *seq, int R, int P, int L, double *result, int32_t *argmax, void *stream);

/* Event-driven frame representation, lens/tools/dvstools.py:279-349 (tool='simple_rep').
 *
 * lens_event_windows: index ranges of the frames.  A frame that starts at time c holds the events with
 * |t - c| <= interval (float64, the reference's own expression); the first later event that is not a hot
 * pixel closes it, is dropped, and its timestamp is the next frame's c.  Frame 0 starts at `start`
 * (events with t < start are skipped), or at t[0] when use_first_event != 0 (the reference's offset == 0
 * case).  The frame still open at the end of the stream is not reported (the reference never saves it).
 *   t [n] f64 device, seconds, ascending; x, y [n] u16 device
 *   lut [sensor_h][sensor_w] i16 device: -2 hot pixel (event ignored entirely), -1 no slot, >= 0 slot
 *   win_begin, win_end [max_windows] i64 device, win_t0 [max_windows] f64 device, n_windows [1] i64 device
 *   unsorted: nullable [1] i32 device, set to 1 when t is not ascending (the ranges are then meaningless)
 * lens_bin_events_lut: frames[w][slot] = (weight * #{events of range w whose pixel maps to slot}) mod 256
 * (`frame_data[index] += accum_factor` on a uint8 vector, dvstools.py:310-322).
 *   counts [n_windows][n_slots] u32 device scratch, frames [n_windows][n_slots] u8 device;
 *   events_per_window_hint only sizes the grid.  x / y 16-byte aligned, n_slots <= 8192.              */
int lens_event_windows(const double *t, const uint16_t *x, const uint16_t *y, int64_t n_events, const int16_t *lut,
                       int sensor_w, int sensor_h, int use_first_event, double start, double interval,
                       int64_t max_windows, int64_t *win_begin, int64_t *win_end, double *win_t0,
                       int64_t *n_windows, int *unsorted, void *stream);
int lens_bin_events_lut(const uint16_t *x, const uint16_t *y, const int16_t *lut, int sensor_w, int sensor_h,
                        const int64_t *win_begin, const int64_t *win_end, int64_t n_windows, int n_slots, int weight,
                        int64_t events_per_window_hint, uint32_t *counts, uint8_t *frames, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LENS_B200_H */
