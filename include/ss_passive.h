/*
 * ss_passive.h -- C ABI of libsspassive.so, the B200 (sm_100a) replacement for the
 * ASW / GSW hot path of decadenza/SimpleStereo.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only -- no Python, numpy or torch types.
 *
 * Conventions
 *   - images are BGR uint8, H x W x 3, C-contiguous (the reference assumes this without
 *     checking, _passive.cpp:333-334);  disparity maps are int16, H x W.
 *   - all functions return SS_OK (0) or a negative SS_ERR_* code; ss_last_error() gives the text.
 *   - "host" entry points take host pointers and do H2D / compute / D2H internally;
 *     "device" entry points take device pointers and a cudaStream_t (passed as void*) and only
 *     enqueue work -- nothing is synchronised.
 *   - rows [row_begin,row_end) select an image-row stripe (rows are independent jobs in the
 *     reference, _passive.cpp:372-374); output buffers of *_rows/_device calls hold only the stripe.
 *   - the library keeps no host pointer after a call returns.  Device scratch is cached per
 *     process and grows on demand (ss_shutdown releases it).
 *   - there is no CPU fallback: without a CUDA device every compute call fails with SS_ERR_CUDA.
 */
#ifndef SS_PASSIVE_H
#define SS_PASSIVE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_OK            0
#define SS_ERR_FORMAT   -1   /* -> ValueError("Invalid input format!")                  _passive.cpp:304,712 */
#define SS_ERR_TYPE     -2   /* -> TypeError("Wrong type input!")                       _passive.cpp:312,720 */
#define SS_ERR_DIMS     -3   /* -> ValueError("Wrong image dimensions!")                _passive.cpp:319,727 */
#define SS_ERR_WINSIZE  -4   /* -> ValueError("winSize must be a positive odd number!") _passive.cpp:323,731 */
#define SS_ERR_PARAM    -5   /* rejected parameter the reference would turn into NaN / out-of-bounds reads
                                (gamma <= 0, minDisparity < 0) -> ValueError */
#define SS_ERR_CUDA     -6   /* CUDA runtime failure or no device -> RuntimeError(ss_last_error()) */
#define SS_ERR_NOMEM    -7

/* ---- process-wide context ------------------------------------------------------------- */

/* Select the CUDA device used by this process (one process per GPU).  Optional: the first
 * compute call initialises device 0 / the current device.  Replaces the reference's thread-pool
 * set-up (std::thread::hardware_concurrency() workers, _passive.cpp:351-355, :751-754). */
int ss_init(int device);

/* Free every cached device buffer.  (The reference never frees anything, _passive.cpp:338-358.) */
int ss_shutdown(void);

/* Text of the last error raised on the calling thread's context. */
const char *ss_last_error(void);

/* ABI version, bumped on any signature change. */
int ss_abi_version(void);

/* ---- whole-call entry points (host buffers) -------------------------------------------- */

/* _passive.computeASW(img1, img2, winSize, maxDisparity, minDisparity, gammaC, gammaP[, consistent])
 * -- _passive.cpp:293-400, called from simplestereo/passive.py:88-90.
 * Disparity candidates are the inclusive range [min_disp, max_disp] (_passive.cpp:56). */
int ss_asw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height,
                   int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                   int consistent, int16_t *out_disp);

/* _passive.computeGSW(img1, img2, winSize, maxDisparity, minDisparity, gamma, fMax, iterations, bins)
 * -- _passive.cpp:703-774, called from simplestereo/passive.py:153-156.
 * gamma is an int and bins is unused, as upstream (_passive.cpp:706, :410). */
int ss_gsw_compute(const uint8_t *img1, const uint8_t *img2, int width, int height,
                   int win_size, int max_disp, int min_disp, int gamma, float f_max,
                   int iterations, int bins, int16_t *out_disp);

/* Row-stripe variants: full images in, rows [row_begin,row_end) out ((row_end-row_begin) x W).
 * One stripe is what one std::thread worker processes in the reference (a row job,
 * _passive.cpp:29-32); it is the multi-GPU sharding unit. */
int ss_asw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height,
                        int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                        int consistent, int row_begin, int row_end, int16_t *out_rows);
int ss_gsw_compute_rows(const uint8_t *img1, const uint8_t *img2, int width, int height,
                        int win_size, int max_disp, int min_disp, int gamma, float f_max,
                        int iterations, int bins, int row_begin, int row_end, int16_t *out_rows);

/* ---- device-resident entry points (device buffers, caller's stream) -------------------- */

/* Same computation; d_img1/d_img2 are device copies of the full BGR images, d_out_rows a device
 * buffer of (row_end-row_begin) x W int16.  Work is enqueued on `stream` (a cudaStream_t). */
int ss_asw_compute_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height,
                          int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                          int consistent, int row_begin, int row_end, int16_t *d_out_rows, void *stream);
int ss_gsw_compute_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height,
                          int win_size, int max_disp, int min_disp, int gamma, float f_max,
                          int iterations, int bins, int row_begin, int row_end, int16_t *d_out_rows,
                          void *stream);

/* ---- disparity-range sharding (device buffers) ----------------------------------------- */

/* Evaluate only disparities [disp_begin,disp_end] (inclusive, clipped to [min_disp,max_disp]) and
 * return the per-pixel winners as packed keys  (float_bits(cost) << 32) | disparity  (smaller key
 * wins, which is also the reference's smallest-disparity tie-break, _passive.cpp:90-93):
 *   d_best_left  uint64[rows x W]  left-reference winners  (_passive.cpp:54-98)
 *   d_best_right uint64[rows x W]  right-reference winners (_passive.cpp:209-248), may be NULL
 *                                  when !consistent.
 * Keys of different shards are merged with an element-wise min (ss_merge_keys_device) and turned
 * into the final map by ss_finalize_keys_device. */
int ss_asw_partial_device(const uint8_t *d_img1, const uint8_t *d_img2, int width, int height,
                          int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                          int consistent, int row_begin, int row_end, int disp_begin, int disp_end,
                          uint64_t *d_best_left, uint64_t *d_best_right, void *stream);
/* d_keys[0] = min over k of d_keys[k], element-wise, n elements each, laid out back to back
 * (the shape an all-gather leaves them in). */
int ss_merge_keys_device(uint64_t *d_keys, int n_shards, long long n, void *stream);
/* winners -> int16 map: WTA decode, L-R invalidation (_passive.cpp:251-252) and occlusion fill
 * (:258-285).  d_best_right may be NULL (no consistency check). */
int ss_finalize_keys_device(const uint64_t *d_best_left, const uint64_t *d_best_right, int width,
                            int rows, int min_disp, int16_t *d_out_rows, void *stream);

/* ---- staged outputs for parity adjudication (host buffers; any output may be NULL) ------ */

/* out_left   int16[H*W]  stage 1, left-reference WTA map
 * out_right  int16[H*W]  stage 2, right-reference map (selected left column - xr)
 * out_invalid uint8[H*W] stage 3, 1 where the L-R check invalidated the pixel
 * out_final  int16[H*W]  stage 4, filled map (== ss_asw_compute output when consistent)
 * out_cost   float[H*W*D] aggregated cost volume, index (y*W + x)*D + (disp - min_disp),
 *            +INF where the pair is not evaluated.  Stages 2-4 need consistent != 0. */
int ss_asw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height,
                  int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                  int consistent, int16_t *out_left, int16_t *out_right, uint8_t *out_invalid,
                  int16_t *out_final, float *out_cost);
/* GSW: out_cost_left / out_cost_right are the left- and right-reference volumes, both indexed by
 * the LEFT column:  (y*W + x)*D + (disp - min_disp). */
int ss_gsw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height,
                  int win_size, int max_disp, int min_disp, int gamma, float f_max,
                  int iterations, int bins, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost_left, float *out_cost_right);

/* ---- instrumentation ------------------------------------------------------------------- */

/* When enabled, every aggregation-kernel launch is bracketed by CUDA events on its stream. */
int ss_profile_enable(int on);
/* Sum of aggregation-kernel durations (ms) and launch counts since the last reset.  Synchronises
 * the recorded events.  total_launches counts every kernel this library launched. */
int ss_profile_read(double *agg_ms, long long *agg_launches, long long *total_launches);
int ss_profile_reset(void);
/* Measured FP32 FFMA issue peak of the current device in TFLOP/s (2 flop per FFMA): the roofline
 * denominator of the aggregation kernel, which is FP32-pipe bound.  Runs a register-only
 * FFMA kernel for a few milliseconds on `stream` and synchronises it. */
int ss_measure_fp32_peak(double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SS_PASSIVE_H */
