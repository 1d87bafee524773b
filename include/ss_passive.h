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
 *     enqueue work on the caller's CURRENT device -- nothing is synchronised.  Calls may use different streams: the cached
 *     scratch is shared per device, so each call's stream first waits (cudaStreamWaitEvent) for the previous call's work.
 *   - rows [row_begin,row_end) select an image-row stripe (rows are independent jobs in the
 *     reference, _passive.cpp:372-374); output buffers of *_rows/_device calls hold only the stripe.
 *   - the library keeps no host pointer after a call returns.  Device scratch is cached per
 *     device and grows on demand (ss_shutdown releases it).  Calls are serialised per device (internal mutex).
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

/* Select the CUDA device the host entry points of this process run on (one process per GPU).  Optional: without it
 * they use the caller's current device.  Replaces the reference's thread-pool set-up
 * (std::thread::hardware_concurrency() workers, _passive.cpp:351-355, :751-754). */
int ss_init(int device);

/* Single-process multi-GPU: the host entry points (ss_asw_compute, ss_gsw_compute, *_rows) shard the image rows over
 * `devices[0..n)` -- one host thread and one context (stream, cached scratch) per device, every device reading the rows it
 * needs from the caller's arrays and writing its stripe straight into the caller's output, no collective -- exactly the
 * unit the reference's workers pop from their queue (a row index, _passive.cpp:372-374).  devices == NULL or n <= 0 selects
 * every visible device.  (The NCCL communicators of ss_*_compute_multi_device are created by its first call:
 * ncclCommInitAll over this list, libnccl.so.2 loaded at run time.)
 * The device-resident entry points always run on the caller's CURRENT device, whatever this list holds. */
int ss_init_devices(const int *devices, int n);
/* number of devices in the list (0 before ss_init / ss_init_devices) */
int ss_device_count(void);

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

/* Single-process multi-GPU, device-resident.  Arrays of n = ss_device_count() pointers, entry k on device k of the
 * ss_init_devices list: d_img1[k] / d_img2[k] copies of the full images, d_out[k] a buffer of n * ceil(H/n) x W int16
 * (equal stripes: the all-gather needs them), streams[k] a cudaStream_t of device k (streams == NULL: the library's own
 * streams, synchronised before returning).  Device k computes rows [k*S, (k+1)*S) into its slot of d_out[k] and ONE
 * grouped in-place ncclAllGather leaves the whole map on every device. */
int ss_asw_compute_multi_device(const uint8_t *const *d_img1, const uint8_t *const *d_img2, int width, int height,
                                int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p, int consistent,
                                int16_t *const *d_out, void *const *streams);
int ss_gsw_compute_multi_device(const uint8_t *const *d_img1, const uint8_t *const *d_img2, int width, int height,
                                int win_size, int max_disp, int min_disp, int gamma, float f_max, int iterations, int bins,
                                int16_t *const *d_out, void *const *streams);

/* ---- disparity-range sharding (device buffers) ----------------------------------------- */

/* The kernel and its disparity-chunk grid are chosen from the CALL's range [min_disp,max_disp] (chunks start at
 * min_disp + k * chunk), never from the sub-range: every cost a shard computes is bit-identical to the unsharded call's,
 * so merged shards reproduce the unsharded map exactly.  Shards aligned to the chunk size (128 disparities when the call
 * spans more than 64) waste no work.
 * Evaluate only disparities [disp_begin,disp_end] (inclusive, clipped to [min_disp,max_disp]) and
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
/* The same for the image-row stripe [row_begin,row_end): every output holds only the stripe's rows (full-size frames are
 * checked against the oracle a few rows at a time). */
int ss_asw_stages_rows(const uint8_t *img1, const uint8_t *img2, int width, int height,
                       int win_size, int max_disp, int min_disp, double gamma_c, double gamma_p,
                       int consistent, int row_begin, int row_end, int16_t *out_left, int16_t *out_right,
                       uint8_t *out_invalid, int16_t *out_final, float *out_cost);
int ss_gsw_stages_rows(const uint8_t *img1, const uint8_t *img2, int width, int height,
                       int win_size, int max_disp, int min_disp, int gamma, float f_max,
                       int iterations, int bins, int row_begin, int row_end, int16_t *out_left,
                       int16_t *out_right, uint8_t *out_invalid, int16_t *out_final,
                       float *out_cost_left, float *out_cost_right);
/* GSW: out_cost_left / out_cost_right are the left- and right-reference volumes, both indexed by
 * the LEFT column:  (y*W + x)*D + (disp - min_disp). */
int ss_gsw_stages(const uint8_t *img1, const uint8_t *img2, int width, int height,
                  int win_size, int max_disp, int min_disp, int gamma, float f_max,
                  int iterations, int bins, int16_t *out_left, int16_t *out_right,
                  uint8_t *out_invalid, int16_t *out_final, float *out_cost_left, float *out_cost_right);

/* BGR -> CIELab exactly as the kernels see it (float32 L,a,b per pixel, H*W*3): the conversion stage of
 * ColorConversion::ImageFromBGR2Lab (headers/colorconversion.hpp:18-86) on its own, for parity tests. */
int ss_debug_lab(const uint8_t *img, int width, int height, float *out_lab);

/* Which aggregation kernel served the last call on `device` (< 0: the device of the host entry points): 1 = k_aggregate_tc
 * (tensor-core denominators), 2 = k_aggregate_ws (all CUDA cores), 0 = none yet; *disp_chunk (may be NULL) receives the
 * disparity chunk it ran with.  Lets the parity tests assert that a case reached the kernel it is meant to cover. */
int ss_debug_last_kernel(int device, int *disp_chunk);

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
