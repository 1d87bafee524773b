/*
 * ss_post.h -- C ABI of the steps either side of the ASW / GSW hot path, exported by the
 * same libsspassive.so as include/ss_passive.h and sharing its conventions (return codes, ss_last_error(),
 * "host" entry points copy in/out and synchronise, "device" entry points take device pointers plus a
 * cudaStream_t passed as void* and only enqueue).  No CPU fallback.
 *
 * The reference implements these steps by calling OpenCV; each entry point names the reference call site it
 * replaces (paths relative to the reference checkout) and reproduces OpenCV's arithmetic bit for bit for the
 * argument types the reference passes.
 */
#ifndef SS_POST_H
#define SS_POST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- disparity -> 3-D points ------------------------------------------------------------- */

/* cv2.reprojectImageTo3D(disparityMap, Q) -- simplestereo/points.py:176 (getAdimensional3DPoints) and
 * simplestereo/_rigs.py:628 (RectifiedStereoRig.get3DPoints).  disp: int16 [H][W]; Q: 16 doubles, row-major 4x4;
 * points: float32 [H][W][3].  Division by w == 0 yields +-inf / nan as OpenCV does (handleMissingValues=False). */
int ss_reproject(const int16_t *disp, int width, int height, const double *Q, float *points);
int ss_reproject_device(const int16_t *d_disp, int width, int height, const double *Q, float *d_points, void *stream);

/* StereoASW.compute (simplestereo/passive.py:72-92) followed by get3DPoints (_rigs.py:569-628) in one call: the
 * disparity map stays on the device between the WTA tail and the reprojection.  out_disp may be NULL. */
int ss_asw_compute_points(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                          int min_disp, double gamma_c, double gamma_p, int consistent, const double *Q,
                          int16_t *out_disp, float *out_points);

/* ---- display post-filter ------------------------------------------------------------------ */

/* cv2.normalize(disp, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1) followed by cv2.applyColorMap --
 * examples/010 StereoMatchingTsukuba.py:44-45 (and every other example script).  lut_bgr: 256 x 3 bytes (the
 * caller supplies the colour map, e.g. COLORMAP_JET); gray [H][W] and bgr [H][W][3] may each be NULL (not both). */
int ss_normalize_colormap(const int16_t *disp, int width, int height, const uint8_t *lut_bgr, uint8_t *gray,
                          uint8_t *bgr);
int ss_normalize_colormap_device(const int16_t *d_disp, int width, int height, const uint8_t *d_lut_bgr,
                                 uint8_t *d_gray, uint8_t *d_bgr, void *stream);

/* ---- rectification remap ------------------------------------------------------------------ */

/* cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR) with the default BORDER_CONSTANT (0) --
 * RectifiedStereoRig.rectifyImages, simplestereo/_rigs.py:564-565.  src: uint8 BGR [src_h][src_w][3];
 * mapx / mapy: float32 [dst_h][dst_w] (cv2.initUndistortRectifyMap(..., cv2.CV_32FC1), _rigs.py:540-541);
 * dst: uint8 BGR [dst_h][dst_w][3]. */
int ss_remap_linear(const uint8_t *src, int src_width, int src_height, const float *mapx, const float *mapy,
                    int dst_width, int dst_height, uint8_t *dst);
int ss_remap_linear_device(const uint8_t *d_src, int src_width, int src_height, const float *d_mapx,
                           const float *d_mapy, int dst_width, int dst_height, uint8_t *d_dst, void *stream);

/* ---- point-cloud export ------------------------------------------------------------------- */

/* simplestereo.points.exportPLY(points3D, filepath, referenceImage=None, precision=6) -- simplestereo/points.py:10-80.
 * Host-side ASCII writer, byte-identical to the reference's per-point Python loop, formatted on all host threads.
 * points: n x 3 float32 (points_are_double == 0) or float64; shape / ndims: the original array shape for the header
 * comment; bgr: optional n x 3 uint8 (written as R G B); intensity: used only when bgr is NULL -- n values, int64
 * (intensity_kind 1) or float64 (2), or n x 3 int64 B G R triples of a non-uint8 integer colour image (3: the
 * reference formats them with "{:d}" whatever their range, points.py:53-55). */
int ss_export_ply(const void *points, int points_are_double, long long n, const long long *shape, int ndims,
                  const uint8_t *bgr, const void *intensity, int intensity_kind, const char *path, int precision);

#ifdef __cplusplus
}
#endif
#endif /* SS_POST_H */
