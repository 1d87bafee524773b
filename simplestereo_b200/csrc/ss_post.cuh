// ss_post.cuh -- the steps either side of the ASW / GSW hot path, part of libsspassive.so.
// Included at the end of ss_passive.cu (same translation unit: shares the context, the scratch cache and the
// error plumbing).  Declared in include/ss_post.h.
//
//   k_reproject           disparity -> 3-D points, cv2.reprojectImageTo3D as called by points.py:176 and
//                         _rigs.py:628 (double homogeneous transform, float32 rounding before AND after the
//                         multiplication by 1/w, exactly OpenCV 4.x).  12 B written per pixel: HBM-write bound.
//   k_minmax_i16 +
//   k_normalize_colormap  cv2.normalize(.., 0, 255, NORM_MINMAX, CV_8UC1) + cv2.applyColorMap (examples/010:44-45).
//   k_remap_linear        cv2.remap(.., INTER_LINEAR), BORDER_CONSTANT 0 (_rigs.py:564-565): 5-bit fixed-point
//                         coordinates, 15-bit weights, exactly OpenCV's arithmetic.  Gather, HBM/L2 bound.
//
// All three are O(W*H) and bit-exact against OpenCV (tests/test_gpu_post.py, oracle/post_oracle.py).

namespace {

struct Q16 {
    double q[16];
};

__device__ __forceinline__ void reproject_px(const Q16 &Q, int x, int y, int disp, float *o) {
    const double d = (double)disp, xd = (double)x, yd = (double)y;
    // Q * (x, y, d, 1): row sums left to right, no contraction (OpenCV's Matx44d * Vec4d)
    auto row = [&](int k) {
        const double *q = Q.q + 4 * k;
        return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], xd), __dmul_rn(q[1], yd)), __dmul_rn(q[2], d)), q[3]);
    };
    const double iw = __ddiv_rn(1.0, row(3));                             // Vec3f /= w is *= 1/w in OpenCV
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float X = __double2float_rn(row(k));                        // Vec3f = Vec3d(homogeneous point)
        o[k] = __double2float_rn(__dmul_rn((double)X, iw));
    }
}

// VEC = 4: four pixels per thread, one 8-byte load and three 16-byte stores (needs W % 4 == 0)
template <int VEC>
__global__ void k_reproject(const int16_t *__restrict__ disp, float *__restrict__ pts, int W, int H, Q16 Q) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    if (VEC == 4) {
        const short4 d4 = *reinterpret_cast<const short4 *>(disp + i);
        float o[12];
        reproject_px(Q, x, y, d4.x, o);
        reproject_px(Q, x + 1, y, d4.y, o + 3);
        reproject_px(Q, x + 2, y, d4.z, o + 6);
        reproject_px(Q, x + 3, y, d4.w, o + 9);
        float4 *dst = reinterpret_cast<float4 *>(pts + 3 * i);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        dst[2] = make_float4(o[8], o[9], o[10], o[11]);
    } else {
        reproject_px(Q, x, y, disp[i], pts + 3 * i);
    }
}

// mm[0] = min(v), mm[1] = min(-v): both start at 0x7f7f7f7f (one cudaMemsetAsync), max = -mm[1]
__global__ void k_minmax_i16(const int16_t *__restrict__ disp, long long n, int *__restrict__ mm, int vec4) {
    int lo = 32767, hi = -32768;
    const long long n4 = vec4 ? n / 4 : 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const short4 d4 = *reinterpret_cast<const short4 *>(disp + 4 * q);
        lo = min(min(lo, (int)d4.x), min(min((int)d4.y, (int)d4.z), (int)d4.w));
        hi = max(max(hi, (int)d4.x), max(max((int)d4.y, (int)d4.z), (int)d4.w));
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = disp[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // one pair of global atomics per block: same-address atomics serialise in L2
    __shared__ int s_lo[32], s_hi[32];
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) { lo = min(lo, s_lo[k]); hi = max(hi, s_hi[k]); }
        atomicMin(mm, lo);
        atomicMin(mm + 1, -hi);
    }
}

__global__ void k_normalize_colormap(const int16_t *__restrict__ disp, long long n, const int *__restrict__ mm,
                                     const uint8_t *__restrict__ lut, uint8_t *__restrict__ gray, uint8_t *__restrict__ bgr,
                                     int vec4) {
    __shared__ uint32_t slut[256];                                 // B | G << 8 | R << 16 per level
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
        slut[k] = (uint32_t)lut[3 * k] | ((uint32_t)lut[3 * k + 1] << 8) | ((uint32_t)lut[3 * k + 2] << 16);
    __syncthreads();
    // cv::normalize: scale = (255 - 0) * (1 / (smax - smin)) or 0, shift = 0 - smin * scale, in double
    const double smin = (double)mm[0], smax = -(double)mm[1];
    const double range = smax - smin;
    const double scale = 255.0 * (range > 2.220446049250313e-16 ? 1.0 / range : 0.0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, b = (float)shift;
    // convertTo: one rounding (fma), round half to even, saturate to uint8
    auto level = [&](int v) { return min(255, max(0, __float2int_rn(fmaf((float)v, a, b)))); };
    const long long n4 = vec4 ? n / 4 : 0;                     // groups of 4 pixels: 8-byte load, 4 + 12-byte stores
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const short4 d4 = *reinterpret_cast<const short4 *>(disp + 4 * q);
        const int g0 = level(d4.x), g1 = level(d4.y), g2 = level(d4.z), g3 = level(d4.w);
        if (gray) *reinterpret_cast<uint32_t *>(gray + 4 * q) = (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)g2 << 16) | ((uint32_t)g3 << 24);
        if (bgr) {
            const uint32_t c0 = slut[g0], c1 = slut[g1], c2 = slut[g2], c3 = slut[g3];
            uint32_t *o = reinterpret_cast<uint32_t *>(bgr + 12 * q);
            o[0] = c0 | (c1 << 24);
            o[1] = (c1 >> 8) | (c2 << 16);
            o[2] = (c2 >> 16) | (c3 << 8);
        }
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int g = level(disp[i]);
        if (gray) gray[i] = (uint8_t)g;
        if (bgr) {
            const uint32_t cc = slut[g];
            bgr[3 * i + 0] = (uint8_t)cc;
            bgr[3 * i + 1] = (uint8_t)(cc >> 8);
            bgr[3 * i + 2] = (uint8_t)(cc >> 16);
        }
    }
}

// cvRound(v * 32) the way OpenCV's remap sees it on x86: round half to even; NaN / inf / out of int range -> INT_MIN
__device__ __forceinline__ int cvround32(float v) {
    const float f = v * 32.0f;
    if (!(fabsf(f) < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rn(f);
}

__global__ void k_remap_linear(const uint8_t *__restrict__ src, int sw, int sh, const float *__restrict__ mapx,
                               const float *__restrict__ mapy, int dw, int dh, uint8_t *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const size_t i = (size_t)y * dw + x;
    const int sx = cvround32(mapx[i]), sy = cvround32(mapy[i]);
    const int ax = sx & 31, ay = sy & 31;
    const int ix = min(32767, max(-32768, sx >> 5)), iy = min(32767, max(-32768, sy >> 5));
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    int acc[3] = {0, 0, 0};
    auto tap = [&](int yy, int xx, int w) {
        if (w == 0 || xx < 0 || xx >= sw || yy < 0 || yy >= sh) return;         // BORDER_CONSTANT, value 0
        const uint8_t *p = src + 3 * ((size_t)yy * sw + xx);
        acc[0] += w * p[0];
        acc[1] += w * p[1];
        acc[2] += w * p[2];
    };
    tap(iy, ix, w00);
    tap(iy, ix + 1, w01);
    tap(iy + 1, ix, w10);
    tap(iy + 1, ix + 1, w11);
    uint8_t *o = dst + 3 * i;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(255, (acc[c] + (1 << 14)) >> 15);
}

int post_reproject_enqueue(Ctx &c, const int16_t *d_disp, int W, int H, const double *Q, float *d_pts, cudaStream_t st) {
    Q16 q;
    for (int k = 0; k < 16; ++k) q.q[k] = Q[k];
    const bool vec = (W % 4 == 0) && ((uintptr_t)d_disp % 8 == 0) && ((uintptr_t)d_pts % 16 == 0);
    if (vec) {
        dim3 b(128), g((W / 4 + 127) / 128, H);
        k_reproject<4><<<g, b, 0, st>>>(d_disp, d_pts, W, H, q);
    } else {
        dim3 b(256), g((W + 255) / 256, H);
        k_reproject<1><<<g, b, 0, st>>>(d_disp, d_pts, W, H, q);
    }
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    return SS_OK;
}

int post_colormap_enqueue(Ctx &c, const int16_t *d_disp, int W, int H, const uint8_t *d_lut, uint8_t *d_gray, uint8_t *d_bgr,
                          cudaStream_t st) {
    int rc;
    if ((rc = ensure(c.post_mm, 2 * sizeof(int)))) return rc;
    CU_TRY(cudaMemsetAsync(c.post_mm.p, 0x7f, 2 * sizeof(int), st));
    const long long n = (long long)W * H;
    const int vec4 = ((uintptr_t)d_disp % 8 == 0) && (!d_gray || (uintptr_t)d_gray % 4 == 0) && (!d_bgr || (uintptr_t)d_bgr % 4 == 0);
    const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, 148 * 16);
    k_minmax_i16<<<blocks, 256, 0, st>>>(d_disp, n, (int *)c.post_mm.p, vec4);
    k_normalize_colormap<<<blocks, 256, 0, st>>>(d_disp, n, (const int *)c.post_mm.p, d_lut, d_gray, d_bgr, vec4);
    CU_TRY(cudaGetLastError());
    c.total_launches += 2;
    return SS_OK;
}

int post_check_dims(int W, int H) {
    if (W <= 0 || H <= 0) return fail(SS_ERR_DIMS, "Wrong image dimensions!");
    return SS_OK;
}

}  // namespace

extern "C" {

int ss_reproject_device(const int16_t *d_disp, int width, int height, const double *Q, float *d_points, void *stream) {
    if (!d_disp || !Q || !d_points) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(width, height);
    if (rc) return rc;
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    return post_reproject_enqueue(c, d_disp, width, height, Q, d_points, (cudaStream_t)stream);
}

int ss_reproject(const int16_t *disp, int width, int height, const double *Q, float *points) {
    if (!disp || !Q || !points) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(width, height);
    if (rc) return rc;
    CtxLock L(default_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    const size_t n = (size_t)width * height;
    if ((rc = ensure(c.out, n * 2 + 2))) return rc;
    if ((rc = ensure(c.post_pts, n * 12))) return rc;
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    CU_TRY(cudaMemcpyAsync(c.out.p, disp, n * 2, cudaMemcpyHostToDevice, st));
    if ((rc = post_reproject_enqueue(c, (const int16_t *)c.out.p, width, height, Q, (float *)c.post_pts.p, st))) return rc;
    CU_TRY(cudaMemcpyAsync(points, c.post_pts.p, n * 12, cudaMemcpyDeviceToHost, st));
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

int ss_asw_compute_points(const uint8_t *img1, const uint8_t *img2, int width, int height, int win_size, int max_disp,
                          int min_disp, double gamma_c, double gamma_p, int consistent, const double *Q, int16_t *out_disp,
                          float *out_points) {
    if (!img1 || !img2 || !Q || !out_points) return fail(SS_ERR_FORMAT, "Invalid input format!");
    const Call q = asw_call(width, height, win_size, max_disp, min_disp, gamma_c, gamma_p, consistent, 0, height);
    int rc = validate(q);
    if (rc) return rc;
    CtxLock L(default_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    const size_t nimg = (size_t)width * height * 3, npx = (size_t)width * height;
    if ((rc = ensure(c.img1, nimg))) return rc;
    if ((rc = ensure(c.img2, nimg))) return rc;
    if ((rc = ensure(c.out, npx * 2 + 2))) return rc;
    if ((rc = ensure(c.post_pts, npx * 12))) return rc;
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    CU_TRY(cudaMemcpyAsync(c.img1.p, img1, nimg, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c.img2.p, img2, nimg, cudaMemcpyHostToDevice, st));
    Outputs o;
    o.d_final = (int16_t *)c.out.p;
    if ((rc = run_device(c, q, (const uint8_t *)c.img1.p, (const uint8_t *)c.img2.p, o, st))) return rc;
    // the disparity map never leaves the device between the WTA tail and the reprojection
    if ((rc = post_reproject_enqueue(c, (const int16_t *)c.out.p, width, height, Q, (float *)c.post_pts.p, st))) return rc;
    if (out_disp) CU_TRY(cudaMemcpyAsync(out_disp, c.out.p, npx * 2, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out_points, c.post_pts.p, npx * 12, cudaMemcpyDeviceToHost, st));
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

int ss_normalize_colormap_device(const int16_t *d_disp, int width, int height, const uint8_t *d_lut_bgr, uint8_t *d_gray,
                                 uint8_t *d_bgr, void *stream) {
    if (!d_disp || !d_lut_bgr || (!d_gray && !d_bgr)) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(width, height);
    if (rc) return rc;
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    if ((rc = scratch_begin(c, (cudaStream_t)stream))) return rc;
    if ((rc = post_colormap_enqueue(c, d_disp, width, height, d_lut_bgr, d_gray, d_bgr, (cudaStream_t)stream))) return rc;
    return scratch_end(c, (cudaStream_t)stream);
}

int ss_normalize_colormap(const int16_t *disp, int width, int height, const uint8_t *lut_bgr, uint8_t *gray, uint8_t *bgr) {
    if (!disp || !lut_bgr || (!gray && !bgr)) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(width, height);
    if (rc) return rc;
    CtxLock L(default_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    const size_t n = (size_t)width * height;
    if ((rc = ensure(c.out, n * 2 + 2))) return rc;
    if ((rc = ensure(c.post_a, n * 4 + 768 + 8))) return rc;      // [lut 768 | bgr 3n, padded to 4 | gray n]
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    uint8_t *d_lut = (uint8_t *)c.post_a.p, *d_bgr = d_lut + 768, *d_gray = d_bgr + ((3 * n + 3) & ~(size_t)3);
    CU_TRY(cudaMemcpyAsync(c.out.p, disp, n * 2, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_lut, lut_bgr, 768, cudaMemcpyHostToDevice, st));
    if ((rc = post_colormap_enqueue(c, (const int16_t *)c.out.p, width, height, d_lut, d_gray, d_bgr, st))) return rc;
    if (gray) CU_TRY(cudaMemcpyAsync(gray, d_gray, n, cudaMemcpyDeviceToHost, st));
    if (bgr) CU_TRY(cudaMemcpyAsync(bgr, d_bgr, 3 * n, cudaMemcpyDeviceToHost, st));
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

int ss_remap_linear_device(const uint8_t *d_src, int src_width, int src_height, const float *d_mapx, const float *d_mapy,
                           int dst_width, int dst_height, uint8_t *d_dst, void *stream) {
    if (!d_src || !d_mapx || !d_mapy || !d_dst) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(src_width, src_height);
    if (rc) return rc;
    if ((rc = post_check_dims(dst_width, dst_height))) return rc;
    CtxLock L(current_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    dim3 b(128), g((dst_width + 127) / 128, dst_height);
    k_remap_linear<<<g, b, 0, (cudaStream_t)stream>>>(d_src, src_width, src_height, d_mapx, d_mapy, dst_width, dst_height, d_dst);
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    return SS_OK;
}

int ss_remap_linear(const uint8_t *src, int src_width, int src_height, const float *mapx, const float *mapy, int dst_width,
                    int dst_height, uint8_t *dst) {
    if (!src || !mapx || !mapy || !dst) return fail(SS_ERR_FORMAT, "Invalid input format!");
    int rc = post_check_dims(src_width, src_height);
    if (rc) return rc;
    if ((rc = post_check_dims(dst_width, dst_height))) return rc;
    CtxLock L(default_device());
    if (L.rc) return L.rc;
    Ctx &c = *L.c;
    const size_t ns = (size_t)src_width * src_height * 3, nd = (size_t)dst_width * dst_height;
    if ((rc = ensure(c.img1, ns))) return rc;
    if ((rc = ensure(c.post_pts, nd * 8))) return rc;             // mapx | mapy
    if ((rc = ensure(c.post_a, nd * 3))) return rc;
    cudaStream_t st = c.stream;
    if ((rc = scratch_begin(c, st))) return rc;
    float *d_mx = (float *)c.post_pts.p, *d_my = d_mx + nd;
    CU_TRY(cudaMemcpyAsync(c.img1.p, src, ns, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_mx, mapx, nd * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_my, mapy, nd * 4, cudaMemcpyHostToDevice, st));
    dim3 b(128), g((dst_width + 127) / 128, dst_height);
    k_remap_linear<<<g, b, 0, st>>>((const uint8_t *)c.img1.p, src_width, src_height, d_mx, d_my, dst_width, dst_height,
                                    (uint8_t *)c.post_a.p);
    CU_TRY(cudaGetLastError());
    c.total_launches += 1;
    CU_TRY(cudaMemcpyAsync(dst, c.post_a.p, nd * 3, cudaMemcpyDeviceToHost, st));
    if ((rc = scratch_end(c, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// ss_export_ply -- ASCII PLY writer, byte-identical to simplestereo.points.exportPLY (points.py:10-80), which
// formats every point in a Python loop (seconds per 4K frame).  Host code: the points come back from the device
// as float32 [n][3]; formatting is split over the host threads, the file is written in order.
// ------------------------------------------------------------------------------------------

namespace {

// Python's "{:.{p}f}".format(v): correctly rounded like glibc's printf, but nan never carries a sign
inline void ply_fmt(std::string &o, double v, int precision, int width) {
    char buf[64];
    const char *fmt = width >= 0 ? "%*f" : "%.*f";                  // "{:{p}f}": width p, default precision 6 (points.py:77)
    const int arg = width >= 0 ? width : precision;
    int n;
    if (std::isnan(v)) n = snprintf(buf, sizeof(buf), "%*s", width >= 0 ? width : 0, "nan");
    else n = snprintf(buf, sizeof(buf), fmt, arg, v);
    if (n < (int)sizeof(buf)) {
        o.append(buf, (size_t)std::max(n, 0));
        return;
    }
    // large magnitudes / precisions (up to 300 decimals of a 1e308 double): format into a buffer of the exact size
    std::string big((size_t)n + 1, '\0');
    if (std::isnan(v)) snprintf(&big[0], big.size(), "%*s", width >= 0 ? width : 0, "nan");
    else snprintf(&big[0], big.size(), fmt, arg, v);
    o.append(big.data(), (size_t)n);
}

}  // namespace

extern "C" int ss_export_ply(const void *points, int points_are_double, long long n, const long long *shape, int ndims,
                             const uint8_t *bgr, const void *intensity, int intensity_kind, const char *path,
                             int precision) {
    if (!points || !path || n < 0 || ndims < 0 || (ndims > 0 && !shape) || precision < 0 || precision > 300)
        return fail(SS_ERR_FORMAT, "Invalid input format!");
    if (intensity && (intensity_kind < 1 || intensity_kind > 3)) return fail(SS_ERR_FORMAT, "Invalid input format!");
    const bool bgr64 = !bgr && intensity && intensity_kind == 3;    // BGR triples of a non-uint8 integer image, as int64
    FILE *f = fopen(path, "w");
    if (!f) return fail(SS_ERR_PARAM, std::string("cannot open ") + path + " for writing");
    std::string hdr = "ply\nformat ascii 1.0\ncomment SimpleStereo point cloud export\ncomment Original array shape ";
    for (int k = 0; k < ndims; ++k) hdr += (k ? "x" : "") + std::to_string(shape[k]);
    hdr += "\nelement vertex " + std::to_string(n) + "\nproperty double x\nproperty double y\nproperty double z\n";
    if (bgr || bgr64) hdr += "property uchar red\nproperty uchar green\nproperty uchar blue\n";
    else if (intensity) hdr += intensity_kind == 1 ? "property int intensity\n" : "property float intensity\n";
    hdr += "end_header\n";
    fwrite(hdr.data(), 1, hdr.size(), f);

    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const long long per = 1 << 16;                                 // points per work item
    const long long nitems = (n + per - 1) / per;
    const unsigned nthreads = (unsigned)std::min<long long>(hw, std::max<long long>(nitems, 1));
    bool ok = true;
    for (long long base = 0; base < nitems && ok; base += nthreads) {
        const unsigned cnt = (unsigned)std::min<long long>(nthreads, nitems - base);
        std::vector<std::string> out(cnt);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < cnt; ++t) {
            th.emplace_back([&, t]() {
                const long long i0 = (base + t) * per, i1 = std::min(n, i0 + per);
                std::string &o = out[t];
                o.reserve((size_t)(i1 - i0) * (3 * (precision + 8) + 16));
                for (long long i = i0; i < i1; ++i) {
                    for (int k = 0; k < 3; ++k) {
                        const double v = points_are_double ? static_cast<const double *>(points)[3 * i + k]
                                                           : (double)static_cast<const float *>(points)[3 * i + k];
                        if (k) o += ' ';
                        ply_fmt(o, v, precision, -1);
                    }
                    if (bgr) {                                     // BGR -> RGB (points.py:55)
                        o += ' '; o += std::to_string((int)bgr[3 * i + 2]);
                        o += ' '; o += std::to_string((int)bgr[3 * i + 1]);
                        o += ' '; o += std::to_string((int)bgr[3 * i + 0]);
                    } else if (bgr64) {                            // "{:d}" of whatever the integers hold (points.py:53-55)
                        const long long *c = static_cast<const long long *>(intensity) + 3 * i;
                        o += ' '; o += std::to_string(c[2]);
                        o += ' '; o += std::to_string(c[1]);
                        o += ' '; o += std::to_string(c[0]);
                    } else if (intensity && intensity_kind == 1) {
                        o += ' '; o += std::to_string(static_cast<const long long *>(intensity)[i]);
                    } else if (intensity) {
                        o += ' ';
                        ply_fmt(o, static_cast<const double *>(intensity)[i], 6, precision);
                    }
                    o += '\n';
                }
            });
        }
        for (auto &x : th) x.join();
        for (unsigned t = 0; t < cnt && ok; ++t) ok = fwrite(out[t].data(), 1, out[t].size(), f) == out[t].size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) return fail(SS_ERR_PARAM, std::string("short write to ") + path);
    return SS_OK;
}
