/* Plain-C caller of the C ABI (include/ss_passive.h, include/ss_post.h): no Python, numpy or torch in the loop.
 * Reads two raw BGR images, runs ASW (+ L-R check), GSW, and the ASW -> 3-D points call, writes raw outputs.
 *   c_abi_smoke <left.bgr> <right.bgr> <W> <H> <out_prefix>
 * Built and checked against the Python mirror by tests/test_gpu_parity.py::test_c_abi_from_plain_c. */
#include <stdio.h>
#include <stdlib.h>
#include "../include/ss_passive.h"
#include "../include/ss_post.h"

static void *slurp(const char *path, size_t n) {
    FILE *f = fopen(path, "rb");
    void *p = malloc(n);
    if (!f || !p || fread(p, 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return p;
}
static void dump(const char *prefix, const char *suffix, const void *p, size_t n) {
    char path[1024];
    snprintf(path, sizeof(path), "%s%s", prefix, suffix);
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(p, 1, n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fclose(f);
}
#define CHECK(call) do { int rc_ = (call); if (rc_ != SS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ss_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s left.bgr right.bgr W H out_prefix\n", argv[0]); return 2; }
    const int W = atoi(argv[3]), H = atoi(argv[4]);
    const size_t npx = (size_t)W * H;
    unsigned char *l = slurp(argv[1], npx * 3), *r = slurp(argv[2], npx * 3);
    int16_t *disp = malloc(npx * 2);
    float *pts = malloc(npx * 12);
    if (ss_abi_version() != 2) return 3;
    CHECK(ss_init(0));
    CHECK(ss_asw_compute(l, r, W, H, 9, 24, 0, 5.0, 17.5, 1, disp));
    dump(argv[5], ".asw.i16", disp, npx * 2);
    CHECK(ss_gsw_compute(l, r, W, H, 7, 24, 0, 10, 120.0f, 3, 20, disp));
    dump(argv[5], ".gsw.i16", disp, npx * 2);
    const double Q[16] = {1, 0, 0, -W / 2.0, 0, 1, 0, -H / 2.0, 0, 0, 0, -(double)W, 0, 0, 1, 0};   /* points.py:147-174 */
    CHECK(ss_asw_compute_points(l, r, W, H, 9, 24, 0, 5.0, 17.5, 0, Q, disp, pts));
    dump(argv[5], ".pts.f32", pts, npx * 12);
    /* error path: even window -> the reference's ValueError("winSize must be a positive odd number!") */
    if (ss_asw_compute(l, r, W, H, 8, 24, 0, 5.0, 17.5, 0, disp) != SS_ERR_WINSIZE) return 4;
    CHECK(ss_shutdown());
    printf("c_abi_smoke ok %dx%d\n", W, H);
    return 0;
}
