/* A plain C caller of the drop-in boundary (include/solid_gpu.h): design a Kaiser low-pass on the device, filter a
 * stream in two calls from ordinary (pageable) host memory, and check the split against a single call -- the
 * streaming contract of the reference's Filter::execute_block (filter/mod.rs:14).
 *   gcc -O2 -Iinclude examples/c/fir_stream.c -Lsolid_dsp_b200/lib -lsolid_gpu -lm -o fir_stream              */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "solid_gpu.h"

#define CHECK(call)                                                                            \
    do {                                                                                       \
        int st_ = (call);                                                                      \
        if (st_ != SGPU_OK) {                                                                  \
            fprintf(stderr, "%s -> %s: %s\n", #call, sgpu_status_name(st_), sgpu_last_error()); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

int main(void) {
    const size_t T = 64, n = 1 << 20, cut = 333333;
    double fc = 0.25, as = 60.0, mu = 0.0;
    double *taps = malloc(T * sizeof(double));
    CHECK(sgpu_firdes_kaiser(T, &fc, &as, &mu, 1, taps, SGPU_HOST, NULL));       /* firdes/mod.rs:278-305 */

    float *x = malloc(2 * n * sizeof(float)), *y = malloc(2 * n * sizeof(float)), *z = malloc(2 * n * sizeof(float));
    unsigned s = 12345u;
    for (size_t i = 0; i < 2 * n; ++i) {
        s = s * 1664525u + 1013904223u;
        x[i] = (float)((s >> 8) * (2.0 / 16777216.0) - 1.0);
    }
    sgpu_fir *f = NULL, *g = NULL;
    CHECK(sgpu_fir_create(taps, T, SGPU_TAPS_REAL, 1, 1.0, 0.0, 0, 0, &f));     /* FIRFilter::new, fir/mod.rs:79 */
    CHECK(sgpu_fir_clone(f, &g));                                                /* #[derive(Clone)] */
    size_t got = 0, got2 = 0;
    CHECK(sgpu_fir_execute_block(f, x, n, n, y, n, &got, SGPU_HOST, NULL));     /* one call */
    CHECK(sgpu_fir_execute_block(g, x, cut, cut, z, cut, &got2, SGPU_HOST, NULL));  /* the same stream in two calls */
    CHECK(sgpu_fir_execute_block(g, x + 2 * cut, n - cut, n - cut, z + 2 * cut, n - cut, &got2, SGPU_HOST, NULL));
    double worst = 0.0, peak = 0.0;
    for (size_t i = 0; i < 2 * n; ++i) {
        worst = fmax(worst, fabs((double)y[i] - (double)z[i]));
        peak = fmax(peak, fabs((double)y[i]));
    }
    printf("%zu outputs, peak %.4f, split-call difference %.3g (%s), kernels launched: %llu\n", got, peak, worst,
           worst <= 1e-6 * peak ? "OK" : "MISMATCH", (unsigned long long)sgpu_launch_count());
    sgpu_fir_destroy(f);
    sgpu_fir_destroy(g);
    free(taps); free(x); free(y); free(z);
    return worst <= 1e-6 * peak ? 0 : 2;
}
