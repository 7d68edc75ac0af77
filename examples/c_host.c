/* A host written in plain C against include/pyatm_b200.h: no CUDA headers, no Python, no torch.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_host.c -o c_host -Lpyatmosphere_b200 -lpyatm_b200 -Wl,-rpath,$PWD/pyatmosphere_b200
 *   ./c_host batch.bin records.bin
 *
 * Reads one Monte-Carlo batch -- grid axes, path geometry, the sparse-spectrum coefficients of every screen -- from a flat
 * binary file (written by tests/test_gpu_c_host.py from the same objects the Python host uses), runs it through
 * pa_simulate_batch (host buffers in, per-realization table out: what simulations/simulation.py:89-114 computes per
 * iteration for BeamResult + PDTResult) and writes the table.  This is the binding a non-Python host of the reference's
 * path would use (cgo / JNI / N-API stubs are one-to-one with these calls).
 *
 * file: int32 n, S, M, B, m_split, degree, method, npupil | double delta, wvl, w0, F0, final_scale, coef_bound |
 *       double legs[S+1], scales[S] | float x[n], y[n] | float fx[S][B][M], fy[S][B][M], coef[S][B][M][2] | float pupils[npupil][3]
 */
#include <stdio.h>
#include <stdlib.h>

#include "pyatm_b200.h"

static void die(const char* what) {
    fprintf(stderr, "%s: %s\n", what, pa_last_error());
    exit(1);
}
static void* slurp(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s batch.bin records.bin\n", argv[0]);
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int* hi = (int*)slurp(f, 8 * sizeof(int));
    const int n = hi[0], S = hi[1], M = hi[2], B = hi[3], npupil = hi[7];
    double* hd = (double*)slurp(f, 6 * sizeof(double));
    double* legs = (double*)slurp(f, (size_t)(S + 1) * sizeof(double));
    double* scales = (double*)slurp(f, (size_t)S * sizeof(double));
    float* x = (float*)slurp(f, (size_t)n * sizeof(float));
    float* y = (float*)slurp(f, (size_t)n * sizeof(float));
    const size_t cnt = (size_t)S * B * M;
    float* fx = (float*)slurp(f, cnt * sizeof(float));
    float* fy = (float*)slurp(f, cnt * sizeof(float));
    float* coef = (float*)slurp(f, cnt * 2 * sizeof(float));
    float* pupils = (float*)slurp(f, (size_t)npupil * 3 * sizeof(float));
    fclose(f);

    int devices = 0;
    if (pa_device_count(&devices) || devices < 1) die("pa_device_count");
    pa_ctx* ctx = NULL;
    if (pa_ctx_create(&ctx, 0, n, PA_C64)) die("pa_ctx_create");
    if (pa_ctx_set_axes(ctx, x, y, hd[0])) die("pa_ctx_set_axes");

    pa_path path;
    path.n_screens = S;
    path.leg_lengths_host = legs;
    path.screen_scale_host = scales;
    path.final_scale = hd[4];
    path.wvl = hd[1];
    path.w0 = hd[2];
    path.F0 = hd[3];
    path.m = M;
    path.m_split = hi[4];
    path.degree = hi[5];
    path.shift_x = path.shift_y = 0.0;
    path.screen_method = hi[6];
    path.coef_bound = hd[5];
    path.from_field = 0;

    const int stride = PA_MEASURE_HEAD + PA_MAX_PUPILS;
    double* table = (double*)calloc((size_t)B * stride, sizeof(double));
    if (pa_simulate_batch(ctx, &path, B, fx, fy, coef, 0, 0, NULL, NULL, pupils, npupil, table, stride, NULL /* default stream */))
        die("pa_simulate_batch");
    printf("pa_version %d, %d realizations, %llu kernel launches; eta of the first: %.9f (aperture: %.9f)\n", pa_version(), B,
           pa_launch_count(0), table[0], npupil ? table[PA_MEASURE_HEAD] : 0.0);
    f = fopen(argv[2], "wb");
    if (!f || fwrite(table, sizeof(double), (size_t)B * stride, f) != (size_t)B * stride) return 2;
    fclose(f);
    pa_ctx_destroy(ctx);
    return 0;
}
