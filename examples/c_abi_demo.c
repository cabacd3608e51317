/* Plain-C caller of libnsc_b200.so: encodes one synthetic ring of points with host buffers
 * through the pipeline entry points (no CUDA headers needed on the caller's side).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lneural_spectral_codec_b200 -lnsc_b200 -lm \
 *       -Wl,-rpath,$PWD/neural_spectral_codec_b200 -o /tmp/c_abi_demo && /tmp/c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "nsc_b200.h"

int main(void) {
    const int n = 20000, n_scans = 2;
    float* pts = (float*)malloc(sizeof(float) * 4 * n * n_scans);
    int64_t offsets[3] = {0, n, 2 * n};
    for (int s = 0; s < n_scans; ++s)
        for (int i = 0; i < n; ++i) {
            const double az = 6.283185307179586 * i / n, el = -0.4 + 0.42 * ((i * 7) % 64) / 64.0;
            const double r = 10.0 + 5.0 * sin(3.0 * az + s);
            float* p = pts + 4 * ((size_t)s * n + i);
            p[0] = (float)(r * cos(el) * cos(az));
            p[1] = (float)(r * cos(el) * sin(az));
            p[2] = (float)(r * sin(el));
            p[3] = 0.5f;
        }
    nsc_params prm;
    nsc_default_params(&prm);
    int32_t lut[NSC_N_FREQS];
    int st = nsc_freq_to_bin(2.0f, &prm, lut);
    if (st != NSC_OK) { fprintf(stderr, "freq_to_bin: %s\n", nsc_strerror(st)); return 1; }
    nsc_pipeline* pl = NULL;
    st = nsc_pipeline_create(1 << 20, 2, 0, &pl);
    if (st != NSC_OK) { fprintf(stderr, "pipeline_create: %s %s\n", nsc_strerror(st), nsc_last_cuda_error()); return 1; }
    float* out = (float*)malloc(sizeof(float) * n_scans * prm.target_rows * prm.n_bins);
    st = nsc_pipeline_encode(pl, pts, 4, offsets[n_scans], offsets, n_scans, &prm, lut, out);
    if (st != NSC_OK) { fprintf(stderr, "encode: %s %s\n", nsc_strerror(st), nsc_last_cuda_error()); return 1; }
    for (int s = 0; s < n_scans; ++s) {
        double sum = 0;
        for (int i = 0; i < prm.target_rows * prm.n_bins; ++i) sum += out[s * 800 + i];
        printf("scan %d: descriptor sum %.6f, first bins %.5f %.5f %.5f\n", s, sum, out[s * 800], out[s * 800 + 1],
               out[s * 800 + 2]);
    }
    nsc_pipeline_destroy(pl);
    free(out);
    free(pts);
    return 0;
}
