// Host check of csrc/nsc_fft.cuh: the three register-butterfly Stockham passes against a
// float64 DFT.   nvcc -O2 -I../neural_spectral_codec_b200/csrc fft_selftest.cu -o _build/fft_selftest
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "nsc_fft.cuh"
using namespace nsc;
int main() {
    const int N = 360;
    std::vector<float2> tw(N), x(N), a(N), b(N);
    for (int m = 0; m < N; ++m) tw[m] = make_float2((float)cos(2 * M_PI * m / N), (float)-sin(2 * M_PI * m / N));
    srand(1);
    double worst = 0, scale = 0;
    for (int trial = 0; trial < 20; ++trial) {
        for (int n = 0; n < N; ++n) x[n] = make_float2(rand() / (float)RAND_MAX * 60.f, rand() / (float)RAND_MAX * 60.f);
        for (int j = 0; j < N / 8; ++j) stockham_butterfly<8, 1>(x.data(), a.data(), tw.data(), j);
        for (int j = 0; j < N / 9; ++j) stockham_butterfly<9, 8>(a.data(), b.data(), tw.data(), j);
        for (int j = 0; j < N / 5; ++j) stockham_butterfly<5, 72>(b.data(), a.data(), tw.data(), j);
        for (int k = 0; k < N; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < N; ++n) {
                double ang = -2 * M_PI * (double)((long)k * n % N) / N;
                re += x[n].x * cos(ang) - x[n].y * sin(ang);
                im += x[n].x * sin(ang) + x[n].y * cos(ang);
            }
            worst = fmax(worst, hypot(a[k].x - re, a[k].y - im));
            scale = fmax(scale, hypot(re, im));
        }
    }
    printf("max |err| = %.3e, max |X| = %.3e, relative %.3e\n", worst, scale, worst / scale);
    return worst / scale < 2e-6 ? 0 : 1;
}
