#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
typedef struct { float g, y2; } st;
static inline void step(st* s, float xr, float xi, float alpha) {
    float yr = xr * s->g, yi = xi * s->g;
    float y2 = yr * yr + yi * yi;
    s->y2 = (1.0 - alpha) * s->y2 + alpha * y2;
    if (s->y2 > 1e-6f) s->g *= expf(-0.5f * alpha * logf(s->y2));
    if (s->g > 1e6f) s->g = 1e6f;
}
static unsigned long long rs = 88172645463325252ull;
static double rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (rs >> 11) * (1.0 / 9007199254740992.0); }
static double gauss(void) { double u = rnd() + 1e-300, v = rnd(); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); }
int main(int argc, char** argv) {
    const int N = 4000000;
    float alpha = argc > 1 ? atof(argv[1]) : 0.01f;
    float* xr = malloc(N * 4), *xi = malloc(N * 4);
    for (int i = 0; i < N; i++) {
        double amp = 0.05 * (1.0 + 0.8 * sin(i * 2e-4)) * (1 + 0.5 * ((i / 50000) % 3));    // slow fades + steps
        double ph = i * 0.37;
        xr[i] = (float)(amp * cos(ph) + 0.01 * gauss());
        xi[i] = (float)(amp * sin(ph) + 0.01 * gauss());
    }
    st* tr = malloc((size_t)N * sizeof(st));
    st s = {1.0f, 1.0f};
    for (int i = 0; i < N; i++) { tr[i] = s; step(&s, xr[i], xi[i], alpha); }   // tr[i] = state BEFORE sample i
    double deltas[] = {1.0, 1e-2, 1e-4, 1e-5, 1e-6, 1e-7};
    for (int d = 0; d < 6; d++) {
        int cnt = 0, mx = 0; long sum = 0; int hist[64] = {0};
        int times[2000];
        for (int k = 0; k < 1000; k++) {
            int s0 = 100000 + k * 3000;
            st p = tr[s0];
            double sg = (rnd() < 0.5 ? -1 : 1), sy = (rnd() < 0.5 ? -1 : 1);
            if (deltas[d] == 1.0) { p.g = 1.0f; p.y2 = 1.0f; }
            else { p.g = (float)(p.g * (1.0 + sg * deltas[d])); p.y2 = (float)(p.y2 * (1.0 + sy * deltas[d])); }
            int n = 0;
            while (s0 + n < N - 1) {
                if (memcmp(&p, &tr[s0 + n], sizeof(st)) == 0) break;
                step(&p, xr[s0 + n], xi[s0 + n], alpha);
                n++;
                if (n > 200000) break;
            }
            times[cnt++] = n; sum += n; if (n > mx) mx = n;
        }
        // percentiles
        for (int a = 0; a < cnt; a++) for (int b = a + 1; b < cnt; b++) if (times[b] < times[a]) { int t = times[a]; times[a] = times[b]; times[b] = t; }
        printf("alpha %g delta %g: mean %.0f p50 %d p90 %d p99 %d max %d\n", alpha, deltas[d], (double)sum / cnt, times[cnt / 2], times[cnt * 9 / 10], times[cnt * 99 / 100], mx);
    }
    return 0;
}
