// CPU experiment behind DESIGN.md section 1: can the exact biquad be restarted speculatively?
// Runs the upstream recursion (f32 state, f64 intermediates) from the true state and from zero state
// at many offsets and reports how long it takes the two trajectories to become bit-identical.
// Result: most restarts never merge within 100,000 samples (they settle a few ulps apart: the slow pole
// leaves a dead band of ~250 ulps), so K0 stays a serial recursion.
// build: gcc -O2 -ffp-contract=off -o biquad_merge biquad_merge.c -lm
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
static const float b_hp[2] = {-2.f, 1.f}, a_hp[2] = {-1.99599f, 0.99600f};
static inline void step(float *m, float xi, float *y) {
  float yi = xi + m[0];
  m[0] = (float)(m[1] + (b_hp[0] * (double)xi - a_hp[0] * (double)yi));
  m[1] = (float)(b_hp[1] * (double)xi - a_hp[1] * (double)yi);
  *y = yi;
}
int main(int argc, char **argv) {
  int N = 48000 * 20;
  float *x = malloc(N * sizeof(float));
  int worst = 0; long total = 0; int cnt = 0; int hist[32] = {0};
  for (int trial = 0; trial < 200; trial++) {
    srand(trial + 1);
    double amp = pow(10.0, -(trial % 5)) * 8000.0; // loud to quiet
    double f0 = 90 + (trial * 7) % 170, ph = 0;
    for (int i = 0; i < N; i++) {
      ph += 2 * M_PI * f0 / 48000.0;
      double env = fmax(0.0, sin(2 * M_PI * 3.0 * i / 48000.0) + 0.35);
      double v = 0; for (int h = 1; h <= 8; h++) v += sin(h * ph) / h;
      double n = ((rand() / (double)RAND_MAX) - 0.5) * 2.0;
      double s = amp * (env * v + 0.3 * n) + ((trial % 7 == 0) ? 500.0 : 0.0); // DC offset on some
      if (trial % 3 == 1) s = rint(s); // int16-like input
      x[i] = (float)s;
    }
    // true trajectory states
    float m[2] = {0, 0}, y;
    float *st = malloc(2 * N * sizeof(float));
    for (int i = 0; i < N; i++) { step(m, x[i], &y); st[2 * i] = m[0]; st[2 * i + 1] = m[1]; }
    // zero-start from several offsets
    for (int start = 48000; start < N - 100000; start += 77777) {
      float w[2] = {0, 0}; int merged = -1;
      for (int i = start; i < start + 100000; i++) {
        step(w, x[i], &y);
        if (w[0] == st[2 * i] && w[1] == st[2 * i + 1] && memcmp(w, &st[2*i], 8) == 0) { merged = i - start + 1; break; }
      }
      if (merged < 0) { printf("trial %d start %d: NOT merged in 100000\n", trial, start); merged = 100000; }
      if (merged > worst) worst = merged;
      total += merged; cnt++;
      hist[merged / 1000 < 31 ? merged / 1000 : 31]++;
    }
    free(st);
  }
  printf("cases %d mean %.0f worst %d\n", cnt, (double)total / cnt, worst);
  for (int i = 0; i < 32; i++) if (hist[i]) printf("  %2dk: %d\n", i, hist[i]);
  return 0;
}
