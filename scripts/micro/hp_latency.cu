// micro-benchmark: latency of the exact biquad recursion variants on one warp (B200)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double f2d_manual(float f) {  // exact for normal floats and zero
  const uint32_t u = __float_as_uint(f);
  const uint32_t e = (u >> 23) & 0xFF;
  if (e == 0 || e == 255) return (double)f;
  const uint32_t hi = (u & 0x80000000u) | (((u >> 3) & 0x0FFFFFFFu) + 0x38000000u);
  const uint32_t lo = u << 29;
  return __hiloint2double((int)hi, (int)lo);
}

template <int V>
__global__ void k(const float *x, float *y, int n, long long *cycles, float *mout) {
  __shared__ float sx[4096];
  for (int i = threadIdx.x; i < 4096; i += 32) sx[i] = x[i];
  __syncwarp();
  float m0 = 0.f, m1 = 0.f;
  const double a0 = (double)-1.99599f, a1 = (double)0.99600f;
  float acc = 0.f, acc2 = 0.f, m0_prev = 0.f;
  int bad = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    const float xi = sx[(i + threadIdx.x) & 4095];
    const float yi = xi + m0;
    if (V == 0) {
      const double xd = (double)xi, yd = (double)yi;
      m0 = (float)((double)m1 + (-2.0 * xd - a0 * yd));
      m1 = (float)(xd - a1 * yd);
    } else if (V == 1) {
      const double xd = (double)xi, yd = (double)yi;
      m0 = (float)((double)m1 + fma(-a0, yd, -2.0 * xd));
      m1 = (float)fma(-a1, yd, xd);
    } else if (V == 2) {
      const double xd = f2d_manual(xi), yd = f2d_manual(yi);
      m0 = (float)(f2d_manual(m1) + fma(-a0, yd, -2.0 * xd));
      m1 = (float)fma(-a1, yd, xd);
    } else if (V == 3) {  // f32 only (NOT exact; latency reference)
      m0 = m1 + (-2.0f * xi - -1.99599f * yi);
      m1 = xi - 0.99600f * yi;
    } else if (V == 5 || V == 6) {
      // K0 since round 2: mem0 speculated in error-free f32 arithmetic, upstream's f64 value computed beside it
      // and compared off the chain (mismatches counted in `bad`); mem1 by upstream's f64 expression.  V == 6 takes
      // the error term of ch + ph from a two-sided FastTwoSum (12 instead of 16 cycles deep).
      const float a0f = 1.99599f;
      const float ph = __fmul_rn(a0f, yi), pl = __fmaf_rn(a0f, yi, -ph);
      const float b = -2.0f * xi;
      const float ch = __fadd_rn(m1, b), cbb = __fsub_rn(ch, m1);
      const float cl = __fadd_rn(__fsub_rn(m1, __fsub_rn(ch, cbb)), __fsub_rn(b, cbb));
      const float s1 = __fadd_rn(ch, ph);
      float e1;
      if (V == 6) {
        e1 = (fabsf(ch) >= fabsf(ph)) ? __fsub_rn(ph, __fsub_rn(s1, ch)) : __fsub_rn(ch, __fsub_rn(s1, ph));
      } else {
        const float sbb = __fsub_rn(s1, ch);
        e1 = __fadd_rn(__fsub_rn(ch, __fsub_rn(s1, sbb)), __fsub_rn(ph, sbb));
      }
      const float m0s = __fadd_rn(s1, __fadd_rn(__fadd_rn(cl, pl), e1));
      const double xd = (double)xi, yd = (double)yi;
      const float m0r = (float)((double)m1 + fma(-a0, yd, -2.0 * xd));
      m1 = (float)fma(-a1, yd, xd);
      bad += (__float_as_uint(m0s) != __float_as_uint(m0r));
      m0 = m0s;
    } else if (V == 15 || V == 16) {
      // mem1 = RN32(x - a1 y) IS one f32 FMA (exact product, one rounding; upstream rounds to 53 bits first, which
      // differs with probability ~2^-29), and -2x folds into FMAs: 19 instructions a sample instead of 30.
      // 15: error terms by TwoSum; 16: by the two-sided FastTwoSum
      const float a0f = 1.99599f, na1f = -0.99600f;
      const float ph = __fmul_rn(a0f, yi), pl = __fmaf_rn(a0f, yi, -ph);
      const float ch = __fmaf_rn(-2.0f, xi, m1);
      const float s1 = __fadd_rn(ch, ph);
      float cl, e1;
      if (V == 15) {
        const float cbb = __fsub_rn(ch, m1), sbb = __fsub_rn(s1, ch);
        cl = __fadd_rn(__fsub_rn(m1, __fsub_rn(ch, cbb)), __fmaf_rn(-2.0f, xi, -cbb));
        e1 = __fadd_rn(__fsub_rn(ch, __fsub_rn(s1, sbb)), __fsub_rn(ph, sbb));
      } else {
        const float b = -2.0f * xi;
        cl = (fabsf(m1) >= fabsf(b)) ? __fsub_rn(b, __fsub_rn(ch, m1)) : __fsub_rn(m1, __fsub_rn(ch, b));
        e1 = (fabsf(ch) >= fabsf(ph)) ? __fsub_rn(ph, __fsub_rn(s1, ch)) : __fsub_rn(ch, __fsub_rn(s1, ph));
      }
      m1 = __fmaf_rn(na1f, yi, xi);
      m0 = __fadd_rn(s1, __fadd_rn(__fadd_rn(cl, pl), e1));
    } else if (V >= 11 && V <= 14) {
      // what bounds the speculation warp?  11: the head fused (s1 = fma(a0, y, ch), error two-sided); 12: as 8 with mem1
      // taken out of the loop (main chain alone); 13: as 8 without e1 (a chain of five: NOT exact, depth probe);
      // 14: as 8 with the chain cut (y from the state two samples back: the same instructions, issue-bound)
      const float a0f = 1.99599f, na1f = -0.99600f;
      const float yy = (V == 14) ? xi + m0_prev : yi;
      const float ph = __fmul_rn(a0f, yy), pl = __fmaf_rn(a0f, yy, -ph);
      const float b = -2.0f * xi;
      const float ch = __fadd_rn(m1, b), cbb = __fsub_rn(ch, m1);
      const float cl = __fadd_rn(__fsub_rn(m1, __fsub_rn(ch, cbb)), __fsub_rn(b, cbb));
      float s1, e1;
      if (V == 11) {
        s1 = __fmaf_rn(a0f, yy, ch);
        const float ea = __fadd_rn(__fsub_rn(ch, s1), ph), eb = __fadd_rn(__fsub_rn(ph, s1), ch);
        e1 = __fadd_rn(fabsf(ch) >= fabsf(ph) ? ea : eb, pl);
      } else {
        s1 = __fadd_rn(ch, ph);
        const float sbb = __fsub_rn(s1, ch);
        e1 = (V == 13) ? 0.f : __fadd_rn(__fsub_rn(ch, __fsub_rn(s1, sbb)), __fsub_rn(ph, sbb));
      }
      const float qh = __fmul_rn(na1f, yy), ql = __fmaf_rn(na1f, yy, -qh);
      const float s2 = __fadd_rn(xi, qh), tbb = __fsub_rn(s2, xi);
      const float e2 = __fadd_rn(__fsub_rn(xi, __fsub_rn(s2, tbb)), __fsub_rn(qh, tbb));
      m0_prev = m0;
      m0 = (V == 11) ? __fadd_rn(s1, __fadd_rn(cl, e1)) : __fadd_rn(s1, __fadd_rn(__fadd_rn(cl, pl), e1));
      if (V == 12) acc2 += __fadd_rn(s2, __fadd_rn(e2, ql)); else m1 = __fadd_rn(s2, __fadd_rn(e2, ql));
    } else if (V == 10) {  // as the kernel now: every error term by the ordered FastTwoSum, m1 - 2x included
      const float a0f = 1.99599f, na1f = -0.99600f;
      const float ph = __fmul_rn(a0f, yi), pl = __fmaf_rn(a0f, yi, -ph);
      const float b = -2.0f * xi;
      const float ch = __fadd_rn(m1, b);
      const bool c0 = fabsf(m1) >= fabsf(b);
      const float cl = __fsub_rn(c0 ? b : m1, __fsub_rn(ch, c0 ? m1 : b));
      const float s1 = __fadd_rn(ch, ph);
      const float qh = __fmul_rn(na1f, yi), ql = __fmaf_rn(na1f, yi, -qh);
      const float s2 = __fadd_rn(xi, qh);
      const bool c1 = fabsf(ch) >= fabsf(ph), c2 = fabsf(xi) >= fabsf(qh);
      const float e1 = __fsub_rn(c1 ? ph : ch, __fsub_rn(s1, c1 ? ch : ph));
      const float e2 = __fsub_rn(c2 ? qh : xi, __fsub_rn(s2, c2 ? xi : qh));
      m0 = __fadd_rn(s1, __fadd_rn(__fadd_rn(cl, pl), e1));
      m1 = __fadd_rn(s2, __fadd_rn(e2, ql));
    } else if (V == 7 || V == 8 || V == 9) {
      // the speculation warp of the time-parallel K0: mem0 AND mem1 in error-free f32 arithmetic, no f64 at all
      // (V == 7: error terms by the two-sided FastTwoSum, as the kernel; V == 8: by TwoSum)
      const float a0f = 1.99599f, na1f = -0.99600f;
      const float ph = __fmul_rn(a0f, yi), pl = __fmaf_rn(a0f, yi, -ph);
      const float b = -2.0f * xi;
      const float ch = __fadd_rn(m1, b), cbb = __fsub_rn(ch, m1);
      const float cl = __fadd_rn(__fsub_rn(m1, __fsub_rn(ch, cbb)), __fsub_rn(b, cbb));
      const float s1 = __fadd_rn(ch, ph);
      const float qh = __fmul_rn(na1f, yi), ql = __fmaf_rn(na1f, yi, -qh);
      const float s2 = __fadd_rn(xi, qh);
      float e1, e2;
      if (V == 9) {  // operands ordered by magnitude before the sum is needed: two dependent operations behind s
        const bool c1 = fabsf(ch) >= fabsf(ph), c2 = fabsf(xi) >= fabsf(qh);
        e1 = __fsub_rn(c1 ? ph : ch, __fsub_rn(s1, c1 ? ch : ph));
        e2 = __fsub_rn(c2 ? qh : xi, __fsub_rn(s2, c2 ? xi : qh));
      } else if (V == 7) {
        e1 = (fabsf(ch) >= fabsf(ph)) ? __fsub_rn(ph, __fsub_rn(s1, ch)) : __fsub_rn(ch, __fsub_rn(s1, ph));
        e2 = (fabsf(xi) >= fabsf(qh)) ? __fsub_rn(qh, __fsub_rn(s2, xi)) : __fsub_rn(xi, __fsub_rn(s2, qh));
      } else {
        const float sbb = __fsub_rn(s1, ch), tbb = __fsub_rn(s2, xi);
        e1 = __fadd_rn(__fsub_rn(ch, __fsub_rn(s1, sbb)), __fsub_rn(ph, sbb));
        e2 = __fadd_rn(__fsub_rn(xi, __fsub_rn(s2, tbb)), __fsub_rn(qh, tbb));
      }
      m0 = __fadd_rn(s1, __fadd_rn(__fadd_rn(cl, pl), e1));
      m1 = __fadd_rn(s2, __fadd_rn(e2, ql));
    } else if (V == 4) {  // pure double state (not exact either): how slow is a DFMA chain
      static double d0, d1;
      const double xd = (double)xi;
      double yd = xd + (double)m0;
      m0 = (float)(0.0 + fma(-a0, yd, -2.0 * xd));
      m1 = 0.f;
    }
    acc += yi;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) {
    cycles[0] = t1 - t0;
  }
  y[threadIdx.x] = acc;
  mout[threadIdx.x] = m0 + m1 + (float)bad + acc2 + m0_prev;
}

int main() {
  float *x, *y, *m;
  long long *c;
  cudaMalloc(&x, 4096 * 4);
  cudaMalloc(&y, 32 * 4);
  cudaMalloc(&m, 32 * 4);
  cudaMalloc(&c, 8);
  float hx[4096];
  for (int i = 0; i < 4096; i++) hx[i] = 1000.f * sinf(0.01f * i) + 13.f * ((i * 7919) % 101);
  cudaMemcpy(x, hx, sizeof(hx), cudaMemcpyHostToDevice);
  const int n = 100000;
  long long hc;
#define RUN(V)                                                   \
  k<V><<<1, 32>>>(x, y, n, c, m);                                \
  k<V><<<1, 32>>>(x, y, n, c, m);                                \
  cudaDeviceSynchronize();                                       \
  cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);                 \
  printf("variant %d: %.1f cycles/sample (%s)\n", V, (double)hc / n, cudaGetErrorString(cudaGetLastError()));
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
  {
    float hm[32];
    cudaMemcpy(hm, m, sizeof(hm), cudaMemcpyDeviceToHost);
    printf("(variant 6 lane 0: m0 + m1 + mismatches = %g)\n", hm[0]);
  }
  return 0;
}
