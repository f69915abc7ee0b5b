// tools/exact_math_sweep.cpp — exhaustive check of cmix_b200/csrc/exact_math.h
// against the host libm (glibc) over ALL 2^32 float bit patterns.
//   g++ -O2 -std=c++17 -ffp-contract=off -fopenmp tools/exact_math_sweep.cpp -o /tmp/sweep -lm && /tmp/sweep
#include "../cmix_b200/csrc/exact_math.h"
#include <math.h>
#include <stdio.h>
#include <atomic>

static bool same(float a, float b) {
  if (isnan(a) && isnan(b)) return true;
  return XM_F2U(a) == XM_F2U(b);
}
int main() {
  std::atomic<long> bad_exp(0), bad_tanh(0), bad_logi(0);
#pragma omp parallel for schedule(dynamic, 1)
  for (long hi = 0; hi < 65536; ++hi) {
    for (long lo = 0; lo < 65536; ++lo) {
      uint32_t u = (uint32_t)((hi << 16) | lo);
      float x = XM_U2F(u);
      float a = xm_expf(x), b = expf(x);
      if (!same(a, b)) { if (bad_exp++ < 5) printf("expf  x=%a (%08x): mine %a libm %a\n", x, u, a, b); }
      a = xm_tanhf(x); b = tanhf(x);
      if (!same(a, b)) { if (bad_tanh++ < 5) printf("tanhf x=%a (%08x): mine %a libm %a\n", x, u, a, b); }
    }
  }
  printf("mismatches: expf %ld tanhf %ld (of 4294967296 each)\n", bad_exp.load(), bad_tanh.load());
  return (bad_exp || bad_tanh) ? 1 : 0;
}
