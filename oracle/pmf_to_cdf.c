/* oracle/pmf_to_cdf.c — C restatement of the reference's pmf -> quantised cdf routine
 * (compressai/cpp_exts/ops/ops.cpp:40-109).  TEST INFRASTRUCTURE: built by
 * `make -C oracle port`, loaded only by tests/, smoke() and bench.py's cpu_baseline leg.
 *
 *   1. every p must be finite and >= 0                               (ops.cpp:46-52)
 *   2. f[i+1] = roundf(p[i] * 2^precision), f[0] = 0                 (ops.cpp:54-58)
 *   3. total = sum f (int); total == 0 is an error                   (ops.cpp:60-64)
 *   4. f[i] = (2^precision * f[i]) / total   (64-bit, truncating)    (ops.cpp:66-70)
 *   5. prefix sum; last = 2^precision                                (ops.cpp:72-73)
 *   6. for each i with cdf[i] == cdf[i+1]: find the bin with the smallest
 *      frequency > 1 (first wins) and shift the boundary run by one  (ops.cpp:75-100)
 */
#include <math.h>
#include <stdint.h>

int masic_oracle_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
  const uint32_t one = (uint32_t)1 << precision;
  int i, j;
  for (i = 0; i < n; ++i)
    if (!(pmf[i] >= 0.0f) || !isfinite(pmf[i])) return 1;
  cdf[0] = 0;
  for (i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)one);
  {
    int total = 0;
    for (i = 0; i <= n; ++i) total += (int)cdf[i];
    if (total == 0) return 2;
    for (i = 0; i <= n; ++i) cdf[i] = (uint32_t)(((uint64_t)one * cdf[i]) / (uint32_t)total);
  }
  for (i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = one;
  for (i = 0; i < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    {
      uint32_t best_freq = 0xFFFFFFFFu;
      int best = -1;
      for (j = 0; j < n; ++j) {
        const uint32_t f = cdf[j + 1] - cdf[j];
        if (f > 1 && f < best_freq) { best_freq = f; best = j; }
      }
      if (best < 0) return 3;
      if (best < i) { for (j = best + 1; j <= i; ++j) cdf[j]--; }
      else          { for (j = i + 1; j <= best; ++j) cdf[j]++; }
    }
  }
  return 0;
}
