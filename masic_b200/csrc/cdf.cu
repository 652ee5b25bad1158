// cdf.cu — host-side integer table construction used by update() (cold path, runs once per
// model).  Replaces the reference's native module compressai._CXX
// (compressai/cpp_exts/ops/ops.cpp:40-109, bound at entropy_models.py:8-9,50-53).
// The tables define the rANS bitstream, so the arithmetic below is integer-exact.
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/masic_b200.h"

// Returns 0, or 1 = negative / non-finite probability, 2 = all probabilities round to zero
// (the two std::domain_error cases of ops.cpp:46-64), MASIC_EINVAL for bad arguments.
extern "C" int masic_pmf_to_quantized_cdf(const float* pmf_host, int n, int precision,
                                          uint32_t* cdf_host) {
  if (!pmf_host || !cdf_host || n <= 0 || precision < 1 || precision > 24) return MASIC_EINVAL;
  for (int i = 0; i < n; ++i)
    if (pmf_host[i] < 0.0f || !isfinite(pmf_host[i])) return 1;
  const uint32_t top = 1u << precision;
  std::vector<uint32_t> c(static_cast<size_t>(n) + 1);
  c[0] = 0;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    c[i + 1] = static_cast<uint32_t>(roundf(pmf_host[i] * static_cast<float>(top)));
    total += static_cast<int>(c[i + 1]);
  }
  if (total == 0) return 2;
  uint32_t run = 0;
  for (int i = 0; i <= n; ++i) {   // rescale so the counts fit in 2^precision, then prefix-sum
    run += static_cast<uint32_t>((static_cast<uint64_t>(top) * c[i]) / static_cast<uint32_t>(total));
    c[i] = run;
  }
  c[n] = top;
  // every symbol needs a non-empty interval: borrow one count from the cheapest donor
  for (int i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    uint32_t donor_freq = UINT32_MAX;
    int donor = -1;
    for (int j = 0; j < n; ++j) {
      const uint32_t f = c[j + 1] - c[j];
      if (f > 1 && f < donor_freq) { donor_freq = f; donor = j; }
    }
    if (donor < 0) return 3;
    if (donor < i) for (int j = donor + 1; j <= i; ++j) --c[j];
    else           for (int j = i + 1; j <= donor; ++j) ++c[j];
  }
  for (int i = 0; i <= n; ++i) cdf_host[i] = c[i];
  return MASIC_OK;
}

// Whole table at once (entropy_models.py:136-142 `_pmf_to_cdf`): row i = pmf[i, :len[i]] ++ tail[i].
// cdf_host is (rows, max_len + 2) int32, zero padded.
extern "C" int masic_pmf_table_to_cdf(const float* pmf_host, int rows, int row_stride,
                                      const float* tail_mass_host, const int32_t* pmf_length_host,
                                      int max_length, int precision, int32_t* cdf_host) {
  if (!pmf_host || !tail_mass_host || !pmf_length_host || !cdf_host || rows <= 0) return MASIC_EINVAL;
  std::vector<float> p;
  std::vector<uint32_t> c;
  for (int r = 0; r < rows; ++r) {
    const int len = pmf_length_host[r];
    if (len < 0 || len > max_length || len > row_stride) return MASIC_EINVAL;
    p.assign(pmf_host + static_cast<size_t>(r) * row_stride, pmf_host + static_cast<size_t>(r) * row_stride + len);
    p.push_back(tail_mass_host[r]);
    c.assign(p.size() + 1, 0);
    const int rc = masic_pmf_to_quantized_cdf(p.data(), static_cast<int>(p.size()), precision, c.data());
    if (rc) return rc;
    int32_t* out = cdf_host + static_cast<size_t>(r) * (max_length + 2);
    for (int j = 0; j < max_length + 2; ++j) out[j] = j < static_cast<int>(c.size()) ? static_cast<int32_t>(c[j]) : 0;
  }
  return MASIC_OK;
}
