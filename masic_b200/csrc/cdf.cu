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

// ------------------------------------------------------------------ range coder of the y bitstream (host)
// HSIC.compress / decompress drive the PyPI package `range_coder` (RangeEncoder.encode(symbols, cumFreq),
// MASIC.py:958,1043,1221), which the reference neither vendors nor pins and which is not installable here.
// This is a plain 32-bit carry-propagating range coder over arbitrary integer cumulative frequencies
// (total <= 2^16 + alphabet size): same interface, own byte format (the y byte string is NOT
// comparable with the package's; lengths agree with the ideal code length to a fraction of a percent).
namespace {
struct RangeEnc {
  uint64_t low = 0;
  uint32_t range = 0xFFFFFFFFu;
  uint8_t cache = 0;
  int64_t cache_size = 1;
  uint8_t* out; int64_t cap; int64_t len = 0; bool overflow = false;
  void put(uint8_t b) { if (len < cap) out[len] = b; else overflow = true; ++len; }
  void shift_low() {
    if (static_cast<uint32_t>(low) < 0xFF000000u || (low >> 32) != 0) {
      uint8_t c = cache;
      do { put(static_cast<uint8_t>(c + static_cast<uint8_t>(low >> 32))); c = 0xFF; } while (--cache_size != 0);
      cache = static_cast<uint8_t>(low >> 24);
    }
    ++cache_size;
    low = (low & 0x00FFFFFFu) << 8;
  }
  void encode(uint32_t cum, uint32_t freq, uint32_t total) {
    const uint32_t r = range / total;
    low += static_cast<uint64_t>(r) * cum;
    range = r * freq;
    while (range < (1u << 24)) { range <<= 8; shift_low(); }
  }
  void finish() { for (int i = 0; i < 5; ++i) shift_low(); }
};
}  // namespace

struct MasicRangeDecoder {
  const uint8_t* data; int64_t len; int64_t pos = 0;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  uint8_t next() { return pos < len ? data[pos++] : 0; }
};

extern "C" int masic_range_encode(const int32_t* intervals_host, int64_t n, uint8_t* out_host, int64_t out_cap,
                                  int64_t* out_len) {
  if ((!intervals_host && n > 0) || !out_host || !out_len || n < 0) return MASIC_EINVAL;
  RangeEnc e;
  e.out = out_host; e.cap = out_cap;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t lo = intervals_host[3 * i], fr = intervals_host[3 * i + 1], tot = intervals_host[3 * i + 2];
    if (lo < 0 || fr <= 0 || tot <= 0 || lo + fr > tot || tot > (1 << 22)) return MASIC_EINVAL;
    e.encode(static_cast<uint32_t>(lo), static_cast<uint32_t>(fr), static_cast<uint32_t>(tot));
  }
  e.finish();
  *out_len = e.len;
  return e.overflow ? MASIC_EINVAL : MASIC_OK;
}

extern "C" int masic_range_decoder_create(const uint8_t* data_host, int64_t len, MasicRangeDecoder** dec_out) {
  if (!data_host || len < 0 || !dec_out) return MASIC_EINVAL;
  MasicRangeDecoder* d = new MasicRangeDecoder();
  d->data = data_host; d->len = len;
  d->next();                                   // the encoder's first byte is always 0 (initial cache)
  for (int i = 0; i < 4; ++i) d->code = (d->code << 8) | d->next();
  *dec_out = d;
  return MASIC_OK;
}

// Decode one symbol per CDF row: rows_host is (n_rows, row_len) int32 with row[0] = 0 and row[row_len-1] = total.
extern "C" int masic_range_decode_rows(MasicRangeDecoder* d, const int32_t* rows_host, int n_rows, int row_len,
                                       int32_t* symbols_host) {
  if (!d || !rows_host || !symbols_host || n_rows < 0 || row_len < 2) return MASIC_EINVAL;
  for (int i = 0; i < n_rows; ++i) {
    const int32_t* row = rows_host + static_cast<size_t>(i) * row_len;
    const uint32_t total = static_cast<uint32_t>(row[row_len - 1]);
    if (total == 0 || total > (1u << 22)) return MASIC_EINVAL;
    const uint32_t r = d->range / total;
    uint32_t v = d->code / r;
    if (v >= total) v = total - 1;
    int lo = 0, hi = row_len - 1;              // largest s with row[s] <= v
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (static_cast<uint32_t>(row[mid]) <= v) lo = mid; else hi = mid; }
    const uint32_t cum = static_cast<uint32_t>(row[lo]), freq = static_cast<uint32_t>(row[lo + 1]) - cum;
    if (freq == 0) return MASIC_EINVAL;
    d->code -= r * cum;
    d->range = r * freq;
    while (d->range < (1u << 24)) { d->code = (d->code << 8) | d->next(); d->range <<= 8; }
    symbols_host[i] = lo;
  }
  return MASIC_OK;
}

extern "C" void masic_range_decoder_destroy(MasicRangeDecoder* d) { delete d; }
