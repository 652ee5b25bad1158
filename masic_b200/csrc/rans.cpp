// rans.cpp — host-side range-ANS serialisation of the z side information.
//
// Replaces, at the boundary, the reference's pybind11 module `compressai.ans`
// (compressai/cpp_exts/rans/rans_interface.cpp:108-173 encode, :205-268 decode, :270-343 streaming decode;
// third_party/ryg_rans/rans64.h for the 64-bit state machine), which is reached from
// EntropyModel.compress / decompress (compressai/entropy_models/entropy_models.py:165-239) through Python lists
// (`.tolist()` of every table and symbol).  Here the same coder takes plain int32 buffers, so the
// tables and symbols go straight from tensor storage to the coder.
//
// Byte format (must stay identical to the reference ext — tests/test_rans_cpu.py compares byte strings):
//   * 64-bit state x, normalisation interval [2^31, 2^63), 32-bit little-endian words, written back to front;
//   * modelled symbol with 16-bit CDF row: if x >= (2^15 * freq) << 32 ... emit low word; x = (x / freq) << 16 |
//     (x % freq) + start;
//   * values outside [0, max_value) of their table are escaped: the row's last interval, then the value as
//     4-bit "bypass" nibbles (count first, in a 15-saturating unary-by-nibble code, then the nibbles LSB first);
//   * final state flushed as two words, low first.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/masic_b200.h"

namespace {

constexpr uint32_t kScaleBits = 16;        // probability resolution of the tables (entropy_coder_precision)
constexpr uint32_t kNibbleBits = 4;        // raw ("bypass") payload granularity
constexpr uint32_t kNibbleMax = (1u << kNibbleBits) - 1;
constexpr uint64_t kLow = 1ull << 31;      // lower end of the normalisation interval

struct Item {                              // one coding step, recorded forward and replayed backward
  uint16_t start;
  uint16_t freq;                           // 0 => raw nibble `start`
};

struct Tables {
  const int32_t* cdfs; int n_tables; int pitch; const int32_t* sizes; const int32_t* offsets;
  bool ok() const { return cdfs && sizes && offsets && n_tables > 0 && pitch >= 2; }
};

inline void raw(std::vector<Item>& q, uint32_t nib) { q.push_back({static_cast<uint16_t>(nib), 0}); }

}  // namespace

struct MasicRansEncoder {
  std::vector<Item> items;
  std::vector<uint32_t> words;
};

struct MasicRansDecoder {
  std::vector<uint32_t> words;
  size_t pos = 0;
  uint64_t x = 0;
  uint32_t next() { return pos < words.size() ? words[pos++] : 0u; }
  uint32_t nibble() {
    const uint32_t v = static_cast<uint32_t>(x) & kNibbleMax;
    x >>= kNibbleBits;
    if (x < kLow) x = (x << 32) | next();
    return v;
  }
};

extern "C" int masic_rans_encoder_create(MasicRansEncoder** enc_out) {
  if (!enc_out) return MASIC_EINVAL;
  *enc_out = new MasicRansEncoder();
  return MASIC_OK;
}

extern "C" void masic_rans_encoder_destroy(MasicRansEncoder* e) { delete e; }

// BufferedRansEncoder.encode_with_indexes (rans_interface.cpp:108-173) on int32 buffers.
extern "C" int masic_rans_encoder_push(MasicRansEncoder* e, const int32_t* symbols_host, const int32_t* indexes_host,
                                       int64_t n, const int32_t* cdfs_host, int n_tables, int row_pitch,
                                       const int32_t* cdf_sizes_host, const int32_t* offsets_host) {
  const Tables t{cdfs_host, n_tables, row_pitch, cdf_sizes_host, offsets_host};
  if (!e || n < 0 || (n > 0 && (!symbols_host || !indexes_host)) || !t.ok()) return MASIC_EINVAL;
  std::vector<Item>& q = e->items;
  q.reserve(q.size() + static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ti = indexes_host[i];
    if (ti < 0 || ti >= t.n_tables) return MASIC_EINVAL;
    const int32_t top = t.sizes[ti] - 2;                 // index of the escape interval
    if (top < 0 || top + 1 >= t.pitch) return MASIC_EINVAL;
    const int32_t* row = t.cdfs + static_cast<size_t>(ti) * t.pitch;
    int32_t v = symbols_host[i] - t.offsets[ti];
    uint32_t escaped = 0;
    if (v < 0) { escaped = static_cast<uint32_t>(-2 * v - 1); v = top; }
    else if (v >= top) { escaped = static_cast<uint32_t>(2 * (v - top)); v = top; }
    q.push_back({static_cast<uint16_t>(row[v]), static_cast<uint16_t>(row[v + 1] - row[v])});
    if (q.back().freq == 0) return MASIC_EINVAL;         // empty interval: a table update() would never produce
    if (v != top) continue;
    int32_t nibbles = 0;
    while (nibbles < 8 && (escaped >> (nibbles * kNibbleBits)) != 0) ++nibbles;
    int32_t count = nibbles;                             // nibble count, 15-saturating
    while (count >= static_cast<int32_t>(kNibbleMax)) { raw(q, kNibbleMax); count -= kNibbleMax; }
    raw(q, static_cast<uint32_t>(count));
    for (int32_t j = 0; j < nibbles; ++j) raw(q, (escaped >> (j * kNibbleBits)) & kNibbleMax);
  }
  return MASIC_OK;
}

// BufferedRansEncoder.flush (rans_interface.cpp:175-200).  *data_out points into the encoder (valid until the next
// call on it); the item queue is cleared.
extern "C" int masic_rans_encoder_flush(MasicRansEncoder* e, const uint8_t** data_out, int64_t* len_out) {
  if (!e || !data_out || !len_out) return MASIC_EINVAL;
  std::vector<uint32_t>& w = e->words;
  w.assign(e->items.size() + 2, 0u);                     // <= one word per step, + the final state
  size_t p = w.size();
  uint64_t x = kLow;
  for (size_t i = e->items.size(); i-- > 0;) {
    const Item it = e->items[i];
    const uint32_t freq = it.freq ? it.freq : (1u << (kScaleBits - kNibbleBits));
    const uint64_t limit = ((kLow >> kScaleBits) << 32) * freq;
    if (x >= limit) { w[--p] = static_cast<uint32_t>(x); x >>= 32; }
    if (it.freq) x = ((x / it.freq) << kScaleBits) + (x % it.freq) + it.start;
    else         x = (x << kNibbleBits) | it.start;
  }
  w[--p] = static_cast<uint32_t>(x >> 32);
  w[--p] = static_cast<uint32_t>(x);
  e->items.clear();
  *data_out = reinterpret_cast<const uint8_t*>(w.data() + p);
  *len_out = static_cast<int64_t>((w.size() - p) * sizeof(uint32_t));
  return MASIC_OK;
}

// RansDecoder.set_stream (rans_interface.cpp:270-276): the stream is copied.
extern "C" int masic_rans_decoder_create(const uint8_t* data_host, int64_t len, MasicRansDecoder** dec_out) {
  if (!data_host || len < 8 || (len & 3) || !dec_out) return MASIC_EINVAL;
  MasicRansDecoder* d = new MasicRansDecoder();
  d->words.resize(static_cast<size_t>(len / 4));
  memcpy(d->words.data(), data_host, static_cast<size_t>(len));
  d->x = static_cast<uint64_t>(d->words[0]) | (static_cast<uint64_t>(d->words[1]) << 32);
  d->pos = 2;
  *dec_out = d;
  return MASIC_OK;
}

extern "C" void masic_rans_decoder_destroy(MasicRansDecoder* d) { delete d; }

// RansDecoder.decode_stream (rans_interface.cpp:278-343) on int32 buffers; may be called repeatedly.
extern "C" int masic_rans_decoder_decode(MasicRansDecoder* d, const int32_t* indexes_host, int64_t n,
                                         const int32_t* cdfs_host, int n_tables, int row_pitch,
                                         const int32_t* cdf_sizes_host, const int32_t* offsets_host,
                                         int32_t* symbols_host) {
  const Tables t{cdfs_host, n_tables, row_pitch, cdf_sizes_host, offsets_host};
  if (!d || n < 0 || (n > 0 && (!indexes_host || !symbols_host)) || !t.ok()) return MASIC_EINVAL;
  const uint64_t mask = (1ull << kScaleBits) - 1;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ti = indexes_host[i];
    if (ti < 0 || ti >= t.n_tables) return MASIC_EINVAL;
    const int32_t len = t.sizes[ti], top = len - 2;
    if (top < 0 || len > t.pitch) return MASIC_EINVAL;
    const int32_t* row = t.cdfs + static_cast<size_t>(ti) * t.pitch;
    const uint32_t cum = static_cast<uint32_t>(d->x & mask);
    int lo = 0, hi = len - 1;                            // largest s with row[s] <= cum (rows are increasing)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (static_cast<uint32_t>(row[mid]) <= cum) lo = mid; else hi = mid; }
    const uint32_t start = static_cast<uint32_t>(row[lo]), freq = static_cast<uint32_t>(row[lo + 1]) - start;
    d->x = freq * (d->x >> kScaleBits) + (d->x & mask) - start;
    if (d->x < kLow) d->x = (d->x << 32) | d->next();
    int32_t v = lo;
    if (v == top) {
      uint32_t nib = d->nibble();
      int32_t nibbles = static_cast<int32_t>(nib);
      while (nib == kNibbleMax) { nib = d->nibble(); nibbles += static_cast<int32_t>(nib); }
      uint32_t escaped = 0;
      for (int32_t j = 0; j < nibbles; ++j) {
        const uint32_t b = d->nibble();
        if (j < 8) escaped |= b << (j * kNibbleBits);
      }
      v = static_cast<int32_t>(escaped >> 1);
      v = (escaped & 1u) ? -v - 1 : v + top;
    }
    symbols_host[i] = v + t.offsets[ti];
  }
  return MASIC_OK;
}
