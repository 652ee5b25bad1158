// wgrad_tc.cu — weight gradients of the conv / transposed-conv layers on the tcgen05 tensor cores
// (training step, BASELINE.json configs[4]; reference: loss.backward() in
// coremasic/mywork/newtrain_codec_real.py:134, i.e. ATen's cudnn/mkldnn convolution_backward_weight
// for every conv()/deconv() of compressai/models/utils.py:128-146).  sm_100a only.
//
//   dW[cl][ch][ky][kx] = sum_{n, p}  LO[n, p, cl] * HI[n, s*p + k - pad, ch]
//
//   conv   (W is (Cout, Cin, k, k)):   LO = dL/d(out),  HI = layer input     -> dW in Conv2d layout
//   deconv (W is (Cin, Cout, k, k)):   LO = layer input, HI = dL/d(out)      -> dW in ConvTranspose2d layout
//
// Both operands are the NHWC bf16 activation / gradient buffers themselves: the reduction runs over
// PIXELS, so the channel-contiguous NHWC rows are "MN-major" UMMA operands — a 16x8-pixel tile that TMA
// drops into shared memory with SWIZZLE_128B (one 128-B line of 64 channels per pixel) is read by
// tcgen05.mma directly, no transposed copies.  The HI operand uses the forward kernel's strips: one
// (16 + taps - 1)-row strip per (kx, row phase) serves every vertical tap through a row-shifted
// descriptor; stride-2 layers read the phase-split 5-D view of the same memory.
//
// Work decomposition: a job TYPE = (128-channel LO tile, one HI strip (or two HI channel tiles for 1x1),
// <= 4 accumulators of 128 TMEM columns); the CTAs of a type split the pixel tiles and accumulate in TMEM
// across their whole range; each CTA writes one fp32 partial, a second kernel adds the partials in a fixed
// order (deterministic) into the torch-layout gradient.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/masic_b200.h"
#include "ptx.cuh"

namespace masic {

constexpr int WG_TILE_W = 8, WG_TILE_H = 16;
constexpr int WG_THREADS = 256;      // warp 0: TMA producer, 1: MMA issuer, 2: TMEM alloc, 4-7: epilogue
constexpr int WG_MAX_ACC = 4;
constexpr int WG_MAX_STAGES = 4;

struct WgType {                       // 48 bytes, lives in global memory
  int16_t lo_c0;                      // first channel of the LO tile (inside the LO view's channel dim)
  int16_t n_b;                        // HI strips loaded per tile (1, or 2 for 1x1 layers)
  int16_t b_c0[2], b_dx[2], b_p2[2], b_dy[2];
  int16_t n_acc;
  int8_t acc_b[WG_MAX_ACC];           // which HI strip feeds accumulator j
  int8_t acc_row[WG_MAX_ACC];         // row shift of the tap inside the strip
  int16_t pad;
};
struct WgCta { int type, tile_begin, tile_end, pad; };
struct WgItem {                       // one (type, accumulator): where its sum goes
  int cta_begin, cta_end, acc, cl0, ch0, tap, pad0, pad1;
};

struct WgParams {
  CUtensorMap tmL, tmH;
  const WgType* types;
  const WgCta* ctas;
  float* partial;                     // [n_ctas][WG_MAX_ACC][128][n_cols]
  int tiles_x, tiles_y;
  int n_cols;                         // MMA N (64 or 128)
  int l_blk_bytes, h_blk_bytes;       // one 64-channel block of the LO tile / of a HI strip
  int stage_bytes, n_stages;
  uint32_t idesc;
};

// MN-major, SWIZZLE_128B operand: 64 MN elements (128 B) contiguous per K index, 8 K indices per 1024-B atom,
// next 8 K indices `sbo` bytes on, next 64 MN elements `lbo` bytes on.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sMisc = smem_base + p.n_stages * p.stage_bytes;
  // misc: full[4] @0, empty[4] @32, acc_full @64, tmem ptr @72
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem_gen + p.n_stages * p.stage_bytes + 72);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const WgCta me = p.ctas[blockIdx.x];
  const WgType ty = p.types[me.type];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmL);
    tma_prefetch_desc(&p.tmH);
    for (int i = 0; i < WG_MAX_STAGES; ++i) {
      mbar_init(sMisc + 8 * i, 1);
      mbar_init(sMisc + 32 + 8 * i, 1);
    }
    mbar_init(sMisc + 64, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(sMisc + 72, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int n_lblk = 2;                                  // LO tile = 128 channels = two 64-channel blocks
  const int n_hblk = p.n_cols / 64;
  const uint32_t l_bytes = n_lblk * p.l_blk_bytes;
  const uint32_t h_bytes = n_hblk * p.h_blk_bytes;

  if (warp == 0) {
    // ===================== producer =====================
    uint32_t st = 0, ph = 0;
    for (int t = me.tile_begin; t < me.tile_end; ++t) {
      const int tx = t % p.tiles_x, r = t / p.tiles_x;
      const int x0 = tx * WG_TILE_W, y0 = (r % p.tiles_y) * WG_TILE_H, n = r / p.tiles_y;
      mbar_wait(sMisc + 32 + 8 * st, ph ^ 1);
      if (elect_one()) {
        const uint32_t full = sMisc + 8 * st;
        const uint32_t base = smem_base + st * p.stage_bytes;
        mbar_expect_tx(full, l_bytes + ty.n_b * h_bytes);
        for (int b = 0; b < n_lblk; ++b)
          tma_load_5d(base + b * p.l_blk_bytes, &p.tmL, full, ty.lo_c0 + 64 * b, x0, 0, y0, n);
        for (int s = 0; s < ty.n_b; ++s)
          for (int b = 0; b < n_hblk; ++b)
            tma_load_5d(base + l_bytes + s * h_bytes + b * p.h_blk_bytes, &p.tmH, full, ty.b_c0[s] + 64 * b,
                        x0 + ty.b_dx[s], ty.b_p2[s], y0 + ty.b_dy[s], n);
      }
      __syncwarp();
      if (++st == (uint32_t)p.n_stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t st = 0, ph = 0;
    uint32_t accum = 0;
    for (int t = me.tile_begin; t < me.tile_end; ++t) {
      mbar_wait(sMisc + 8 * st, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_base + st * p.stage_bytes;
        // LBO = distance between the two 64-channel blocks, SBO = 1024 B between 8-pixel rows
        const uint64_t adesc0 = umma_desc_mn_sw128(base, (uint32_t)p.l_blk_bytes, 1024u);
        for (int j = 0; j < ty.n_acc; ++j) {
          const uint32_t baddr = base + l_bytes + ty.acc_b[j] * h_bytes + ty.acc_row[j] * 1024u;
          const uint64_t bdesc0 = umma_desc_mn_sw128(baddr, (uint32_t)p.h_blk_bytes, 1024u);
          const uint32_t d = tmem_base + j * 128;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)              // K = 16 pixels = two 8-pixel rows = 2048 B
            umma_bf16(d, adesc0 + ks * 128u, bdesc0 + ks * 128u, p.idesc, (ks | accum) ? 1u : 0u);
        }
        umma_commit(sMisc + 32 + 8 * st);
        if (t + 1 == me.tile_end) umma_commit(sMisc + 64);
      }
      __syncwarp();
      accum = 1u;
      if (++st == (uint32_t)p.n_stages) { st = 0; ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> fp32 partial =====================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    float* out = p.partial + (size_t)blockIdx.x * WG_MAX_ACC * 128 * p.n_cols;
    if (me.tile_end > me.tile_begin) {
      mbar_wait(sMisc + 64, 0);
      tc_fence_after();
      for (int j = 0; j < ty.n_acc; ++j) {
        float* orow = out + ((size_t)j * 128 + row) * p.n_cols;
        for (int c = 0; c < p.n_cols; c += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + j * 128 + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(orow + c)[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                  __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        }
      }
    } else {
      for (int j = 0; j < ty.n_acc; ++j)
        for (int c = 0; c < p.n_cols; ++c) out[((size_t)j * 128 + row) * p.n_cols + c] = 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dW[cl][ch][tap] (+)= sum over the CTAs of the item's type
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, const WgItem* __restrict__ items, int n_cols, int c_lo, int c_hi,
                    int ktaps, int accumulate, float* __restrict__ dw) {
  const WgItem it = items[blockIdx.x];
  const int e = blockIdx.y * blockDim.x + threadIdx.x;           // one (row, column) of the 128 x n_cols tile per thread
  if (e >= 128 * n_cols) return;
  const int r = e / n_cols, c = e - r * n_cols;
  const int cl = it.cl0 + r, ch = it.ch0 + c;
  if (cl >= c_lo || ch >= c_hi) return;
  const size_t stride = (size_t)WG_MAX_ACC * 128 * n_cols;
  const float* p = partial + ((size_t)it.cta_begin * WG_MAX_ACC + it.acc) * 128 * n_cols + e;
  const int n = it.cta_end - it.cta_begin;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;              // fixed association -> deterministic
  int b = 0;
  for (; b + 4 <= n; b += 4) {
    s0 += p[(size_t)b * stride]; s1 += p[(size_t)(b + 1) * stride];
    s2 += p[(size_t)(b + 2) * stride]; s3 += p[(size_t)(b + 3) * stride];
  }
  for (; b < n; ++b) s0 += p[(size_t)b * stride];
  const float s = (s0 + s1) + (s2 + s3);
  float* o = dw + ((size_t)cl * c_hi + ch) * ktaps + it.tap;
  *o = accumulate ? *o + s : s;
}

// The same sum for items with MANY partials (>= 32 CTAs per type).  A block owns 32 consecutive elements of the 128 x n_cols
// tile; its 8 warps each sum every 8th CTA's partial (coalesced 128-byte rows, 4 loads in flight per thread) and the
// 8 part sums are combined in a fixed order: deterministic, and 8x the memory-level parallelism of one thread per
// element (the 1x1 GDN gradients, 148 partials for 16 K elements, took 23-30 us in a 64-block grid).
constexpr int WG_RED_PARTS = 8;

__global__ void __launch_bounds__(256)
wgrad_reduce8_kernel(const float* __restrict__ partial, const WgItem* __restrict__ items, int n_cols, int c_lo, int c_hi,
                     int ktaps, int accumulate, float* __restrict__ dw) {
  __shared__ float part_sum[WG_RED_PARTS][33];
  const WgItem it = items[blockIdx.x];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int e = blockIdx.y * 32 + lane;                          // one (row, column) of the 128 x n_cols tile
  const bool in_tile = e < 128 * n_cols;
  const size_t stride = (size_t)WG_MAX_ACC * 128 * n_cols;
  const int n = it.cta_end - it.cta_begin;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;              // fixed association -> deterministic
  if (in_tile) {
    const float* p = partial + ((size_t)it.cta_begin * WG_MAX_ACC + it.acc) * 128 * n_cols + e;
    int b = part;
    for (; b + 3 * WG_RED_PARTS < n; b += 4 * WG_RED_PARTS) {
      s0 += p[(size_t)b * stride]; s1 += p[(size_t)(b + WG_RED_PARTS) * stride];
      s2 += p[(size_t)(b + 2 * WG_RED_PARTS) * stride]; s3 += p[(size_t)(b + 3 * WG_RED_PARTS) * stride];
    }
    for (; b < n; b += WG_RED_PARTS) s0 += p[(size_t)b * stride];
  }
  part_sum[part][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (part != 0 || !in_tile) return;
  const int r = e / n_cols, c = e - r * n_cols;
  const int cl = it.cl0 + r, ch = it.ch0 + c;
  if (cl >= c_lo || ch >= c_hi) return;
  const float s = ((part_sum[0][lane] + part_sum[1][lane]) + (part_sum[2][lane] + part_sum[3][lane])) +
                  ((part_sum[4][lane] + part_sum[5][lane]) + (part_sum[6][lane] + part_sum[7][lane]));
  float* o = dw + ((size_t)cl * c_hi + ch) * ktaps + it.tap;
  *o = accumulate ? *o + s : s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}
// (s*Cp, W/s, s, H/s, N) view of an NHWC bf16 buffer, box = 64 channels x 8 pixels x `rows` rows, SWIZZLE_128B
static int wg_encode_view(CUtensorMap* tm, const void* base, int n, int h, int w, int cp, int split, int rows) {
  EncodeTiledFn enc = wg_encode_fn();
  if (!enc) return MASIC_EDRIVER;
  const int s = split ? 2 : 1;
  if (h % s || w % s || cp % 8 || reinterpret_cast<uintptr_t>(base) % 16) return MASIC_EINVAL;
  cuuint64_t dims[5] = {(cuuint64_t)s * cp, (cuuint64_t)(w / s), (cuuint64_t)s, (cuuint64_t)(h / s), (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)s * cp * 2, (cuuint64_t)w * cp * 2, (cuuint64_t)s * w * cp * 2,
                           (cuuint64_t)h * w * cp * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)WG_TILE_W, 1, (cuuint32_t)rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MASIC_OK : MASIC_EDRIVER;
}

}  // namespace masic

using namespace masic;

struct MasicWgradPlan {
  WgParams kp;
  void* d_tables = nullptr;
  WgItem* d_items = nullptr;
  int n_ctas = 0, n_items = 0, smem_bytes = 0;
  int max_ctas_per_item = 0;          // partials the reduce kernel sums per output element
  int c_lo = 0, c_hi = 0, ktaps = 0, accumulate = 0;
  float* dw = nullptr;
  double flops = 0;
  int64_t ws_bytes = 0;
};

namespace {
inline int fdiv2(int t) { return t >= 0 ? t / 2 : -((-t + 1) / 2); }
}

extern "C" int masic_wgrad_plan_create(const MasicWgradDesc* dp, MasicWgradPlan** out) {
  if (!dp || !out) return MASIC_EINVAL;
  const MasicWgradDesc& d = *dp;
  *out = nullptr;
  if (d.ksize != 1 && d.ksize != 3 && d.ksize != 5) return MASIC_EINVAL;
  if (d.stride != 1 && d.stride != 2) return MASIC_EINVAL;
  if (d.c_lo < 16 || d.c_hi < 16) return MASIC_ENOSUP;      // small-channel layers: masic_wgrad_small
  const int c_hi_pad = (d.c_hi + 63) / 64 * 64;               // channels read beyond the real count are ignored
  if (d.lo_cpitch % 8 || d.hi_cpitch % 8 || d.lo_coff % 8 || d.hi_coff % 8) return MASIC_EINVAL;
  if (!d.lo || !d.hi || !d.dw || d.n <= 0 || d.h_lo <= 0 || d.w_lo <= 0) return MASIC_EINVAL;
  const int k = d.ksize, pad = k / 2, s = d.stride;
  const uint32_t mask = d.tap_mask ? d.tap_mask : 0xFFFFFFFFu;
  const int n_cols = c_hi_pad >= 128 ? 128 : 64;
  const int max_acc = WG_MAX_ACC;

  // HI strips: (c0 relative to hi_coff incl. the column phase, dx, p2, dy) and their live taps (row, tap id)
  struct SSpec { int c0, dx, p2, dy; std::vector<std::pair<int, int>> taps; };
  std::vector<SSpec> strips;
  int rows = WG_TILE_H;
  if (s == 1) {
    rows = WG_TILE_H + k - 1;
    for (int kx = 0; kx < k; ++kx) {
      SSpec sp{0, kx - pad, 0, -pad, {}};
      for (int ky = 0; ky < k; ++ky)
        if (mask & (1u << (ky * k + kx))) sp.taps.push_back({ky, ky * k + kx});
      if (!sp.taps.empty()) strips.push_back(sp);
    }
  } else {
    const int hmin = fdiv2(-pad), hmax = fdiv2(k - 1 - pad);
    rows = WG_TILE_H + (hmax - hmin);
    for (int kx = 0; kx < k; ++kx) {
      const int tx = kx - pad, hx = fdiv2(tx), px = tx - 2 * hx;
      for (int py = 0; py < 2; ++py) {
        SSpec sp{px * d.hi_cpitch, hx, py, hmin, {}};
        for (int ky = 0; ky < k; ++ky) {
          const int t = ky - pad, hy = fdiv2(t);
          if (t - 2 * hy != py) continue;
          if (mask & (1u << (ky * k + kx))) sp.taps.push_back({hy - hmin, ky * k + kx});
        }
        if (!sp.taps.empty()) strips.push_back(sp);
      }
    }
  }
  if (strips.empty()) return MASIC_EINVAL;

  const int l_blk = WG_TILE_H * 1024, h_blk = rows * 1024;
  const int n_hblk = n_cols / 64;
  const int cl_tiles = (d.c_lo + 127) / 128, ch_tiles = (c_hi_pad + n_cols - 1) / n_cols;
  const bool one_by_one = (k == 1);
  const int max_b = one_by_one ? 2 : 1;
  const int stage_bytes = 2 * l_blk + max_b * n_hblk * h_blk;
  int n_stages = (227 * 1024 - 2048) / stage_bytes;
  if (n_stages > WG_MAX_STAGES) n_stages = WG_MAX_STAGES;
  if (n_stages < 2) return MASIC_ENOSUP;

  std::vector<WgType> types;
  std::vector<std::vector<std::pair<int, int>>> type_out;   // per type, per acc: (ch0, tap)
  std::vector<int> type_cl0;
  std::vector<int> type_weight;
  for (int ct = 0; ct < cl_tiles; ++ct) {
    if (one_by_one) {
      for (int h0 = 0; h0 < ch_tiles; h0 += 2) {
        WgType t; memset(&t, 0, sizeof(t));
        t.lo_c0 = (int16_t)(d.lo_coff + ct * 128);
        t.n_b = (int16_t)((h0 + 1 < ch_tiles) ? 2 : 1);
        std::vector<std::pair<int, int>> o;
        for (int b = 0; b < t.n_b; ++b) {
          t.b_c0[b] = (int16_t)(d.hi_coff + (h0 + b) * n_cols);
          t.b_dx[b] = 0; t.b_p2[b] = 0; t.b_dy[b] = 0;
          t.acc_b[b] = (int8_t)b; t.acc_row[b] = 0;
          o.push_back({(h0 + b) * n_cols, 0});
        }
        t.n_acc = t.n_b;
        types.push_back(t); type_out.push_back(o); type_cl0.push_back(ct * 128); type_weight.push_back(t.n_acc);
      }
    } else {
      for (int ht = 0; ht < ch_tiles; ++ht)
        for (const auto& sp : strips)
          for (size_t a0 = 0; a0 < sp.taps.size(); a0 += max_acc) {
            WgType t; memset(&t, 0, sizeof(t));
            t.lo_c0 = (int16_t)(d.lo_coff + ct * 128);
            t.n_b = 1;
            t.b_c0[0] = (int16_t)(d.hi_coff + sp.c0 + ht * n_cols);
            t.b_dx[0] = (int16_t)sp.dx; t.b_p2[0] = (int16_t)sp.p2; t.b_dy[0] = (int16_t)sp.dy;
            std::vector<std::pair<int, int>> o;
            int na = 0;
            for (size_t a = a0; a < sp.taps.size() && na < max_acc; ++a, ++na) {
              t.acc_b[na] = 0; t.acc_row[na] = (int8_t)sp.taps[a].first;
              o.push_back({ht * n_cols, sp.taps[a].second});
            }
            t.n_acc = (int16_t)na;
            types.push_back(t); type_out.push_back(o); type_cl0.push_back(ct * 128); type_weight.push_back(na);
          }
    }
  }

  // CTAs: every type gets a share of ~2 waves of 148 proportional to its accumulator count, capped by the tiles
  const int tiles_x = (d.w_lo + WG_TILE_W - 1) / WG_TILE_W, tiles_y = (d.h_lo + WG_TILE_H - 1) / WG_TILE_H;
  const int n_tiles = tiles_x * tiles_y * d.n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long wsum = 0;
  for (int w : type_weight) wsum += w;
  const long target = (long)types.size() >= sms ? (long)types.size() : ((long)types.size() > sms / 2 ? 2L * sms : sms);
  std::vector<WgCta> ctas;
  std::vector<WgItem> items;
  for (size_t ti = 0; ti < types.size(); ++ti) {
    long share = (target * type_weight[ti] + wsum / 2) / wsum;
    if (share < 1) share = 1;
    if (share > n_tiles) share = n_tiles;
    const int c0 = (int)ctas.size();
    for (long c = 0; c < share; ++c) {
      WgCta e; e.type = (int)ti; e.pad = 0;
      e.tile_begin = (int)((long)n_tiles * c / share);
      e.tile_end = (int)((long)n_tiles * (c + 1) / share);
      ctas.push_back(e);
    }
    for (int a = 0; a < types[ti].n_acc; ++a) {
      WgItem it; memset(&it, 0, sizeof(it));
      it.cta_begin = c0; it.cta_end = (int)ctas.size(); it.acc = a;
      it.cl0 = type_cl0[ti]; it.ch0 = type_out[ti][a].first; it.tap = type_out[ti][a].second;
      items.push_back(it);
    }
  }

  MasicWgradPlan* pl = new MasicWgradPlan();
  WgParams& kp = pl->kp;
  memset(&kp, 0, sizeof(kp));
  int rc = wg_encode_view(&kp.tmL, d.lo, d.n, d.h_lo, d.w_lo, d.lo_cpitch, 0, WG_TILE_H);
  if (!rc) rc = wg_encode_view(&kp.tmH, d.hi, d.n, d.h_lo * s, d.w_lo * s, d.hi_cpitch, s == 2, rows);
  if (rc) { delete pl; return rc; }
  const size_t tb = types.size() * sizeof(WgType), cb = ctas.size() * sizeof(WgCta), ib = items.size() * sizeof(WgItem);
  const size_t tb_al = (tb + 255) & ~size_t(255), cb_al = (cb + 255) & ~size_t(255);
  cudaError_t ce = cudaMalloc(&pl->d_tables, tb_al + cb_al + ib);
  if (ce != cudaSuccess) { delete pl; return (int)ce; }
  uint8_t* base = static_cast<uint8_t*>(pl->d_tables);
  cudaMemcpy(base, types.data(), tb, cudaMemcpyHostToDevice);
  cudaMemcpy(base + tb_al, ctas.data(), cb, cudaMemcpyHostToDevice);
  cudaMemcpy(base + tb_al + cb_al, items.data(), ib, cudaMemcpyHostToDevice);
  kp.types = reinterpret_cast<const WgType*>(base);
  kp.ctas = reinterpret_cast<const WgCta*>(base + tb_al);
  pl->d_items = reinterpret_cast<WgItem*>(base + tb_al + cb_al);
  kp.tiles_x = tiles_x; kp.tiles_y = tiles_y;
  kp.n_cols = n_cols;
  kp.l_blk_bytes = l_blk; kp.h_blk_bytes = h_blk;
  kp.stage_bytes = stage_bytes; kp.n_stages = n_stages;
  // kind::f16, bf16 x bf16 -> fp32, A and B both MN-major (bits 15, 16), M = 128, N = n_cols
  kp.idesc = umma_idesc_bf16(n_cols) | (1u << 15) | (1u << 16);
  pl->n_ctas = (int)ctas.size(); pl->n_items = (int)items.size();
  for (const WgItem& it : items) pl->max_ctas_per_item = std::max(pl->max_ctas_per_item, it.cta_end - it.cta_begin);
  pl->smem_bytes = n_stages * stage_bytes + 2048;
  pl->c_lo = d.c_lo; pl->c_hi = d.c_hi; pl->ktaps = k * k; pl->accumulate = d.accumulate; pl->dw = d.dw;
  pl->ws_bytes = (int64_t)pl->n_ctas * WG_MAX_ACC * 128 * n_cols * 4;
  int live = 0;
  for (int i = 0; i < k * k; ++i) live += (mask >> i) & 1;
  pl->flops = 2.0 * d.n * d.h_lo * d.w_lo * (double)live * d.c_lo * d.c_hi;
  static bool attr_set = false;
  if (!attr_set) {
    ce = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) { cudaFree(pl->d_tables); delete pl; return (int)ce; }
    attr_set = true;
  }
  *out = pl;
  return MASIC_OK;
}

extern "C" int64_t masic_wgrad_plan_workspace_bytes(const MasicWgradPlan* pl) { return pl ? pl->ws_bytes : 0; }

extern "C" int masic_wgrad_plan_info(const MasicWgradPlan* pl, double* flops, int* n_ctas) {
  if (!pl) return MASIC_EINVAL;
  if (flops) *flops = pl->flops;
  if (n_ctas) *n_ctas = pl->n_ctas;
  return MASIC_OK;
}

extern "C" int masic_wgrad_plan_launch(const MasicWgradPlan* pl, void* workspace, void* stream) {
  if (!pl || !workspace) return MASIC_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  WgParams kp = pl->kp;
  kp.partial = static_cast<float*>(workspace);
  int smem = pl->smem_bytes < 120 * 1024 ? 120 * 1024 : pl->smem_bytes;     // 1 CTA / SM: 512 TMEM columns each
  wgrad_tc_kernel<<<pl->n_ctas, WG_THREADS, smem, s>>>(kp);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return (int)ce;
  if (pl->max_ctas_per_item >= 32) {
    dim3 grid(pl->n_items, (128 * kp.n_cols + 31) / 32);
    wgrad_reduce8_kernel<<<grid, 256, 0, s>>>(kp.partial, pl->d_items, kp.n_cols, pl->c_lo, pl->c_hi, pl->ktaps,
                                              pl->accumulate, pl->dw);
  } else {
    dim3 grid(pl->n_items, (128 * kp.n_cols + 255) / 256);
    wgrad_reduce_kernel<<<grid, 256, 0, s>>>(kp.partial, pl->d_items, kp.n_cols, pl->c_lo, pl->c_hi, pl->ktaps,
                                             pl->accumulate, pl->dw);
  }
  return (int)cudaGetLastError();
}

extern "C" void masic_wgrad_plan_destroy(MasicWgradPlan* pl) {
  if (!pl) return;
  if (pl->d_tables) cudaFree(pl->d_tables);
  delete pl;
}
